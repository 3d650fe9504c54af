"""ORACLE - ctypes loader for oracle/libgxoracle.so (the C plan replay). Test infrastructure only."""
from __future__ import annotations

import ctypes
import os
import subprocess
from pathlib import Path
from typing import Optional, Sequence, Tuple

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB = None


def build() -> Path:
    subprocess.run(["make", "-C", str(_HERE), "-s"], check=True)
    return _HERE / "libgxoracle.so"


def lib():
    global _LIB
    if _LIB is None:
        path = _HERE / "libgxoracle.so"
        if not path.exists():
            build()
        _LIB = ctypes.CDLL(str(path))
        _LIB.gxo_replay.restype = ctypes.c_int
        _LIB.gxo_replay.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int64,
                                    ctypes.POINTER(ctypes.c_void_p), ctypes.c_void_p, ctypes.c_void_p,
                                    ctypes.c_int]
        _LIB.gxo_opcode_name.restype = ctypes.c_char_p
    return _LIB


def opcode_names():
    l = lib()
    return [l.gxo_opcode_name(i).decode() for i in range(l.gxo_opcode_count())]


def replay(blob: bytes, inputs: Sequence[np.ndarray], n_rows: int, n_cols: int, has_primal: bool = True,
           n_threads: Optional[int] = None) -> Tuple[np.ndarray, Optional[np.ndarray]]:
    """Run the plan on the CPU. inputs[i] is [batch, size_i] float32; returns (jac [B,rows,cols], primal)."""
    ins = [np.ascontiguousarray(x, dtype=np.float32).reshape(len(x), -1) for x in inputs]
    batch = ins[0].shape[0]
    jac = np.empty((batch, n_rows, n_cols), dtype=np.float32)
    primal = np.empty((batch, n_rows), dtype=np.float32) if has_primal else None
    ptrs = (ctypes.c_void_p * len(ins))(*[x.ctypes.data for x in ins])
    if n_threads is None:
        n_threads = len(os.sched_getaffinity(0))
    buf = ctypes.create_string_buffer(blob, len(blob))
    rc = lib().gxo_replay(buf, len(blob), batch, ptrs, jac.ctypes.data,
                          primal.ctypes.data if primal is not None else None, n_threads)
    if rc != 0:
        raise ValueError(f"gxo_replay rejected the plan blob (code {rc})")
    return jac, primal
