"""Shared test helpers (the only place besides smoke()/bench.py that touches oracle/)."""
from __future__ import annotations

import json
from pathlib import Path
from typing import List, Sequence

import numpy as np

GOLDEN = Path(__file__).resolve().parent / "golden"


def pack_blocks(jac, batch: int, in_sizes: Sequence[int]) -> np.ndarray:
    """oracle / API block pytree -> packed [B, rows, cols] tile."""
    rows = []
    for row in jac:
        row = row if isinstance(row, (tuple, list)) else [row]
        blocks = []
        for b, n_in in zip(row, in_sizes):
            b = b.detach().cpu().numpy() if hasattr(b, "detach") else np.asarray(b)
            blocks.append(b.reshape(batch, -1, n_in))
        rows.append(np.concatenate(blocks, axis=2))
    return np.concatenate(rows, axis=1)


def bits(a: np.ndarray) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def parity_bound(ref64: np.ndarray) -> np.ndarray:
    """|dJ| <= 1e-5 * max|J_sample| + 1e-5 * |J|  (BASELINE.md section 5)."""
    mx = np.abs(ref64).max(axis=(1, 2), keepdims=True)
    return 1e-5 * mx + 1e-5 * np.abs(ref64)


def load_random_g():
    from graphax_b200.tracer import jnp
    d = json.loads((GOLDEN / "random_g.json").read_text())

    def g(*xs):
        env = dict(zip(d["args"], xs))
        for dst, fn, args in d["ops"]:
            env[dst] = jnp.ones(()) if fn == "ones" else getattr(jnp, fn)(*[env[a] for a in args])
        return tuple(env[o] for o in d["outputs"])
    return g, d


def run_engine(plan, inputs: List[np.ndarray], kernel: str = "auto", device: int = 0, with_primal: bool = True,
               samples_per_thread: int = 0):
    """Launch through the C ABI on device buffers; returns (jac [B,r,c], primal [B,r], kernel name)."""
    import torch
    from graphax_b200 import runtime
    dp = runtime.DevicePlan(plan, device, jit=False)
    if kernel == "interpreter":
        dp.select_kernel(runtime.KERNEL_INTERPRETER)
    elif kernel == "jit":
        dp.specialize(samples_per_thread=samples_per_thread)
        dp.select_kernel(runtime.KERNEL_JIT)
    elif kernel == "aot":
        dp.select_kernel(runtime.KERNEL_AOT)
    batch = len(inputs[0])
    dev = [torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).cuda(device) for x in inputs]
    jac = torch.full((batch, plan.n_rows, plan.n_cols), float("nan"), device=f"cuda:{device}")
    primal = torch.full((batch, plan.n_rows), float("nan"), device=f"cuda:{device}") if with_primal else None
    dp.launch(batch, [t.data_ptr() for t in dev], jac.data_ptr(),
              primal.data_ptr() if primal is not None else None, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    name = dp.kernel
    out = jac.cpu().numpy(), (primal.cpu().numpy() if primal is not None else None), name
    dp.close()
    return out
