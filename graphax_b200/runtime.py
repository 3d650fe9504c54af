"""ctypes binding of the C ABI (include/graphax_b200.h) + device-side plan handles.

PyTorch is used for what the contract allows - device memory and streams; the
compute is the library's CUDA kernels.  There is no CPU fallback: if the shared
library has not been built, or no CUDA device is present, calls fail loudly.
"""
from __future__ import annotations

import ctypes
import os
import threading
import warnings
from pathlib import Path
from typing import Dict, List, Optional, Sequence, Tuple

from .plan import Plan

# GX_LIB: another build of the library (A/B measurements of ahead-of-time kernels); the default is the in-tree one
LIB_PATH = Path(os.environ.get("GX_LIB") or Path(__file__).resolve().parent / "libgraphax_b200.so")
GX_MAX_ARRAYS = 32
KERNEL_INTERPRETER, KERNEL_AOT, KERNEL_JIT = 0, 1, 2
KERNEL_NAMES = {0: "interpreter", 1: "aot", 2: "jit"}


class GxInfo(ctypes.Structure):
    _fields_ = [(n, ctypes.c_uint32) for n in (
        "abi_version", "n_in_arrays", "n_out_arrays", "n_in_scalars", "n_rows", "n_cols", "n_ops", "n_slots",
        "has_primal", "kernel_kind", "threads", "smem_bytes", "ctas_per_sm", "regs_per_thread")] + [
        ("in_sizes", ctypes.c_uint32 * GX_MAX_ARRAYS), ("out_sizes", ctypes.c_uint32 * GX_MAX_ARRAYS),
        ("plan_key", ctypes.c_uint64), ("launches", ctypes.c_uint64)]


EXPORTS = ["gx_plan_create", "gx_plan_destroy", "gx_plan_info", "gx_plan_specialize", "gx_plan_specialize_opts",
           "gx_plan_select_kernel",
           "gx_jacve_launch", "gx_jacve_host", "gx_last_error", "gx_abi_version", "gx_aot_count", "gx_aot_key",
           "gx_hash_bytes", "gx_gemm_bf16_tn", "gx_gemm_bf16", "gx_gemm_bf16_ld", "gx_transpose_bf16", "gx_gemm_last_error", "gx_gemm_set_pair", "gx_gemm_set_trace", "gx_gemm_zero_with_next_launch", "gx_gemm_bf16_dual_ld",
           "gx_jit_compile", "gx_jit_destroy", "gx_jit_launch", "gx_colsum_f32", "gx_gemm_fused_launch",
           "gx_split3_bf16", "gx_gemm_bf16x3_ld"]

_lib = None
_lock = threading.Lock()


class GraphaxB200Error(RuntimeError):
    pass


def load_library() -> ctypes.CDLL:
    """dlopen the engine. Raises if it was never built - nothing else can run the hot path."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not LIB_PATH.exists():
            raise GraphaxB200Error(
                f"{LIB_PATH} is missing: build the CUDA extension first "
                "(python -c 'import __graft_entry__ as g; g.build()'). graphax_b200 has no CPU fallback.")
        lib = ctypes.CDLL(str(LIB_PATH))
        vp, i64, sz = ctypes.c_void_p, ctypes.c_int64, ctypes.c_size_t
        lib.gx_plan_create.argtypes = [vp, sz, ctypes.c_int, ctypes.POINTER(vp)]
        lib.gx_plan_create.restype = ctypes.c_int
        lib.gx_plan_destroy.argtypes = [vp]
        lib.gx_plan_destroy.restype = None
        lib.gx_plan_info.argtypes = [vp, ctypes.POINTER(GxInfo)]
        lib.gx_plan_info.restype = ctypes.c_int
        lib.gx_plan_specialize.argtypes = [vp, ctypes.c_char_p, sz, ctypes.c_char_p]
        lib.gx_plan_specialize.restype = ctypes.c_int
        lib.gx_plan_specialize_opts.argtypes = [vp, ctypes.c_char_p, sz, ctypes.c_char_p, ctypes.c_int, ctypes.c_int,
                                                ctypes.c_int, ctypes.c_int]
        lib.gx_plan_specialize_opts.restype = ctypes.c_int
        lib.gx_plan_select_kernel.argtypes = [vp, ctypes.c_int]
        lib.gx_plan_select_kernel.restype = ctypes.c_int
        lib.gx_jacve_launch.argtypes = [vp, i64, ctypes.POINTER(vp), vp, vp, vp]
        lib.gx_jacve_launch.restype = ctypes.c_int
        lib.gx_jacve_host.argtypes = [vp, i64, ctypes.POINTER(vp), vp, vp, ctypes.c_int]
        lib.gx_jacve_host.restype = ctypes.c_int
        lib.gx_last_error.restype = ctypes.c_char_p
        lib.gx_abi_version.restype = ctypes.c_int
        lib.gx_aot_count.restype = ctypes.c_int
        lib.gx_aot_key.argtypes = [ctypes.c_int]
        lib.gx_aot_key.restype = ctypes.c_uint64
        lib.gx_hash_bytes.argtypes = [vp, sz]
        lib.gx_hash_bytes.restype = ctypes.c_uint64
        lib.gx_gemm_bf16_tn.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, vp, vp, ctypes.c_int, ctypes.c_int, vp]
        lib.gx_gemm_bf16_tn.restype = ctypes.c_int
        lib.gx_gemm_bf16.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, ctypes.c_int, vp, ctypes.c_int, vp,
                                     ctypes.c_int, ctypes.c_int, vp]
        lib.gx_gemm_bf16.restype = ctypes.c_int
        lib.gx_transpose_bf16.argtypes = [ctypes.c_int, ctypes.c_int, vp, vp, vp]
        lib.gx_transpose_bf16.restype = ctypes.c_int
        lib.gx_gemm_last_error.restype = ctypes.c_char_p
        lib.gx_gemm_set_pair.argtypes = [ctypes.c_int]
        lib.gx_gemm_set_pair.restype = ctypes.c_int
        lib.gx_gemm_set_trace.argtypes = [vp, ctypes.c_int, ctypes.c_int]
        lib.gx_gemm_set_trace.restype = ctypes.c_int
        lib.gx_gemm_zero_with_next_launch.argtypes = [vp, ctypes.c_longlong]
        lib.gx_gemm_zero_with_next_launch.restype = ctypes.c_int
        ip, pp = ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_void_p)
        lib.gx_gemm_bf16_dual_ld.argtypes = [ip, ip, ip, pp, ip, ip, pp, ip, ip, pp, ip, ip, ip, vp]
        lib.gx_gemm_bf16_dual_ld.restype = ctypes.c_int
        lib.gx_jit_compile.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int, ctypes.POINTER(vp)]
        lib.gx_jit_compile.restype = ctypes.c_int
        lib.gx_jit_destroy.argtypes = [vp]
        lib.gx_jit_destroy.restype = None
        lib.gx_jit_launch.argtypes = [vp, ctypes.c_uint, ctypes.c_uint, ctypes.c_uint, ctypes.POINTER(vp), vp]
        lib.gx_jit_launch.restype = ctypes.c_int
        lib.gx_gemm_fused_launch.argtypes = [vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, ctypes.c_int, ctypes.c_int,
                                             vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, ctypes.c_int, vp]
        lib.gx_gemm_bf16_ld.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, ctypes.c_int, ctypes.c_int, vp,
                                        ctypes.c_int, ctypes.c_int, vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp]
        lib.gx_gemm_bf16_ld.restype = ctypes.c_int
        lib.gx_gemm_fused_launch.restype = ctypes.c_int
        lib.gx_colsum_f32.argtypes = [vp, vp, ctypes.c_int, ctypes.c_int, vp]
        lib.gx_colsum_f32.restype = ctypes.c_int
        lib.gx_split3_bf16.argtypes = [vp, vp, vp, vp, ctypes.c_longlong, vp]
        lib.gx_split3_bf16.restype = ctypes.c_int
        lib.gx_gemm_bf16x3_ld.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(vp), ctypes.c_int,
                                          ctypes.c_int, ctypes.POINTER(vp), ctypes.c_int, ctypes.c_int, vp, ctypes.c_int,
                                          ctypes.c_int, vp]
        lib.gx_gemm_bf16x3_ld.restype = ctypes.c_int
        _lib = lib
        return lib


def _check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load_library().gx_last_error().decode(errors="replace")
        raise GraphaxB200Error(f"{what} failed ({rc}): {msg}")


def aot_keys() -> List[int]:
    lib = load_library()
    return [lib.gx_aot_key(i) for i in range(lib.gx_aot_count())]


def jit_max_ops() -> int:
    """Longest plan that is compiled at run time (straight-line code: NVRTC time grows faster than linearly
    with the body: 0.3 s for 1 k operations, 5 s for 6 k, ~25 s for 23 k); longer plans run on the generic interpreter kernel.  ``GX_JIT_MAX_OPS`` overrides."""
    return int(os.environ.get("GX_JIT_MAX_OPS", "30000"))


class DevicePlan:
    """A plan uploaded to one GPU (wraps ``gx_plan*``)."""

    def __init__(self, plan: Plan, device: int = 0, jit: bool = True):
        self.plan = plan
        self.device = device
        self._lib = load_library()
        self._h = ctypes.c_void_p()
        buf = ctypes.create_string_buffer(plan.blob, len(plan.blob))
        _check(self._lib.gx_plan_create(buf, len(plan.blob), device, ctypes.byref(self._h)), "gx_plan_create")
        if self.info().kernel_kind == KERNEL_INTERPRETER and jit and plan.stats.get("n_ops", 0) <= jit_max_ops():
            try:
                self.specialize()
            except ValueError:             # staging does not fit in shared memory: interpreter kernel by design
                pass
            except GraphaxB200Error as e:  # stay on the (GPU) interpreter kernel, but say so
                warnings.warn(f"plan {plan.key}: run-time specialisation unavailable, using the generic "
                              f"interpreter kernel: {e}")

    def specialize(self, arch: str = "sm_100a", threads: int = 0, min_blocks: int = 0, out_stages: int = 0,
                   samples_per_thread: int = 0) -> None:
        """Compile this plan's straight-line kernel with NVRTC and attach it (threads/min_blocks 0 = defaults)."""
        from .codegen import emit_source, launch_shape
        d_threads, d_blocks, d_stages, d_spt = launch_shape(self.plan)
        src = emit_source(self.plan, nvrtc=True).encode()
        _check(self._lib.gx_plan_specialize_opts(self._h, src, len(src), arch.encode(), threads or d_threads,
                                                 min_blocks or d_blocks, out_stages or d_stages,
                                                 samples_per_thread or d_spt), "gx_plan_specialize")

    def select_kernel(self, kind: int) -> None:
        _check(self._lib.gx_plan_select_kernel(self._h, kind), "gx_plan_select_kernel")

    def info(self) -> GxInfo:
        info = GxInfo()
        _check(self._lib.gx_plan_info(self._h, ctypes.byref(info)), "gx_plan_info")
        return info

    @property
    def kernel(self) -> str:
        return KERNEL_NAMES[self.info().kernel_kind]

    def launch(self, batch: int, in_ptrs: Sequence[int], jac_ptr: int, primal_ptr: Optional[int], stream: int) -> None:
        arr = (ctypes.c_void_p * len(in_ptrs))(*in_ptrs)
        _check(self._lib.gx_jacve_launch(self._h, batch, arr, jac_ptr, primal_ptr, stream), "gx_jacve_launch")

    def run_host(self, batch: int, in_ptrs: Sequence[int], jac_ptr: int, primal_ptr: Optional[int],
                 pinned: bool) -> None:
        arr = (ctypes.c_void_p * len(in_ptrs))(*in_ptrs)
        _check(self._lib.gx_jacve_host(self._h, batch, arr, jac_ptr, primal_ptr, 1 if pinned else 0),
               "gx_jacve_host")

    def close(self) -> None:
        if self._h:
            self._lib.gx_plan_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def gemm_bf16_tn(a, b, out=None, out_dtype=None, k_splits: int = 1, a_mn: bool = False, b_mn: bool = False,
                 zero: bool = True):
    """``A @ B.T`` on the tcgen05 tensor cores (bf16 operands, fp32 accumulate); torch tensors in, torch tensor
    out.  ``a`` is A[M,K] (or, with ``a_mn``, A stored as [K,M]); ``b`` is B[N,K] (with ``b_mn`` stored [K,N]).
    ``k_splits > 1`` splits the contraction axis over CTAs (fp32 output, atomics)."""
    import torch
    lib = load_library()
    assert a.is_cuda and b.is_cuda and a.dtype == b.dtype == torch.bfloat16 and a.is_contiguous() and b.is_contiguous()
    (m, k), (n, k2) = (a.shape[::-1] if a_mn else a.shape), (b.shape[::-1] if b_mn else b.shape)
    assert k == k2
    out_dtype = out_dtype or (out.dtype if out is not None else torch.float32)
    if out is None:
        out = (torch.zeros if k_splits > 1 else torch.empty)((m, n), dtype=out_dtype, device=a.device)
    elif k_splits > 1 and zero:      # zero=False: the caller has already cleared the split-K accumulator
        out.zero_()
    rc = lib.gx_gemm_bf16(m, n, k, a.data_ptr(), 1 if a_mn else 0, b.data_ptr(), 1 if b_mn else 0, out.data_ptr(),
                          1 if out.dtype == torch.bfloat16 else 0, k_splits, torch.cuda.current_stream().cuda_stream)
    if rc != 0:
        raise GraphaxB200Error(f"gx_gemm_bf16_tn failed ({rc}): {lib.gx_gemm_last_error().decode()}")
    return out


def transpose_bf16(x, out=None):
    import torch
    lib = load_library()
    assert x.is_cuda and x.dtype == torch.bfloat16 and x.is_contiguous() and x.ndim == 2
    rows, cols = x.shape
    if out is None:
        out = torch.empty((cols, rows), dtype=torch.bfloat16, device=x.device)
    rc = lib.gx_transpose_bf16(rows, cols, x.data_ptr(), out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    if rc != 0:
        raise GraphaxB200Error(f"gx_transpose_bf16 failed ({rc}): {lib.gx_gemm_last_error().decode()}")
    return out
