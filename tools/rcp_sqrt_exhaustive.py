#!/usr/bin/env python
"""Exhaustive check of the branch-free reciprocal / square root (codegen._RCP_SQRT_HELPERS) against __frcp_rn /
__fsqrt_rn over all 2^32 float32 bit patterns: wherever the range flag stays clear the results must be
bit-identical.  Prints the number of operands inside the range and the number of mismatches (must be 0)."""
import ctypes
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from graphax_b200 import runtime
from graphax_b200.codegen import _RCP_SQRT_HELPERS

SRC = _RCP_SQRT_HELPERS + r"""
extern "C" __global__ void gx_check(unsigned long long* out) {
  unsigned long long safe_r = 0, bad_r = 0, safe_s = 0, bad_s = 0;
  const unsigned long long n = 1ull << 32;
  for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n;
       i += (unsigned long long)gridDim.x * blockDim.x) {
    const float v = __uint_as_float((unsigned)i);
    unsigned u = 0;
    const float fr = gx_rcp<false>(v, u);
    if (!u) { ++safe_r; if (__float_as_uint(fr) != __float_as_uint(__frcp_rn(v))) { ++bad_r; out[4] = i; } }
    u = 0;
    const float fs = gx_sqrt<false>(v, u);
    if (!u) { ++safe_s; if (__float_as_uint(fs) != __float_as_uint(__fsqrt_rn(v))) { ++bad_s; out[5] = i; } }
  }
  atomicAdd(out + 0, safe_r); atomicAdd(out + 1, bad_r); atomicAdd(out + 2, safe_s); atomicAdd(out + 3, bad_s);
}
"""
lib = runtime.load_library()
h = ctypes.c_void_p()
rc = lib.gx_jit_compile(SRC.encode(), b"gx_check", 0, ctypes.byref(h))
assert rc == 0, lib.gx_last_error().decode()
out = torch.zeros(8, dtype=torch.int64, device="cuda")
ptr = ctypes.c_void_p(out.data_ptr())
argv = (ctypes.c_void_p * 1)(ctypes.cast(ctypes.pointer(ptr), ctypes.c_void_p))
rc = lib.gx_jit_launch(h, 148 * 16, 1, 256, argv, torch.cuda.current_stream().cuda_stream)
assert rc == 0, lib.gx_last_error().decode()
torch.cuda.synchronize()
r = out.cpu().tolist()
print(f"rcp : {r[0]} operands in range ({r[0] / 2**32:.1%} of all bit patterns), mismatches {r[1]}" + (f" (e.g. 0x{r[4]:08x})" if r[1] else ""))
print(f"sqrt: {r[2]} operands in range ({r[2] / 2**32:.1%}), mismatches {r[3]}" + (f" (e.g. 0x{r[5]:08x})" if r[3] else ""))
sys.exit(0 if r[1] == 0 and r[3] == 0 else 1)
