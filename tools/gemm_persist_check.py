#!/usr/bin/env python
"""Persistent GEMM variant (grids of several waves) against torch.matmul, all operand layouts, fp32 and bf16 out."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from graphax_b200 import runtime as rt
torch.manual_seed(0)
ok = True
for (m, n, k) in [(4096, 4096, 1024), (8192, 2048, 576), (2560, 3840, 2048), (4096, 3968, 512)]:
    for a_mn in (False, True):
        for b_mn in (False, True):
            A = (torch.randn(m, k, device="cuda") * 0.1).bfloat16()
            B = (torch.randn(n, k, device="cuda") * 0.1).bfloat16()
            ref = A.float() @ B.float().T
            a = A.T.contiguous() if a_mn else A
            b = B.T.contiguous() if b_mn else B
            for dt in (torch.float32, torch.bfloat16):
                c = rt.gemm_bf16_tn(a, b, out_dtype=dt, a_mn=a_mn, b_mn=b_mn)
                torch.cuda.synchronize()
                err = float((c.float() - ref).abs().max() / ref.abs().max())
                good = err < (1e-5 if dt == torch.float32 else 2 ** -8)
                ok &= good
                print(f"M{m} N{n} K{k} a_mn={a_mn} b_mn={b_mn} out={dt}: rel err {err:.2e} {'ok' if good else 'FAIL'}", flush=True)
print("persistent gemm check", "passed" if ok else "FAILED")
sys.exit(0 if ok else 1)
