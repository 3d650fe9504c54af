#!/usr/bin/env python
"""Correctness + speed of the tcgen05 GEMM against torch.matmul (cuBLAS) on the shapes of config 5."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from graphax_b200 import runtime as rt

torch.manual_seed(0)
shapes = [(128, 128, 64, 1), (128, 128, 512, 1), (256, 384, 1024, 1), (4096, 1024, 512, 1), (4096, 512, 1024, 1),
          (512, 1024, 4096, 8), (1024, 512, 4096, 8), (4096, 1024, 512, 1)]
for (m, n, k, splits) in shapes:
    a = (torch.randn(m, k, device="cuda") * 0.5).bfloat16()
    b = (torch.randn(n, k, device="cuda") * 0.5).bfloat16()
    ref = a.float() @ b.float().t()
    for dt in ((torch.float32, torch.bfloat16) if splits == 1 else (torch.float32,)):
        c = rt.gemm_bf16_tn(a, b, out_dtype=dt, k_splits=splits)
        torch.cuda.synchronize()
        err = (c.float() - ref).abs().max().item()
        tol = 1e-3 * k ** 0.5 if dt == torch.float32 else 0.02 * ref.abs().max().item()
        ok = err <= tol and bool(torch.isfinite(c.float()).all())
        print(f"M{m} N{n} K{k} splits{splits} {str(dt)[6:]}: max err {err:.3e} (tol {tol:.2e}) {'ok' if ok else 'FAIL'}", flush=True)
        assert ok
    c = torch.empty((m, n), dtype=torch.float32, device="cuda")
    for fn, name in ((lambda: rt.gemm_bf16_tn(a, b, out=c, k_splits=splits), "tcgen05"), (lambda: torch.matmul(a, b.t()), "cublas")):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            fn()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 50 * 1e3
        print(f"    {name}: {us:.1f} us  {2 * m * n * k / us / 1e6:.1f} TFLOP/s")
for (m, n, k) in [(128, 128, 64), (256, 128, 256), (512, 1024, 4096), (4096, 1024, 512)]:
    at = (torch.randn(k, m, device="cuda") * 0.5).bfloat16()      # A stored [K, M]
    bt = (torch.randn(k, n, device="cuda") * 0.5).bfloat16()      # B stored [K, N]
    a, b = at.t().contiguous(), bt.t().contiguous()
    ref = a.float() @ b.float().t()
    for amn, bmn in ((True, False), (False, True), (True, True)):
        c = rt.gemm_bf16_tn(at if amn else a, bt if bmn else b, a_mn=amn, b_mn=bmn)
        torch.cuda.synchronize()
        err = (c - ref).abs().max().item()
        print(f"MN-major A={amn} B={bmn} M{m} N{n} K{k}: max err {err:.3e}", "ok" if err < 1e-3 * k ** 0.5 else "FAIL", flush=True)
x = torch.randn(4096, 1024, device="cuda").bfloat16()
assert torch.equal(rt.transpose_bf16(x), x.t().contiguous())
print("gemm_check done")
