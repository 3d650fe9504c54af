#!/usr/bin/env python
"""tcgen05 GEMM kernel at sizes where launch latency does not matter (achieved TFLOP/s vs cuBLAS)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from graphax_b200 import runtime as rt
for (m, n, k) in [(4096, 4096, 4096), (8192, 8192, 8192), (16384, 4096, 2048)]:
    a = (torch.randn(m, k, device="cuda") * 0.1).bfloat16()
    b = (torch.randn(n, k, device="cuda") * 0.1).bfloat16()
    c = torch.empty((m, n), dtype=torch.bfloat16, device="cuda")
    for fn, name in ((lambda: rt.gemm_bf16_tn(a, b, out=c), "tcgen05 (bf16 out)"), (lambda: torch.matmul(a, b.t(), out=c), "cublas")):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            fn()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 20 * 1e3
        print(f"M{m} N{n} K{k} {name}: {us:.1f} us  {2 * m * n * k / us / 1e6:.1f} TFLOP/s", flush=True)
