#!/usr/bin/env python
"""The five contractions of the MLP step (512->1024->512, batch 4096), one by one: the tcgen05 kernel against
torch.matmul (cuBLAS), each replayed from a CUDA graph of 20 launches so that launch overhead is amortised."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from graphax_b200 import runtime as rt

torch.manual_seed(0)
# (M, N, K, k_splits, a_mn, b_mn, bf16 out): forward x W1^T, a1 W2^T; backward g2 W2, g2^T a1, g1^T x
shapes = [(4096, 1024, 512, 1, False, False, True), (4096, 512, 1024, 1, False, False, True),
          (4096, 1024, 512, 1, False, True, True), (512, 1024, 4096, 8, True, True, False),
          (1024, 512, 4096, 8, True, True, False)]


def timed(fn, reps=20, rounds=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn()
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            for _ in range(reps):
                fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g.replay()
    e0.record()
    for _ in range(rounds):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (reps * rounds) * 1e3


for (m, n, k, splits, amn, bmn, bf16) in shapes:
    a = (torch.randn((k, m) if amn else (m, k), device="cuda") * 0.5).bfloat16()
    b = (torch.randn((k, n) if bmn else (n, k), device="cuda") * 0.5).bfloat16()
    c = torch.zeros((m, n), dtype=torch.bfloat16 if bf16 else torch.float32, device="cuda")
    A = a.t() if amn else a
    Bt = b if bmn else b.t()
    mine = timed(lambda: rt.gemm_bf16_tn(a, b, out=c, k_splits=splits, a_mn=amn, b_mn=bmn, zero=False))
    ref_out = torch.empty((m, n), dtype=torch.bfloat16, device="cuda")
    cub = timed(lambda: torch.matmul(A, Bt, out=ref_out))
    fl = 2.0 * m * n * k
    print(f"M{m} N{n} K{k} splits{splits} a_mn={amn} b_mn={bmn}: tcgen05 {mine:.2f} us ({fl / mine / 1e6:.0f} TF/s)   "
          f"cuBLAS {cub:.2f} us ({fl / cub / 1e6:.0f} TF/s)", flush=True)
