#!/usr/bin/env python
"""Config 5: weight Jacobians of the MLP loss over a batch (reverse-order elimination), timed on the GPU.
Reports per-step time, achieved TFLOP/s over the five GEMMs (2 forward, 3 backward) and a torch.autograd
reference (cuBLAS) for context."""
import argparse, json, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch
from graphax_b200.dense import dense_jacve, DenseExecutor, build_dense_plan
from graphax_b200.examples import MLP_loss
from graphax_b200.tracer import make_jaxpr

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=4096)
ap.add_argument("--widths", default="512,1024,512")
ap.add_argument("--steps", type=int, default=200)
ap.add_argument("--graph", action="store_true")
ap.add_argument("--profile", action="store_true", help="per-kernel times (serialised, warm) to stderr")
a = ap.parse_args()
I, H, O = map(int, a.widths.split(","))
B = a.batch
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(B, I, device="cuda", generator=g).bfloat16()
W1 = (torch.randn(H, I, device="cuda", generator=g) * I ** -0.5).bfloat16()
b1 = torch.randn(H, device="cuda", generator=g) * 0.1
W2 = (torch.randn(O, H, device="cuda", generator=g) * H ** -0.5).bfloat16()
b2 = torch.randn(O, device="cuda", generator=g) * 0.1
y = torch.randn(B, O, device="cuda", generator=g).bfloat16()
args = (x, W1, b1, W2, b2, y)
plan = build_dense_plan(make_jaxpr(MLP_loss, [tuple(t.shape) for t in args]), "rev", (1, 2, 3, 4))
ex = DenseExecutor(plan, 0)
print('materialised:', [(n, plan.dag.nodes[n].kind, plan.dag.nodes[n].shape) for n in ex.order], file=sys.stderr)
for _ in range(3):
    grads, loss = ex.run(args)
torch.cuda.synchronize()
l0 = ex.launches
ex.run(args)
per_step_launches = ex.launches - l0
if a.profile:
    # warm GPU-side time of every kernel group: a CUDA graph with 10 back-to-back copies of the node's launches
    side = torch.cuda.Stream()
    tot = 0.0
    for n in ex.order:
        node = plan.dag.nodes[n]
        with torch.cuda.stream(side):
            ex._launch_node(n)
            torch.cuda.synchronize()
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr, stream=side):
                for _ in range(10):
                    ex._launch_node(n)
            gr.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                gr.replay()
            e1.record()
            torch.cuda.synchronize()
        us_n = e0.elapsed_time(e1) * 1e3 / 100
        tot += us_n
        print(f"  {n}:{node.kind}:{'x'.join(map(str, node.shape))} tx={getattr(node, 'tx', None)} ty={getattr(node, 'ty', None)}  {us_n:7.2f} us", file=sys.stderr)
    print(f"  sum {tot:.1f} us", file=sys.stderr)
runner = lambda: ex.run(args)
if a.graph:
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        ex.run(args)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=s):
            ex.run(args)
    runner = graph.replay
for _ in range(5):
    runner()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.steps):
    runner()
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) / a.steps * 1e3
flops = 5 * 2 * B * I * H if I == O else 2 * B * (2 * I * H + 3 * H * O)
# torch.autograd with cuBLAS, same precision policy, for context
tx = [t.clone().float().requires_grad_(t.ndim >= 1 and i in (1, 2, 3, 4)) for i, t in enumerate(args)]
def torch_step():
    with torch.autocast("cuda", dtype=torch.bfloat16):
        d = torch.tanh(torch.tanh(tx[0] @ tx[1].T + tx[2]) @ tx[3].T + tx[4]) - tx[5]
        L = .5 * (d.float() ** 2).sum()
    return torch.autograd.grad(L, [tx[1], tx[2], tx[3], tx[4]])
for _ in range(5):
    torch_step()
torch.cuda.synchronize()
e0.record()
for _ in range(a.steps):
    torch_step()
e1.record()
torch.cuda.synchronize()
us_t = e0.elapsed_time(e1) / a.steps * 1e3
ref = torch_step()
err = max(float((gi - ri).abs().max() / ri.abs().max()) for gi, ri in zip(grads, ref))
print(json.dumps({"workload": f"MLP {I}->{H}->{O}, batch {B}, bf16 operands / fp32 accumulate, order rev",
                  "us_per_step": us, "tflops": flops / us / 1e6, "launches_per_step": per_step_launches,
                  "cuda_graph": a.graph, "torch_autograd_us": us_t, "max_rel_diff_vs_torch_autocast": err}))
