#!/usr/bin/env python
"""Where does a single-tile contraction spend its time?  Phase stamps (GemmArgs::trace, %globaltimer) of the five
launches of the MLP step (config 5), replayed from the same CUDA graph the bench uses.

    python tools/gemm_trace.py [--pair]

Per launch: when its CTAs enter / pass the dependency wait / see their first operand stage / complete the
accumulator / finish the epilogue / leave, relative to the first entry of the first launch (min / median / max over CTAs)."""
import argparse, ctypes, os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
ap = argparse.ArgumentParser()
ap.add_argument("--pair", action="store_true")
ap.add_argument("--smid", action="store_true", help="also print which SMs every launch used and the main-loop time by SM range")
ap.add_argument("--batch", type=int, default=4096)
ap.add_argument("--widths", default="512,1024,512")
a = ap.parse_args()
if a.pair:
    os.environ["GX_GEMM_PAIR"] = "1"
import numpy as np
import torch
from graphax_b200.dense import DenseExecutor, build_dense_plan
from graphax_b200.examples import MLP_loss
from graphax_b200.tracer import make_jaxpr

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
for _ in range(3):
    ex.run(args)
torch.cuda.synchronize()
MAXL, MAXC = 8, 512
buf = torch.zeros((MAXL, MAXC, 8), dtype=torch.int64, device="cuda")
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    ex.run(args)
    torch.cuda.synchronize()
    ex.lib.gx_gemm_set_trace(ctypes.c_void_p(buf.data_ptr()), MAXL, MAXC)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=s):
        ex.run(args)
    ex.lib.gx_gemm_set_trace(None, 0, 0)
for _ in range(20):
    graph.replay()
torch.cuda.synchronize()
t = buf.cpu().numpy()
names = [f"{n}:{'x'.join(map(str, plan.dag.nodes[n].shape))}" for n in ex.order if plan.dag.nodes[n].kind == "gemm"]
prev_end = None
ticks = np.unique(np.diff(np.unique(t[t > 0])))
print("globaltimer granularity (smallest distinct steps, ns):", ticks[:4])
for l in range(MAXL):
    rows = t[l][t[l][:, 0] > 0]
    if not len(rows):
        continue
    base = rows[:, 0].min() if prev_end is None else base0
    base0 = base
    rel = (rows - base) / 1e3
    cols = ["entry", "wait passed", "first stage", "acc complete", "epilogue done", "exit"]
    print(f"launch {l} ({names[l] if l < len(names) else '?'}), {len(rows)} CTAs; us after the first entry of the first launch")
    for c, name in enumerate(cols):
        v = rel[:, c]
        print(f"    {name:14s} min {v.min():7.2f}  median {np.median(v):7.2f}  max {v.max():7.2f}")
    d = np.diff(rows[:, :6], axis=1) / 1e3
    print("    phases (median us): wait", round(float(np.median(d[:, 0])), 2), " fill", round(float(np.median(d[:, 1])), 2),
          " main loop", round(float(np.median(d[:, 2])), 2), " epilogue", round(float(np.median(d[:, 3])), 2),
          " drain", round(float(np.median(d[:, 4])), 2), "| span", round(float((rows[:, 5].max() - rows[:, 0].min()) / 1e3), 2))
    if a.smid:
        sm = rows[:, 6].astype(int)
        ml = (rows[:, 3] - rows[:, 2]) / 1e3
        print("    SMs used:", len(set(sm.tolist())), "of ids", sm.min(), "..", sm.max(), " idle ids:",
              sorted(set(range(int(sm.max()) + 1)) - set(sm.tolist()))[:40])
        for lo in range(0, int(sm.max()) + 1, 16):
            sel = (sm >= lo) & (sm < lo + 16)
            if sel.any():
                print(f"      SM {lo:3d}-{lo + 15:3d}: {int(sel.sum()):3d} CTAs, main loop median {np.median(ml[sel]):.2f} us")
    prev_end = rows[:, 5].max()
