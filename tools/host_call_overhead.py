#!/usr/bin/env python
"""Fixed cost of one host-array call of the public API (gx.vmap(jacve(...))(*pinned_host_arrays, out=...)): a batch
small enough that copies and kernel are negligible, so the time is Python + ctypes + the pipeline's own
synchronisation.  Prints microseconds per call and a cProfile of 200 calls."""
import cProfile, pstats, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch
import graphax_b200 as gx
from graphax_b200.configs import WORKLOADS

w = WORKLOADS["roe3d"]
jf = gx.jacve(w.fun, w.order("mixed"), w.argnums)
for batch in (1024, 1 << 17):
    host = [torch.from_numpy(x).pin_memory() for x in w.inputs(batch)]
    plan = jf.lower(w.arg_shapes)
    out = torch.empty((batch, plan.n_rows, plan.n_cols), dtype=torch.float32).pin_memory()
    for _ in range(5):
        jf.batched(*host, out=out)
    torch.cuda.synchronize()
    n = 300
    t0 = time.perf_counter()
    for _ in range(n):
        jf.batched(*host, out=out)
    dt = (time.perf_counter() - t0) / n
    print(f"batch {batch}: {dt * 1e6:.1f} us per call")
host = [torch.from_numpy(x).pin_memory() for x in w.inputs(1024)]
out = torch.empty((1024, plan.n_rows, plan.n_cols), dtype=torch.float32).pin_memory()
pr = cProfile.Profile()
pr.enable()
for _ in range(200):
    jf.batched(*host, out=out)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
