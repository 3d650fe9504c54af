#!/usr/bin/env python
"""Where does the end-to-end (host buffers) time go? alloc vs gx_jacve_host vs packaging."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import graphax_b200 as gx
from graphax_b200.configs import WORKLOADS
w = WORKLOADS["roe3d"]
batch = 1 << 20
host_in = [torch.from_numpy(x).pin_memory() for x in w.inputs(batch)]
jf = gx.jacve(w.fun, order=w.order("mixed"), argnums=w.argnums)
plan = jf.lower(w.arg_shapes)
dp = jf._device_plan(plan, 0)
jac = torch.empty((batch, 5, 10), dtype=torch.float32, pin_memory=True)
for _ in range(3):
    dp.run_host(batch, [t.data_ptr() for t in host_in], jac.data_ptr(), None, True)
t0 = time.perf_counter()
for _ in range(10):
    dp.run_host(batch, [t.data_ptr() for t in host_in], jac.data_ptr(), None, True)
print("run_host (preallocated pinned out): %.2f ms" % ((time.perf_counter() - t0) / 10 * 1e3))
t0 = time.perf_counter()
for _ in range(10):
    o = torch.empty((batch, 5, 10), dtype=torch.float32, pin_memory=True)
print("torch.empty pinned 210MB: %.2f ms" % ((time.perf_counter() - t0) / 10 * 1e3))
keep = []
t0 = time.perf_counter()
for _ in range(10):
    keep.append(jf.batched(*host_in))
    keep = keep[-1:]
print("jf.batched(host): %.2f ms" % ((time.perf_counter() - t0) / 10 * 1e3))
out = torch.empty((batch, 5, 10), dtype=torch.float32).pin_memory()
for _ in range(3):
    jf.batched(*host_in, out=out)
t0 = time.perf_counter()
for _ in range(20):
    jf.batched(*host_in, out=out)
print("jf.batched(host, out=pinned): %.2f ms" % ((time.perf_counter() - t0) / 20 * 1e3))
import cProfile, pstats
pr = cProfile.Profile()
pr.enable()
for _ in range(20):
    jf.batched(*host_in, out=out)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
import os
for chunk in (1 << 15, 1 << 16, 1 << 17, 1 << 18):
    os.environ["GX_HOST_CHUNK"] = str(chunk)
    jf2 = gx.jacve(w.fun, order=w.order("mixed"), argnums=w.argnums)
    dp2 = jf2._device_plan(jf2.lower(w.arg_shapes), 0)
    for _ in range(3):
        dp2.run_host(batch, [t.data_ptr() for t in host_in], jac.data_ptr(), None, True)
    t0 = time.perf_counter()
    for _ in range(10):
        dp2.run_host(batch, [t.data_ptr() for t in host_in], jac.data_ptr(), None, True)
    print("chunk %d: run_host %.2f ms" % (chunk, (time.perf_counter() - t0) / 10 * 1e3))
