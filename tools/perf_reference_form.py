#!/usr/bin/env python
"""The reference's own timing protocol (performance_measurements/perf.py:17-76, performance_measurements.py:105-154)
applied to this engine: call jacve(f, order, argnums)(*arrays) 1000 times, wall-clock every call around a device
synchronise, drop the first 10, report the median and the 2.5 % / 97.5 % quantiles.  Workloads as in the reference:
RoeFlux_1d on six (600,) arrays and RobotArm_6DOF on six (1000,) arrays, orders fwd / rev / Markowitz / AlphaGrad."""
import json
import sys
import time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch
import graphax_b200 as gx
from graphax_b200.configs import WORKLOADS

out = {}
for wname, n in (("roe1d", 600), ("robotarm", 1000)):
    w = WORKLOADS[wname]
    xs = [torch.from_numpy(x).cuda() for x in w.inputs(n)]
    for order in ("fwd", "rev", "markowitz", "alphagrad"):
        for sparse in (False, True):
            jf = gx.jacve(w.fun, order=w.order(order), argnums=w.argnums, sparse_representation=sparse)
            ts = []
            for i in range(1000):
                t0 = time.perf_counter()
                res = jf(*xs)
                torch.cuda.synchronize()
                ts.append(time.perf_counter() - t0)
            ts = np.array(ts[10:]) * 1e3
            key = f"{wname}[{n}]:{order}:{'sparse' if sparse else 'dense'}"
            out[key] = {"median_ms": float(np.median(ts)), "q2.5_ms": float(np.quantile(ts, .025)),
                        "q97.5_ms": float(np.quantile(ts, .975))}
            print(key, out[key], flush=True)
print(json.dumps(out))
