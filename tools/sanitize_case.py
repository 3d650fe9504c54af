#!/usr/bin/env python
"""Small driver for compute-sanitizer: AOT (TMA path + ragged tail), NVRTC and interpreter kernels once each."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
import numpy as np
from graphax_b200.configs import WORKLOADS
from helpers import bits, run_engine
from oracle import replay

for name, order, batch in (("roe3d", "mixed", 1000), ("roe1d", "rev", 333), ("robotarm", "rev", 200)):
    w = WORKLOADS[name]
    plan = w.plan(order)
    xs = w.inputs(batch)
    ref, _ = replay.replay(plan.blob, xs, plan.n_rows, plan.n_cols)
    for kernel in ("aot", "jit", "interpreter"):
        jac, primal, used = run_engine(plan, xs, kernel)
        ok = np.array_equal(bits(jac), bits(ref)) if name != "robotarm" else np.allclose(jac, ref, rtol=1e-4, atol=1e-4)
        print(name, order, kernel, used, "ok" if ok else "MISMATCH", flush=True)
        assert ok
print("sanitize_case done")
