#!/usr/bin/env python
"""Per-sample engine on the reference's transformer-style examples (4x4 attention blocks): Jacobians/s of
vmap(jacve(f, 'rev', argnums)) over a batch, per kernel kind."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
import numpy as np
import torch
import graphax_b200 as gx
from graphax_b200 import runtime
from test_deep_learning import CASES, _inputs

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
for name in ("Perceptron", "Encoder", "EncoderDecoder"):
    fn, shapes, argnums = CASES[name]
    xs = [torch.from_numpy(x.astype(np.float32)).cuda() for x in _inputs(name, batch)]
    jf = gx.jacve(fn, order="rev", argnums=argnums)
    t0 = time.perf_counter()
    plan = jf.lower(shapes)
    t_lower = time.perf_counter() - t0
    t0 = time.perf_counter()
    out = gx.vmap(jf)(*xs)
    torch.cuda.synchronize()
    t_first = time.perf_counter() - t0
    dp = jf._device_plan(plan, 0)
    info = dp.info()
    reps = 5
    t0 = time.perf_counter()
    for _ in range(reps):
        out = gx.vmap(jf)(*xs)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    print(f"{name}: {plan.stats['n_ops']} ops, {plan.n_slots} live, tile {plan.n_rows}x{plan.n_cols}, lower {t_lower:.1f} s, "
          f"first call {t_first:.1f} s, kernel={runtime.KERNEL_NAMES[info.kernel_kind]} threads={info.threads} regs={info.regs_per_thread} "
          f"-> {dt * 1e3:.2f} ms per batch of {batch} = {batch / dt / 1e6:.2f} M Jacobians/s", flush=True)
