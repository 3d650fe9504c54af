#!/usr/bin/env python
"""Launch one named plan's ahead-of-time kernel a few times (a target for ncu).

    ncu --set full --import-source on -k regex:gx_jacve_kernel -s 3 -c 1 -o out python tools/launch_plan.py roe1d rev 1048576
"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from graphax_b200 import runtime                      # noqa: E402
from graphax_b200.configs import WORKLOADS            # noqa: E402


def main():
    name, order, batch = sys.argv[1], sys.argv[2], int(sys.argv[3])
    rounding = sys.argv[4] if len(sys.argv) > 4 else "reference"
    w = WORKLOADS[name]
    plan = w.plan(order, rounding=rounding)
    dp = runtime.DevicePlan(plan, 0, jit=False)
    dev = [torch.from_numpy(x).cuda() for x in w.inputs(batch)]
    jac = torch.empty((batch, plan.n_rows, plan.n_cols), device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    for _ in range(6):
        dp.launch(batch, [t.data_ptr() for t in dev], jac.data_ptr(), None, s)
    torch.cuda.synchronize()
    print(dp.kernel, dp.info().threads, dp.info().ctas_per_sm, dp.info().regs_per_thread)


if __name__ == "__main__":
    main()
