#!/usr/bin/env python
"""Does the per-sample kernel reach the copy engines' PCIe rate when its TMA stores go straight to pinned host memory?

    python tools/zero_copy_probe.py

Times (CUDA events) the RoeFlux_3d mixed-order kernel on 2^20 device-resident samples with the Jacobian / primal
buffers (a) in HBM, followed by a cudaMemcpyAsync D2H of the same bytes, and (b) in pinned host memory (UVA pointer
handed to gx_jacve_launch: every 6400-byte bulk store crosses PCIe), (c) with the inputs read from pinned host memory
too.  Checks that (b) and (c) equal (a) bit for bit."""
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from graphax_b200 import runtime                      # noqa: E402
from graphax_b200.configs import WORKLOADS            # noqa: E402


def main():
    w = WORKLOADS["roe3d"]
    batch = 1 << 20
    plan = w.plan("mixed", rounding="reference")
    dp = runtime.DevicePlan(plan, 0, jit=False)
    dev = [torch.from_numpy(x).cuda() for x in w.inputs(batch)]
    ptrs = [t.data_ptr() for t in dev]
    s = torch.cuda.current_stream()
    d_j, d_p = torch.empty((batch, 5, 10), device="cuda"), torch.empty((batch, 5), device="cuda")
    h_j, h_p = torch.empty((batch, 5, 10)).pin_memory(), torch.empty((batch, 5)).pin_memory()
    z_j, z_p = torch.zeros((batch, 5, 10)).pin_memory(), torch.zeros((batch, 5)).pin_memory()

    def timed(fn, reps=5):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(reps):
            fn()
        e1.record(s)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    def staged():
        dp.launch(batch, ptrs, d_j.data_ptr(), d_p.data_ptr(), s.cuda_stream)
        h_j.copy_(d_j, non_blocking=True)
        h_p.copy_(d_p, non_blocking=True)

    def direct():
        dp.launch(batch, ptrs, z_j.data_ptr(), z_p.data_ptr(), s.cuda_stream)

    h_in = [torch.from_numpy(x).pin_memory() for x in w.inputs(batch)]
    hptrs = [t.data_ptr() for t in h_in]
    y_j, y_p = torch.zeros((batch, 5, 10)).pin_memory(), torch.zeros((batch, 5)).pin_memory()

    def direct_both():          # (c) inputs read from pinned host memory as well: the whole call is one launch
        dp.launch(batch, hptrs, y_j.data_ptr(), y_p.data_ptr(), s.cuda_stream)

    def staged_both():          # H2D + kernel + D2H back to back on one stream (no overlap): the naive reference point
        for d, h in zip(dev, h_in):
            d.copy_(h, non_blocking=True)
        staged()

    out = {"bytes": 4 * batch * 55, "bytes_in": 4 * batch * 10}
    out["h2d_kernel_d2h_serial_ms"] = timed(staged_both)
    out["kernel_host_in_host_out_ms"] = timed(direct_both)
    out["equal_both"] = bool(torch.equal(d_j.cpu(), y_j) and torch.equal(d_p.cpu(), y_p))
    out["kernel_then_d2h_ms"] = timed(staged)
    out["kernel_into_pinned_host_ms"] = timed(direct)
    out["equal"] = bool(torch.equal(h_j, z_j) and torch.equal(h_p, z_p))
    for k in ("kernel_then_d2h_ms", "kernel_into_pinned_host_ms"):
        out[k.replace("_ms", "_gbs")] = out["bytes"] / out[k] / 1e6
    print(json.dumps(out))


if __name__ == "__main__":
    main()
