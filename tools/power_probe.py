#!/usr/bin/env python
"""Is a plan-specialised kernel bound by the GPU's power cap?  (B200: 1000 W)

    python tools/power_probe.py [--workload roe3d] [--order mixed] [--rounding reference] [--seconds 3]

Launches the kernel back to back for a few seconds.  CUDA events every WINDOW launches give the step time as a
function of time since the start; an NVML thread records SM clock, board power and the clock-event reasons every
few milliseconds.  If the step time of the first milliseconds (clocks still at their maximum) is shorter than the
sustained one and the sustained windows coincide with `sw_power_cap` and a lower SM clock, the kernel runs at the
power cap: removing stall cycles then buys nothing, only fewer joules per Jacobian do.  The kernel also counts its
own SM cycles (clock64 around the launch loop is not available from outside, so the clock is taken from NVML).
"""
import argparse
import json
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

from graphax_b200 import runtime  # noqa: E402
from graphax_b200.configs import WORKLOADS  # noqa: E402


class Sampler(threading.Thread):
    def __init__(self, period=0.002):
        super().__init__(daemon=True)
        import pynvml
        pynvml.nvmlInit()
        self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(0)
        self.period, self.rows, self._halt = period, [], threading.Event()

    def run(self):
        nv = self.nv
        while not self._halt.is_set():
            try:
                self.rows.append((time.perf_counter(), nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM),
                                  nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0,
                                  nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)))
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._halt.set()
        self.join(timeout=2)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="roe3d")
    ap.add_argument("--order", default="mixed")
    ap.add_argument("--rounding", default="reference")
    ap.add_argument("--seconds", type=float, default=3.0)
    ap.add_argument("--window", type=int, default=50)
    ap.add_argument("--idle", type=float, default=1.0, help="seconds of idle before the run (lets the board cool / clocks rise)")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    batch = min(w.batch, 1 << 20)
    plan = w.plan(args.order, rounding=args.rounding)
    host = w.inputs(batch)
    sets = [([torch.from_numpy(x).cuda() for x in host], torch.empty((batch, plan.n_rows, plan.n_cols), device="cuda"))
            for _ in range(4)]
    dp = runtime.DevicePlan(plan, 0)
    stream = torch.cuda.current_stream()

    def launch(i):
        ins, out = sets[i % 4]
        dp.launch(batch, [t.data_ptr() for t in ins], out.data_ptr(), None, stream.cuda_stream)

    for i in range(8):
        launch(i)
    torch.cuda.synchronize()
    time.sleep(args.idle)
    sampler = Sampler()
    sampler.start()
    time.sleep(0.05)
    events = [torch.cuda.Event(enable_timing=True)]
    t_start = time.perf_counter()
    events[0].record(stream)
    n = 0
    while time.perf_counter() - t_start < args.seconds:
        for _ in range(args.window):
            launch(n)
            n += 1
        e = torch.cuda.Event(enable_timing=True)
        e.record(stream)
        events.append(e)
        if len(events) % 8 == 0:
            events[-6].synchronize()          # keep the queue a few windows deep, not unbounded
    torch.cuda.synchronize()
    t_end = time.perf_counter()
    sampler.stop()
    us = [events[i].elapsed_time(events[i + 1]) * 1e3 / args.window for i in range(len(events) - 1)]
    t_of = [events[0].elapsed_time(events[i + 1]) for i in range(len(events) - 1)]     # ms since start
    rows = [(t - t_start, mhz, watts, mask) for t, mhz, watts, mask in sampler.rows if t_start - 0.05 <= t <= t_end]

    def stat(lo_ms, hi_ms):
        sel = [u for u, t in zip(us, t_of) if lo_ms <= t < hi_ms]
        smp = [r for r in rows if lo_ms <= r[0] * 1e3 < hi_ms]
        if not sel:
            return None
        return {"window_ms": [lo_ms, hi_ms], "us_per_step": round(sum(sel) / len(sel), 2), "min_us": round(min(sel), 2),
                "sm_mhz": round(sum(r[1] for r in smp) / max(1, len(smp))), "watts": round(sum(r[2] for r in smp) / max(1, len(smp))),
                "sw_power_cap": round(sum(1 for r in smp if r[3] & 0x4) / max(1, len(smp)), 2),
                "other_reasons": sorted({hex(r[3] & ~0x5) for r in smp if r[3] & ~0x5})}

    edges = [0, 5, 10, 20, 50, 100, 200, 500, 1000, 2000, 3000, 5000, 10000]
    out = {"workload": args.workload, "order": args.order, "rounding": args.rounding, "kernel": dp.kernel,
           "regs": dp.info().regs_per_thread, "launches": n,
           "windows": [s for s in (stat(a, b) for a, b in zip(edges, edges[1:])) if s]}
    text = json.dumps(out, indent=1)
    print(text)
    if args.out:
        Path(args.out).write_text(text)


if __name__ == "__main__":
    main()
