#!/usr/bin/env python
"""Measure launch-shape variants of the plan-specialised kernels on the GPU (NVRTC), write tuning.json.

    python tools/tune.py [--workloads roe3d,...] [--out gpurun_out/tune.json] [--write]

For every named plan: threads per CTA x resident CTAs/SM (register bound) x output stages, timed with
CUDA events over the BASELINE batch with rotating buffer sets.  --write stores the winners in
graphax_b200/tuning.json (read by codegen.launch_shape at build / specialise time).
"""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

from graphax_b200 import runtime  # noqa: E402
from graphax_b200.configs import WORKLOADS  # noqa: E402


def time_variant(dp, sets, batch, plan, steps):
    if plan.bytes_per_jacobian * batch < (64 << 20):
        return time_variant_graph(dp, sets, batch)
    stream = torch.cuda.current_stream()
    for i in range(5):
        ins, out = sets[i % len(sets)]
        dp.launch(batch, [t.data_ptr() for t in ins], out.data_ptr(), None, stream.cuda_stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(steps):
        ins, out = sets[i % len(sets)]
        dp.launch(batch, [t.data_ptr() for t in ins], out.data_ptr(), None, stream.cuda_stream)
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps * 1e3   # us


def time_variant_graph(dp, sets, batch, per_graph=40, replays=25):
    """Launch-bound sizes: 40 consecutive steps captured into one CUDA graph and replayed (like bench.py), so that
    the kernel's own start-up cost decides and not the host's launch rate."""
    cap = torch.cuda.Stream()
    cap.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(cap):
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=cap):
            cs = torch.cuda.current_stream().cuda_stream
            for i in range(per_graph):
                ins, out = sets[i % len(sets)]
                dp.launch(batch, [t.data_ptr() for t in ins], out.data_ptr(), None, cs)
    torch.cuda.current_stream().wait_stream(cap)
    for _ in range(3):
        graph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(replays):
        graph.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (replays * per_graph) * 1e3   # us


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workloads", default="roe3d")
    ap.add_argument("--orders", default="")
    ap.add_argument("--divisions", default="reciprocal,ieee",
                    help="comma list of reciprocal / ieee (fma rounding) / reference (the reference's own rounding)")
    ap.add_argument("--threads", default="64,128,256")
    ap.add_argument("--min-blocks", default="2,3,4,5,6")
    ap.add_argument("--out-stages", default="2,1")
    ap.add_argument("--spt", default="1,2")
    ap.add_argument("--schedules", default="", help="comma list of plan schedules to compare (default: the plan default)")
    ap.add_argument("--contract", type=int, default=0)
    ap.add_argument("--fast", default="", help="comma list of GX_CHECKED_FAST settings to compare, e.g. 0,1")
    ap.add_argument("--slp", default="", help="comma list of GX_SLP settings to compare (f32x2 pairing, slp.py), e.g. 0,1")
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--out", default="gpurun_out/tune.json")
    ap.add_argument("--write", action="store_true")
    args = ap.parse_args()
    results, best = {}, {}
    for wname in args.workloads.split(","):
        w = WORKLOADS[wname]
        batch = args.batch or min(w.batch, 1 << 20)
        host = w.inputs(batch)
        for order in (args.orders.split(",") if args.orders else w.order_names()):
            for division, sched in [(d, sc) for d in args.divisions.split(",") for sc in (args.schedules.split(",") if args.schedules else [""])]:
                if sched or args.contract:
                    from graphax_b200.plan import build_plan
                    plan = build_plan(w.jaxpr(), w.order(order), w.argnums, schedule=sched or "pressure",
                                      contract=bool(args.contract),
                                      **({"rounding": "reference"} if division == "reference" else
                                         {"division": division, "rounding": "fma"}))
                elif division == "reference":
                    plan = w.plan(order, rounding="reference")
                else:
                    plan = w.plan(order, division=division, rounding="fma")
                nsets = 4
                sets = [([torch.from_numpy(x).cuda() for x in host],
                         torch.empty((batch, plan.n_rows, plan.n_cols), device="cuda")) for _ in range(nsets)]
                dp = runtime.DevicePlan(plan, 0, jit=False)
                name = f"{wname}:{order}:{division}" + (f":{sched}" if sched else "") + (":contract" if args.contract else "")
                rows = []
                if dp.info().kernel_kind == runtime.KERNEL_AOT:
                    i = dp.info()
                    rows.append({"variant": "aot", "threads": i.threads, "ctas_per_sm": i.ctas_per_sm,
                                 "regs": i.regs_per_thread, "us": time_variant(dp, sets, batch, plan, args.steps)})
                # reference result for a bitwise check of every variant (interpreter kernel, ragged batch)
                nchk = min(4096 + 77, batch)       # ragged where the batch allows it
                chk_in = [t[:nchk].contiguous() for t in sets[0][0]]
                dp.select_kernel(runtime.KERNEL_INTERPRETER)
                ref = torch.empty((nchk, plan.n_rows, plan.n_cols), device="cuda")
                dp.launch(nchk, [t.data_ptr() for t in chk_in], ref.data_ptr(), None,
                          torch.cuda.current_stream().cuda_stream)
                torch.cuda.synchronize()
                for fast, slp in [(f, s) for f in (args.fast.split(",") if args.fast else [""])
                                  for s in (args.slp.split(",") if args.slp else [""])]:
                 import os
                 if fast:
                    os.environ["GX_CHECKED_FAST"] = fast
                 if slp:
                    os.environ["GX_SLP"] = slp
                 for spt in map(int, args.spt.split(",")):
                  for threads in map(int, args.threads.split(",")):
                    for mb in map(int, args.min_blocks.split(",")):
                        if threads * mb * spt > 2048 or threads * mb * spt < 192:
                            continue
                        for st in map(int, args.out_stages.split(",")):
                            cfg = {"variant": "jit", "threads": threads, "min_blocks": mb, "out_stages": st,
                                   "samples_per_thread": spt}
                            if fast:
                                cfg["checked_fast"] = int(fast)
                            if slp:
                                cfg["slp"] = int(slp)
                            try:
                                t0 = time.time()
                                dp.specialize(threads=threads, min_blocks=mb, out_stages=st, samples_per_thread=spt)
                                i = dp.info()
                                got = torch.full_like(ref, float("nan"))
                                dp.launch(nchk, [t.data_ptr() for t in chk_in], got.data_ptr(), None,
                                          torch.cuda.current_stream().cuda_stream)
                                torch.cuda.synchronize()
                                exact = bool(torch.equal(got.view(torch.int32), ref.view(torch.int32)))
                                us = time_variant(dp, sets, batch, plan, args.steps)
                                rows.append(dict(cfg, ctas_per_sm=i.ctas_per_sm, regs=i.regs_per_thread, us=us,
                                                 bit_exact=exact, compile_s=round(time.time() - t0, 2)))
                            except runtime.GraphaxB200Error as e:
                                rows.append(dict(cfg, error=str(e)[:300]))
                ok = [r for r in rows if "us" in r and r["variant"] == "jit"]
                if not ok:
                    raise SystemExit(f"{name}: every variant failed, first error: {rows[-1].get('error')}")
                b = min(ok, key=lambda r: r["us"])
                best[plan.key] = {"name": name, "threads": b["threads"], "min_blocks": b["min_blocks"],
                                  "out_stages": b["out_stages"], "samples_per_thread": b["samples_per_thread"],
                                  "bit_exact": b["bit_exact"], "us": b["us"], "regs": b["regs"],
                                  **({"checked_fast": b["checked_fast"]} if "checked_fast" in b else {}),
                                  "gjac_per_s": batch / b["us"] / 1e3}
                results[name] = rows
                Path(args.out).parent.mkdir(parents=True, exist_ok=True)
                Path(args.out).write_text(json.dumps({"results": results, "best": best}, indent=1))
                print(name, plan.key, "best", b, flush=True)
                for r in sorted(ok, key=lambda r: r["us"])[:8]:
                    print("    ", r, flush=True)
                dp.close()
                del sets
                torch.cuda.empty_cache()
    Path(args.out).parent.mkdir(parents=True, exist_ok=True)
    Path(args.out).write_text(json.dumps({"results": results, "best": best}, indent=1))
    if args.write:
        path = ROOT / "graphax_b200" / "tuning.json"
        merged = json.loads(path.read_text()) if path.exists() else {}
        for key, entry in best.items():           # keep what this sweep does not measure (e.g. "input_staged")
            merged[key] = {**merged.get(key, {}), **entry}
        path.write_text(json.dumps(merged, indent=1))


if __name__ == "__main__":
    main()
