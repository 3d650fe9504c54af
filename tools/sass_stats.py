#!/usr/bin/env python
"""Compile a plan-specialised kernel here (no GPU) and print its SASS instruction mix.

    python tools/sass_stats.py roe3d mixed [--min-blocks 3] [--threads 128] [--keep /tmp/k.cu]

Counts are for the whole kernel: with checked fast math the body appears twice (fast path + exact fallback).
"""
import argparse
import collections
import re
import subprocess
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("workload")
    ap.add_argument("order")
    ap.add_argument("--min-blocks", type=int, default=0)
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--keep", default="")
    ap.add_argument("--rounding", default="fma", choices=["fma", "reference"])
    ap.add_argument("--checked-fast", type=int, default=-1)
    args = ap.parse_args()
    from graphax_b200 import codegen
    from graphax_b200.configs import get_workload
    plan = get_workload(args.workload).plan(args.order, rounding=args.rounding)
    if args.checked_fast >= 0:
        import os
        os.environ["GX_CHECKED_FAST"] = str(args.checked_fast)
    src = codegen.emit_source(plan, nvrtc=False, threads=args.threads, min_blocks=args.min_blocks)
    with tempfile.TemporaryDirectory() as d:
        cu = Path(d) / "k.cu"
        cu.write_text(src)
        if args.keep:
            Path(args.keep).write_text(src)
        out = subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-Xptxas", "-v",
                              "-I", str(codegen.CSRC), "-cubin", "-o", str(Path(d) / "k.cubin"), str(cu)],
                             capture_output=True, text=True)
        print("\n".join(l for l in out.stderr.splitlines() if "registers" in l or "spill" in l or "error" in l))
        sass = subprocess.run(["cuobjdump", "-sass", str(Path(d) / "k.cubin")], capture_output=True, text=True).stdout
    ops = collections.Counter()
    for line in sass.splitlines():
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            ops[m.group(2).split(".")[0]] += 1
    total = sum(ops.values())
    fp = {k: ops[k] for k in ("FMUL", "FFMA", "FADD", "FMUL2", "FFMA2", "FADD2", "MOV", "IMAD", "PRMT", "MUFU", "STS", "LDS")}
    print(f"plan ops {len(plan.ssa)}  SASS instructions {total}  {fp}")
    print("  fp issue slots:", sum(ops[k] for k in ("FMUL", "FFMA", "FADD", "FMUL2", "FFMA2", "FADD2")),
          " fp lane-ops:", sum(ops[k] for k in ("FMUL", "FFMA", "FADD")) + 2 * sum(ops[k] for k in ("FMUL2", "FFMA2", "FADD2")))


if __name__ == "__main__":
    main()
