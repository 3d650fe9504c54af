#!/usr/bin/env python
"""Opcode histogram of kernels inside the built library (no GPU needed): the evidence behind the SASS mnemonics
DESIGN.md quotes (UBLKCP = cp.async.bulk / TMA, UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG = TMA tensor loads).

    python tools/sass_histogram.py <kernel-name regex> [--lib graphax_b200/libgraphax_b200.so] [--top 40]
"""
import argparse
import collections
import re
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("pattern")
    ap.add_argument("--lib", default=str(ROOT / "graphax_b200" / "libgraphax_b200.so"))
    ap.add_argument("--top", type=int, default=60)
    args = ap.parse_args()
    sass = subprocess.run(["cuobjdump", "-sass", args.lib], capture_output=True, text=True, check=True).stdout
    kernels, name = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.match(r"\s+Function : (\S+)", line)
        if m:
            name = m.group(1)
            kernels[name] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and name:
            kernels[name][m.group(2)] += 1
    for kname, ops in kernels.items():
        if not re.search(args.pattern, kname):
            continue
        total = sum(ops.values())
        base = collections.Counter()
        for op, n in ops.items():
            base[op.split(".")[0]] += n
        print(f"== {kname}: {total} SASS instructions")
        print("   by mnemonic: " + ", ".join(f"{op} {n}" for op, n in base.most_common(args.top)))
        marks = [op for op in ops if re.match(r"(UBLKCP|UTMALDG|UTMASTG|UTCHMMA|UTCBAR|LDTM|STTM|SYNCS|ELECT|FMNMX3|REDG|ACQBULK)", op)]
        print("   Blackwell-specific forms: " + (", ".join(f"{op} x{ops[op]}" for op in sorted(marks)) or "none"))


if __name__ == "__main__":
    main()
