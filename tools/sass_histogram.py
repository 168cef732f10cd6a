#!/usr/bin/env python
"""SASS opcode histogram of librsrec.so per kernel (cuobjdump -sass): the evidence that the hot kernels use the FP64 tensor
pipe (DMMA.8x8x4), TMA bulk copies (UBLKCP) and mbarriers (SYNCS), and nothing from another architecture.
Usage: python tools/sass_histogram.py > profiles/rNN_sass_opcodes.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "rslmtoasa_b200", "librsrec.so")
WATCH = ["DMMA", "DFMA", "DADD", "DMUL", "UBLKCP", "UBLKPF", "SYNCS", "LDS", "STS", "LDG", "STG", "BAR", "MUFU", "HMMA", "UTCHMMA", "UTMALDG", "LDGSTS"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = re.sub(r"\(.*", "", cur)
            per[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur:
            per[cur][m.group(1)] += 1
            per[cur]["_total"] += 1
    tot = collections.Counter()
    for c in per.values():
        tot.update(c)
    print(f"# {os.path.relpath(LIB, ROOT)}: {len(per)} kernels, {tot['_total']} SASS instructions (sm_100a)")
    print("# totals: " + "  ".join(f"{k}={tot[k]}" for k in WATCH if tot[k]))
    print(f"{'kernel':90s} {'instr':>7s} " + " ".join(f"{k:>6s}" for k in WATCH[:9]))
    for name, c in sorted(per.items(), key=lambda kv: -kv[1]["DMMA"]):
        print(f"{name[:90]:90s} {c['_total']:7d} " + " ".join(f"{c[k]:6d}" for k in WATCH[:9]))


if __name__ == "__main__":
    sys.exit(main())
