#!/usr/bin/env python
"""Per-kernel SASS evidence: counts of the Blackwell-native mnemonics (tcgen05 MMA = UTCHMMA, TMEM loads/stores =
LDTM/STTM, tcgen05.commit / mbarrier = UTCBAR / SYNCS, bulk async copies = UBLKCP, plus FFMA / REDG / ATOMG / LDS / LDG
for context) in the in-tree libhoir_b200.so.   python profiles/sass_summary.py > profiles/sass_summary.txt
(the mnemonics that prove tcgen05 / TMA are listed in /opt/skills/guides/B200_PROFILING.md)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "high-order-implicit-representation_b200", "libhoir_b200.so")
KEYS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTCBAR", "UTCCP", "UBLKCP", "UTMALDG", "SYNCS", "FFMA", "FMUL", "FADD",
        "MUFU", "REDG", "RED", "ATOMG", "ATOMS", "LDS", "STS", "LDG", "STG", "LDGSTS", "BAR", "SHFL", "BRX"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    demangle = {}
    kernels = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m:
            op = m.group(1)
            kernels[cur]["_total"] += 1
            kernels[cur][op] += 1
    names = list(kernels)
    out = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()
    for n, d in zip(names, out):
        demangle[n] = d
    print(f"# {os.path.relpath(LIB, ROOT)}: {len(kernels)} kernels; columns = SASS instruction counts (static)")
    print("# " + " ".join(f"{k:>7s}" for k in ["total"] + KEYS) + "  kernel")
    tot = collections.Counter()
    for n, c in kernels.items():
        row = [c["_total"]] + [c[k] for k in KEYS]
        for k in KEYS:
            tot[k] += c[k]
        name = re.sub(r"\(anonymous namespace\)::", "", demangle.get(n, n))
        name = re.sub(r"^void ", "", name)
        print("  " + " ".join(f"{v:7d}" for v in row) + "  " + name[:150])
    print("# library totals: " + ", ".join(f"{k} {tot[k]}" for k in KEYS if tot[k]))


if __name__ == "__main__":
    sys.exit(main())
