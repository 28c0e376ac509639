#!/usr/bin/env python
"""Aggregate an ncu SASS source page by CUDA source line.
usage: ncu_by_line.py <report.ncu-rep> <lib.so> <kernel-substring> [top]
Joins `ncu --page source --csv` (per-SASS-instruction counters, in program order) with
`nvdisasm -g` line info of the same function (same instruction order)."""
import csv, io, re, subprocess, sys, tempfile, os, glob, collections
rep, lib, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ci = {h: i for i, h in enumerate(hdr)}
inst = rows[hi + 1:]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
lines = None
for cub in glob.glob(tmp + "/*.cubin"):
    txt = subprocess.run(["nvdisasm", "-g", "-c", cub], capture_output=True, text=True).stdout
    secs = re.split(r"//-+ \.text\.", txt)
    for s in secs[1:]:
        name = s.split(" ", 1)[0]
        if kern in name:
            n_inst = len(re.findall(r"^\s+/\*[0-9a-f]{4,}\*/", s, flags=re.M))
            if n_inst == len(inst):
                lines = s
                break
    if lines:
        break
if lines is None:
    sys.exit("no function with %d instructions matching %r" % (len(inst), kern))
cur = ("?", 0)
per_inst = []
for l in lines.splitlines():
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"^\s+/\*[0-9a-f]{4,}\*/", l):
        per_inst.append(cur)
agg = collections.defaultdict(lambda: [0, 0, 0, 0, 0])
tot = [0, 0, 0, 0, 0]


def num(r, key):
    try:
        return int(float(r[ci[key]] or 0)) if key in ci else 0
    except ValueError:
        return 0


for r, loc in zip(inst, per_inst):
    ex = num(r, "Instructions Executed")
    smp = num(r, "Warp Stall Sampling (All Samples)")
    wv = num(r, "L1 Wavefronts Shared")
    wx = num(r, "L1 Wavefronts Shared Excessive")
    agg[loc][0] += ex; agg[loc][1] += smp; agg[loc][2] += 1; agg[loc][3] += wv; agg[loc][4] += wx
    tot[0] += ex; tot[1] += smp; tot[3] += wv; tot[4] += wx
print(f"total warp-instructions {tot[0]}, stall samples {tot[1]}, SASS instructions {len(inst)}, "
      f"shared wavefronts {tot[3]} (excessive = bank conflicts: {tot[4]})")
# BY_INST=1: sort by executed instructions; BY_CONFLICT=1: by excessive shared-memory wavefronts; default: stall samples
key = 0 if os.environ.get("BY_INST") else (4 if os.environ.get("BY_CONFLICT") else 1)
print(f"{'file:line':32s} {'inst%':>7s} {'samples%':>9s} {'sass':>6s} {'smem wavefronts':>16s} {'excessive':>10s}")
for loc, v in sorted(agg.items(), key=lambda kv: -kv[1][key])[:top]:
    print(f"{loc[0]+':'+str(loc[1]):32s} {100*v[0]/max(tot[0],1):7.2f} {100*v[1]/max(tot[1],1):9.2f} {v[2]:6d} {v[3]:16d} {v[4]:10d}")
