#!/usr/bin/env python
"""Hot code footprint per stage: python tools/ncu_hot.py report.ncu-rep cubin kernel-substring name:lo:hi,...
(line ranges of the kernel's own source file).  A SASS instruction is "hot" when it is executed at least 1/1000 as often
as the most executed one."""
import csv
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from ncu_lines import sass_lines, stage_of

rep, cubin, ksub, rng = sys.argv[1:5]
ranges = [(n, int(lo), int(hi)) for n, lo, hi in (p.split(":") for p in rng.split(","))]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
col = {h: i for i, h in enumerate(rows[hi])}
prof = [(int(r[0], 16), int(r[col["Instructions Executed"]] or 0), int(r[col["# Samples"]] or 0), int(r[col["stall_no_inst"]] or 0))
        for r in rows[hi + 1:] if r and r[0].startswith("0x")]
base = prof[0][0]
by_off = {o: loc for o, txt, loc in sass_lines(cubin, ksub)}
mx = max(p[1] for p in prof)
st = {}
for a, ie, smp, noi in prof:
    k = stage_of(by_off.get(a - base, "?"), ranges)
    x = st.setdefault(k, [0, 0, 0, 0, 0])
    x[0] += 1
    x[1] += 1 if ie * 1000 >= mx else 0
    x[2] += ie
    x[3] += smp
    x[4] += noi
tot = [sum(v[i] for v in st.values()) for i in range(5)]
print(f"{'stage':24s} {'sass':>6} {'hot':>6} {'warp-inst %':>11} {'samples %':>9} {'no_inst % of its samples':>10}")
for k, v in sorted(st.items(), key=lambda kv: -kv[1][2]):
    print(f"{k:24s} {v[0]:6d} {v[1]:6d} {100*v[2]/tot[2]:11.1f} {100*v[3]/tot[3]:9.1f} {100*v[4]/max(v[3],1):10.1f}")
print(f"{'total':24s} {tot[0]:6d} {tot[1]:6d}   hot bytes {tot[1]*16}")

if len(sys.argv) > 5:      # per innermost source line of one stage: hot SASS count and executed share
    want = sys.argv[5]
    per = {}
    for a, ie, smp, noi in prof:
        loc = by_off.get(a - base, "?")
        if stage_of(loc, ranges) != want:
            continue
        inner = loc.split(" <- ")[0] + "  @" + loc.split(" <- ")[-1]
        x = per.setdefault(inner, [0, 0, 0])
        x[0] += 1
        x[1] += 1 if ie * 1000 >= mx else 0
        x[2] += ie
    print(f"-- {want}: innermost line @ kernel line: sass / hot / warp-inst %")
    for k, v in sorted(per.items(), key=lambda kv: -kv[1][1])[:45]:
        print(f"   {v[0]:5d} {v[1]:5d} {100*v[2]/tot[2]:6.2f}  {k}")
