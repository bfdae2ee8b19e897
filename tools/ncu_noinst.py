#!/usr/bin/env python
"""Where the instruction-fetch stalls are: python tools/ncu_noinst.py report.ncu-rep [bucket_bytes]
Prints, per code-address bucket, warp instructions executed and stall_no_inst / all stall samples (ncu --page source)."""
import csv, subprocess, sys
rep = sys.argv[1]
bucket = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[1]
ia, ii, isamp, ino = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("stall_no_inst")
base = None
agg = {}
tot = [0, 0, 0]
n_exec = 0
for r in rows[2:]:
    if len(r) <= ino:
        continue
    a = int(r[ia], 16)
    if base is None:
        base = a
    b = (a - base) // bucket
    x = agg.setdefault(b, [0, 0, 0, 0])
    ie, s, no = int(r[ii] or 0), int(r[isamp] or 0), int(r[ino] or 0)
    x[0] += ie; x[1] += s; x[2] += no; x[3] += 1 if ie else 0
    tot[0] += ie; tot[1] += s; tot[2] += no
    n_exec += 1 if ie else 0
print(f"total inst {tot[0]:.3e} samples {tot[1]} no_inst {tot[2]} ({100*tot[2]/max(tot[1],1):.1f} %), SASS instructions executed at least once: {n_exec} of {len(rows)-2}")
for b in sorted(agg):
    x = agg[b]
    if x[0]:
        print(f"  +{b*bucket:6d}: inst {x[0]:.3e} ({100*x[0]/tot[0]:5.1f} %)  samples {x[1]:7d}  no_inst {x[2]:6d} ({100*x[2]/max(x[1],1):4.1f} %)  live instrs {x[3]}")
