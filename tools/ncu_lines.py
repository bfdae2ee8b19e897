#!/usr/bin/env python
"""Attribute an ncu SASS-level profile to source lines.

    python tools/ncu_lines.py <report.ncu-rep> <cubin> <kernel-substring> [top]

Reads `ncu --page source --csv` (per-SASS-instruction counters), takes line info from
`nvdisasm -g` of the same cubin, and prints per source line: warp instructions executed,
thread instructions executed, average active threads, and stall samples.
"""
import csv
import re
import subprocess
import sys
from collections import defaultdict


def sass_lines(cubin, kernel_sub):
    out = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
    res = []          # per instruction: (file:line chain)
    cur_fn = None
    loc = "?"
    take = False
    for ln in out.splitlines():
        m = re.match(r"\s*\.text\.(\S+):", ln)
        if m:
            cur_fn = m.group(1)
            take = kernel_sub in cur_fn
            continue
        if not take:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
        if m:
            f = m.group(1).split("/")[-1]
            loc = f"{f}:{m.group(2)}"
            inl = re.findall(r'inlined at "([^"]+)", line (\d+)', m.group(3))
            if inl:
                loc += " <- " + " <- ".join(f"{a.split('/')[-1]}:{b}" for a, b in inl)
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            res.append((int(m.group(1), 16), m.group(2).strip(), loc))
    return res


def main():
    rep, cubin, ksub = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    col = {h: i for i, h in enumerate(hdr)}
    prof = []
    for r in rows[hdr_i + 1:]:
        if len(r) < len(hdr) or not r[0].startswith("0x"):
            continue
        prof.append((int(r[0], 16), r[col["Source"]].strip(), int(r[col["Instructions Executed"]] or 0),
                     int(r[col["Thread Instructions Executed"]] or 0), int(r[col["# Samples"]] or 0)))
    base = prof[0][0]
    sl = sass_lines(cubin, ksub)
    by_off = {o: (txt, loc) for o, txt, loc in sl}
    agg = defaultdict(lambda: [0, 0, 0, 0])
    tot = [0, 0, 0]
    miss = 0
    for addr, txt, ie, tie, smp in prof:
        e = by_off.get(addr - base)
        loc = e[1] if e else "?"
        if not e:
            miss += 1
        outer = loc.split(" <- ")[-1]
        key = loc if "--full" in sys.argv else loc.split(" <- ")[0] + ("  @" + outer if " <- " in loc else "")
        a = agg[key]
        a[0] += ie
        a[1] += tie
        a[2] += smp
        a[3] += 1
        tot[0] += ie
        tot[1] += tie
        tot[2] += smp
    print(f"total warp-inst {tot[0]:.4g}  thread-inst {tot[1]:.4g}  avg threads {tot[1] / max(tot[0], 1):.2f}  samples {tot[2]}  (unmapped sass: {miss})")
    print(f"{'warp-inst%':>10} {'thr-inst%':>9} {'avg thr':>7} {'samples%':>8} {'#sass':>5}  line")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{100 * a[0] / tot[0]:10.2f} {100 * a[1] / tot[1]:9.2f} {a[1] / max(a[0], 1):7.2f} {100 * a[2] / max(tot[2], 1):8.2f} {a[3]:5d}  {k}")


if __name__ == "__main__":
    main()
