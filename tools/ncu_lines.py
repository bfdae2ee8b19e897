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
    """[(offset, sass text, "inner.cu:12 <- mid.cuh:3 <- outer.cu:99")] for the kernel whose name contains kernel_sub.
    nvdisasm -gi prints the inline chain as consecutive //## lines, innermost first."""
    out = subprocess.run(["nvdisasm", "-gi", "-c", cubin], capture_output=True, text=True).stdout
    res = []
    chain, pending = [], []
    take = False
    for ln in out.splitlines():
        m = re.match(r"\s*\.text\.(\S+):", ln)
        if m:
            take = kernel_sub in m.group(1)
            continue
        if not take:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            pending.append(f"{m.group(1).split('/')[-1]}:{m.group(2)}")
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            if pending:
                chain, pending = pending, []
            res.append((int(m.group(1), 16), m.group(2).strip(), " <- ".join(chain) if chain else "?"))
    return res


def stage_of(loc, ranges):
    """Map the outermost frame (a line of the kernel's own file) to a named source range."""
    outer = loc.split(" <- ")[-1]
    m = re.match(r"([\w.]+):(\d+)", outer)
    if not m:
        return outer
    line = int(m.group(2))
    for name, lo, hi in ranges:
        if lo <= line < hi:
            return name
    return outer


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
    ranges = []
    for a in sys.argv:
        if a.startswith("--ranges="):     # e.g. --ranges=vote:190:260,NODE:260:310
            for part in a[9:].split(","):
                n, lo, hi = part.split(":")
                ranges.append((n, int(lo), int(hi)))
    if ranges:
        st = defaultdict(lambda: [0, 0, 0])
        for addr, txt, ie, tie, smp in prof:
            e = by_off.get(addr - base)
            k = stage_of(e[1] if e else "?", ranges)
            st[k][0] += ie
            st[k][1] += tie
            st[k][2] += smp
        print("per source range (outermost frame):")
        for k, a in sorted(st.items(), key=lambda kv: -kv[1][0])[:20]:
            print(f"  {k:28s} warp-inst {a[0] / 1e9:8.2f} G ({100 * a[0] / tot[0]:5.1f} %)  avg thr {a[1] / max(a[0], 1):5.2f}  samples {100 * a[2] / max(tot[2], 1):5.1f} %")
    print(f"total warp-inst {tot[0]:.4g}  thread-inst {tot[1]:.4g}  avg threads {tot[1] / max(tot[0], 1):.2f}  samples {tot[2]}  (unmapped sass: {miss})")
    print(f"{'warp-inst%':>10} {'thr-inst%':>9} {'avg thr':>7} {'samples%':>8} {'#sass':>5}  line")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{100 * a[0] / tot[0]:10.2f} {100 * a[1] / tot[1]:9.2f} {a[1] / max(a[0], 1):7.2f} {100 * a[2] / max(tot[2], 1):8.2f} {a[3]:5d}  {k}")


if __name__ == "__main__":
    main()
