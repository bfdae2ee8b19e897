#!/usr/bin/env python
"""Writes profiles/facts.json: per workload the per-launch DRAM bytes, warp instructions and kernel time of the committed
`ncu --set full` summary (profiles/*_ncu_summary.txt, produced by tools/ncu_keys.py).  bench.py reads the file for
roofline.traffic and the issue-slot roofline, so those numbers are reproducible from profiles/ by arithmetic alone.

    python tools/ncu_facts.py c2=profiles/r2_c2_ncu_summary.txt c3=... c4=...
"""
import json
import os
import sys

UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def parse(path):
    vals = {}
    for line in open(path):
        parts = line.split()
        if len(parts) < 2 or parts[0] == "Kernel":
            continue
        name, value = parts[0], parts[-1]
        unit = parts[1] if len(parts) >= 3 else ""
        try:
            vals[name] = (float(value.replace(",", "")), unit)
        except ValueError:
            pass
    dram = sum(vals[k][0] * UNIT[vals[k][1]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    ms = vals["gpu__time_duration.sum"]
    ms = ms[0] * {"ms": 1.0, "us": 1e-3, "s": 1e3, "ns": 1e-6}[ms[1]]
    return {"file": os.path.relpath(path, ROOT), "dram_bytes": int(round(dram)), "warp_inst": int(vals["smsp__inst_executed.sum"][0]),
            "ncu_kernel_ms": ms, "threads_per_inst": vals["smsp__thread_inst_executed_per_inst_executed.ratio"][0],
            "issue_active_pct": vals["smsp__issue_active.avg.pct_of_peak_sustained_active"][0]}


def main():
    out_path = os.path.join(ROOT, "profiles", "facts.json")
    facts = json.load(open(out_path)) if os.path.exists(out_path) else {}
    for arg in sys.argv[1:]:
        key, path = arg.split("=", 1)
        facts[key] = parse(os.path.join(ROOT, path) if not os.path.isabs(path) else path)
    json.dump(facts, open(out_path, "w"), indent=1, sort_keys=True)
    print(json.dumps(facts, indent=1, sort_keys=True))


if __name__ == "__main__":
    main()
