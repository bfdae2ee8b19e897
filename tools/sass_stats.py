#!/usr/bin/env python
"""Static SASS statistics of one kernel: python tools/sass_stats.py lib.so <substring of the mangled name> [--hist]"""
import collections, re, subprocess, sys
lib, key = sys.argv[1], sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
name, ops = None, None
res = []
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        if name and ops is not None:
            res.append((name, ops))
        name, ops = m.group(1), []
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
    if m and ops is not None:
        ins = m.group(1)
        ins = re.sub(r"^@!?U?P\d+\s+", "", ins)
        ops.append(ins.split()[0])
if name and ops is not None:
    res.append((name, ops))
for name, ops in res:
    if key in name:
        c = collections.Counter(o.split(".")[0] for o in ops)
        print(name[-70:], "inst", len(ops), "LDL", c["LDL"], "STL", c["STL"], "LDS", c["LDS"], "STS", c["STS"], "BSSY", c["BSSY"], "BRA", c["BRA"], "CALL", c["CALL"])
        if "--hist" in sys.argv:
            print("   ", c.most_common(25))
