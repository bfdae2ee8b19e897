#!/bin/bash
# A/B helper for gpurun: runs bench.py for a list of "label|dir|env|args" cases and prints one compact line per case.
# usage: tools/ab_run.sh out_prefix  case...     (dir relative to the repo root; env like PTB_TUNE=team=0 or "-")
out=$1; shift
for c in "$@"; do
  IFS='|' read -r label dir envs args <<< "$c"
  f=gpurun_out/${out}_${label}.json
  ( cd "$dir" && if [ "$envs" != "-" ]; then export $envs; fi; python bench.py $args --no-cpu-baseline > "$OLDPWD/$f" 2> "$OLDPWD/${f%.json}.err" )
  python - "$f" "$label" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = d.get("roofline") or {}
    print(f"{sys.argv[2]:28s} value {d['value']:9.1f}  ms/step {d['ms_per_step']:8.3f}  e2e {d['e2e']['value'] if d.get('e2e') else 0:9.1f}  clk {d['clocks']['sm_mhz'] if d.get('clocks') else None}  launch {d['config']['pipeline']}")
except Exception as e:
    print(f"{sys.argv[2]:28s} FAILED {e}")
PY
done
