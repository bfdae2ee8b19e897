# usage: bash tools/sweep_knobs.sh <workload> <steps> "tune1" "tune2" ...
cd "$(dirname "$0")/.."
w=$1; steps=$2; shift 2
for t in "$@"; do echo "== $w PTB_TUNE=$t"; PTB_TUNE=$t timeout 300 python bench.py --workload $w --steps $steps --warmup 2 --no-cpu-baseline 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('RESULT', round(d['value'],1), round(d['ms_per_step'],2), round(d['e2e']['value'],1), d['config']['pipeline'])
    else: print(l.rstrip()[-300:])
"; done
