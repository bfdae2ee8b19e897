# usage: bash tools/sweep_r1.sh "node_votes=6,tail_cols=40" "shortlist=0" ...   (one bench run per PTB_TUNE string)
cd "$(dirname "$0")/.."
run() { echo "== PTB_TUNE=$1"; PTB_TUNE=$1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline ${BENCH_ARGS:-} 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('RESULT', round(d['value'],1), round(d['ms_per_step'],2), round(d['e2e']['value'],1), d['config']['pipeline']); print('SCHED', json.dumps(d.get('sched_rank0')))
    else: print(l.rstrip()[-300:])
"; }
for a in "$@"; do run "$a"; done
