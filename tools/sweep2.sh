cd /root/repo
run() { echo "== $1 PTB_TUNE=$2"; PTB_TUNE=$2 timeout 300 python bench.py --workload $1 --steps $3 --warmup 2 --no-cpu-baseline 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('RESULT', round(d['value'],1), round(d['ms_per_step'],2), round(d['e2e']['value'],1), d['config']['pipeline'])
    else: print(l.rstrip()[-300:])
"; }
run c3 idle_ns=64 2
run c3 idle_ns=500 2
run c3 idle_ns=2000 2
run c3 block=640 2
run c3 block=512 2
run c3 tail_cols=8 2
run c3 full_cols=24 2
run c3 node_votes=12 2
