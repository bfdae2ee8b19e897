cd /root/repo
run() { echo "== PTB_WF=$1"; PTB_WF=$1 timeout 120 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('RESULT', round(d['value'],1), round(d['ms_per_step'],2), round(d['e2e']['value'],1), d['config']['pipeline']); print('SCHED', json.dumps(d.get('sched_rank0')))
    else: print(l.rstrip()[-300:])
"; }
for a in "$@"; do run "$a"; done
