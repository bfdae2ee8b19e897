cd /root/repo
run() { echo "== $*"; env "${@:2}" timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --emulate-ranks $1 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('RESULT', round(d['value'],1), round(d['ms_per_step'],2), d['config']['pipeline'], {k:round(v['lanes_per_batch'],1) for k,v in d['sched_rank0'].items() if isinstance(v,dict)})
    else: print(l.rstrip()[-300:])
"; }
run 8 PTB_WF_TAIL=0
run 8 PTB_WF_TAIL=8
run 8 PTB_WF_TAIL=24
run 8 PTB_WF_TAIL=40
run 8 PTB_WF_TAIL=64
run 8 PTB_WF_TAIL=96
run 0 PTB_WF_TAIL=24
run 0 PTB_WF_TAIL=64
