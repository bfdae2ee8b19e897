cd /root/repo
run() { echo "== $*"; env "${@:2}" timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --emulate-ranks $1 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('RESULT', round(d['value'],1), round(d['ms_per_step'],2), d['config']['pipeline'])
    else: print(l.rstrip()[-300:])
"; }
run 0 PTB_WF_MISSLOOP=0
run 0 PTB_WF_MISSLOOP=1
run 0 PTB_WF_MISSLOOP=4
run 0 PTB_WF_MISSLOOP=16
run 0 PTB_WF_MISSLOOP=128
