cd /root/repo
run() { echo "== $*"; env "${@:2}" timeout 600 python bench.py --workload $1 --steps 2 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('RESULT', round(d['value'],1), 'Mps/s', round(d['ms_per_step'],2),'ms', d['config']['pipeline'], 'tris', d['config']['triangles'])
    else: print(l.rstrip()[-300:])
"; }
run c4 A=1
run c4 PTB_WF_BLOCK=640
run c4 PTB_WF_BLOCK=736
run c4 PTB_WF_SMEM_KB=224
run c4 PTB_WF_SMEM_KB=224 PTB_WF_BLOCK=768
run c3 A=1
run c3 PTB_WF_SMEM_KB=224
