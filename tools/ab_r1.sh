cd /root/repo
run() { echo "== $*"; env "$@" timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('RESULT', round(d['value'],1), round(d['ms_per_step'],2), round(d['e2e']['value'],1), d['config']['pipeline'])
    else: print(l.rstrip()[-300:])
"; }
run PTB_TILE_ORDER=0
run PTB_TILE_ORDER=1
