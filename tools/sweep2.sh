cd /root/repo
run() { echo "== $1 $2"; timeout 300 python bench.py --workload $1 --steps 2 --warmup 2 --no-cpu-baseline $2 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('RESULT', round(d['value'],1), round(d['ms_per_step'],2), round(d['e2e']['value'],1), d['config']['pipeline'])
    else: print(l.rstrip()[-300:])
"; }
run c3 "--pipeline 1"
run c4 "--pipeline 1"
run c2 "--pipeline 1"
