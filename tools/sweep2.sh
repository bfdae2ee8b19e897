cd /root/repo
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "not c4 and not c5 and not converged" > gpurun_out/pytest_sah.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_sah.log
tail -5 gpurun_out/pytest_sah.log
run() { echo "== $1 $2"; timeout 300 python bench.py --workload $1 --steps $3 --warmup 3 --no-cpu-baseline $2 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('RESULT', round(d['value'],1), round(d['ms_per_step'],2), round(d['e2e']['value'],1), d['config']['pipeline'], d['counters']['n_node'], d['roofline']['frac'])
    else: print(l.rstrip()[-300:])
"; }
run c2 "--tree 1" 8
run c2 "--tree 2" 8
run c3 "--tree 1" 2
run c3 "--tree 2" 2
run c1 "--tree 1" 8
run c1 "--tree 2" 8
