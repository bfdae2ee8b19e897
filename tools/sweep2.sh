cd /root/repo
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_ploc.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_ploc.log
tail -5 gpurun_out/pytest_ploc.log
run() { echo "== $1 $2"; timeout 300 python bench.py --workload $1 --steps $3 --warmup 2 --no-cpu-baseline $2 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('RESULT', round(d['value'],1), round(d['ms_per_step'],2), round(d['e2e']['value'],1), d['config']['pipeline'], d['counters']['n_node'], 'build_ms', d['config']['lbvh_build_ms'])
    else: print(l.rstrip()[-300:])
"; }
run c4 "--tree 1" 2
run c4 "--tree 3" 2
run c4 "" 2
run c3 "--tree 3" 2
run c2 "--tree 3" 4
