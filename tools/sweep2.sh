cd /root/repo
PTB_TUNE=stack_smem=4 timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "wide or sah or c4 or c3 or render_matches or uncounted or soups" > gpurun_out/pytest_spill.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_spill.log
tail -5 gpurun_out/pytest_spill.log
run() { echo "== $1 PTB_TUNE=$2"; PTB_TUNE=$2 timeout 300 python bench.py --workload $1 --steps $3 --warmup 2 --no-cpu-baseline 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('RESULT', round(d['value'],1), round(d['ms_per_step'],2), round(d['e2e']['value'],1), d['config']['pipeline'])
    else: print(l.rstrip()[-300:])
"; }
run c4 stack_smem=0 2
run c4 stack_smem=16 2
run c4 stack_smem=12 2
run c4 stack_smem=8 2
run c3 stack_smem=0 2
run c3 stack_smem=12 2
run c3 stack_smem=8 2
run c2 stack_smem=0 4
