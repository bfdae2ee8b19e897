cd /root/repo
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "wide or uncounted or c4 or c3 or render_matches or traversal_bit" > gpurun_out/pytest_wide.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_wide.log
tail -5 gpurun_out/pytest_wide.log
run() { echo "== $1 PTB_TUNE=$2"; PTB_TUNE=$2 timeout 300 python bench.py --workload $1 --steps 2 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('RESULT', round(d['value'],1), round(d['ms_per_step'],2), d['config']['pipeline'])
    else: print(l.rstrip()[-300:])
"; }
run c3 wide=0
run c3 wide=1
run c4 wide=0
run c4 wide=1
run c4 wide=1,smem_kb=220
