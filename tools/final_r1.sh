# Round-end evidence run (one B200): full GPU test suite, smoke, bench (both arms), other workloads, launch list, ncu full capture.
cd /root/repo
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_final.log
tail -4 gpurun_out/pytest_final.log
timeout 300 python -c "import sys; sys.path.insert(0, 'tests'); import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1; tail -2 gpurun_out/smoke_final.log
timeout 600 python bench.py > gpurun_out/bench_final_n1.json 2> gpurun_out/bench_final_n1.err; tail -c 300 gpurun_out/bench_final_n1.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_final_ref.json 2> gpurun_out/bench_final_ref.err; tail -c 300 gpurun_out/bench_final_ref.json
for w in c1 c3 c4 c5; do timeout 600 python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_final_$w.json 2> gpurun_out/bench_final_$w.err; python -c "
import json,sys
d=json.loads(open('gpurun_out/bench_final_$w.json').read().strip().splitlines()[-1]); print('$w', round(d['value'],1), round(d['ms_per_step'],3), d['config']['pipeline'], d['config']['lbvh_build_ms'])"; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1_final.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_l_final.log 2>&1; grep -c render_wavefront gpurun_out/launches_r1_final.csv
