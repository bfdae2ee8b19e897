# Round-end evidence run (one B200): full GPU test suite, smoke, bench (both arms), launch list, ncu full capture.
cd /root/repo
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_final.log
tail -4 gpurun_out/pytest_final.log
timeout 300 python -c "import sys; sys.path.insert(0, 'tests'); import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1; tail -2 gpurun_out/smoke_final.log
timeout 600 python bench.py > gpurun_out/bench_final_n1.json 2> gpurun_out/bench_final_n1.err; tail -c 600 gpurun_out/bench_final_n1.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_final_ref.json 2> gpurun_out/bench_final_ref.err; tail -c 400 gpurun_out/bench_final_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1_sah.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_l_sah.log 2>&1; tail -3 gpurun_out/launches_r1_sah.csv
timeout 800 ncu --set full --clock-control none --import-source on -k regex:render_wavefront -c 1 -s 1 -o gpurun_out/prof_c2_sah python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_c2_sah.log 2>&1; ls -la gpurun_out/prof_c2_sah.ncu-rep
