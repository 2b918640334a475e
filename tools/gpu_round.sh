# Full round evidence in one box: parity tests, smoke, bench (both arms), ncu launch list + full captures.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q --maxfail=40 --timeout=300 > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
tail -5 gpurun_out/pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
timeout 900 python bench.py --steps 400 --warmup 40 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 30 --warmup 3 > gpurun_out/bench_reference.json 2>> gpurun_out/bench.err
timeout 600 python tools/prof_kernels.py cfg2 cfg3 cfg4 cfg5 > gpurun_out/prof_kernels.log 2>&1; tail -40 gpurun_out/prof_kernels.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 16 --warmup 3 --no-graph --no-e2e --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:detloss -s 10 -c 2 -f -o gpurun_out/prof_detloss python bench.py --steps 8 --warmup 3 --no-graph --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode -s 10 -c 2 -f -o gpurun_out/prof_decode python bench.py --steps 8 --warmup 3 --no-graph --no-e2e --no-cpu-baseline > gpurun_out/ncu_full2.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:decode -s 4 -c 1 -f -o gpurun_out/prof_decode_cfg5 python bench.py --config cfg5 --steps 4 --warmup 3 --no-graph --no-e2e --no-cpu-baseline > gpurun_out/ncu_full4.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:detloss -s 4 -c 1 -f -o gpurun_out/prof_detloss_cfg5 python bench.py --config cfg5 --steps 4 --warmup 3 --no-graph --no-e2e --no-cpu-baseline > gpurun_out/ncu_full3.log 2>&1
cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
