# Round-2 evidence in one box (1 GPU): parity tests, smoke, both bench arms, per-kernel timings, ncu launch list + full captures.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q --maxfail=40 --timeout=300 -rs -s > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
tail -4 gpurun_out/pytest.log; grep "\[elementwise\|\[chain" gpurun_out/pytest.log | cut -c1-260
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py --steps 400 --warmup 40 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 30 --warmup 5 > gpurun_out/bench_reference.json 2>> gpurun_out/bench.err
timeout 600 python bench.py --config cfg5 --steps 200 --warmup 20 --no-extra --no-e2e --no-cpu-baseline > gpurun_out/bench_cfg5.json 2>> gpurun_out/bench.err
timeout 600 python tools/prof_kernels.py cfg2 cfg5 > gpurun_out/prof_kernels.log 2>&1; tail -24 gpurun_out/prof_kernels.log | cut -c1-200
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 16 --warmup 3 --no-graph --no-e2e --no-cpu-baseline --no-extra > gpurun_out/ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:detloss_stash -s 10 -c 2 -f -o gpurun_out/prof_detloss python bench.py --steps 8 --warmup 3 --no-graph --no-e2e --no-cpu-baseline --no-extra > gpurun_out/ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_cluster -s 10 -c 2 -f -o gpurun_out/prof_decode python bench.py --steps 8 --warmup 3 --no-graph --no-e2e --no-cpu-baseline --no-extra > gpurun_out/ncu_full2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_stream -s 4 -c 1 -f -o gpurun_out/prof_decode_cfg5 python bench.py --config cfg5 --steps 4 --warmup 3 --no-graph --no-e2e --no-cpu-baseline --no-extra > gpurun_out/ncu_full4.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:detloss_stream -s 4 -c 1 -f -o gpurun_out/prof_detloss_cfg5 python bench.py --config cfg5 --steps 4 --warmup 3 --no-graph --no-e2e --no-cpu-baseline --no-extra > gpurun_out/ncu_full3.log 2>&1
cat gpurun_out/bench.json | cut -c1-600
