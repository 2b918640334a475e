# Round evidence: bench line, ncu launch list of the same command, full captures of the two hot kernels.
mkdir -p gpurun_out
timeout 900 python bench.py --steps 400 --warmup 40 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 30 --warmup 3 > gpurun_out/bench_reference.json 2>> gpurun_out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 16 --warmup 3 --no-graph --no-e2e --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:detloss_kernel -s 10 -c 2 -o gpurun_out/prof_detloss python bench.py --steps 8 --warmup 3 --no-graph --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode_kernel -s 10 -c 2 -o gpurun_out/prof_decode python bench.py --steps 8 --warmup 3 --no-graph --no-e2e --no-cpu-baseline > gpurun_out/ncu_full2.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:detloss_kernel -s 4 -c 1 -o gpurun_out/prof_detloss_cfg5 python bench.py --config cfg5 --steps 4 --warmup 3 --no-graph --no-e2e --no-cpu-baseline > gpurun_out/ncu_full3.log 2>&1
cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
