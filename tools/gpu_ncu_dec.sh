mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode_cluster -s 10 -c 1 -f -o gpurun_out/prof_decode_cluster python bench.py --steps 8 --warmup 3 --no-graph --no-e2e --no-cpu-baseline > gpurun_out/ncu_dc.log 2>&1; tail -3 gpurun_out/ncu_dc.log
