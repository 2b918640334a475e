mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode_cluster -s 4 -c 1 -f -o gpurun_out/prof_decode_cluster_cfg5 python bench.py --config cfg5 --steps 4 --warmup 3 --no-graph --no-e2e --no-cpu-baseline > gpurun_out/ncu_dc5.log 2>&1; tail -2 gpurun_out/ncu_dc5.log | cut -c1-300
