mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode_cluster -s 4 -c 1 -f -o gpurun_out/r02_decode_cfg5 python bench.py --config cfg5 --steps 4 --warmup 3 --no-graph --no-e2e --no-cpu-baseline --no-extra > gpurun_out/ncu_dec5.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu_dec5.log
