mkdir -p gpurun_out
timeout 600 python tools/prof_kernels.py cfg4 2>&1 | grep -E "entropy|max_square|full_step|decode|fused_stash "
timeout 600 python bench.py --config cfg4 --steps 400 --warmup 40 --no-cpu-baseline > gpurun_out/bench_cfg4.json 2> gpurun_out/bench_cfg4.err; echo "bench rc=$?"; cut -c1-900 gpurun_out/bench_cfg4.json; tail -5 gpurun_out/bench_cfg4.err
