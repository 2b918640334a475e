mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=10 --timeout=120 > gpurun_out/pytest_dec.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/pytest_dec.log
timeout 200 python tools/stage_times.py cfg2 2>&1 | grep -A12 "decode cfg2 rep1" | cut -c1-180
timeout 200 python tools/stage_times.py cfg5 2>&1 | grep -A12 "decode cfg5 rep1" | cut -c1-180
timeout 600 python tools/prof_kernels.py cfg2 cfg5 2>&1 | grep -E "decode|full_step|scale|fused_stash "
timeout 600 python bench.py --steps 400 --warmup 40 > gpurun_out/bench_dec.json 2> gpurun_out/bench_dec.err; echo "bench rc=$?"; cat gpurun_out/bench_dec.json; tail -5 gpurun_out/bench_dec.err
