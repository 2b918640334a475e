mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q --maxfail=10 --timeout=60 -x > gpurun_out/pytest_t5.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/pytest_t5.log
timeout 300 python tools/stage_times.py cfg5 2>&1 | head -8 | cut -c1-200
timeout 600 python tools/prof_kernels.py cfg2 cfg5 2>&1 | grep -E "fused_stash |fused_precount|full_step|fwd_only|main|count"
