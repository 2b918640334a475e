mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=10 --timeout=120 > gpurun_out/pytest_q.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_q.log
timeout 600 python tools/prof_kernels.py cfg1 cfg2 2>&1 | grep -E "fwd_only|fused_stash |full_step"
