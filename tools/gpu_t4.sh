mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=10 --timeout=300 -x > gpurun_out/pytest_t4.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_t4.log
timeout 300 python tools/stage_times.py cfg2 2>&1 | head -12 | cut -c1-200
timeout 600 python tools/prof_kernels.py cfg2 2>&1 | grep -E "fused_stash |full_step|fwd_only|main"
for d in 200 400 800; do echo "--- x delay $d"; CNH_X_DELAY_NS=$d timeout 600 python tools/prof_kernels.py cfg2 2>&1 | grep -E "fused_stash "; done
