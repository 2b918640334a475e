mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=10 --timeout=300 -x > gpurun_out/pytest_t1.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_t1.log
timeout 300 python tools/stage_times.py cfg2 > gpurun_out/stage_cfg2.log 2>&1; echo "stage2 rc=$?"
head -12 gpurun_out/stage_cfg2.log
timeout 600 python tools/prof_kernels.py cfg2 cfg5 > gpurun_out/prof_kernels.log 2>&1; echo "prof rc=$?"
grep -E "fused|fwd_only|count|main|scale|full" gpurun_out/prof_kernels.log
