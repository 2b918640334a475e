mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --maxfail=60 --timeout=300 -k "detection_loss or schedules or precount or totals" > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
grep -E "^(FAILED|ERROR)|passed|failed|rc=" gpurun_out/pytest.log | tail -30
timeout 300 python tools/stage_times.py cfg2 2>&1 | grep -A8 "detloss cfg2 rep"
timeout 600 python tools/prof_kernels.py cfg2 2>&1 | grep -E "fused_stash |main|full_step"
