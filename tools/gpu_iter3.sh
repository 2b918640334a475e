mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --maxfail=60 --timeout=300 -k "decode or full_size" > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
grep -E "^(FAILED|ERROR)|passed|failed|rc=|^E  " gpurun_out/pytest.log | tail -30
timeout 300 python tools/stage_times.py cfg2 2>&1 | grep -A16 "decode cfg2 rep1"
timeout 300 python tools/stage_times.py cfg5 2>&1 | grep -A16 "decode cfg5 rep1"
timeout 600 python tools/prof_kernels.py cfg2 cfg5 2>&1 | grep -E "decode|full_step"
