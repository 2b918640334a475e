mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --maxfail=60 --timeout=300 > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
grep -E "^(FAILED|ERROR)|passed|failed|rc=|^E  " gpurun_out/pytest.log | tail -20
timeout 600 python tools/prof_kernels.py cfg2 cfg5 2>&1 | tail -22
