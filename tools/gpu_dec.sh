mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "decode" --maxfail=10 --timeout=120 > gpurun_out/pytest_dec.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_dec.log
timeout 200 python tools/stage_times.py cfg2 2>&1 | grep -A11 "decode cfg2 rep1" | cut -c1-180
timeout 600 python tools/prof_kernels.py cfg2 cfg5 2>&1 | grep -E "decode|full_step"
