mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_parity.py tests/test_callers.py -m gpu -q --maxfail=10 --timeout=120 2>&1 | tail -3
timeout 200 python tools/dec_time.py cfg2 cfg5 2>&1 | tail -1
timeout 200 python tools/stage_times.py cfg5 2>&1 | grep -A12 "decode cfg5 rep1" | cut -c1-150
