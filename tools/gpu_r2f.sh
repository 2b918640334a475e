# streaming decode: parity on every launch shape, timings
mkdir -p gpurun_out
timeout 420 python -m pytest tests/test_gpu_parity.py -m gpu -v -k "decode or full_size" --maxfail=10 --timeout=120 > gpurun_out/pytest_dec.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/pytest_dec.log | cut -c1-200
( CNH_DECODE_STREAM=0 timeout 300 python tools/dec_time.py cfg2 cfg5 2>&1 | sed "s/^/STREAM=0 /"; timeout 300 python tools/dec_time.py cfg2 cfg5 2>&1 | sed "s/^/auto /"; CNH_DECODE_STREAM=1 timeout 300 python tools/dec_time.py cfg2 2>&1 | sed "s/^/STREAM=1 /" ) | grep -v "Warning\|co-resident" | tee gpurun_out/dec_time.log
timeout 200 python tools/stage_times.py cfg5 2>&1 | grep -A12 "decode cfg5 rep1" | cut -c1-150
