# streaming decode: parity on every launch shape, timings
mkdir -p gpurun_out
timeout 420 python -m pytest tests/test_gpu_parity.py -m gpu -v -k "decode or full_size" --maxfail=10 --timeout=120 > gpurun_out/pytest_dec.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_dec.log | cut -c1-200
( timeout 300 python tools/dec_time.py cfg5 2>&1 | sed "s/^/auto /"; CNH_DECODE_FALLBACK_LAST=1 timeout 300 python tools/dec_time.py cfg5 2>&1 | sed "s/^/fallback_last /";  CNH_DECODE_FALLBACK_LAST=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "decode_vs_oracle and stream" 2>&1 | tail -1 ) | grep -v "Warning\|co-resident" | tee gpurun_out/dec_time.log
timeout 200 python tools/stage_times.py cfg5 2>&1 | grep -A10 "decode cfg5 rep1" | cut -c1-150
