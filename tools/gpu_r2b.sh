# decode with cluster-wide cuts: parity on every launch shape, timings per cluster-size cap
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "decode or full_size" --maxfail=10 --timeout=120 > gpurun_out/pytest_dec.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_dec.log
( timeout 300 python tools/dec_time.py cfg2 cfg5 cfg1; for cs in 6 8 9 10 12 16; do CNH_DECODE_CS=$cs timeout 300 python tools/dec_time.py cfg2 cfg5; done ) 2>&1 | grep -v Warning | tee gpurun_out/dec_time.log
timeout 200 python tools/stage_times.py cfg2 2>&1 | grep -A14 "decode cfg2 rep1" | cut -c1-180
timeout 200 python tools/stage_times.py cfg5 2>&1 | grep -A14 "decode cfg5 rep1" | cut -c1-180
