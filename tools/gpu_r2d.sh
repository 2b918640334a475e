# decode v3 (L2 prefetch, cluster cuts) + limb-length term: parity, then timing sweeps
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "decode or full_size or keypoints or fixture or schedules" --maxfail=10 --timeout=120 > gpurun_out/pytest_dec.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_dec.log
( for pf in 0 1 2 4; do CNH_DECODE_PF=$pf timeout 300 python tools/dec_time.py cfg2 cfg5 2>&1 | grep -v Warning | sed "s/^/PF=$pf /"; done
  for cm in 0x0 0x1 0x9 0x8b 0xff; do CNH_DECODE_CUTMASK=$cm timeout 300 python tools/dec_time.py cfg5 2>&1 | grep -v Warning | sed "s/^/CUT=$cm /"; done ) | grep -v "co-resident" | tee gpurun_out/dec_time.log
timeout 200 python tools/stage_times.py cfg5 2>&1 | grep -A12 "decode cfg5 rep1" | cut -c1-150
timeout 200 python tools/stage_times.py cfg2 2>&1 | grep -A12 "decode cfg2 rep1" | cut -c1-150
