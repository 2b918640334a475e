mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python tools/debug_decode.py > gpurun_out/sanitizer.log 2>&1
grep -vE "^$" gpurun_out/sanitizer.log | head -60
for k in uda advent full_size "decode_after or fused_sigmoid or K_larger"; do
  timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "$k" --timeout=300 2>&1 | grep -E "^(FAILED|ERROR)|passed|failed|^E  " | head -30
done
