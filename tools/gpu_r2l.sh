mkdir -p gpurun_out
for k in 1 2 3; do
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "candidates" --timeout=60 2>&1 | grep -v "^$" | tail -25 | cut -c1-220
done
