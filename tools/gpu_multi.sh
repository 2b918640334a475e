mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout=300 2>&1 | tee gpurun_out/pytest_gpu_2gpus.log | tail -6
