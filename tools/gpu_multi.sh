mkdir -p gpurun_out
nvidia-smi -L
timeout 400 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout=150 -k graphed 2>&1 | tee gpurun_out/pytest_gpu_2gpus_graphed.log | tail -25
