# 2 GPUs: the sharded schedules against one device (NCCL, in-kernel exchange, deferred totals) and the graphed host step
# around the sharded loss.  (Keep the outer limit tight: a hung collective keeps the box until the limit.)
mkdir -p gpurun_out
nvidia-smi -L
timeout 400 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout=150 2>&1 | tee gpurun_out/pytest_gpu_2gpus.log | tail -6
