timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout=300 2>&1 | tail -3
for mode in "X=1" "CNH_BENCH_NO_EXCHANGE=1"; do
  env $mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2000 --warmup 100 --no-e2e 2>gpurun_out/m2.err | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$mode', 'N2 value', d['value'], 'ms', d['ms_per_step'], d['config']['launch'])"
done
tail -3 gpurun_out/m2.err
