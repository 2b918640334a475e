mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout=300 2>&1 | tail -15
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2000 --warmup 100 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench2 rc=$?"; cat gpurun_out/bench_n2.json; tail -5 gpurun_out/bench_n2.err
timeout 600 python bench.py --gpus 1 --steps 2000 --warmup 100 --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; cat gpurun_out/bench_n1.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('N1 value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'])"
