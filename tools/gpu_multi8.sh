N=$(nvidia-smi -L | wc -l)
echo "gpus: $N"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 2000 --warmup 100 2>gpurun_out/m8.err > gpurun_out/bench_n$N.json; echo "rc=$?"
python -c "import json,sys; d=json.loads(open('gpurun_out/bench_n$N.json').read().strip().splitlines()[-1]); print('N', d['n_gpus'], 'value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'] if d['e2e'] else None, d['config']['parallelism'][:60])"
tail -3 gpurun_out/m8.err
