N=$(nvidia-smi -L | wc -l)
for st in 20 20 200; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps $st --warmup 5 --no-extra --no-e2e 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('steps', d['steps'], 'us/step', round(d['ms_per_step']*1e3,2), 'value', round(d['value']))"
done
