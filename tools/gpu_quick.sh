mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=10 --timeout=120 > gpurun_out/pytest_q.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_q.log
python tools/e2e_profile.py 2>&1 | head -1
for i in 1 2; do timeout 600 python bench.py --steps 400 --warmup 40 --no-cpu-baseline 2> gpurun_out/bench_q.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('value', round(d['value']), 'e2e', round(d['e2e']['value']), d['e2e']['ms_per_step'], round(d['e2e']['h2d_GBps'],1), 'boxes', round(d['e2e_boxes']['value']), d['e2e_boxes']['ms_per_step'])"; done
tail -3 gpurun_out/bench_q.err
