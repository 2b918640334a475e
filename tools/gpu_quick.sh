mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=10 --timeout=120 > gpurun_out/pytest_q.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_q.log
for i in 1 2; do timeout 600 python bench.py --steps 2000 --warmup 100 --no-cpu-baseline --no-e2e 2> gpurun_out/bench_q.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('value', round(d['value']), d['ms_per_step'], 'loss us', d['roofline']['us_per_launch'])"; done
CNH_NO_PDL=1 timeout 600 python bench.py --steps 2000 --warmup 100 --no-cpu-baseline --no-e2e 2> gpurun_out/bench_q.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('NO_PDL value', round(d['value']), d['ms_per_step'], 'loss us', d['roofline']['us_per_launch'])"
timeout 300 python bench.py --config cfg5 --steps 300 --warmup 30 --no-cpu-baseline --no-e2e 2>> gpurun_out/bench_q.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('cfg5 value', round(d['value']), d['ms_per_step'])"
tail -3 gpurun_out/bench_q.err
