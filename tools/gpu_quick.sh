mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=10 --timeout=120 > gpurun_out/pytest_q.log 2>&1; echo "pytest rc=$?"
tail -12 gpurun_out/pytest_q.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 600 python bench.py --steps 400 --warmup 40 > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err; echo "bench rc=$?"; cat gpurun_out/bench_q.json; tail -5 gpurun_out/bench_q.err
