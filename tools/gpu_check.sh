mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q --maxfail=40 --timeout=300 > gpurun_out/pytest1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest1.log
tail -5 gpurun_out/pytest1.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke1.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke1.log
timeout 600 python bench.py --steps 200 --warmup 20 > gpurun_out/bench1.json 2> gpurun_out/bench1.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/bench1.json; tail -5 gpurun_out/bench1.err
