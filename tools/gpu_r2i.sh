mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "candidates" --maxfail=10 --timeout=200 2>&1 | tail -2
timeout 300 python tools/step_events.py cfg5 2>&1 | tail -1
timeout 300 python tools/step_events.py cfg5 --no-fuse 2>&1 | tail -1
timeout 600 python bench.py --steps 200 --warmup 20 --no-e2e --no-cpu-baseline > gpurun_out/bench_fuse.json 2> gpurun_out/bench_fuse.err; echo "rc=$?"; tail -3 gpurun_out/bench_fuse.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_fuse.json").read().strip().splitlines()[-1])
print("cfg2 step", round(d["ms_per_step"]*1e3, 2), "us", round(d["step_hbm_frac"], 3))
c5 = d["cfg5"]; print("  cfg5 step", round(c5["ms_per_step"]*1e3, 1), "us hbm", round(c5["step_hbm_frac"], 3), c5["kernels"])
PY
