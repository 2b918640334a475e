mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_callers.py -m gpu -q --maxfail=5 --timeout=120 2>&1 | tail -3
for f in "" "--fuse"; do
timeout 600 python bench.py --steps 200 --warmup 20 --no-e2e --no-cpu-baseline $f > gpurun_out/bench_fuse$f.json 2> gpurun_out/bench_fuse$f.err; echo "rc=$? fuse='$f'"; tail -2 gpurun_out/bench_fuse$f.err
python - "$f" <<'PY'
import json, sys
d = json.loads(open("gpurun_out/bench_fuse%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
print("cfg2 step", round(d["ms_per_step"]*1e3, 2), "us", round(d["step_hbm_frac"], 3))
c5 = d["cfg5"]; print("  cfg5 step", round(c5["ms_per_step"]*1e3, 1), "us hbm", round(c5["step_hbm_frac"], 3), c5["kernels"])
PY
done
