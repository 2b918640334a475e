timeout 200 python tools/emit_dump.py 2>&1 | grep -c "missing 0"
timeout 200 python tools/emit_check.py cfg5 2>&1 | grep -v missing | cut -c1-200 | head
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "candidates" --timeout=200 2>&1 | tail -5
for f in auto off; do
timeout 600 python bench.py --steps 200 --warmup 20 --no-e2e --no-cpu-baseline --fuse $f > gpurun_out/bench_fuse_$f.json 2> gpurun_out/bench_fuse_$f.err; echo "rc=$? fuse=$f"; tail -2 gpurun_out/bench_fuse_$f.err
python - $f <<'PY'
import json, sys
d = json.loads(open("gpurun_out/bench_fuse_%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
print("cfg2 step", round(d["ms_per_step"]*1e3, 2), "us", round(d["step_hbm_frac"], 3))
c5 = d["cfg5"]; print("  cfg5 step", round(c5["ms_per_step"]*1e3, 1), "us hbm", round(c5["step_hbm_frac"], 3), c5["kernels"], c5.get("candidate_emission"))
PY
done
