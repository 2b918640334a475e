mkdir -p gpurun_out
timeout 600 python bench.py --steps 200 --warmup 20 --no-e2e --no-cpu-baseline > gpurun_out/bench_fuse.json 2> gpurun_out/bench_fuse.err; echo "rc=$?"; tail -3 gpurun_out/bench_fuse.err
timeout 600 python bench.py --steps 200 --warmup 20 --no-e2e --no-cpu-baseline --no-fuse > gpurun_out/bench_nofuse.json 2>> gpurun_out/bench_fuse.err
python - <<'PY'
import json
for n in ("bench_fuse", "bench_nofuse"):
    try: d = json.loads(open(f"gpurun_out/{n}.json").read().strip().splitlines()[-1])
    except Exception as e: print(n, "unreadable", e); continue
    print(n, "cfg2 step", round(d["ms_per_step"]*1e3, 2), "us", round(d["step_hbm_frac"], 3))
    c5 = d["cfg5"]; print("  cfg5 step", round(c5["ms_per_step"]*1e3, 1), "us hbm", round(c5["step_hbm_frac"], 3), c5["kernels"]); print("  ", c5["schedule"][:200])
PY
