timeout 200 python tools/emit_check.py cfg5 2>&1 | grep -v missing | cut -c1-200 | head -8
timeout 600 python bench.py --steps 400 --warmup 40 --no-e2e --no-cpu-baseline > gpurun_out/bench_fork.json 2> gpurun_out/bench_fork.err; echo "rc=$?"; tail -3 gpurun_out/bench_fork.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_fork.json").read().strip().splitlines()[-1])
print("value", round(d["value"]), "us", round(d["ms_per_step"]*1e3, 2), "frac", round(d["step_hbm_frac"], 3), d["run"]["schedule"])
c5 = d["cfg5"]; print("cfg5", round(c5["ms_per_step"]*1e3, 1), "us", round(c5["step_hbm_frac"], 3), c5.get("emission_setting"), c5.get("other_emission_setting"))
print({k: round(v["ms_per_step"]*1e3, 2) for k, v in d["shapes"].items()})
PY
