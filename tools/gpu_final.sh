# final verification of the tree as committed: GPU tests, smoke, both bench arms, cfg5 line, per-kernel timings
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --maxfail=40 --timeout=300 -rs -s > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
tail -3 gpurun_out/pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 900 python bench.py --steps 400 --warmup 40 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err
timeout 900 python bench.py > gpurun_out/bench_default.json 2>> gpurun_out/bench.err; echo "bench default rc=$?"
timeout 600 python bench.py --impl reference --steps 30 --warmup 5 > gpurun_out/bench_reference.json 2>> gpurun_out/bench.err
timeout 600 python bench.py --config cfg5 --steps 200 --warmup 20 --no-extra --no-e2e --no-cpu-baseline > gpurun_out/bench_cfg5.json 2>> gpurun_out/bench.err
timeout 600 python tools/prof_kernels.py cfg2 cfg5 > gpurun_out/prof_kernels.log 2>&1
python - <<'PY'
import json
for f in ("bench", "bench_default"):
    d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
    print(f, "value", round(d["value"]), "us", round(d["ms_per_step"]*1e3, 2), "frac", round(d["step_hbm_frac"], 3), "roof", round(d["roofline"]["frac"], 3), round(d["roofline"]["us_per_launch"], 2))
    for k in ("e2e", "e2e_dense_targets", "e2e_targets_only", "e2e_eager"):
        print("  ", k, round(d[k]["value"]), round(d[k]["ms_per_step"]*1e3, 1), "us", round(d[k]["h2d_GBps"], 1), "GB/s")
    c5 = d["cfg5"]; print("   cfg5", round(c5["ms_per_step"]*1e3, 1), "us", round(c5["step_hbm_frac"], 3), c5.get("emission_setting"), c5.get("other_emission_setting"))
    print("  ", {k: round(v["ms_per_step"]*1e3, 2) for k, v in d["shapes"].items()}, d["cpu_baseline"]["value"])
d = json.loads(open("gpurun_out/bench_cfg5.json").read().strip().splitlines()[-1])
print("cfg5 main", round(d["ms_per_step"]*1e3, 1), round(d["step_hbm_frac"], 3), d.get("candidate_emission"), d["roofline"]["frac"])
PY
