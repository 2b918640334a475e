timeout 600 ncu --set full --clock-control none -k regex:detloss_stream -s 2 -c 1 -f -o gpurun_out/prof_detloss_cfg5 python tools/emit_check.py cfg5 > gpurun_out/ncu_full3.log 2>&1; tail -2 gpurun_out/ncu_full3.log
timeout 600 python bench.py --steps 100 --warmup 10 --no-e2e --no-cpu-baseline > gpurun_out/bench_ab.json 2> gpurun_out/bench_ab.err; echo "rc=$?"; tail -3 gpurun_out/bench_ab.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_ab.json").read().strip().splitlines()[-1])
c5 = d["cfg5"]; print("cfg5", round(c5["ms_per_step"]*1e3, 1), c5.get("candidate_emission"), c5.get("other_emission_setting"))
PY
