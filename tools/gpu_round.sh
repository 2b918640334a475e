# Round-2 final evidence in one box (1 GPU): parity tests, smoke, both bench arms, per-kernel timings, ncu launch list + full captures.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q --maxfail=40 --timeout=300 -rs -s > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
tail -4 gpurun_out/pytest.log; grep "\[elementwise\|\[chain" gpurun_out/pytest.log | cut -c1-260
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py --steps 400 --warmup 40 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err
timeout 900 python bench.py > gpurun_out/bench_default.json 2>> gpurun_out/bench.err; echo "bench default rc=$?"
timeout 600 python bench.py --impl reference --steps 30 --warmup 5 > gpurun_out/bench_reference.json 2>> gpurun_out/bench.err
timeout 600 python bench.py --config cfg5 --steps 200 --warmup 20 --no-extra --no-e2e --no-cpu-baseline > gpurun_out/bench_cfg5.json 2>> gpurun_out/bench.err
timeout 600 python tools/prof_kernels.py cfg2 cfg5 > gpurun_out/prof_kernels.log 2>&1; tail -26 gpurun_out/prof_kernels.log | cut -c1-200
B="python bench.py --no-graph --no-e2e --no-cpu-baseline --no-extra"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $B --steps 16 --warmup 3 > gpurun_out/ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:detloss_stash -s 10 -c 2 -f -o gpurun_out/prof_detloss $B --steps 8 --warmup 3 > gpurun_out/ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_cluster -s 10 -c 2 -f -o gpurun_out/prof_decode $B --steps 8 --warmup 3 > gpurun_out/ncu_full2.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:detloss_stream -s 4 -c 1 -f -o gpurun_out/prof_detloss_cfg5 $B --config cfg5 --steps 4 --warmup 3 > gpurun_out/ncu_full3.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:decode_cluster -s 4 -c 1 -f -o gpurun_out/prof_finish_cfg5 $B --config cfg5 --steps 4 --warmup 3 > gpurun_out/ncu_full5.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:detloss_stream -s 4 -c 1 -f -o gpurun_out/prof_detloss_cfg5_plain $B --config cfg5 --fuse off --steps 4 --warmup 3 > gpurun_out/ncu_full6.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_stream -s 4 -c 1 -f -o gpurun_out/prof_decode_cfg5 $B --config cfg5 --fuse off --steps 4 --warmup 3 > gpurun_out/ncu_full4.log 2>&1
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench.json").read().strip().splitlines()[-1])
print("value", round(d["value"]), "us", round(d["ms_per_step"]*1e3, 2), "frac", round(d["step_hbm_frac"], 3), "roof", d["roofline"]["frac"], d["roofline"]["us_per_launch"])
for k in ("e2e", "e2e_dense_targets", "e2e_targets_only", "e2e_eager"):
    print(k, round(d[k]["value"]), round(d[k]["ms_per_step"]*1e3, 1), "us", round(d[k]["h2d_GBps"], 1), "GB/s")
c5 = d["cfg5"]; print("cfg5", round(c5["ms_per_step"]*1e3, 1), "us", round(c5["step_hbm_frac"], 3), c5["kernels"], c5.get("candidate_emission"))
print({k: round(v["ms_per_step"]*1e3, 2) for k, v in d["shapes"].items()}); print(d["cpu_baseline"])
PY
