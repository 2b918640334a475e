# round-2 first GPU call (2 GPUs): whole GPU test suite incl. the 2-GPU parity tests (log kept), smoke, bench N=1/N=2
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --maxfail=30 --timeout=600 -rs > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench1 rc=$?"; tail -5 gpurun_out/bench_n1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench2 rc=$?"; tail -8 gpurun_out/bench_n2.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref.json 2>> gpurun_out/bench_n1.err
python - <<'PY'
import json
for n in ("bench_n1", "bench_n2"):
    try:
        d = json.loads(open(f"gpurun_out/{n}.json").read().strip().splitlines()[-1])
    except Exception as e:
        print(n, "unreadable", e); continue
    print(n, "value", round(d["value"]), "ms", round(d["ms_per_step"]*1e3, 2), "us; hbm", round(d["step_hbm_frac"], 3),
          "e2e", d["e2e"] and round(d["e2e"]["value"]), "tonly", d.get("e2e_targets_only") and round(d["e2e_targets_only"]["value"]))
    c5 = d.get("cfg5")
    if c5:
        print("  cfg5", round(c5["ms_per_step"]*1e3, 1), "us hbm", round(c5["step_hbm_frac"], 3), c5.get("kernels"))
        if "nccl_schedule" in c5: print("  cfg5 nccl", round(c5["nccl_schedule"]["ms_per_step"]*1e3, 1), c5["nccl_schedule"]["kernels"])
    print("  shapes", {k: (round(v["ms_per_step"]*1e3, 2), round(v["step_hbm_frac"], 3)) for k, v in d.get("shapes", {}).items()})
    print("  parity", d.get("sharded_parity"))
PY
