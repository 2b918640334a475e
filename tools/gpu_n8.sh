mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench$N rc=$?"; tail -5 gpurun_out/bench_n$N.err | cut -c1-300
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_n$N.json").read().strip().splitlines()[-1])
print("N=$N value", round(d["value"]), "ms", round(d["ms_per_step"]*1e3, 2), "us; e2e", round(d["e2e"]["value"]), "dense", round(d["e2e_dense_targets"]["value"]), "tonly", round(d["e2e_targets_only"]["value"]), "eager", round(d["e2e_eager"]["value"]))
c5 = d["cfg5"]; print("  cfg5", round(c5["ms_per_step"]*1e3, 1), "us value", round(c5["value"]), "hbm", round(c5["step_hbm_frac"], 3), c5.get("emission_setting"), c5.get("other_emission_setting"), "| nccl", round(c5["nccl_schedule"]["ms_per_step"]*1e3, 1), c5["nccl_schedule"]["kernels"])
print("  shapes", {k: round(v["ms_per_step"]*1e3, 2) for k, v in d["shapes"].items()})
print("  parity", {k: (v["schedule"], v["scalars"], v["grad_hm"], v["dets"]) for k, v in d["sharded_parity"].items()})
PY
