"""Turn the ncu outputs brought back in gpurun_out/ into the tracked summaries under profiles/.
usage: python tools/ncu_summary.py <round-tag>"""
import collections, csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles"); GO = os.path.join(ROOT, "gpurun_out")
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
os.makedirs(OUT, exist_ok=True)

def launches():
    rows = [r for r in csv.reader(open(os.path.join(GO, "launches.csv"))) if len(r) > 5]
    hdr = rows[0]; ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try: v = float(r[vi].replace(",", ""))
        except ValueError: continue
        v *= {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3}.get(r[ui], 1.0)
        agg.setdefault(r[ki], []).append(v)
    tot = sum(sum(v) for v in agg.values())
    lines = ["# ncu --metrics gpu__time_duration.sum --clock-control none ... python bench.py --steps 16 --warmup 3 --no-graph --no-e2e --no-cpu-baseline",
             "# per-launch times are cold-cache and serialised under the profiler: compare SHARES, not absolutes",
             f"{'launches':>8} {'avg_us':>10} {'share':>7}  kernel"]
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        lines.append(f"{len(v):8d} {sum(v)/len(v):10.2f} {100*sum(v)/tot:6.1f}%  {k[:110]}")
    open(os.path.join(OUT, f"{tag}_launches.txt"), "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor", "smsp__inst_executed.sum",
        "sm__cycles_elapsed.max", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
        "smsp__warp_issue_stalled_membar_per_warp_active.pct", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct"]

def full(rep, name):
    path = os.path.join(GO, rep)
    if not os.path.exists(path): return None
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ki = hdr.index("Kernel Name")
    lines = [f"# ncu --set full --clock-control none --import-source on: {name} ({len(data)} launches captured)"]
    res = {}
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            vals = [r[i] for r in data]
            lines.append(f"{w:75s} [{units[i]}] {vals}")
            res[w] = (units[i], vals)
    lines.insert(1, f"# kernel: {data[0][ki][:120]}")
    open(os.path.join(OUT, f"{tag}_{name}.txt"), "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))
    return res

def to_bytes(unit, v):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)

launches()
traffic = {}
for rep, name, cfg in (("prof_detloss.ncu-rep", "detloss_cfg2", "cfg2"), ("prof_decode.ncu-rep", "decode_cfg2", None),
                       ("prof_detloss_cfg5.ncu-rep", "detloss_cfg5_emit", "cfg5_emit"), ("prof_decode_cfg5.ncu-rep", "decode_cfg5", None),
                       ("prof_finish_cfg5.ncu-rep", "finish_cfg5", None), ("prof_detloss_cfg5_plain.ncu-rep", "detloss_cfg5", "cfg5")):
    r = full(rep, name)
    if r and cfg and "dram__bytes_read.sum" in r:
        ur, vr = r["dram__bytes_read.sum"]; uw, vw = r["dram__bytes_write.sum"]
        traffic[cfg] = sum(to_bytes(ur, a) + to_bytes(uw, b) for a, b in zip(vr, vw)) / len(vr)
json.dump(traffic, open(os.path.join(OUT, "traffic.json"), "w"), indent=1)
print("traffic (dram read+write bytes per launch):", traffic)
