mkdir -p gpurun_out
timeout 600 python tools/prof_kernels.py cfg2 cfg5 2>&1 | tail -30
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 16 --warmup 3 --no-graph --no-e2e --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/launches.csv')) if len(r) > 5]
hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value'); ui = hdr.index('Metric Unit')
agg = collections.defaultdict(list)
for r in rows[1:]:
    try: agg[r[ki][:70]].append(float(r[vi].replace(',', '')) * (1e-3 if r[ui] in ('ns', 'nsecond') else 1))
    except Exception: pass
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{len(v):4d} x {sum(v)/len(v):9.2f} us  {k}")
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:detloss_kernel -s 6 -c 2 -o gpurun_out/prof_detloss python bench.py --steps 8 --warmup 3 --no-graph --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode_kernel -s 6 -c 2 -o gpurun_out/prof_decode python bench.py --steps 8 --warmup 3 --no-graph --no-e2e --no-cpu-baseline > gpurun_out/ncu_full2.log 2>&1
ls -la gpurun_out/*.ncu-rep
