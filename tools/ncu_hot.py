"""Summarise an ncu report's SASS page: top instructions by warp-stall samples with stall reasons.
usage: python tools/ncu_hot.py <report.ncu-rep> [top_n]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
# several kernels may be concatenated: take the first block
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
end = next((i for i in range(hdr_i + 1, len(rows)) if rows[i] and rows[i][0] == "Kernel Name"), len(rows))
body = rows[hdr_i + 1:end]
si, ai, src = hdr.index("# Samples"), hdr.index("Address"), hdr.index("Source")
stalls = [(j, c) for j, c in enumerate(hdr) if c.startswith("stall_") and "Not Issued" not in c]
def f(x):
    try: return float(x.replace(",", ""))
    except Exception: return 0.0
tot = sum(f(r[si]) for r in body)
print("kernel:", rows[hdr_i - 1][1][:100], "| instructions:", len(body), "| samples:", tot)
agg = {}
for j, c in stalls:
    agg[c] = sum(f(r[j]) for r in body)
print("stall mix:", ", ".join(f"{c[6:]}={100*v/max(1,sum(agg.values())):.0f}%" for c, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
for idx, r in sorted(enumerate(body), key=lambda ir: -f(ir[1][si]))[:top]:
    why = sorted(((f(r[j]), c[6:]) for j, c in stalls), reverse=True)[:2]
    print(f"{f(r[si]):7.0f} {100*f(r[si])/tot:5.1f}%  #{idx:5d} {r[src][:90]:90s} {why[0][1]}:{why[0][0]:.0f} {why[1][1]}:{why[1][0]:.0f}")
