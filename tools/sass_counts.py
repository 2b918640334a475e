"""Per-kernel counts of the SASS mnemonics that show what the library is built from (cuobjdump -sass of the in-tree .so):
UBLKCP / UTMALDG / UBLKPF (TMA bulk copies, tensor-map loads, L2 bulk prefetch), SYNCS (mbarrier), packed fp32
(FMUL2 / FADD2 / FFMA2), MUFU, cluster barriers (UCGABAR), RED / ATOM, and -- must stay zero: nothing here is a
contraction -- tensor-core mnemonics (UTC*MMA, HMMA, LDTM).  usage: python tools/sass_counts.py > profiles/r02_sass.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "centernet-uda_b200", "lib", "libcnhead_sm100.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
pats = collections.OrderedDict([("UBLKCP", r"\bUBLKCP"), ("UTMALDG", r"\bUTMALDG"), ("UBLKPF", r"\bUBLKPF"), ("SYNCS", r"\bSYNCS"),
                                ("F*2 packed", r"\bF(MUL|ADD|FMA)2\b"), ("MUFU", r"\bMUFU"), ("UCGABAR", r"\bUCGABAR"),
                                ("RED", r"\bRED\b|\bREDG"), ("ATOM(S/G)", r"\bATOM"), ("LDS.128", r"\bLDS\.128"),
                                ("tensor (UTC*MMA|HMMA|LDTM)", r"UTC\w*MMA|\bHMMA|\bLDTM")])
archs = sorted(set(re.findall(r"arch = (sm_\w+)", txt)))
print("# cuobjdump -sass centernet-uda_b200/lib/libcnhead_sm100.so | per-kernel mnemonic counts")
print("# cubin architectures in the library:", ", ".join(archs))
cur, rows = None, collections.OrderedDict()
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", name)
        rows[cur] = collections.Counter()
        continue
    if cur and "/*" in line:
        rows[cur]["instr"] += 1
        for k, p in pats.items():
            if re.search(p, line):
                rows[cur][k] += 1
hdr = ["instr"] + list(pats)
print(("%-74s" % "kernel") + " ".join("%12s" % h[:12] for h in hdr))
tot = collections.Counter()
for k, c in rows.items():
    print(("%-74s" % k[:74]) + " ".join("%12d" % c[h] for h in hdr))
    tot.update(c)
print(("%-74s" % "TOTAL") + " ".join("%12d" % tot[h] for h in hdr))
