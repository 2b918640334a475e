mkdir -p gpurun_out
for cm in 0x1 0x8b; do for pf in 0 2; do echo "== CUT=$cm PF=$pf"; CNH_DECODE_CUTMASK=$cm CNH_DECODE_PF=$pf timeout 200 python tools/stage_times.py cfg5 2>&1 | grep -A22 "decode cfg5 rep1" | grep "cluster loop\|kernel span\|local cuts" | cut -c1-300; done; done
echo "== cfg2"; timeout 200 python tools/stage_times.py cfg2 2>&1 | grep -A22 "decode cfg2 rep1" | grep "cluster loop\|kernel span\|local cuts" | cut -c1-300
