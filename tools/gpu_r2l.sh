for r in 1 2; do
for v in hint nohint; do
if [ $v = nohint ]; then export CNH_LIB_PATH=$PWD/tools/ubench/alt/libcnhead_nohint.so; else unset CNH_LIB_PATH; fi
timeout 600 python tools/prof_kernels.py cfg2 cfg5 2>&1 | grep -E "fused_stash |decode |loss_emitting|full_step|fused_precount " | sed "s/^/$v /" | cut -c1-150
done
done
