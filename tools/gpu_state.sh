mkdir -p gpurun_out
timeout 300 python tools/stage_times.py cfg2 > gpurun_out/stage_cfg2.log 2>&1; echo "stage2 rc=$?"
timeout 300 python tools/stage_times.py cfg5 > gpurun_out/stage_cfg5.log 2>&1; echo "stage5 rc=$?"
timeout 600 python tools/prof_kernels.py cfg2 cfg5 > gpurun_out/prof_kernels.log 2>&1; echo "prof rc=$?"
cat gpurun_out/stage_cfg2.log | head -80
