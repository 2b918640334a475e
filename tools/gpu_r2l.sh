mkdir -p gpurun_out
timeout 100 python tools/step_events.py cfg5 --fuse 2>&1 | tail -1
timeout 100 python tools/step_events.py cfg5 --fuse 2>&1 | tail -1
timeout 100 python tools/stage_times.py cfg5 --emit 2>&1 | grep -A11 "detloss+emit cfg5 rep1" | cut -c1-600
