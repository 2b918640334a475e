timeout 200 python tools/stage_times.py cfg2 2>&1 | grep -A18 "decode cfg2 rep1" | cut -c1-180
timeout 200 python tools/stage_times.py cfg5 2>&1 | grep -A18 "decode cfg5 rep1" | cut -c1-180
