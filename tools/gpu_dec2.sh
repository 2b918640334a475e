CNH_DECODE_ROWS=32 timeout 200 python tools/stage_times.py cfg5 2>&1 | grep -A18 "decode cfg5 rep1" | cut -c1-180
