timeout 150 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "candidates" --timeout=60 2>&1 | tail -3
timeout 100 python tools/stage_times.py cfg5 --emit 2>&1 | grep -A7 "detloss+emit cfg5 rep1" | cut -c1-420
timeout 100 python tools/step_events.py cfg5 2>&1 | tail -1
timeout 100 python tools/dec_time.py cfg5 2>&1 | tail -1
