# compute-sanitizer over the decode / rasteriser parity tests (memcheck) and the cluster decode (racecheck)
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "decode_vs_oracle or raster or epilogue or fixture" --timeout=600 > gpurun_out/sanitize.log 2>&1; echo "memcheck rc=$?"
grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitize.log | head -4
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "test_decode_vs_oracle and cluster and not rows" --timeout=600 > gpurun_out/racecheck.log 2>&1; echo "racecheck rc=$?"
grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/racecheck.log | head -4
grep -E "Error: Race|Warning: Race" gpurun_out/racecheck.log | sed -e 's/+0x[0-9a-f]*//' | cut -c1-150 | sort | uniq -c | sort -rn | head -12
