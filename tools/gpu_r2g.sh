# fused candidates: parity of decode-from-candidates, loss parity (log series), timings
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -v -k "candidates or detection_loss or schedules or precount or elementwise" --maxfail=10 --timeout=200 > gpurun_out/pytest_cand.log 2>&1; echo "pytest rc=$?"
grep -c PASSED gpurun_out/pytest_cand.log; grep "FAILED\|elementwise cfg" gpurun_out/pytest_cand.log | cut -c1-250 | head -20
tail -3 gpurun_out/pytest_cand.log | cut -c1-200
