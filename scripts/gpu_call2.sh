mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -x -q -p no:warnings > gpurun_out/pytest_gpu2.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu2.log | cut -c1-600
timeout 300 python bench/stress_bucket_protocol.py --epochs 300 2>&1 | tail -3
timeout 300 python bench/stress_bucket_protocol.py --epochs 100 --mode trmean --buckets 3 2>&1 | tail -2
cat gpurun_out/mismatch_*.txt 2>/dev/null | head -20
