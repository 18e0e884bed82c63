mkdir -p gpurun_out
for v in "--sleep-after-first-rank" "--sleep-after-first-rank --spin 0" ; do echo "== $v"; timeout 80 python bench/stress_bucket_protocol.py --epochs 10 $v 2>&1 | tail -2 | cut -c1-330; done
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_control_plane.py tests/test_gpu_engine.py -x -q -p no:warnings -k "tma or control or preagg or capturable or pre_aggregate_on_cuda or bucket or vmm or gram" > gpurun_out/pytest_dbg.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_dbg.log | cut -c1-400
python benchmarks/gram_bench.py 2>&1 | tail -15
