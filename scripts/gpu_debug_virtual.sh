mkdir -p gpurun_out
nproc; lscpu | grep -E "Model name|Flags" | cut -c1-200 | head -2
timeout 300 python bench/debug_virtual.py --repeats 1500 2>&1 | tail -15
timeout 200 python bench/debug_virtual.py --repeats 800 --threads 1 2>&1 | tail -6
for i in 1 2 3 4 5 6 7 8; do
  timeout 300 python -m pytest tests/test_gpu_engine.py tests/test_gpu_kernels.py -q -p no:warnings -k "attacks_cuda or virtual_rows or cw_select or colstat or little" 2>&1 | tail -2 | cut -c1-200
done
cat gpurun_out/mismatch_*.txt 2>/dev/null
