mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -p no:warnings > gpurun_out/pytest20.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/pytest20.log | cut -c1-300
timeout 300 python benchmarks/gram_bench.py > gpurun_out/gram_bench20.log 2>&1; tail -5 gpurun_out/gram_bench20.log | cut -c1-330
timeout 200 python benchmarks/agg_sweep.py --n 8 --dims 1e7,1e8 --out gpurun_out/agg_sweep20_n8.json > gpurun_out/agg_sweep20_n8.log 2>&1; grep -h "geometric\|centered\|multi" gpurun_out/agg_sweep20_n8.log | cut -c1-220
timeout 200 python benchmarks/agg_sweep.py --n 64 --f 8 --dims 1e7 --out gpurun_out/agg_sweep20_n64.json > gpurun_out/agg_sweep20_n64.log 2>&1; grep -h "geometric\|centered\|multi" gpurun_out/agg_sweep20_n64.log | cut -c1-220
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench20_n1.log 2>&1; tail -1 gpurun_out/bench20_n1.log | cut -c1-200
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gram_umma_kernel -s 1 -c 1 -f -o gpurun_out/gram_umma_v4_n64 python bench/ncu_ops.py --op gram_umma --n 64 --d 4194304 > gpurun_out/ncu_umma4.log 2>&1
ls -la gpurun_out/gram_umma_v4_n64.ncu-rep
