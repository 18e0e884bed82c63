mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -p no:warnings > gpurun_out/pytest10.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/pytest10.log
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench10_n1.log 2>&1; tail -1 gpurun_out/bench10_n1.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-overlap-wgrad > gpurun_out/bench10_n1_noov.log 2>&1; tail -1 gpurun_out/bench10_n1_noov.log | cut -c1-200
timeout 300 python bench.py --steps 20 --warmup 5 --no-direct-grads > gpurun_out/bench10_n1_nodirect.log 2>&1; tail -1 gpurun_out/bench10_n1_nodirect.log | cut -c1-200
timeout 300 python bench.py --steps 20 --warmup 5 --worker-streams 1 > gpurun_out/bench10_n1_ws1.log 2>&1; tail -1 gpurun_out/bench10_n1_ws1.log | cut -c1-200
timeout 300 python bench.py --steps 20 --warmup 5 --worker-streams 1 --no-overlap-wgrad > gpurun_out/bench10_n1_ws1_noov.log 2>&1; tail -1 gpurun_out/bench10_n1_ws1_noov.log | cut -c1-200
timeout 200 python bench/profile_step.py --streams 1 --out gpurun_out/profile_step10.txt > gpurun_out/prof10.log 2>&1
timeout 100 python bench/conv_stem.py 2>&1 | tail -4
timeout 200 python benchmarks/agg_sweep.py --n 8 --dims 1e7,1e8 --out gpurun_out/agg_sweep10_n8.json > gpurun_out/agg_sweep10_n8.log 2>&1; grep -h "median\|trimmed" gpurun_out/agg_sweep10_n8.log | cut -c1-220
timeout 200 python benchmarks/agg_sweep.py --n 16 --f 3 --dims 1e7 --out gpurun_out/agg_sweep10_n16.json > gpurun_out/agg_sweep10_n16.log 2>&1; grep -h "median\|trimmed" gpurun_out/agg_sweep10_n16.log | cut -c1-220
timeout 200 python benchmarks/agg_sweep.py --n 64 --f 8 --dims 1e7 --out gpurun_out/agg_sweep10_n64.json > gpurun_out/agg_sweep10_n64.log 2>&1; grep -h "median\|trimmed" gpurun_out/agg_sweep10_n64.log | cut -c1-220
