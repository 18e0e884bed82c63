mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -p no:warnings > gpurun_out/pytest16.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/pytest16.log | cut -c1-300
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench16_n1.log 2>&1; tail -1 gpurun_out/bench16_n1.log | cut -c1-200
timeout 300 python bench.py --steps 20 --warmup 5 --worker-streams 1 > gpurun_out/bench16_n1_ws1.log 2>&1; tail -1 gpurun_out/bench16_n1_ws1.log | cut -c1-200
timeout 300 python bench.py --steps 20 --warmup 5 --worker-streams 8 > gpurun_out/bench16_n1_ws8.log 2>&1; tail -1 gpurun_out/bench16_n1_ws8.log | cut -c1-200
timeout 200 python bench/bn_layers.py 2>&1 | tail -12
timeout 200 python bench/profile_step.py --streams 1 --out gpurun_out/profile_step16.txt > gpurun_out/prof16.log 2>&1; tail -2 gpurun_out/prof16.log
