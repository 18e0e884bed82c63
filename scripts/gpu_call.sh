mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -p no:warnings -k "maxpool or direct or resnet18 or batchnorm or krsc or normalize" > gpurun_out/pytest12.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest12.log
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench12_n1.log 2>&1; tail -1 gpurun_out/bench12_n1.log | cut -c1-200
timeout 300 python bench.py --steps 20 --warmup 5 --worker-streams 1 > gpurun_out/bench12_n1_ws1.log 2>&1; tail -1 gpurun_out/bench12_n1_ws1.log | cut -c1-200
timeout 300 python bench.py --steps 20 --warmup 5 --worker-streams 8 > gpurun_out/bench12_n1_ws8.log 2>&1; tail -1 gpurun_out/bench12_n1_ws8.log | cut -c1-200
timeout 200 python bench/profile_step.py --streams 1 --out gpurun_out/profile_step12.txt > gpurun_out/prof12.log 2>&1; tail -2 gpurun_out/prof12.log
timeout 200 python bench/bn_layers.py 2>&1 | tail -12
