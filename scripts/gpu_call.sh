mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -p no:warnings -k "batchnorm or resnet" 2>&1 | tail -3
timeout 200 python bench/bn_layers.py 2>&1 | tail -10
timeout 300 python bench.py --steps 20 --warmup 5 --worker-streams 1 > gpurun_out/bench31_n1_ws1.log 2>&1; tail -1 gpurun_out/bench31_n1_ws1.log | cut -c1-200
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench31_n1.log 2>&1; tail -1 gpurun_out/bench31_n1.log | cut -c1-200
