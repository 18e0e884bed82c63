mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_engine.py -m gpu -x -q -p no:warnings -k "resnet or direct or branch or device_round or prefetch" 2>&1 | tail -3
timeout 300 python bench.py --steps 20 --warmup 5 --worker-streams 1 > gpurun_out/bench34_n1_ws1.log 2>&1; tail -1 gpurun_out/bench34_n1_ws1.log | cut -c1-200
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench34_n1.log 2>&1; tail -1 gpurun_out/bench34_n1.log | cut -c1-200
