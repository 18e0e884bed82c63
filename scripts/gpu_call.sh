mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -p no:warnings -k "maxpool or resnet" 2>&1 | tail -3
timeout 300 python bench.py --steps 20 --warmup 5 --worker-streams 1 > gpurun_out/bench33_n1_ws1.log 2>&1; tail -1 gpurun_out/bench33_n1_ws1.log | cut -c1-200
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench33_n1.log 2>&1; tail -1 gpurun_out/bench33_n1.log | cut -c1-200
timeout 200 python bench/profile_step.py --streams 1 --out gpurun_out/profile_step33.txt > gpurun_out/prof33.log 2>&1; grep -i "maxpool\|reduce_partial\|apply_kernel\|finalize" gpurun_out/profile_step33.txt | cut -c1-60,180-330 | head -12
