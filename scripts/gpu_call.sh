mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_engine.py -m gpu -x -q -p no:warnings -k "bert or p2p" 2>&1 | tail -4 | cut -c1-300
timeout 400 python benchmarks/training_configs.py --config 4 --steps 10 > gpurun_out/cfg4_n1c.log 2>&1; echo "cfg4 n1 rc=$?"; tail -1 gpurun_out/cfg4_n1c.log | cut -c1-300
