mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -p no:warnings 2>&1 | tail -12 | cut -c1-300
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 benchmarks/training_configs.py --config 3 --steps 10 > gpurun_out/cfg3_n2c.log 2>&1; echo "cfg3 n2 rc=$?"; tail -1 gpurun_out/cfg3_n2c.log | cut -c1-300
