mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_engine.py -m gpu -x -q -p no:warnings > gpurun_out/pytest24.log 2>&1; echo "pytest multi+engine rc=$?"; tail -6 gpurun_out/pytest24.log | cut -c1-400
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -p no:warnings -k "branch_stream or maxpool or direct_gradients or fused_median or geometric_median" > gpurun_out/pytest24b.log 2>&1; echo "pytest kernels rc=$?"; tail -4 gpurun_out/pytest24b.log | cut -c1-400
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29552 benchmarks/training_configs.py --config 4 --steps 10 > gpurun_out/cfg4_n2b.log 2>&1; echo "cfg4 n2 rc=$?"; tail -1 gpurun_out/cfg4_n2b.log | cut -c1-300
timeout 400 python benchmarks/training_configs.py --config 4 --steps 10 > gpurun_out/cfg4_n1b.log 2>&1; echo "cfg4 n1 rc=$?"; tail -1 gpurun_out/cfg4_n1b.log | cut -c1-300
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench24_n2.log 2>&1; tail -1 gpurun_out/bench24_n2.log | cut -c1-200
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 20 --warmup 5 --no-branch-streams > gpurun_out/bench24_n2_nobr.log 2>&1; tail -1 gpurun_out/bench24_n2_nobr.log | cut -c1-200
timeout 300 python bench.py --steps 20 --warmup 5 --worker-streams 1 > gpurun_out/bench24_n1_ws1.log 2>&1; tail -1 gpurun_out/bench24_n1_ws1.log | cut -c1-200
timeout 300 python bench.py --steps 20 --warmup 5 --worker-streams 1 --no-branch-streams > gpurun_out/bench24_n1_ws1_nobr.log 2>&1; tail -1 gpurun_out/bench24_n1_ws1_nobr.log | cut -c1-200
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench24_n1.log 2>&1; tail -1 gpurun_out/bench24_n1.log | cut -c1-200
