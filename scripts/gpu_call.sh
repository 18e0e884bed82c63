mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29552 benchmarks/training_configs.py --config 4 --steps 10 > gpurun_out/cfg4_n4.log 2>&1; echo "cfg4 n4 rc=$?"; tail -1 gpurun_out/cfg4_n4.log | cut -c1-300
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29551 benchmarks/training_configs.py --config 3 --steps 10 > gpurun_out/cfg3_n4.log 2>&1; echo "cfg3 n4 rc=$?"; tail -1 gpurun_out/cfg3_n4.log | cut -c1-300
