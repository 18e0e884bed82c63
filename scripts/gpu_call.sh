mkdir -p gpurun_out
timeout 400 python benchmarks/training_configs.py --config 3 --steps 10 > gpurun_out/cfg3_n1.log 2>&1; echo "cfg3 n1 rc=$?"; tail -2 gpurun_out/cfg3_n1.log | cut -c1-400
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 benchmarks/training_configs.py --config 3 --steps 10 > gpurun_out/cfg3_n2.log 2>&1; echo "cfg3 n2 rc=$?"; tail -2 gpurun_out/cfg3_n2.log | cut -c1-400
timeout 400 python benchmarks/training_configs.py --config 4 --steps 10 > gpurun_out/cfg4_n1.log 2>&1; echo "cfg4 n1 rc=$?"; tail -2 gpurun_out/cfg4_n1.log | cut -c1-400
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29552 benchmarks/training_configs.py --config 4 --steps 10 > gpurun_out/cfg4_n2.log 2>&1; echo "cfg4 n2 rc=$?"; tail -2 gpurun_out/cfg4_n2.log | cut -c1-400
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29553 benchmarks/agg_sweep_multi.py --dims 1e7,1e8 > gpurun_out/cfg5_n2.log 2>&1; echo "cfg5 n2 rc=$?"; grep "^{" gpurun_out/cfg5_n2.log | cut -c1-300
