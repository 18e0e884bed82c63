mkdir -p gpurun_out
nvidia-smi -L | wc -l
for agg in median multikrum; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 tests/multi_gpu/check_fused_round.py --agg $agg > gpurun_out/mg19_$agg.log 2>&1; echo "check $agg rc=$?"; tail -1 gpurun_out/mg19_$agg.log | cut -c1-200
done
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench19_n8.log 2>&1; tail -1 gpurun_out/bench19_n8.log | cut -c1-1500
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/bench19_n4.log 2>&1; tail -1 gpurun_out/bench19_n4.log | cut -c1-300
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 8 --steps 20 --warmup 5 --no-overlap-wgrad > gpurun_out/bench19_n8_noov.log 2>&1; tail -1 gpurun_out/bench19_n8_noov.log | cut -c1-300
