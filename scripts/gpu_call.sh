mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 1500 python -m pytest tests -m gpu -x -q -p no:warnings > gpurun_out/pytest18.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/pytest18.log | cut -c1-300
for agg in median multikrum; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/multi_gpu/check_fused_round.py --agg $agg --graph > gpurun_out/mg18_$agg.log 2>&1; echo "check $agg rc=$?"; tail -2 gpurun_out/mg18_$agg.log | cut -c1-200
done
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench18_n2.log 2>&1; tail -1 gpurun_out/bench18_n2.log | cut -c1-1500
