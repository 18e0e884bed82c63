mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/multi_gpu/check_fused_round.py --agg median --workers 3 --attack little > gpurun_out/mg36.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/mg36.log | cut -c1-250
