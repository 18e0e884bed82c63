mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -p no:warnings > gpurun_out/pytest_gpu2.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu2.log | cut -c1-400
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561"
timeout 200 $TR tests/multi_gpu/check_fused_round.py --agg median --buckets 3 --multicast 1 2>&1 | tail -4
timeout 200 $TR tests/multi_gpu/check_fused_round.py --agg multikrum --multicast 1 2>&1 | tail -3
for v in "" "--buckets 1" "--multicast 0" "--buckets 1 --multicast 0"; do
  echo "== bench N=2 $v"; timeout 300 $TR bench.py --gpus 2 --steps 200 $v 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d.get('round'))"
done
