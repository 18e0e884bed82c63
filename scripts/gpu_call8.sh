mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29571"
timeout 200 $TR --nproc-per-node 8 tests/multi_gpu/check_fused_round.py --agg median --buckets 3 --multicast 1 2>&1 | tail -3
timeout 200 $TR --nproc-per-node 8 tests/multi_gpu/check_fused_round.py --agg multikrum --multicast 1 2>&1 | tail -2
for v in "" "--buckets 1" "--multicast 0" "--buckets 1 --multicast 0"; do
  echo "== bench N=8 $v"; timeout 300 $TR --nproc-per-node 8 bench.py --gpus 8 --steps 300 $v 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d.get('round'), d['clocks'])"
done
echo "== bench N=4"; timeout 300 $TR --nproc-per-node 4 bench.py --gpus 4 --steps 300 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d.get('round'))"
echo "== bench N=4 --buckets 1"; timeout 300 $TR --nproc-per-node 4 bench.py --gpus 4 --steps 300 --buckets 1 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d.get('round'))"
