mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29581 --nproc-per-node 8"
timeout 100 $TR tests/multi_gpu/check_fused_round.py --agg median --buckets 3 2>&1 | tail -2
for v in "" "--overlap-grid 16" "--overlap-grid 74" "--buckets 1"; do
  tag=$(echo "$v" | tr -d ' -')
  echo "== N=8 $v"; timeout 200 $TR bench.py --gpus 8 --steps 300 --trace $v 2>gpurun_out/trace_n8_$tag.txt | tail -1 | cut -c104-140; grep -A5 "timeline rank [07]" gpurun_out/trace_n8_$tag.txt | cut -c1-230
done
