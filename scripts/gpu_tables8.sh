#!/usr/bin/env bash
# Round tables on an 8-GPU box: every BASELINE.json config at N = 8 (ours, and the reference arm where it is
# affordable), then N = 4 / N = 2 runs packed side by side on disjoint GPU subsets.  Writes gpurun_out/tables/*.json.
mkdir -p gpurun_out/tables
T=gpurun_out/tables
port=29600
run() {  # run <name> <ngpus> <visible> <timeout> bench args...
  name=$1; n=$2; vis=$3; to=$4; shift 4
  port=$((port+1))
  if [ "$n" = "1" ]; then
    CUDA_VISIBLE_DEVICES=$vis timeout $to python bench.py --gpus 1 "$@" > $T/$name.json 2> $T/$name.err
  else
    CUDA_VISIBLE_DEVICES=$vis timeout $to python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n "$@" > $T/$name.json 2> $T/$name.err
  fi
  echo "$name rc=$? $(tail -c 300 $T/$name.json | head -c 300 | tr '\n' ' ' | cut -c1-160)"
}
ALL=0,1,2,3,4,5,6,7
# ---- correctness at world 8 on the final tree (logs kept under profiles/)
for a in "median --buckets 3 --multicast 1" "multikrum --multicast 1" "trmean --attack little --workers 8"; do
  port=$((port+1)); timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $port tests/multi_gpu/check_fused_round.py --agg $a 2>&1 | grep -E "MULTI_GPU|buckets=" | tee -a $T/check_world8.log
done
port=$((port+1)); timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $port tests/multi_gpu/check_fused_round.py --agg multikrum --workers 20 2>&1 | grep -E "MULTI_GPU" | tee -a $T/check_world8.log
# ---- N = 8
run c2_ours_n8 8 $ALL 200 --steps 300
run c3_ours_n8 8 $ALL 240 --config 3 --steps 100
run c4_ours_n8 8 $ALL 240 --config 4 --steps 20
run c5_ours_n8 8 $ALL 240 --config 5 --sweep-reps 20
run c3_ref_n8 8 $ALL 240 --config 3 --steps 8 --warmup 3 --impl reference
run c5_ref_n8 8 $ALL 240 --config 5 --sweep-reps 5 --sweep-dims 1e6,1e7,1e8 --impl reference
run c4_ref_n8 8 $ALL 200 --config 4 --steps 2 --impl reference
# ---- N = 4 (two jobs side by side), N = 2 (four side by side)
run c2_ours_n4 4 0,1,2,3 200 --steps 300 &
run c3_ours_n4 4 4,5,6,7 240 --config 3 --steps 100 &
wait
run c4_ours_n4 4 0,1,2,3 240 --config 4 --steps 20 &
run c5_ours_n4 4 4,5,6,7 240 --config 5 --sweep-reps 20 &
wait
run c2_ours_n2 2 0,1 200 --steps 300 &
run c3_ours_n2 2 2,3 240 --config 3 --steps 60 &
run c4_ours_n2 2 4,5 240 --config 4 --steps 20 &
run c5_ours_n2 2 6,7 240 --config 5 --sweep-reps 20 &
wait
run c2_ours_n1 1 0 200 --steps 200 &
run c3_ours_n1 1 1 240 --config 3 --steps 40 &
run c4_ours_n1 1 2 240 --config 4 --steps 20 &
run c5_ours_n1 1 3 240 --config 5 --sweep-reps 20 &
run c2_ref_n1 1 4 240 --steps 20 --warmup 3 --impl reference &
run c3_ref_n1 1 5 240 --config 3 --steps 8 --warmup 3 --impl reference &
run c5_ref_n1 1 6 240 --config 5 --sweep-reps 5 --sweep-dims 1e6,1e7,1e8 --impl reference &
wait
ls $T | wc -l
