#!/usr/bin/env bash
# First GPU call of the NEXT session (the round-2 budget ran out before these could run): the late tests, the A/B of
# the warp-tiled selection kernel, then -- on an 8-GPU box -- scripts/gpu_tables8.sh.
#   gpurun --timeout 1500 -- 'bash scripts/gpu_next.sh'
mkdir -p gpurun_out
timeout 900 python -m pytest tests/gpu_late/late_cases.py -m gpu -q -p no:warnings -p no:cacheprovider > gpurun_out/late_tests.log 2>&1
echo "late tests rc=$?"; tail -5 gpurun_out/late_tests.log
for n in 32 64 128; do
  timeout 300 python benchmarks/agg_sweep.py --cw-variants --n $n --f 8 --dims 1e7 --out gpurun_out/cw_variants_n$n.json 2>&1 | tail -9
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cw_select_tiled_kernel -s 2 -c 1 -f \
  -o gpurun_out/ncu_cw_tiled python benchmarks/agg_sweep.py --cw-variants --n 64 --dims 1e7 > gpurun_out/ncu_cw_tiled.log 2>&1
ncu -i gpurun_out/ncu_cw_tiled.ncu-rep --page raw --csv > gpurun_out/cw_tiled.raw.csv 2>/dev/null
echo "ncu rc=$?"
