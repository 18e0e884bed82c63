#!/usr/bin/env bash
# compute-sanitizer passes over the hand-written kernels (run on a GPU box:
#   gpurun --timeout 1500 -- 'bash scripts/sanitize.sh' ; logs under gpurun_out/sanitize_*.log).
#   memcheck   out-of-bounds / misaligned global + shared accesses
#   racecheck  shared-memory hazards (selection networks' staging ring, Gram conversion ring, n-space solvers)
#   synccheck  divergent barriers / invalid mbarrier use (tcgen05 pipeline, wait_all)
#   initcheck  reads of uninitialised global memory (scratch / partial buffers)
# The cross-GPU flag protocol is exercised by bench/stress_bucket_protocol.py with both ranks on one device
# (sanitizers cannot follow peer mappings); its own per-epoch oracle check is the race detector there.
set -u
mkdir -p gpurun_out
CS=${CS:-compute-sanitizer}
KTESTS='tests/test_gpu_kernels.py -k "cw_select_matches_reference or gram or weighted_sum or colstat or scale or virtual"'
rc=0
for tool in memcheck racecheck synccheck initcheck; do
  echo "== $tool: bucket protocol stress"
  timeout 900 $CS --tool $tool --error-exitcode 9 python bench/stress_bucket_protocol.py --epochs 6 --d 32768 \
      > gpurun_out/sanitize_${tool}_protocol.log 2>&1; r=$?; tail -3 gpurun_out/sanitize_${tool}_protocol.log; [ $r -ne 0 ] && rc=1
done
for tool in memcheck racecheck; do
  echo "== $tool: kernel numerics tests (subset)"
  timeout 1500 bash -c "$CS --tool $tool --error-exitcode 9 python -m pytest $KTESTS -x -q -p no:warnings -p no:cacheprovider" \
      > gpurun_out/sanitize_${tool}_kernels.log 2>&1; r=$?; tail -3 gpurun_out/sanitize_${tool}_kernels.log; [ $r -ne 0 ] && rc=1
done
echo "sanitize rc=$rc"
exit $rc
