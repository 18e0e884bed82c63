mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -p no:warnings > gpurun_out/pytest17.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/pytest17.log | cut -c1-300
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench17_n1.log 2>&1; tail -1 gpurun_out/bench17_n1.log | cut -c1-1600
BYZPY_B200_NO_PDL=1 timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench17_n1_nopdl.log 2>&1; tail -1 gpurun_out/bench17_n1_nopdl.log | cut -c1-200
timeout 300 python bench.py --steps 20 --warmup 5 --worker-streams 1 > gpurun_out/bench17_n1_ws1.log 2>&1; tail -1 gpurun_out/bench17_n1_ws1.log | cut -c1-200
BYZPY_B200_NO_PDL=1 timeout 300 python bench.py --steps 20 --warmup 5 --worker-streams 1 > gpurun_out/bench17_n1_ws1_nopdl.log 2>&1; tail -1 gpurun_out/bench17_n1_ws1_nopdl.log | cut -c1-200
timeout 300 python bench.py --steps 20 --warmup 5 --worker-streams 8 > gpurun_out/bench17_n1_ws8.log 2>&1; tail -1 gpurun_out/bench17_n1_ws8.log | cut -c1-200
timeout 200 python bench/bn_layers.py 2>&1 | tail -12
