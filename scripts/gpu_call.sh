mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -p no:warnings > gpurun_out/pytest11.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/pytest11.log
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench11_n1.log 2>&1; tail -1 gpurun_out/bench11_n1.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-overlap-wgrad > gpurun_out/bench11_n1_noov.log 2>&1; tail -1 gpurun_out/bench11_n1_noov.log | cut -c1-200
timeout 300 python bench.py --steps 20 --warmup 5 --no-direct-grads > gpurun_out/bench11_n1_nodirect.log 2>&1; tail -1 gpurun_out/bench11_n1_nodirect.log | cut -c1-200
timeout 300 python bench.py --steps 20 --warmup 5 --worker-streams 1 > gpurun_out/bench11_n1_ws1.log 2>&1; tail -1 gpurun_out/bench11_n1_ws1.log | cut -c1-200
timeout 300 python bench.py --steps 20 --warmup 5 --worker-streams 1 --no-overlap-wgrad > gpurun_out/bench11_n1_ws1_noov.log 2>&1; tail -1 gpurun_out/bench11_n1_ws1_noov.log | cut -c1-200
timeout 300 python bench.py --steps 20 --warmup 5 --worker-streams 2 > gpurun_out/bench11_n1_ws2.log 2>&1; tail -1 gpurun_out/bench11_n1_ws2.log | cut -c1-200
timeout 300 python bench.py --steps 20 --warmup 5 --worker-streams 8 > gpurun_out/bench11_n1_ws8.log 2>&1; tail -1 gpurun_out/bench11_n1_ws8.log | cut -c1-200
timeout 200 python bench/profile_step.py --streams 1 --out gpurun_out/profile_step11.txt > gpurun_out/prof11.log 2>&1
