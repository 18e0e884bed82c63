mkdir -p gpurun_out
timeout 300 python bench/profile_step.py --workers 3 --streams 1 --out gpurun_out/profile_step_w3.txt 2>&1 | tail -3
timeout 300 python bench.py --steps 50 --trace 2>gpurun_out/trace_n1.txt | tail -1 | cut -c1-200; cat gpurun_out/trace_n1.txt | tail -8
