mkdir -p gpurun_out
timeout 400 python bench.py --impl reference --gpus 1 --steps 5 --warmup 3 > gpurun_out/bench29_ref.log 2>&1; echo "ref rc=$?"; tail -1 gpurun_out/bench29_ref.log | cut -c1-300
timeout 1200 python -m pytest tests -m gpu -x -q -p no:warnings > gpurun_out/pytest29.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest29.log | cut -c1-300
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"reduce_partial|apply_kernel|finalize_kernel" -s 30 -c 6 -f -o gpurun_out/bn_stem python bench/bn_layers.py --stem-only > gpurun_out/ncu_bn.log 2>&1; ls -la gpurun_out/bn_stem.ncu-rep
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
