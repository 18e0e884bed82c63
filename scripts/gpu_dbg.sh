echo "== bn_layers cluster<=8MB"; BYZPY_BN_CLUSTER_MAX_MB=8 python bench/bn_layers.py 2>&1 | tail -6 | head -4 | cut -c1-140
echo "== bn_layers cluster<=16MB"; BYZPY_BN_CLUSTER_MAX_MB=16 python bench/bn_layers.py 2>&1 | tail -8 | head -4 | cut -c1-140
for v in 4 8 16; do echo "== bench ws=1 cluster max $v MB"; BYZPY_BN_CLUSTER_MAX_MB=$v timeout 300 python bench.py --steps 60 --worker-streams 1 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'])"; done
