mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_engine.py -x -q -p no:warnings -k "batchnorm or resnet or weighted_sum or capturable or bn" > gpurun_out/pytest_dbg.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_dbg.log | cut -c1-300
echo "== bn_layers cluster"; python bench/bn_layers.py 2>&1 | tail -8 | cut -c1-200
echo "== bn_layers two-launch"; BYZPY_BN_CLUSTER=0 python bench/bn_layers.py 2>&1 | tail -8 | cut -c1-200
for v in 1 0; do echo "== bench ws=1 cluster=$v"; BYZPY_BN_CLUSTER=$v timeout 300 python bench.py --steps 60 --worker-streams 1 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'])"; done
python - <<'PY'
import torch
from byzpy_b200 import ops
dev=torch.device("cuda",0)
def timeit(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/reps
for n,m,d in [(64,64,10_000_000),(32,32,10_000_000),(128,128,4_000_000),(64,32,10_000_000),(48,48,10_000_000)]:
    d -= d % 128
    X=torch.randn(n,d,device=dev); W=torch.randn(m,n,device=dev); rows=list(X.unbind(0)); out=torch.empty(m,d,device=dev)
    t1=timeit(lambda: ops.weighted_sum(rows,W,out=out,multi_impl="multi")); t2=timeit(lambda: ops.weighted_sum(rows,W,out=out,multi_impl="passes"))
    byte_ms=(n+m)*d*4/6.4827e9
    print(f"wsum n={n} m={m} d={d}: one-pass {t1:.3f} ms ({t1/byte_ms:.2f}x bytes, {m*n*d/t1/1e9:.1f} TFMA/s), 8-row passes {t2:.3f} ms ({t2/byte_ms:.2f}x)")
    del X,W,out,rows
PY
