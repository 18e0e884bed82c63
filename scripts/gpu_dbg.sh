mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_control_plane.py tests/test_gpu_engine.py -x -q -p no:warnings > gpurun_out/pytest_dbg.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_dbg.log | cut -c1-400
python benchmarks/gram_bench.py --ns 8,16 2>&1 | tail -2 | cut -c1-300
BYZPY_GRAM_SMALL_IMPL=1 python benchmarks/gram_bench.py --ns 8,16 2>&1 | tail -2 | cut -c1-300
python - <<'PY'
import torch, time
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
for n,m,d in [(64,64,10_000_000),(32,32,10_000_000),(128,128,4_000_000),(64,32,10_000_000)]:
    d -= d % 128
    X=torch.randn(n,d,device=dev); W=torch.randn(m,n,device=dev); rows=list(X.unbind(0)); out=torch.empty(m,d,device=dev)
    t1=timeit(lambda: ops.weighted_sum(rows,W,out=out)); t2=timeit(lambda: ops.weighted_sum(rows,W,out=out,multi_impl="passes"))
    byte_ms=(n+m)*d*4/6.4827e9
    print(f"wsum n={n} m={m} d={d}: one-pass {t1:.3f} ms ({t1/byte_ms:.2f}x single-pass bytes), 8-row passes {t2:.3f} ms ({t2/byte_ms:.2f}x)")
    del X,W,out,rows
PY
