"""Multi-Krum: direct call vs NodeScheduler(pool=None) vs ActorPool x {2,4,6} workers
(counterpart of the reference's benchmarks/pytorch/multikrum_actor_pool.py).  Thin front-end of
benchmarks/operator_pool_bench.py with the operator fixed; all of its flags apply
(--num-grads --grad-dim --f --pool-workers --pool-backend --device --warmup --repeat --seed).

    python benchmarks/pytorch/multikrum_actor_pool.py --num-grads 64 --grad-dim 65536 --pool-workers 2,4,6
"""
import asyncio
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from benchmarks.operator_pool_bench import main  # noqa: E402

if __name__ == "__main__":
    if "--op" not in sys.argv:
        sys.argv += ["--op", "multi-krum"]
    asyncio.run(main())
