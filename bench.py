#!/usr/bin/env python
"""Benchmarks of the BASELINE.json configurations.  Default = the headline (config 2).

    python bench.py --gpus 1                               # headline, N = 1
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29511 bench.py --gpus 8              # headline, N = 8
    python bench.py --impl reference ...                   # the unmodified reference (baseline/_ref)
    python bench.py --config 3|4|5 ...                     # the other BASELINE.json configs

  --config 2  (default, the headline metric) PS steps/sec, device-timed, max over ranks -- ResNet-18,
              8 workers (6 honest + 2 SignFlip Byzantine), CoordinateWiseMedian.
  --config 3  PS steps/sec -- ResNet-50, 8 rows (6 honest replicas + 2 Little rows),
              Bucketing(2) -> Multi-Krum(f=1, q=2).
  --config 4  P2P rounds/sec -- BERT-base (110 M parameters), 8 peers (7 honest + 1 Empire),
              complete topology, GeometricMedian (Weiszfeld).
  --config 5  aggregator sweep: 8 gradients x d floats (d = 1e5 .. 1e8), Median / TrimmedMean /
              Krum / CenteredClipping, ms per aggregate and fraction of the NVLink / HBM roofline.

The workers / peers are a fixed job ("strong" scaling): each of the N ranks hosts 8/N of them.
Both arms run the same workload: synthetic inputs of the named shape in pinned host memory,
random-init weights of the named architecture, bf16 autocast fwd/bwd, fp32 master weights /
gradients / aggregation.

  ours       byzpy_b200 ParameterServer / PeerToPeer over device nodes: one CUDA-graph launch per
             round = the replicas' fwd/bwd + the fused sm_100a aggregation kernels (P2P gather over
             NVLink, selection network or tcgen05 Gram + n-space solve + weighted sum, NVLS multicast
             broadcast, SGD), aggregation buckets enqueued from inside backward.
  reference  byzpy (baseline/_ref, unmodified) node actors + attacks + (pre-)aggregators called
             through their public API on CUDA tensors behind an NCCL all_gather (the reference has no
             GPU collective path; its own ParameterServer.round() raises TypeError with
             CoordinateWiseMedian at this commit and its P2P runner discards the aggregate --
             SURVEY.md 0.4 -- so a round makes the same public node / attack / aggregator calls those
             orchestrators make).

`value` is device-timed with CUDA events around exactly K rounds (ours: `replay()`, inputs resident
on the device; the reference's node API copies its batch H2D inside the round).  `e2e` is the same
metric through the public API including, every round, the H2D copy of that round's inputs from
pinned host memory and a D2H read of the round's losses.
"""
from __future__ import annotations

import argparse
import asyncio
import json
import os
import random
import sys
import threading
import time
from typing import List, Optional

import torch
import torch.distributed as dist

REPO = os.path.dirname(os.path.abspath(__file__))
N_WORKERS = 8
BASELINE_PUBLISHED = None  # BASELINE.json "published": {} -- no reference number for these configs

CONFIGS = {
    2: dict(kind="ps", model="resnet18", n_honest=6, n_byz_workers=2, n_virtual=0, attack="SignFlip",
            aggregator="CoordinateWiseMedian", pre=None, update_byzantines=True,
            metric="PS steps/sec (device-timed, max over ranks) ResNet-18 + CoordinateWiseMedian, 2 Byzantine"),
    3: dict(kind="ps", model="resnet50", n_honest=6, n_byz_workers=0, n_virtual=2, attack="Little(f=2)",
            aggregator="MultiKrum(f=1,q=2)", pre="Bucketing(2)", update_byzantines=False,
            metric="PS steps/sec (device-timed, max over ranks) ResNet-50 + Bucketing->Multi-Krum, 2 Little"),
    4: dict(kind="p2p", model="bert-base", n_honest=7, n_byz=1, attack="Empire(-1)", aggregator="GeometricMedian",
            metric="P2P rounds/sec (device-timed, max over ranks) BERT-base + GeometricMedian, 1 Empire"),
    5: dict(kind="sweep", metric="aggregate latency, 8 gradients x 1e8 floats, CoordinateWiseMedian (ms)"),
}


# --------------------------------------------------------------------------- helpers
def env_int(name: str, default: int) -> int:
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


class ClockSampler:
    """Samples SM clocks / throttle reasons of the local GPU during the timed region."""

    def __init__(self, index: int, period: float = 0.1):
        self.index, self.period = index, period
        self.samples: List[int] = []
        self.reasons = set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._thr: Optional[threading.Thread] = None
        self._nv = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self._nv = None

    _REASONS = {
        0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown",
        0x4: "sw_power_cap", 0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting",
    }

    def _loop(self):
        nv = self._nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                for bit, name in self._REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(self.period)

    def __enter__(self):
        if self._nv is not None:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thr is not None:
            self._thr.join(timeout=2)

    def summary(self) -> dict:
        s = sorted(self.samples)
        med = s[len(s) // 2] if s else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def make_image_pool(n_local: int, batch: int, image: int, classes: int, pool: int, seed: int):
    """Pinned host pool of synthetic uint8 NHWC image batches + labels, per local worker."""
    g = torch.Generator().manual_seed(seed)
    xs, ys = [], []
    for _ in range(n_local):
        xs.append([torch.randint(0, 256, (batch, image, image, 3), dtype=torch.uint8, generator=g).pin_memory()
                   for _ in range(pool)])
        ys.append([torch.randint(0, classes, (batch,), dtype=torch.int64, generator=g).pin_memory()
                   for _ in range(pool)])
    return xs, ys


make_pool = make_image_pool      # (older scripts under bench/ import this name)


def make_token_pool(n_local: int, batch: int, seq: int, vocab: int, pool: int, seed: int):
    """Pinned host pool of synthetic token-id batches (inputs double as MLM targets)."""
    g = torch.Generator().manual_seed(seed)
    return [[torch.randint(0, vocab, (batch, seq), dtype=torch.int64, generator=g).pin_memory() for _ in range(pool)]
            for _ in range(n_local)]


def preprocess_uint8_nhwc(x: torch.Tensor) -> torch.Tensor:
    # uint8 NHWC -> normalised float NCHW view with channels_last strides (no transpose copy)
    return (x.permute(0, 3, 1, 2).float() - 127.5) * (1.0 / 127.5)


def preprocess_fused(x: torch.Tensor) -> torch.Tensor:
    # same map, one streaming kernel of this library, bf16 output (what the first conv consumes)
    from byzpy_b200.ops import normalize_uint8_nhwc

    return normalize_uint8_nhwc(x, 127.5, 127.5, s2d=True)


def max_over_ranks(value: float, device) -> float:
    if dist.is_initialized():
        t = torch.tensor([value], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    return value


def sum_over_ranks(values, device):
    if dist.is_initialized():
        t = torch.tensor(list(values), dtype=torch.float64, device=device)
        dist.all_reduce(t)
        return [int(v) for v in t.tolist()]
    return [int(v) for v in values]


def barrier_sync(device):
    if dist.is_initialized():
        dist.barrier()
    torch.cuda.synchronize(device)


def ps_layout(cfg, world):
    from byzpy_b200.parallel.device_ps import RowLayout

    n_w = cfg["n_honest"] + cfg["n_byz_workers"]
    if n_w % world == 0:
        return RowLayout.block(cfg["n_honest"], cfg["n_byz_workers"], world, cfg["n_virtual"])
    return RowLayout.spread(cfg["n_honest"], cfg["n_byz_workers"], world, cfg["n_virtual"])


def timed_rounds(one_round, steps: int, device, clock_index: int):
    """CUDA events around exactly `steps` calls, barrier + synchronize on both sides, max over ranks."""
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(clock_index) as clk:
        barrier_sync(device)
        e0.record()
        for _ in range(steps):
            one_round()
        e1.record()
        torch.cuda.synchronize(device)
        barrier_sync(device)
    return max_over_ranks(e0.elapsed_time(e1), device), clk.summary()


# ------------------------------------------------------------------------ ours: PS (2, 3)
def run_ours_ps(args, cfg, rank, world, device):
    sys.path.insert(0, REPO)
    from byzpy_b200.aggregators.coordinate_wise import CoordinateWiseMedian
    from byzpy_b200.aggregators.geometric_wise import MultiKrum
    from byzpy_b200.attacks import LittleAttack, SignFlipAttack
    from byzpy_b200.engine.node.device import DeviceByzantineNode, DeviceHonestNode
    from byzpy_b200.engine.parameter_server.ps import ParameterServer
    from byzpy_b200.models import build_model
    from byzpy_b200.pre_aggregators import Bucketing

    layout = ps_layout(cfg, world)
    gids = layout.local_ids(rank)
    L = len(gids)
    n_honest = cfg["n_honest"]
    xs, ys = make_image_pool(L, args.batch, args.image, args.classes, args.pool, seed=1234 + rank)
    honest, byz = [], []
    cursor = [0] * L

    def source(slot):
        # the node's data source: this worker's next pinned host batch (cycles through the pool)
        def nxt():
            k = cursor[slot] % args.pool
            cursor[slot] += 1
            return xs[slot][k], ys[slot][k]
        return nxt

    for slot, g in enumerate(gids):
        torch.manual_seed(0)  # identical init on every replica, like a PS that broadcasts the model
        model = build_model(cfg["model"], num_classes=args.classes)
        kw = dict(lr=args.lr, momentum=0.9, device=str(device), preprocess=preprocess_fused, data=source(slot))
        if g < n_honest:
            honest.append(DeviceHonestNode(model, name=f"honest{g}", **kw))
        else:
            byz.append(DeviceByzantineNode(SignFlipAttack(), model=model, name=f"byz{g}", **kw))
    for v in range(cfg["n_virtual"]):       # omniscient rows synthesised in-kernel; every rank declares them
        byz.append(DeviceByzantineNode(LittleAttack(f=cfg["n_virtual"]), device=str(device), name=f"little{v}"))
    if cfg["aggregator"].startswith("MultiKrum"):
        agg, pre = MultiKrum(f=1, q=2), Bucketing(bucket_size=2, rng=random.Random(args.seed))
    else:
        agg, pre = CoordinateWiseMedian(), None
    ps = ParameterServer(honest, byz, agg, pre_aggregator=pre, update_byzantines=cfg["update_byzantines"],
                         layout=layout, amp_dtype=torch.bfloat16, use_cuda_graph=not args.no_graph,
                         worker_streams=args.worker_streams, fused=True, lr=args.lr, momentum=0.9,
                         direct_grads=not args.no_direct_grads, overlap_wgrad=not args.no_overlap_wgrad,
                         branch_streams=not args.no_branch_streams, buckets=args.buckets,
                         multicast=None if args.multicast < 0 else bool(args.multicast),
                         device_options=dict(trace=args.trace, overlap_grid=args.overlap_grid))
    rnd = ps.device_round

    def batches(i):
        k = i % args.pool
        return [(xs[s][k], ys[s][k]) for s in range(L)]

    # warm-up (includes bucket validation and graph capture)
    for i in range(max(args.warmup, 3)):
        ps.step(batches(i))
    rnd.read_losses()
    barrier_sync(device)

    # ---- device-timed region: exactly K rounds on resident inputs ----
    ms, clocks = timed_rounds(rnd.replay, args.steps, device, device.index)
    rnd.check_status()
    if args.trace:
        print_timeline(rnd, rank, world)

    # ---- end-to-end through the public API: H2D inputs + round + D2H losses every step ----
    # ps.step() pulls every worker's next pinned host batch from its data source; the H2D copy of
    # batch k+1 is issued right after round k is launched (double-buffered inputs), the loss of
    # round k is read back synchronously every step.
    ps.step()                       # fills the prefetch pipeline (and captures the second graph)
    ps.step()
    rnd.read_losses()
    barrier_sync(device)
    t0 = time.perf_counter()
    for i in range(args.steps):
        ps.step()
        losses = rnd.read_losses()
    torch.cuda.synchronize(device)
    e2e_s = max_over_ranks(time.perf_counter() - t0, device)
    barrier_sync(device)
    h2d = sum(x[0].numel() * x[0].element_size() + y[0].numel() * y[0].element_size() for x, y in zip(xs, ys))
    h2d, d2h = sum_over_ranks([h2d, rnd.losses.numel() * 4], device)
    mean_loss = float(losses.mean().item()) if L else float("nan")
    result = dict(ms=ms, e2e_s=e2e_s, h2d=h2d, d2h=d2h, clocks=clocks,
                  launches=(rnd.launches_per_step + rnd.model_launches_per_step) * args.steps,
                  loss=mean_loss, d=rnd.d,
                  round=dict(buckets=rnd.n_buckets if rnd._use_buckets else 1, multicast=bool(rnd._agg_mc),
                             heap=rnd.sym.kind, overlap_grid=rnd.overlap_grid, bounds=list(rnd._bounds)))
    asyncio.run(ps.shutdown())
    return result


def print_timeline(rnd, rank, world):
    tl = rnd.timeline()
    allt = [None] * world
    if dist.is_initialized():
        dist.all_gather_object(allt, tl)
    else:
        allt = [tl]
    if rank == 0:
        for r, t in enumerate(allt):
            print(f"[timeline rank {r}] backward_done={t['backward_done']} round_done={t['round_done']} us", file=sys.stderr)
            for k, b in enumerate(t["buckets"]):
                print(f"    bucket {k} ({b['elements']} el): produced {b['produced']}  start {b['start']}  ready-wait-> "
                      f"{b['ready_wait_done']}  phase1(block0) {b['block0_phase1_done']}  phase1(all) "
                      f"{b['all_ctas_phase1_done']}  delivery-wait-> {b['delivery_wait_done']}  sgd {b['sgd_done']}",
                      file=sys.stderr)


# ------------------------------------------------------------------- reference: PS (2, 3)
def _install_reference(ref_root):
    """baseline/_ref is git-ignored, so a re-created checkout has lost it: re-do the one offline install the
    task allows when the reference source is present.  One rank installs (file lock), the
    others wait on the lock and find it done."""
    import fcntl
    import shutil
    import subprocess
    import tempfile

    src = "/root/reference/python"
    if not os.path.isdir(src):
        return
    os.makedirs(os.path.dirname(ref_root), exist_ok=True)
    with open(os.path.join(os.path.dirname(ref_root), ".install.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        if os.path.isdir(os.path.join(ref_root, "byzpy")):
            return
        tmp = tempfile.mkdtemp(prefix="refsrc_")
        try:
            shutil.copytree(src, os.path.join(tmp, "src"))          # the reference tree is read-only
            subprocess.run([sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps",
                            "--find-links", "/opt/wheelhouse", "--target", ref_root, os.path.join(tmp, "src")],
                           stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=600, check=False)
        finally:
            shutil.rmtree(tmp, ignore_errors=True)


def _ref_imports():
    ref_root = os.path.join(REPO, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref_root, "byzpy")):
        try:
            _install_reference(ref_root)
        except Exception:
            pass
    if not os.path.isdir(os.path.join(ref_root, "byzpy")):
        return "baseline/_ref/byzpy not installed"
    if ref_root not in sys.path:
        sys.path.insert(0, ref_root)
    return None


def run_reference_ps(args, cfg, rank, world, device):
    err = _ref_imports()
    if err:
        return {"unavailable": err}
    try:
        import torchvision
        from byzpy.aggregators.coordinate_wise import CoordinateWiseMedian
        from byzpy.aggregators.geometric_wise import MultiKrum
        from byzpy.attacks import LittleAttack, SignFlipAttack
        from byzpy.configs.actor import set_actor
        from byzpy.engine.graph.pool import ActorPoolConfig
        from byzpy.engine.node.actors import ByzantineNodeActor, HonestNodeActor
        from byzpy.engine.node.distributed import DistributedByzantineNode, DistributedHonestNode
        from byzpy.pre_aggregators import Bucketing
    except Exception as exc:  # pragma: no cover
        return {"unavailable": f"reference import failed: {exc!r}"}

    layout = ps_layout(cfg, world)          # the same placement of the 8 rows over the ranks
    gids = layout.local_ids(rank)
    L = len(gids)
    per = max(1, layout.max_local())
    n_honest, n_virtual = cfg["n_honest"], cfg["n_virtual"]
    n_workers = cfg["n_honest"] + cfg["n_byz_workers"]
    xs, ys = make_image_pool(L, args.batch, args.image, args.classes, args.pool, seed=1234 + rank)
    classes, lr = args.classes, args.lr
    tv_model = getattr(torchvision.models, cfg["model"])

    def flatten(model):
        return torch.cat([(torch.zeros_like(p) if p.grad is None else p.grad).reshape(-1)
                          for p in model.parameters()])

    def write_grads(model, vec):
        off = 0
        for p in model.parameters():
            n = p.numel()
            chunk = vec[off:off + n].view(p.shape)
            if p.grad is None:
                p.grad = chunk.clone()
            else:
                p.grad.copy_(chunk)
            off += n

    class _Common:
        def _setup(self, slot):
            torch.manual_seed(0)
            self.model = tv_model(num_classes=classes).to(device).to(memory_format=torch.channels_last)
            self.opt = torch.optim.SGD(self.model.parameters(), lr=lr, momentum=0.9)
            self.crit = torch.nn.CrossEntropyLoss()
            self.slot, self.it, self.last_loss = slot, 0, None

        def next_batch(self):
            k = self.it % len(xs[self.slot])
            self.it += 1
            return (xs[self.slot][k].to(device, non_blocking=True),
                    ys[self.slot][k].to(device, non_blocking=True))

        def _grad(self, x, y):
            self.model.zero_grad(set_to_none=True)
            with torch.autocast("cuda", dtype=torch.bfloat16):
                loss = self.crit(self.model(preprocess_uint8_nhwc(x)), y)
            loss.backward()
            self.last_loss = loss.detach()
            return flatten(self.model)

        def apply_server_gradient(self, g):
            write_grads(self.model, g.to(device))
            self.opt.step()

        def loss_value(self):
            return self.last_loss

    class RefHonest(_Common, DistributedHonestNode):
        def __init__(self, slot):
            DistributedHonestNode.__init__(
                self, actor_pool=[ActorPoolConfig(backend="gpu", count=1, name="worker")],
                aggregator=CoordinateWiseMedian(), name=f"honest{slot}")
            self._setup(slot)

        def local_honest_gradient(self, *, x, y):
            return self._grad(x, y)

    class RefSignFlip(_Common, DistributedByzantineNode):
        def __init__(self, slot):
            DistributedByzantineNode.__init__(
                self, actor_pool=[ActorPoolConfig(backend="gpu", count=1, name="worker")],
                attack=SignFlipAttack(), name=f"byz{slot}")
            self._setup(slot)

        def prepare_attack_inputs(self, *, x=None, y=None, honest_grads=None, base_grad=None, model=None):
            # SignFlip needs the node's own gradient: compute it on this node's batch
            xb, yb = self.next_batch()
            return {"base_grad": self._grad(xb, yb)}

    class RefLittle(DistributedByzantineNode):
        """Omniscient node: no model, its vector is LittleAttack.apply(honest_grads=...)."""

        def __init__(self, slot):
            DistributedByzantineNode.__init__(
                self, actor_pool=[ActorPoolConfig(backend="gpu", count=1, name="worker")],
                attack=LittleAttack(f=n_virtual), name=f"little{slot}")

        def next_batch(self):
            return torch.empty(0), torch.empty(0, dtype=torch.long)

        def apply_server_gradient(self, g):
            return None

    async def build():
        hon, byz, lit = [], [], []
        for slot, g in enumerate(gids):
            if g < n_honest:
                hon.append(await HonestNodeActor.spawn(RefHonest, backend=set_actor("thread"), kwargs=dict(slot=slot)))
            else:
                byz.append(await ByzantineNodeActor.spawn(RefSignFlip, backend=set_actor("thread"),
                                                          kwargs=dict(slot=slot)))
        for v in range(n_virtual):      # stateless omniscient nodes: every rank runs its own copy on the gathered rows
            lit.append(await ByzantineNodeActor.spawn(RefLittle, backend=set_actor("thread"), kwargs=dict(slot=v)))
        return hon, byz, lit

    if cfg["aggregator"].startswith("MultiKrum"):
        agg, pre = MultiKrum(f=1, q=2), Bucketing(bucket_size=2, rng=random.Random(args.seed))
    else:
        agg, pre = CoordinateWiseMedian(), None
    d_model = sum(p.numel() for p in tv_model(num_classes=classes).parameters())

    async def one_round(hon, byz, lit):
        grads = list(await asyncio.gather(*[h.honest_gradient_for_next_batch() for h in hon]))
        grads += list(await asyncio.gather(*[b.byzantine_gradient_for_next_batch(tuple(grads)) for b in byz]))
        if world > 1:
            local = torch.zeros((per, d_model), dtype=torch.float32, device=device)
            for k, gvec in enumerate(grads):
                local[k] = gvec
            full = torch.empty((world, per, d_model), dtype=torch.float32, device=device)
            dist.all_gather_into_tensor(full.view(-1), local.view(-1))
            grads = [full[layout.rank_of[g], layout.slot_of[g]] for g in range(n_workers)]
        if lit:
            honest_rows = tuple(grads[:n_honest])
            grads = list(grads) + list(await asyncio.gather(
                *[b.byzantine_gradient_for_next_batch(honest_rows) for b in lit]))
        if pre is not None:
            grads = list(pre.pre_aggregate(grads))
        g = agg.aggregate(grads)
        targets = hon + (byz if cfg["update_byzantines"] else [])
        await asyncio.gather(*[n.apply_server_gradient(g) for n in targets])
        return g

    async def losses(hon, byz):
        vals = await asyncio.gather(*[n.loss_value() for n in hon + byz])
        return torch.stack([v.float() for v in vals]).cpu() if vals else torch.zeros(0)

    async def main():
        hon, byz, lit = await build()
        for _ in range(max(min(args.warmup, 5), 3)):
            await one_round(hon, byz, lit)
        barrier_sync(device)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ClockSampler(device.index) as clk:
            barrier_sync(device)
            e0.record()
            for _ in range(args.steps):
                await one_round(hon, byz, lit)
            e1.record()
            torch.cuda.synchronize(device)
            barrier_sync(device)
        ms = max_over_ranks(e0.elapsed_time(e1), device)
        barrier_sync(device)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            await one_round(hon, byz, lit)
            lv = await losses(hon, byz)
        torch.cuda.synchronize(device)
        e2e_s = max_over_ranks(time.perf_counter() - t0, device)
        h2d = sum(x[0].numel() * x[0].element_size() + y[0].numel() * y[0].element_size() for x, y in zip(xs, ys))
        h2d, d2h = sum_over_ranks([h2d, 4 * (len(hon) + len(byz))], device)
        res = dict(ms=ms, e2e_s=e2e_s, h2d=h2d, d2h=d2h, clocks=clk.summary(), launches=0,
                   loss=float(lv.mean().item()) if lv.numel() else float("nan"), d=d_model)
        for n in hon + byz + lit:
            await n._ref._backend.close()
        return res

    return asyncio.run(main())


# ------------------------------------------------------------------------ ours: P2P (4)
def _mlm_loss(out, y):
    return torch.nn.functional.cross_entropy(out.flatten(0, 1), y.flatten())


def run_ours_p2p(args, cfg, rank, world, device):
    sys.path.insert(0, REPO)
    from byzpy_b200.aggregators.geometric_wise import GeometricMedian
    from byzpy_b200.attacks import EmpireAttack
    from byzpy_b200.engine.node.device import DeviceP2PByzantineNode, DeviceP2PHonestNode
    from byzpy_b200.engine.peer_to_peer.topology import Topology
    from byzpy_b200.engine.peer_to_peer.train import PeerToPeer
    from byzpy_b200.models import build_model
    from byzpy_b200.parallel.device_p2p import PeerLayout

    peers = cfg["n_honest"] + cfg["n_byz"]
    layout = PeerLayout(cfg["n_honest"], cfg["n_byz"], world)
    gids = layout.local_ids(rank)
    hon_ids = [g for g in gids if g < layout.n_honest]
    toks = make_token_pool(len(hon_ids), args.p2p_batch, args.seq, 30522, args.pool, seed=77 + rank)
    cursor = [0] * len(hon_ids)

    def source(slot):
        def nxt():
            k = cursor[slot] % args.pool
            cursor[slot] += 1
            return toks[slot][k], toks[slot][k]
        return nxt

    hon, byz = [], []
    for g in gids:
        if g < layout.n_honest:
            torch.manual_seed(0)
            hon.append(DeviceP2PHonestNode(build_model(cfg["model"]), GeometricMedian(), loss_fn=_mlm_loss,
                                           data=source(len(hon)), device=str(device), name=f"peer{g}"))
        else:
            byz.append(DeviceP2PByzantineNode(EmpireAttack(scale=-1.0), device=str(device), name=f"empire{g}"))
    p2p = PeerToPeer(hon, byz, Topology.complete(peers), lr=args.p2p_lr, layout=layout, fused=True,
                     amp_dtype=torch.bfloat16, use_cuda_graph=not args.no_graph)
    rnd = p2p.device_round
    for _ in range(max(min(args.warmup, 5), 3)):
        p2p.step()
    rnd.read_losses()
    barrier_sync(device)
    ms, clocks = timed_rounds(rnd.replay, args.steps, device, device.index)
    rnd.check_status()
    barrier_sync(device)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        p2p.step()                  # H2D of every honest peer's token batch from pinned memory + the round
        losses = rnd.read_losses()
    torch.cuda.synchronize(device)
    e2e_s = max_over_ranks(time.perf_counter() - t0, device)
    h2d = sum(2 * t[0].numel() * t[0].element_size() for t in toks)
    h2d, d2h = sum_over_ranks([h2d, rnd.losses.numel() * 4], device)
    hl = losses[: len(hon)] if len(hon) else losses
    res = dict(ms=ms, e2e_s=e2e_s, h2d=h2d, d2h=d2h, clocks=clocks, launches=rnd.launches_per_step * args.steps,
               loss=float(hl.mean().item()) if hl.numel() else float("nan"), d=rnd.d,
               round=dict(heap=rnd.sym.kind, staged_remote_vectors=len(rnd._staged_ids)))
    asyncio.run(p2p.shutdown())
    return res


# ------------------------------------------------------------------- reference: P2P (4)
def run_reference_p2p(args, cfg, rank, world, device):
    err = _ref_imports()
    if err:
        return {"unavailable": err}
    try:
        import transformers
        from byzpy.aggregators.geometric_wise import GeometricMedian
        from byzpy.attacks import EmpireAttack
        from byzpy.engine.node.mixin import P2PByzantineMixin, P2PHonestMixin
    except Exception as exc:  # pragma: no cover
        return {"unavailable": f"reference import failed: {exc!r}"}

    peers, n_h = cfg["n_honest"] + cfg["n_byz"], cfg["n_honest"]
    per = peers // world
    gids = [g for g in range(peers) if g // per == rank]
    hon_ids = [g for g in gids if g < n_h]
    toks = make_token_pool(len(hon_ids), args.p2p_batch, args.seq, 30522, args.pool, seed=77 + rank)
    lr = args.p2p_lr

    class RefPeer(P2PHonestMixin):
        """The reference's P2P step functions (engine/node/mixin.py:59-80) on a stock BERT-base."""

        def __init__(self, slot):
            torch.manual_seed(0)
            self.device = device
            self.model = transformers.BertForMaskedLM(transformers.BertConfig()).to(device)
            self.model.train()
            self.criterion = self._loss
            self.p2p_agg = GeometricMedian()
            self.p2p_pre = None
            self.slot, self.it, self.last_loss = slot, 0, None

        def _loss(self, out, y):
            loss = torch.nn.functional.cross_entropy(out.logits.flatten(0, 1).float(), y.flatten())
            self.last_loss = loss.detach()
            return loss

        def next_batch(self):
            k = self.it % len(toks[self.slot])
            self.it += 1
            t = toks[self.slot][k].to(device, non_blocking=True)
            return t, t

        def p2p_half_step(self, lr):                # same call, under the benchmark's bf16 autocast
            with torch.autocast("cuda", dtype=torch.bfloat16):
                return P2PHonestMixin.p2p_half_step(self, lr)

    class RefEmpire(P2PByzantineMixin):
        def __init__(self):
            self.device = device
            self.attack = EmpireAttack(scale=-1.0)

    hon = [RefPeer(s) for s in range(len(hon_ids))]
    d_model = sum(p.numel() for p in (hon[0].model.parameters() if hon else []))
    if world > 1:
        t = torch.tensor([d_model], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        d_model = int(t.item())
    empire = RefEmpire()

    def one_round():
        halves = torch.zeros((per, d_model), dtype=torch.float32, device=device)
        for slot, node in enumerate(hon):
            halves[gids.index(hon_ids[slot])] = node.p2p_half_step(lr)
        if world > 1:
            full = torch.empty((world, per, d_model), dtype=torch.float32, device=device)
            dist.all_gather_into_tensor(full.view(-1), halves.view(-1))
        else:
            full = halves.view(1, per, d_model)
        vec = {g: full[g // per, g % per] for g in range(peers)}
        honest_vecs = [vec[g] for g in range(n_h)]
        mal = None
        if n_h < peers:         # complete topology: the Empire peer sees every honest vector (every rank recomputes it)
            mal = empire.p2p_broadcast_vector(neighbor_vectors=honest_vecs, like=honest_vecs[0])
        for slot, node in enumerate(hon):
            g = hon_ids[slot]
            others = [vec[j] for j in range(n_h) if j != g] + ([mal] if mal is not None else [])
            node.p2p_aggregate_and_set(vec[g], others)

    for _ in range(2):
        one_round()
    barrier_sync(device)
    ms, clocks = timed_rounds(one_round, args.steps, device, device.index)
    barrier_sync(device)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        one_round()
        lv = torch.stack([n.last_loss.float() for n in hon]).cpu() if hon else torch.zeros(0)
    torch.cuda.synchronize(device)
    e2e_s = max_over_ranks(time.perf_counter() - t0, device)
    h2d = sum(t[0].numel() * t[0].element_size() for t in toks)
    h2d, d2h = sum_over_ranks([h2d, 4 * len(hon)], device)
    return dict(ms=ms, e2e_s=e2e_s, h2d=h2d, d2h=d2h, clocks=clocks, launches=0,
                loss=float(lv.mean().item()) if lv.numel() else float("nan"), d=d_model)


# ------------------------------------------------------------------------- sweep (5)
SWEEP_AGGS = ("median", "trmean", "krum", "cclip")


def run_sweep(args, rank, world, device, impl):
    """8 gradient rows x d floats, block-distributed over the ranks.  ours: the fused round without
    replicas (symmetric rows, P2P gather + aggregate + multicast broadcast, no SGD); reference: NCCL
    all_gather of the rows + the reference operator's direct path on the gathered CUDA tensors."""
    sys.path.insert(0, REPO)
    if impl == "reference":
        err = _ref_imports()
        if err:
            return {"unavailable": err}
        from byzpy.aggregators.coordinate_wise import CoordinateWiseMedian, CoordinateWiseTrimmedMean
        from byzpy.aggregators.geometric_wise import Krum
        from byzpy.aggregators.norm_wise import CenteredClipping
    else:
        from byzpy_b200.aggregators.coordinate_wise import CoordinateWiseMedian, CoordinateWiseTrimmedMean
        from byzpy_b200.aggregators.geometric_wise import Krum
        from byzpy_b200.aggregators.norm_wise import CenteredClipping
    mk = {"median": lambda: CoordinateWiseMedian(), "trmean": lambda: CoordinateWiseTrimmedMean(f=2),
          "krum": lambda: Krum(f=2), "cclip": lambda: CenteredClipping(c_tau=10.0, M=10)}
    n = N_WORKERS
    per = n // world
    rows_table = []
    dims = [int(float(x)) for x in args.sweep_dims.split(",")]
    for d in dims:
        g = torch.Generator(device=device).manual_seed(100 + rank)
        local = torch.randn((per, d), generator=g, device=device)
        for name in SWEEP_AGGS:
            if impl == "reference":
                agg = mk[name]()
                full = torch.empty((world, per, d), device=device) if world > 1 else None

                def once(agg=agg, full=full):
                    if world > 1:
                        dist.all_gather_into_tensor(full.view(-1), local.view(-1))
                        rows = list(full.view(n, d).unbind(0))
                    else:
                        rows = list(local.unbind(0))
                    return agg.aggregate(rows)
                closer = None
            else:
                once, closer = _sweep_ours(mk[name](), local, per, d, rank, world, device, args)
            for _ in range(3):
                once()
            reps = max(3, min(args.sweep_reps, int(2e9 // (n * d)) + 3))
            ms, _clk = timed_rounds(once, reps, device, device.index)
            ms /= reps
            if closer is not None:
                closer()
            # bytes a rank must move: its shard of the 8 rows (7/8 of them over NVLink at world > 1;
            # the Gram family reads them twice) + the broadcast of its shard of the result
            passes = 2 if name in ("krum", "cclip") else 1
            local_bytes = passes * n * d * 4 / world + d * 4
            remote_bytes = passes * (world - 1) / world * n * d * 4 / world + (world - 1) / world * d * 4
            roof_ms = max(local_bytes / (args.hbm_gbs * 1e6), remote_bytes / (args.nvlink_gbs * 1e6))
            rows_table.append(dict(agg=name, d=d, ms=round(ms, 4), roofline_ms=round(roof_ms, 4),
                                   frac=round(roof_ms / ms, 3) if ms > 0 else None))
            del once
        del local
        torch.cuda.empty_cache()
    head = next((r for r in rows_table if r["agg"] == "median" and r["d"] == max(dims)), rows_table[0])
    return dict(sweep=rows_table, ms=head["ms"], d=head["d"])


def _sweep_ours(agg, local, per, d, rank, world, device, args):
    """A DeviceRound without model replicas: the rows are written into its symmetric gradient rows."""
    from byzpy_b200.parallel.device_ps import DeviceRound, DeviceWorker, RowLayout

    class _Vec(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.v = torch.nn.Parameter(torch.zeros(d))

    layout = RowLayout.block(N_WORKERS, 0, world)
    workers = [DeviceWorker(_Vec(), lambda o, y: o, role="honest") for _ in range(per)]
    rnd = DeviceRound(workers, layout, agg.fused_plan(N_WORKERS), lr=0.0, momentum=0.0, device=device,
                      amp_dtype=None, use_cuda_graph=False, direct_grads=False, buckets=1,
                      multicast=None if args.multicast < 0 else bool(args.multicast))
    rnd.grads[:, :d].copy_(local)
    rnd._upd_params, rnd._upd_moms = [], []     # aggregate + broadcast only (the sweep has no optimizer step)
    torch.cuda.synchronize(device)
    return rnd.launch_aggregate, rnd.close


# ------------------------------------------------------------------------------ main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None, help="timed rounds (default 200; 20 for config 4)")
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--config", type=int, choices=sorted(CONFIGS), default=2)
    ap.add_argument("--model", default=None, help="override the config's model (PS configs)")
    ap.add_argument("--batch", type=int, default=32, help="per-worker batch (PS configs)")
    ap.add_argument("--image", type=int, default=224)
    ap.add_argument("--classes", type=int, default=1000)
    ap.add_argument("--pool", type=int, default=4, help="distinct pinned host batches per worker")
    ap.add_argument("--lr", type=float, default=0.05)
    ap.add_argument("--p2p-batch", type=int, default=8, help="per-peer batch (config 4)")
    ap.add_argument("--seq", type=int, default=128, help="tokens per sequence (config 4)")
    ap.add_argument("--p2p-lr", type=float, default=0.01)
    ap.add_argument("--seed", type=int, default=1234, help="shared seed of randomised operators (Bucketing)")
    ap.add_argument("--sweep-dims", default="1e5,1e6,1e7,1e8", help="config 5: gradient sizes")
    ap.add_argument("--sweep-reps", type=int, default=50)
    ap.add_argument("--hbm-gbs", type=float, default=None, help="roofline denominators (default MEASURED_PEAKS.json)")
    ap.add_argument("--nvlink-gbs", type=float, default=770.0)
    ap.add_argument("--worker-streams", type=int, default=8)
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--buckets", type=int, default=None,
                    help="gradient buckets of the fused round (default: automatic; 1 = one launch after backward)")
    ap.add_argument("--overlap-grid", type=int, default=None,
                    help="CTAs of a bucket launch that overlaps backward (default: a quarter of the SMs)")
    ap.add_argument("--trace", action="store_true",
                    help="print every rank's device-side timeline of the last timed round (stderr)")
    ap.add_argument("--multicast", type=int, default=-1,
                    help="NVLS multicast broadcast: -1 when supported, 0 peer stores, 1 required")
    ap.add_argument("--no-direct-grads", action="store_true",
                    help="A/B: stock autograd gradient accumulation instead of in-place arena gradients")
    ap.add_argument("--no-branch-streams", action="store_true",
                    help="A/B: projection shortcuts of the residual blocks on the worker stream")
    ap.add_argument("--no-overlap-wgrad", action="store_true",
                    help="A/B: weight-gradient GEMMs on the worker stream instead of a side stream")
    args = ap.parse_args()
    cfg = dict(CONFIGS[args.config])
    if args.model and cfg["kind"] == "ps":
        cfg["model"] = args.model
    if args.steps is None:
        args.steps = 20 if args.config == 4 else 200
    if args.hbm_gbs is None:
        try:
            args.hbm_gbs = float(json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))["hbm_gbs"])
        except Exception:
            args.hbm_gbs = 6650.0

    world = env_int("WORLD_SIZE", 1)
    rank = env_int("RANK", 0)
    local_rank = env_int("LOCAL_RANK", 0)
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if args.gpus > 1 and world == 1:
        raise SystemExit("launch multi-GPU runs with torch.distributed.run (one rank per GPU)")
    if N_WORKERS % world != 0:
        raise SystemExit("--gpus must divide 8")
    if not torch.cuda.is_available():
        if args.impl == "reference":
            print(json.dumps({"impl": "reference", "unavailable": "no CUDA device"}))
            return
        raise SystemExit("bench.py needs a CUDA device (run under gpurun)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)
    torch.backends.cudnn.benchmark = True

    if cfg["kind"] == "ps":
        res = (run_ours_ps if args.impl == "ours" else run_reference_ps)(args, cfg, rank, world, device)
    elif cfg["kind"] == "p2p":
        res = (run_ours_p2p if args.impl == "ours" else run_reference_p2p)(args, cfg, rank, world, device)
    else:
        res = run_sweep(args, rank, world, device, args.impl)
    if rank == 0:
        if "unavailable" in res:
            print(json.dumps({"impl": "reference", "unavailable": res["unavailable"]}))
        elif cfg["kind"] == "sweep":
            print(json.dumps({
                "metric": cfg["metric"], "value": res["ms"], "unit": "ms", "n_gpus": world, "steps": args.sweep_reps,
                "warmup": 3, "ms_per_step": res["ms"], "higher_is_better": False, "scaling": "strong",
                "vs_baseline": None, "dtype": "fp32", "data": "synthetic", "impl": args.impl,
                "config": {"rows": N_WORKERS, "dims": args.sweep_dims, "aggregators": list(SWEEP_AGGS),
                           "roofline": f"max(HBM bytes / {args.hbm_gbs} GB/s, NVLink bytes / {args.nvlink_gbs} GB/s)",
                           "l2": "gradient rows of 1e7+ floats exceed the 126 MB L2; smaller sizes are L2 resident"},
                "sweep": res["sweep"]}))
        else:
            sps = args.steps / (res["ms"] / 1e3)
            e2e = args.steps / res["e2e_s"]
            n_rows = N_WORKERS
            conf = {"model": cfg["model"], "aggregator": cfg["aggregator"], "attack": cfg["attack"],
                    "grad_dim": res["d"], "optimizer": "fp32 master weights",
                    "timing": "CUDA events around exactly K rounds, barrier+synchronize both sides, max over ranks",
                    "l2": "per-round working set (fp32 gradient rows + parameters + activations) exceeds the "
                          "126 MB L2; no explicit flush"}
            if cfg["kind"] == "ps":
                conf.update({"workers": n_rows, "byzantine": cfg["n_byz_workers"] + cfg["n_virtual"],
                             "pre_aggregator": cfg["pre"], "per_worker_batch": args.batch,
                             "global_batch": args.batch * (cfg["n_honest"] + cfg["n_byz_workers"]),
                             "image": [3, args.image, args.image], "classes": args.classes,
                             "parallelism": f"byzantine-dp{n_rows} over {world} gpu(s)",
                             "optimizer": f"SGD(momentum=0.9), fp32 master weights, "
                                          f"update_byzantines={cfg['update_byzantines']}"})
            else:
                conf.update({"peers": n_rows, "byzantine": cfg["n_byz"], "topology": "complete",
                             "per_peer_batch": args.p2p_batch, "seq_len": args.seq,
                             "global_batch": args.p2p_batch * cfg["n_honest"],
                             "parallelism": f"gossip-dp{n_rows} over {world} gpu(s)",
                             "optimizer": f"local SGD half step (lr={args.p2p_lr}) + robust aggregation of parameters"})
            out = {
                "metric": cfg["metric"],
                "value": round(sps, 3), "unit": "steps/s" if cfg["kind"] == "ps" else "rounds/s", "n_gpus": world,
                "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": round(res["ms"] / args.steps, 4),
                "higher_is_better": True, "scaling": "strong",
                "vs_baseline": (sps / BASELINE_PUBLISHED) if BASELINE_PUBLISHED else None,
                "dtype": "bf16", "data": "synthetic",
                "impl": args.impl,
                "config": conf,
                "clocks": res["clocks"],
                "e2e": {"value": round(e2e, 3), "unit": "steps/s" if cfg["kind"] == "ps" else "rounds/s",
                        "h2d_bytes_per_step": res["h2d"], "d2h_bytes_per_step": res["d2h"]},
                "gpu_launches": res["launches"],
                "round": res.get("round"),
                "final_loss": round(res["loss"], 4),
            }
            print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
