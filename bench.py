#!/usr/bin/env python
"""Headline benchmark: Byzantine-robust parameter-server training throughput.

Metric (BASELINE.json): PS steps/sec, device-timed, max over ranks -- ResNet-18, 8 workers
(6 honest + 2 SignFlip Byzantine), CoordinateWiseMedian, on N = 1/2/4/8 B200 of one node.
The 8 workers are a fixed job ("strong" scaling): each of the N ranks hosts 8/N replicas.

    python bench.py --gpus 1 --steps 20 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5
    python bench.py --impl reference ...        # the unmodified reference (baseline/_ref)

Both arms run the same workload: synthetic ImageNet-shaped uint8 batches (per-worker batch 32,
3x224x224, 1000 classes) in pinned host memory, random-init ResNet-18, bf16 autocast +
channels_last fwd/bwd, fp32 master weights / gradients, SGD(momentum 0.9), every node
(Byzantine included) applies the aggregated gradient.

  ours       byzpy_b200.engine.parameter_server.ParameterServer over device nodes: per round one
             CUDA-graph launch = 8/N x (fwd/bwd) + ONE fused sm_100a kernel (P2P gather over
             NVLink + median selection network + P2P broadcast + SGD on all local replicas).
  reference  byzpy (baseline/_ref) node actors + byzpy SignFlipAttack + byzpy
             CoordinateWiseMedian.aggregate on CUDA tensors behind an NCCL all_gather (the
             reference has no GPU collective path and its own ParameterServer.round() raises
             TypeError with this aggregator at this commit -- SURVEY.md 0.4 -- so the round is
             driven through the same public node/attack/aggregator calls round() makes).

`value` is device-timed with CUDA events around exactly K rounds (inputs resident on the
device for ours; the reference's node API copies its batch H2D inside the round).  `e2e` is the
same metric through the public API including, every round, the H2D copy of that round's inputs
from pinned host memory and a D2H read of the round's losses.
"""
from __future__ import annotations

import argparse
import asyncio
import json
import os
import sys
import threading
import time
from typing import List, Optional

import torch
import torch.distributed as dist

REPO = os.path.dirname(os.path.abspath(__file__))
N_WORKERS, N_BYZ = 8, 2
BASELINE_PUBLISHED = None  # BASELINE.json "published": {} -- no reference number for this config


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


def make_pool(n_local: int, batch: int, image: int, classes: int, pool: int, seed: int):
    """Pinned host pool of synthetic uint8 NHWC image batches + labels, per local worker."""
    g = torch.Generator().manual_seed(seed)
    xs, ys = [], []
    for _ in range(n_local):
        xs.append([torch.randint(0, 256, (batch, image, image, 3), dtype=torch.uint8, generator=g).pin_memory()
                   for _ in range(pool)])
        ys.append([torch.randint(0, classes, (batch,), dtype=torch.int64, generator=g).pin_memory()
                   for _ in range(pool)])
    return xs, ys


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


def barrier_sync(device):
    if dist.is_initialized():
        dist.barrier()
    torch.cuda.synchronize(device)


# ------------------------------------------------------------------------------ ours
def run_ours(args, rank, world, device):
    sys.path.insert(0, REPO)
    from byzpy_b200.aggregators.coordinate_wise import CoordinateWiseMedian
    from byzpy_b200.attacks import SignFlipAttack
    from byzpy_b200.engine.node.device import DeviceByzantineNode, DeviceHonestNode
    from byzpy_b200.engine.parameter_server.ps import ParameterServer
    from byzpy_b200.models import build_model
    from byzpy_b200.parallel.device_ps import RowLayout

    L = N_WORKERS // world
    n_honest = N_WORKERS - N_BYZ
    layout = RowLayout.block(n_honest, N_BYZ, world)
    gids = layout.local_ids(rank)
    xs, ys = make_pool(L, args.batch, args.image, args.classes, args.pool, seed=1234 + rank)
    torch.manual_seed(0)  # identical init on every replica, like a PS that broadcasts the model
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
        torch.manual_seed(0)
        model = build_model(args.model, num_classes=args.classes)
        kw = dict(lr=args.lr, momentum=0.9, device=str(device), preprocess=preprocess_fused, data=source(slot))
        if g < n_honest:
            honest.append(DeviceHonestNode(model, name=f"honest{g}", **kw))
        else:
            byz.append(DeviceByzantineNode(SignFlipAttack(), model=model, name=f"byz{g}", **kw))
    ps = ParameterServer(honest, byz, CoordinateWiseMedian(), update_byzantines=True,
                         layout=layout, amp_dtype=torch.bfloat16, use_cuda_graph=not args.no_graph,
                         worker_streams=args.worker_streams, fused=True,
                         direct_grads=not args.no_direct_grads, overlap_wgrad=not args.no_overlap_wgrad,
                         branch_streams=not args.no_branch_streams, buckets=args.buckets,
                         multicast=None if args.multicast < 0 else bool(args.multicast),
                         device_options=dict(trace=args.trace, overlap_grid=args.overlap_grid))
    rnd = ps.device_round

    def batches(i):
        k = i % args.pool
        return [(xs[s][k], ys[s][k]) for s in range(L)]

    # warm-up (includes graph capture)
    for i in range(max(args.warmup, 3)):
        ps.step(batches(i))
    rnd.read_losses()
    rnd.check_status()
    barrier_sync(device)

    # ---- device-timed region: exactly K rounds, CUDA events on the launching stream ----
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(device.index) as clk:
        barrier_sync(device)
        e0.record()
        for i in range(args.steps):
            if rnd.use_cuda_graph:
                rnd._graph.replay()
            else:
                rnd._body()
        e1.record()
        torch.cuda.synchronize(device)
        barrier_sync(device)
    ms = e0.elapsed_time(e1)
    ms = max_over_ranks(ms, device)
    rnd.check_status()
    if args.trace:
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
    h2d = sum(x[0].numel() * x[0].element_size() + y[0].numel() * y[0].element_size()
              for x, y in zip(xs, ys))
    d2h = rnd.losses.numel() * 4
    if dist.is_initialized():
        tot = torch.tensor([h2d, d2h], dtype=torch.float64, device=device)
        dist.all_reduce(tot)
        h2d, d2h = int(tot[0].item()), int(tot[1].item())
    result = dict(ms=ms, e2e_s=e2e_s, h2d=h2d, d2h=d2h, clocks=clk.summary(),
                  launches=(rnd.launches_per_step + rnd.model_launches_per_step) * args.steps,
                  loss=float(losses.mean().item()), d=rnd.d,
                  round=dict(buckets=rnd.n_buckets if rnd._use_buckets else 1, multicast=bool(rnd._agg_mc),
                             heap=rnd.sym.kind, bounds=list(rnd._bounds)))
    asyncio.run(ps.shutdown())
    return result


# ------------------------------------------------------------------------- reference
def run_reference(args, rank, world, device):
    ref_root = os.path.join(REPO, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref_root, "byzpy")):
        return {"unavailable": "baseline/_ref/byzpy not installed"}
    sys.path.insert(0, ref_root)
    try:
        import torchvision
        from byzpy.aggregators.coordinate_wise import CoordinateWiseMedian
        from byzpy.attacks import SignFlipAttack
        from byzpy.configs.actor import set_actor
        from byzpy.engine.graph.pool import ActorPoolConfig
        from byzpy.engine.node.actors import ByzantineNodeActor, HonestNodeActor
        from byzpy.engine.node.distributed import DistributedByzantineNode, DistributedHonestNode
    except Exception as exc:  # pragma: no cover
        return {"unavailable": f"reference import failed: {exc!r}"}

    L = N_WORKERS // world
    n_honest = N_WORKERS - N_BYZ
    per = N_WORKERS // world
    gids = [g for g in range(N_WORKERS) if g // per == rank]
    xs, ys = make_pool(L, args.batch, args.image, args.classes, args.pool, seed=1234 + rank)
    classes, lr = args.classes, args.lr

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
            self.model = torchvision.models.resnet18(num_classes=classes).to(device).to(
                memory_format=torch.channels_last)
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

    class RefByz(_Common, DistributedByzantineNode):
        def __init__(self, slot):
            DistributedByzantineNode.__init__(
                self, actor_pool=[ActorPoolConfig(backend="gpu", count=1, name="worker")],
                attack=SignFlipAttack(), name=f"byz{slot}")
            self._setup(slot)

        def prepare_attack_inputs(self, *, x=None, y=None, honest_grads=None, base_grad=None, model=None):
            # SignFlip needs the node's own gradient: compute it on this node's batch
            xb, yb = self.next_batch()
            return {"base_grad": self._grad(xb, yb)}

    async def build():
        hon, byz = [], []
        for slot, g in enumerate(gids):
            if g < n_honest:
                hon.append(await HonestNodeActor.spawn(RefHonest, backend=set_actor("thread"),
                                                       kwargs=dict(slot=slot)))
            else:
                byz.append(await ByzantineNodeActor.spawn(RefByz, backend=set_actor("thread"),
                                                          kwargs=dict(slot=slot)))
        return hon, byz

    agg = CoordinateWiseMedian()

    async def one_round(hon, byz):
        grads = list(await asyncio.gather(*[h.honest_gradient_for_next_batch() for h in hon]))
        grads += list(await asyncio.gather(*[b.byzantine_gradient_for_next_batch(tuple(grads)) for b in byz]))
        if world > 1:
            local = torch.stack(grads)
            full = torch.empty((world,) + tuple(local.shape), dtype=local.dtype, device=device)
            dist.all_gather_into_tensor(full.view(-1), local.view(-1))
            grads = list(full.view(N_WORKERS, -1).unbind(0))
        g = agg.aggregate(grads)
        await asyncio.gather(*[n.apply_server_gradient(g) for n in hon + byz])
        return g

    async def losses(hon, byz):
        vals = await asyncio.gather(*[n.loss_value() for n in hon + byz])
        return torch.stack([v.float() for v in vals]).cpu()

    async def main():
        hon, byz = await build()
        for _ in range(max(args.warmup, 3)):
            await one_round(hon, byz)
        barrier_sync(device)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ClockSampler(device.index) as clk:
            barrier_sync(device)
            e0.record()
            for _ in range(args.steps):
                await one_round(hon, byz)
            e1.record()
            torch.cuda.synchronize(device)
            barrier_sync(device)
        ms = max_over_ranks(e0.elapsed_time(e1), device)
        barrier_sync(device)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            await one_round(hon, byz)
            lv = await losses(hon, byz)
        torch.cuda.synchronize(device)
        e2e_s = max_over_ranks(time.perf_counter() - t0, device)
        d = sum(p.numel() for p in torchvision.models.resnet18(num_classes=classes).parameters())
        h2d = sum(x[0].numel() * x[0].element_size() + y[0].numel() * y[0].element_size()
                  for x, y in zip(xs, ys)) * world
        res = dict(ms=ms, e2e_s=e2e_s, h2d=h2d, d2h=4 * N_WORKERS, clocks=clk.summary(), launches=0,
                   loss=float(lv.mean().item()), d=d)
        for n in hon + byz:
            await n._ref._backend.close()
        return res

    return asyncio.run(main())


# ------------------------------------------------------------------------------ main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--model", default="resnet18")
    ap.add_argument("--batch", type=int, default=32, help="per-worker batch")
    ap.add_argument("--image", type=int, default=224)
    ap.add_argument("--classes", type=int, default=1000)
    ap.add_argument("--pool", type=int, default=4, help="distinct pinned host batches per worker")
    ap.add_argument("--lr", type=float, default=0.05)
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

    res = run_ours(args, rank, world, device) if args.impl == "ours" else run_reference(args, rank, world, device)
    if rank == 0:
        if "unavailable" in res:
            print(json.dumps({"impl": "reference", "unavailable": res["unavailable"]}))
        else:
            sps = args.steps / (res["ms"] / 1e3)
            e2e = args.steps / res["e2e_s"]
            out = {
                "metric": "PS steps/sec (device-timed, max over ranks) ResNet-18 + CoordinateWiseMedian, 2 Byzantine",
                "value": round(sps, 3), "unit": "steps/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": round(res["ms"] / args.steps, 4),
                "higher_is_better": True, "scaling": "strong",
                "vs_baseline": (sps / BASELINE_PUBLISHED) if BASELINE_PUBLISHED else None,
                "dtype": "bf16", "data": "synthetic",
                "impl": args.impl,
                "config": {
                    "model": args.model, "workers": N_WORKERS, "byzantine": N_BYZ, "attack": "SignFlip",
                    "aggregator": "CoordinateWiseMedian", "per_worker_batch": args.batch,
                    "global_batch": args.batch * N_WORKERS, "image": [3, args.image, args.image],
                    "classes": args.classes, "grad_dim": res["d"],
                    "parallelism": f"byzantine-dp{N_WORKERS} over {world} gpu(s), {N_WORKERS // world} replicas/gpu",
                    "optimizer": "SGD(momentum=0.9), fp32 master weights, update_byzantines=True",
                    "l2": "per-round working set (8/N x (46.8 MB grads + 93.6 MB params/momentum) + "
                          "GBs of activations) exceeds the 126 MB L2; no explicit flush",
                    "timing": "CUDA events around exactly K rounds, barrier+synchronize both sides, max over ranks",
                },
                "clocks": res["clocks"],
                "e2e": {"value": round(e2e, 3), "unit": "steps/s", "h2d_bytes_per_step": res["h2d"],
                        "d2h_bytes_per_step": res["d2h"]},
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
