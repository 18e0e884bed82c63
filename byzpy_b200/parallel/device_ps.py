"""B200-native parameter-server round: the CUDA runtime behind
:class:`byzpy_b200.engine.parameter_server.ps.ParameterServer` when its nodes
are device-resident (:class:`DeviceWorker`).

One process per GPU.  A rank hosts ``L = n / world`` worker replicas; every
replica's parameters, momentum and gradient live in flat fp32 arenas, and the
gradient rows sit in CUDA-IPC symmetric memory so that any rank's kernel can
read them over NVLink.  A training round is

    for each local worker:  H2D batch -> fwd/bwd (PyTorch / cuDNN, bf16 autocast)
    ONE fused kernel:       gather(P2P ld) + robust aggregate + broadcast(P2P st)
                            + SGD(momentum) on all local replicas

captured end-to-end in a CUDA graph (epoch counter lives on the device), so the
steady state is a single ``cudaGraphLaunch`` per round per rank and there is no
NCCL call on the path.  Semantics mirror the reference round (reference
engine/parameter_server/ps.py:103-144): honest gradients, then Byzantine
vectors appended after them, aggregate, apply to every
honest node (and Byzantine ones when ``update_byzantines``).
"""
from __future__ import annotations

import contextlib
import os
from dataclasses import dataclass, field
from typing import Callable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist
import torch.nn as nn

from .. import ops
from .arena import ParamArena, padded_size
from .symmetric import SymmetricBuffer, _dist_on, tensor_from_ptr


PAD_SYNC = 48       # flag words of the recovery / resynchronisation barriers (csrc/fused_ps.h uses 0-39)


# --------------------------------------------------------------------------- plans
@dataclass
class CwPlan:
    """Coordinate-wise aggregation plan (median / trimmed mean / meamed / mean)."""

    mode: int
    f: int = 0


@dataclass
class GramPlan:
    """Gram-family plan: pass 1 Gram, n-space solve -> weights, pass 2 weighted sum.

    ``solver(G)`` maps the fp64 Gram of (real rows + aux rows) ON THE DEVICE to the fp32 weight
    vector over the same rows.  ``aux`` lists auxiliary rows the solver expects after the real
    ones: ``"median"`` = coordinate-wise median of the real rows (the Weiszfeld / centered-clipping
    start point), ``("const", make)`` = a constant vector ``make(d, like)`` resident on the device
    (CAF's fixed power-iteration start direction).  ``capturable`` is False when the solver synchronises with the host (MDA, SMEA,
    CAF searches), which rules out CUDA-graph capture of the round.  ``refresh`` (optional) is
    called on the host before every round (e.g. to draw a new bucketing permutation).
    """

    solver: Callable[[torch.Tensor], torch.Tensor]
    name: str = "gram"
    aux: Tuple[str, ...] = ()
    capturable: bool = True
    refresh: Optional[Callable[[], None]] = None


@dataclass
class MapCwPlan:
    """A linear pre-aggregator followed by a coordinate-wise aggregator (Bucketing -> median, NNM ->
    trimmed mean, ...).  The composition is not linear in n-space (the selection runs per coordinate
    on the MIXED rows), so unlike :class:`GramPlan` the ``m`` pre-aggregated rows do exist -- but only
    as each rank's coordinate shard, in local HBM: ``Y[:, shard] = W_p X[:, shard]`` is produced by
    the weighted-sum kernels straight from the (peer) gradient rows, then the fused gather / select /
    deliver / SGD kernel runs over those ``m`` local rows.

    ``weights(G)`` returns the ``(m, n)`` map ON THE DEVICE; ``G`` is the fp64 Gram of the real rows
    when ``needs_gram`` (Clipping / ARC / NNM: computed like in the Gram round, partials all-reduced
    through the switch) and ``None`` otherwise (Bucketing: a constant matrix, ``refresh`` redraws it
    on the host before the round)."""

    cw: CwPlan
    m: int
    weights: Callable[[Optional[torch.Tensor]], torch.Tensor]
    needs_gram: bool = False
    name: str = "map+cw"
    capturable: bool = True
    refresh: Optional[Callable[[], None]] = None


@dataclass
class RowFold:
    """How a Byzantine row is produced without materialising it."""

    kind: str  # "scale" (own gradient * scale) | "virtual" (a*mean+b*std of honest rows) | "alias"
    scale: float = 1.0
    a: float = 0.0
    b: float = 0.0
    index: int = 0


# ------------------------------------------------------------------------- workers
class DeviceWorker:
    """One model replica resident on the local GPU.

    ``role``: ``"honest"`` or ``"byzantine"``; a Byzantine worker with a
    ``scale`` fold (SignFlip) still computes its own gradient (its ``base_grad``)
    and the sign/scale is folded into the aggregation kernel's row load.
    """

    def __init__(self, model: nn.Module, loss_fn: Callable, *, role: str = "honest",
                 fold: Optional[RowFold] = None, name: Optional[str] = None,
                 preprocess: Optional[Callable[[torch.Tensor], torch.Tensor]] = None,
                 data: Optional[Callable[[], Tuple[torch.Tensor, torch.Tensor]]] = None):
        self.model = model
        self.loss_fn = loss_fn
        self.role = role
        self.fold = fold
        self.name = name or role
        self.preprocess = preprocess
        self.data = data
        self.arena: Optional[ParamArena] = None
        self.mom: Optional[torch.Tensor] = None
        # two sets of static input buffers: the round engine reads set ``buf`` while the next
        # step's batch is copied H2D into the other one (DeviceRound.step, double-buffered graphs)
        self._static: List[Optional[Tuple[torch.Tensor, torch.Tensor]]] = [None, None]
        self.buf = 0
        self.loss_slot: Optional[torch.Tensor] = None
        self.sink = None            # ops.fused_layers.GradSink once direct gradients are on
        self._mark_handles: list = []
        self._mark_sink: Optional[Callable] = None   # DeviceRound callback (worker, bucket, events)

    # filled in by the engine
    def bind(self, flat_params: torch.Tensor, flat_grads: torch.Tensor, mom: Optional[torch.Tensor],
             loss_slot: torch.Tensor) -> None:
        self.arena = ParamArena(self.model, flat_params=flat_params, flat_grads=flat_grads)
        self.mom = mom
        self.loss_slot = loss_slot

    def enable_direct_grads(self, side_stream: Optional["torch.cuda.Stream"] = None,
                            branch_stream: Optional["torch.cuda.Stream"] = None) -> None:
        """Let the replica's own layer types (models/resnet.py) write their parameter gradients
        straight into the arena row, weight gradients on ``side_stream`` (ops/fused_layers.py);
        projection shortcuts of residual blocks run on ``branch_stream``."""
        from ..ops.fused_layers import enable_direct_grads

        self.sink = enable_direct_grads(self.model, side_stream=side_stream, branch_stream=branch_stream)

    # ---- gradient buckets (DeviceRound overlaps the aggregation of a bucket with the rest of backward)
    def bucket_candidates(self):
        from ..ops.fused_layers import bucket_candidates

        return bucket_candidates(self.model, self.arena.offsets)

    def install_marks(self, marks, sink_cb: Callable) -> None:
        """``marks`` = [(module, bucket index)]; ``sink_cb(worker, bucket, events)`` is invoked from
        inside backward when every parameter gradient at or beyond the module has been issued."""
        from ..ops.fused_layers import install_bucket_marks

        self.remove_marks()
        self._mark_sink = sink_cb
        self._mark_handles = install_bucket_marks(marks, self._on_mark)

    def remove_marks(self) -> None:
        for h in self._mark_handles:
            h.remove()
        self._mark_handles = []
        self._mark_sink = None

    def _on_mark(self, index: int) -> None:
        if self._mark_sink is None:
            return
        dev = self.arena.flat_params.device
        events = []
        sink = self.sink
        if sink is not None:
            sink.flush_pending(force=True)      # weight gradients produced so far -> arena row (side stream)
            if sink._forked and sink.side_stream is not None:
                ev = torch.cuda.Event()
                ev.record(sink.side_stream)
                events.append(ev)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(dev))
        events.append(ev)
        self._mark_sink(self, index, events)

    @property
    def static_x(self) -> Optional[torch.Tensor]:
        cur = self._static[self.buf]
        return None if cur is None else cur[0]

    @property
    def static_y(self) -> Optional[torch.Tensor]:
        cur = self._static[self.buf]
        return None if cur is None else cur[1]

    def stage_batch(self, x: torch.Tensor, y: torch.Tensor, buf: Optional[int] = None) -> bool:
        """Copy a step's inputs (typically pinned host tensors) into static device buffer set ``buf``
        (default: the active one) on the current stream.  Returns True when the buffers had to be
        (re)allocated -- a captured graph that baked the old addresses is then stale."""
        b = self.buf if buf is None else buf
        dev = self.arena.flat_params.device
        cur = self._static[b]
        fresh = (cur is None or cur[0].shape != x.shape or cur[0].dtype != x.dtype
                 or cur[1].shape != y.shape or cur[1].dtype != y.dtype)
        if fresh:
            cur = (torch.empty(x.shape, dtype=x.dtype, device=dev), torch.empty(y.shape, dtype=y.dtype, device=dev))
            self._static[b] = cur
        cur[0].copy_(x, non_blocking=True)
        cur[1].copy_(y, non_blocking=True)
        return fresh

    def forward_backward(self, amp_dtype: Optional[torch.dtype]) -> None:
        self.arena.flat_grads.zero_()
        x = self.static_x
        if self.preprocess is not None:
            x = self.preprocess(x)
        ctx = (torch.autocast("cuda", dtype=amp_dtype) if amp_dtype is not None
               else contextlib.nullcontext())
        if self.sink is not None and amp_dtype == torch.bfloat16:
            self.sink.refresh_shadows()     # all conv weights -> bf16 channels-last, one launch
        with ctx:
            out = self.model(x)
            loss = self.loss_fn(out, self.static_y)
        loss.backward()
        if self.sink is not None:
            self.sink.join()        # weight gradients produced on the side stream are now ordered
        self.loss_slot.copy_(loss.detach().float())

    def state_dict_cpu(self):
        """Checkpoint mapping, identical to the reference nodes' ``dump_state_dict``
        (reference examples/ps/nodes.py:127-128)."""
        return {k: v.detach().cpu() for k, v in self.model.state_dict().items()}


# -------------------------------------------------------------------------- layout
@dataclass
class RowLayout:
    """Global row order of a round: honest workers first, then Byzantine ones."""

    n_honest: int
    n_byz_workers: int          # Byzantine rows backed by a worker replica (scale folds)
    n_virtual: int = 0          # synthesised Byzantine rows (Little / Empire)
    world: int = 1
    rank_of: List[int] = field(default_factory=list)
    slot_of: List[int] = field(default_factory=list)

    @property
    def n_workers(self) -> int:
        return self.n_honest + self.n_byz_workers

    @classmethod
    def block(cls, n_honest: int, n_byz_workers: int, world: int, n_virtual: int = 0) -> "RowLayout":
        n = n_honest + n_byz_workers
        if n % world != 0:
            raise ValueError(f"{n} workers do not divide evenly over {world} ranks")
        per = n // world
        return cls(n_honest, n_byz_workers, n_virtual, world,
                   [g // per for g in range(n)], [g % per for g in range(n)])

    @classmethod
    def spread(cls, n_honest: int, n_byz_workers: int, world: int, n_virtual: int = 0) -> "RowLayout":
        """Like :meth:`block` for worker counts that do not divide the world size: contiguous blocks
        whose sizes differ by at most one; with fewer workers than ranks the trailing ranks host no
        replica and only take part in the aggregation of their coordinate shard."""
        n = n_honest + n_byz_workers
        base, extra = divmod(n, world)
        rank_of, slot_of = [], []
        for r in range(world):
            for s in range(base + (1 if r < extra else 0)):
                rank_of.append(r)
                slot_of.append(s)
        return cls(n_honest, n_byz_workers, n_virtual, world, rank_of, slot_of)

    def local_ids(self, rank: int) -> List[int]:
        return [g for g in range(self.n_workers) if self.rank_of[g] == rank]

    def max_local(self) -> int:
        return max((self.slot_of[g] + 1 for g in range(self.n_workers)), default=0)


# ------------------------------------------------------------------------- buckets
def pick_bucket_offsets(block_offsets: Sequence[int], d: int, cuts: Sequence[float],
                        buckets: Optional[int] = None) -> List[int]:
    """Block offsets (descending) at which the gradient arena is cut into buckets.

    ``block_offsets``: first flat offset of every block of the model, ascending (execution order).
    Walking from the tail of the arena -- the layers backward finishes first -- a cut is placed at
    the first block input after which at least ``cuts[i]`` of the gradient lies behind it."""
    offs = list(block_offsets)
    want = list(cuts)
    if buckets is not None:
        want = want[: max(0, buckets - 1)]
    picks: List[int] = []
    ci = len(offs) - 1
    for frac in want:
        while ci >= 1 and (d - offs[ci]) < frac * d:
            ci -= 1
        if ci < 1:
            break
        picks.append(offs[ci])
        ci -= 1
    return picks


def bucket_bounds(cut_offsets: Sequence[int], d_pad: int, min_bucket: int, align: int = 1024) -> List[int]:
    """``[d_pad, b1, b2, ..., 0]``: bucket k covers ``[bounds[k+1], bounds[k])``.  Cuts are rounded UP
    to ``align`` elements (the few leading elements of the block at a boundary then belong to the
    next, later bucket -- always safe); cuts that would leave a bucket under ``min_bucket`` elements
    are dropped."""
    bounds = [d_pad]
    for off in cut_offsets:                                    # descending
        b = padded_size(off, align)
        if b <= 0 or bounds[-1] - b < min_bucket or b < min_bucket:
            continue
        bounds.append(b)
    bounds.append(0)
    return bounds


# -------------------------------------------------------------------------- engine
class DeviceRound:
    """Owns the symmetric arenas and launches the fused round on this rank."""

    def __init__(self, workers: Sequence[DeviceWorker], layout: RowLayout, plan, *,
                 lr: float, momentum: float = 0.0, weight_decay: float = 0.0,
                 update_byzantines: bool = False, device: Optional[torch.device] = None,
                 group=None, amp_dtype: Optional[torch.dtype] = torch.bfloat16,
                 use_cuda_graph: bool = True, worker_streams: int = 1,
                 virtual_fold: Optional[RowFold] = None, direct_grads: bool = True,
                 overlap_wgrad: bool = True, branch_streams: bool = True,
                 buckets: Optional[int] = None, bucket_cuts: Sequence[float] = (0.40, 0.75, 0.93),
                 min_bucket: int = 1 << 16, overlap_grid: Optional[int] = None,
                 spin_seconds: float = 0.0, multicast: Optional[bool] = None, trace: bool = False):
        if not torch.cuda.is_available():
            raise RuntimeError("DeviceRound needs a CUDA device (B200)")
        # replica shapes are static for the lifetime of a round engine (they are baked into its CUDA
        # graphs), so let cuDNN pick its fastest kernels once
        torch.backends.cudnn.benchmark = True
        self.ext = ops.require_ext()
        self.workers = list(workers)
        self.layout = layout
        self.plan = plan
        self.lr, self.momentum, self.weight_decay = float(lr), float(momentum), float(weight_decay)
        self.update_byzantines = update_byzantines
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self.group = group
        self.amp_dtype = amp_dtype
        self.use_cuda_graph = use_cuda_graph
        self.virtual_fold = virtual_fold
        self.rank = dist.get_rank(group) if _dist_on() else 0
        self.world = dist.get_world_size(group) if _dist_on() else 1
        if layout.world != self.world:
            raise ValueError("layout.world does not match the process group")
        self.local_ids = layout.local_ids(self.rank)
        if len(self.local_ids) != len(self.workers):
            raise ValueError(f"rank {self.rank} hosts {len(self.local_ids)} rows but got "
                             f"{len(self.workers)} workers")
        L = len(self.workers)
        self.L = L
        d = sum(p.numel() for p in self.workers[0].model.parameters()) if L else 0
        if self.world > 1:      # ranks without a replica (RowLayout.spread) learn d from the others
            t = torch.tensor([d], dtype=torch.int64, device=self.device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
            d = int(t.item())
        if d <= 0:
            raise ValueError("no rank hosts a model replica")
        self.d = d
        L_sym = max(L, layout.max_local())      # uniform symmetric layout on every rank
        # shards must be multiples of 4 elements on every rank
        self.d_pad = padded_size(d, 1024)
        self.sm = ops.sm_count(self.device)

        # --- symmetric region: [grads L*d_pad | agg d_pad | pad words | counter,status,epoch]
        f4 = 4
        self._off_grads = 0
        self._off_agg = L_sym * self.d_pad * f4
        self._off_pad = self._off_agg + self.d_pad * f4
        self._off_ctl = self._off_pad + 256
        self._off_gslots = self._off_ctl + 256
        self.nt_max = 144                       # rows + auxiliary rows of a Gram plan
        gram_bytes = (max(2, self.world) * self.nt_max * self.nt_max * 8
                      if isinstance(plan, (GramPlan, MapCwPlan)) else 0)
        nbytes = self._off_gslots + gram_bytes
        self.sym = SymmetricBuffer(nbytes, self.device, group, multicast=multicast)
        self.grads = self.sym.view(torch.float32, L_sym * self.d_pad, self._off_grads).view(L_sym, self.d_pad)
        self.agg = self.sym.view(torch.float32, self.d_pad, self._off_agg)
        self.pad = self.sym.view(torch.int32, 64, self._off_pad)
        self.ctl = self.sym.view(torch.int32, 64, self._off_ctl)  # [0]=counter [1]=status [2]=epoch
        # local (non-symmetric) replica state
        self.params = torch.zeros((L, self.d_pad), dtype=torch.float32, device=self.device)
        self.moms = (torch.zeros((L, self.d_pad), dtype=torch.float32, device=self.device)
                     if self.momentum != 0.0 else None)
        self.losses = torch.zeros(L, dtype=torch.float32, device=self.device)
        self.losses_host = torch.zeros(L, dtype=torch.float32).pin_memory()
        self._status_host = torch.zeros(1, dtype=torch.int32).pin_memory()
        self._fault: Optional[str] = None
        self._recoveries = 0
        self._resync_seq = 0
        self._resync_ctr = torch.zeros(1, dtype=torch.int32, device=self.device)   # epoch word of the resync barriers (stays 0)
        self._round_launches = 0
        self._use_buckets = False
        self._in_round = False
        for i, w in enumerate(self.workers):
            w.model.to(self.device)
            w.bind(self.params[i], self.grads[i], None if self.moms is None else self.moms[i],
                   self.losses[i])
        self._rows, self._scales = self._row_table()
        self._upd_params, self._upd_moms = self._update_table()
        sh = self.d_pad // self.world
        sh -= sh % 4
        self.shard_off = self.rank * sh
        self.shard_len = sh if self.rank < self.world - 1 else self.d_pad - sh * (self.world - 1)
        # one captured graph per static input buffer set (they share one memory pool)
        self._graphs: List[Optional[torch.cuda.CUDAGraph]] = [None, None]
        self._buf = 0
        self._copy_stream = torch.cuda.Stream(self.device)
        self._h2d_done = [torch.cuda.Event(), torch.cuda.Event()]
        self._consumed = [torch.cuda.Event(), torch.cuda.Event()]
        self._prefetched = [False, False]
        worker_streams = max(1, min(int(worker_streams), max(L, 1)))   # never more streams than local replicas
        # model work runs on high-priority streams (the aggregation stream keeps the default, lowest
        # priority): when both have CTAs pending, the block scheduler serves backward first
        hp = dict(priority=-1)
        self._capture_stream = torch.cuda.Stream(self.device, **hp)
        self._side_streams = [torch.cuda.Stream(self.device, **hp) for _ in range(worker_streams - 1)]
        # one weight-gradient stream per worker stream: backward's dgrad chain stays on the worker
        # stream, the wgrad GEMMs of the same replica overlap with it
        self._wgrad_streams = ([torch.cuda.Stream(self.device, **hp) for _ in range(1 + len(self._side_streams))]
                               if (direct_grads and overlap_wgrad) else [])
        self._branch_streams = ([torch.cuda.Stream(self.device, **hp) for _ in range(1 + len(self._side_streams))]
                                if (direct_grads and branch_streams) else [])
        if direct_grads:
            for i, w in enumerate(self.workers):
                w.enable_direct_grads(
                    self._wgrad_streams[i % len(self._wgrad_streams)] if self._wgrad_streams else None,
                    self._branch_streams[i % len(self._branch_streams)] if self._branch_streams else None)
        self._gram_ws = None
        self.launches_per_step = 0
        self.model_launches_per_step = 0
        self.spin_seconds = float(spin_seconds)
        # device-side round timeline (%globaltimer stamps written by tiny kernels and by the fused
        # kernels themselves): [0] round start, [1] backward done, [2] round done, [16+k] bucket k
        # produced locally, [32+8k ..] the 6 stamps of bucket k's launch (csrc/fused_ps.h)
        self._trace = torch.zeros(128, dtype=torch.int64, device=self.device) if trace else None
        self.live_mask = 0                      # 0 = every rank takes part (see recover())
        self._agg_stream = torch.cuda.Stream(self.device)
        self._agg_mc = self.sym.mc_ptr(self._off_agg) if (multicast is not False) else 0
        if multicast and not self._agg_mc:
            raise RuntimeError("multicast=True but the symmetric heap has no NVLS multicast mapping")
        # Overlapped bucket launches share the GPU with backward.  A CTA of the fused kernel parked on
        # an SM (it waits for the peers' flags, then streams over NVLink) keeps cuDNN kernels that need
        # a whole SM's registers / shared memory off that SM, so an overlapped launch only takes a
        # QUARTER of the SMs: one CTA per SM on all of them delayed backward by 250 us per round at
        # 8 GPUs, more than the overlap saved (profiles/round_timeline.md).  The head bucket, launched
        # after backward, uses the full grid.
        env_grid = os.environ.get("BYZPY_OVERLAP_GRID")
        if overlap_grid is None and env_grid:
            overlap_grid = int(env_grid)
        self.overlap_grid = int(overlap_grid) if overlap_grid is not None else max(8, self.sm // 4)
        self._bounds: List[int] = [0, self.d_pad]     # bucket k = [bounds[k+1], bounds[k]) walking from the tail
        self._buckets_validated = True
        if isinstance(plan, CwPlan) and (buckets is None or buckets > 1):
            self._plan_buckets(buckets, bucket_cuts, min_bucket)
        if isinstance(plan, GramPlan):
            self._setup_gram_plan()
            if not plan.capturable:
                self.use_cuda_graph = False
        if isinstance(plan, MapCwPlan):
            self._setup_mapcw_plan()
            if not plan.capturable:
                self.use_cuda_graph = False
        if self.world > 1:
            dist.barrier(group=group)
        torch.cuda.synchronize(self.device)

    # ----------------------------------------------------------------- buckets
    @property
    def n_buckets(self) -> int:
        return len(self._bounds) - 1

    def bucket_range(self, k: int) -> Tuple[int, int]:
        """(offset, length) of bucket k; bucket 0 is the TAIL of the arena (the layers whose
        gradients backward produces first), the last bucket starts at offset 0."""
        hi, lo = self._bounds[k], self._bounds[k + 1]
        return lo, hi - lo

    def _plan_buckets(self, buckets: Optional[int], cuts: Sequence[float], min_bucket: int) -> None:
        """Choose bucket boundaries at block inputs of the replica model and install the backward
        marks.  ``cuts`` are cumulative fractions of the gradient (from the tail) after which a
        boundary is placed at the next block input; ``buckets`` caps their number."""
        local = None
        if self.L:
            offs = [fo for fo, _ in self.workers[0].bucket_candidates()]     # ascending block offsets
            local = pick_bucket_offsets(offs, self.d, cuts, buckets)
        if self.world > 1:
            allb: List[Optional[list]] = [None] * self.world
            dist.all_gather_object(allb, local, group=self.group)
            known = [b for b in allb if b is not None]
            if any(b != known[0] for b in known):
                raise RuntimeError("ranks disagree on gradient bucket boundaries (different replica models?)")
            chosen = known[0]
        else:
            chosen = local or []
        bounds = bucket_bounds(chosen, self.d_pad, min_bucket)
        self._bounds = bounds
        if self.n_buckets > 1:
            self._buckets_validated = False
            if self.L:
                # a mark sits at the input of the FIRST block of a bucket's range: when its backward
                # fires, everything at or beyond the (rounded-up) boundary has been produced
                by_bound: dict = {}
                for fo in chosen:                   # descending: the later block wins a shared boundary
                    by_bound.setdefault(padded_size(fo, 1024), fo)
                for w in self.workers:
                    cw = w.bucket_candidates()
                    marks = []
                    for k in range(self.n_buckets - 1):
                        fo = by_bound[self._bounds[k + 1]]
                        mod = next(m for off, m in cw if off == fo)
                        marks.append((mod, k))
                    w.install_marks(marks, self._on_mark)
        self._bk_events: List[list] = [[] for _ in range(self.n_buckets)]
        self._bk_count = [0] * self.n_buckets
        self._next_bucket = 0
        self._in_round = False
        self._use_buckets = self.n_buckets > 1

    def _stamp(self, slot: int) -> None:
        if self._trace is not None:
            self.ext.stamp(self._trace.data_ptr() + 8 * slot, torch.cuda.current_stream(self.device).cuda_stream)

    def timeline(self) -> dict:
        """The last round's device timeline in microseconds relative to the round's start
        (``trace=True``): when backward finished, when each bucket was produced locally, and the
        phases of every bucket launch."""
        if self._trace is None:
            raise RuntimeError("construct the round with trace=True")
        t = self._trace.cpu().tolist()
        t0 = t[0]
        us = lambda v: None if v == 0 else round((v - t0) / 1e3, 1)      # noqa: E731
        out = {"backward_done": us(t[1]), "round_done": us(t[2]), "buckets": []}
        nb = self.n_buckets if self._use_buckets else 1
        for k in range(nb):
            b = t[32 + 8 * k: 32 + 8 * k + 6]
            off, ln = self.bucket_range(k) if nb > 1 else (0, self.d_pad)
            out["buckets"].append({"elements": ln, "produced": us(t[16 + k]) if nb > 1 else us(t[1]),
                                   "start": us(b[0]), "ready_wait_done": us(b[1]), "block0_phase1_done": us(b[2]),
                                   "all_ctas_phase1_done": us(b[3]), "delivery_wait_done": us(b[4]),
                                   "sgd_done": us(b[5])})
        return out

    def _on_mark(self, worker: DeviceWorker, k: int, events: list) -> None:
        if not self._in_round or not self._use_buckets:
            return
        self._bk_events[k].extend(events)
        self._bk_count[k] += 1
        if self._trace is not None and self._bk_count[k] >= self.L:
            self._stamp(16 + k)
        self._launch_ready_buckets()

    def _launch_ready_buckets(self) -> None:
        """Enqueue, in bucket order, every bucket all local replicas have produced."""
        agg = self._agg_stream
        while self._next_bucket < self.n_buckets - 1 and self._bk_count[self._next_bucket] >= self.L:
            k = self._next_bucket
            for ev in self._bk_events[k]:
                agg.wait_event(ev)
            with torch.cuda.stream(agg):
                self._launch_cw_bucket(k, self.overlap_grid)
            self._next_bucket += 1

    # ------------------------------------------------------------------ tables
    def _row_table(self) -> Tuple[List[int], List[float]]:
        lay = self.layout
        rows, scales = [], []
        for g in range(lay.n_workers):
            r, s = lay.rank_of[g], lay.slot_of[g]
            rows.append(self.sym.peer_ptr(r, self._off_grads + s * self.d_pad * 4))
            scales.append(1.0)
        # folds of Byzantine workers hosted anywhere: every rank derives the same scale
        # from the (replicated) fold description of its local workers; remote ones are exchanged.
        local = {g: (w.fold.scale if (w.fold is not None and w.fold.kind == "scale") else 1.0)
                 for g, w in zip(self.local_ids, self.workers)}
        if self.world > 1:
            allf: List[Optional[dict]] = [None] * self.world
            dist.all_gather_object(allf, local, group=self.group)
            for part in allf:
                for g, sc in part.items():
                    scales[g] = sc
        else:
            for g, sc in local.items():
                scales[g] = sc
        return rows, scales

    def _update_table(self) -> Tuple[List[int], List[int]]:
        ps, ms = [], []
        for i, w in enumerate(self.workers):
            if w.role == "honest" or self.update_byzantines:
                ps.append(self.params[i].data_ptr())
                if self.moms is not None:
                    ms.append(self.moms[i].data_ptr())
        return ps, ms

    # ------------------------------------------------------------------- launch
    def _virtual(self) -> Tuple[int, int, float, float]:
        lay = self.layout
        if lay.n_virtual and self.virtual_fold is not None:
            return lay.n_virtual, lay.n_honest, self.virtual_fold.a, self.virtual_fold.b
        return 0, 0, 0.0, 0.0

    def _shard(self, off: int, ln: int) -> Tuple[int, int]:
        """This rank's share of the coordinate range [off, off + ln) (multiples of 4 elements)."""
        live = [r for r in range(self.world) if self.live_mask == 0 or (self.live_mask >> r) & 1]
        idx = live.index(self.rank)
        sh = (ln // len(live)) // 4 * 4
        s_off = off + idx * sh
        s_len = sh if idx < len(live) - 1 else ln - sh * (len(live) - 1)
        return s_off, s_len

    def _launch_cw_bucket(self, k: int, grid_limit: int = 0, *, whole: bool = False) -> None:
        """One fused launch (gather + select + broadcast + SGD) over bucket k -- or over the whole
        arena with ``whole`` -- on the current stream.  The epoch word must already be bumped."""
        plan = self.plan
        stream = torch.cuda.current_stream(self.device).cuda_stream
        ctl = self.ctl.data_ptr()
        nv, nh, va, vb = self._virtual()
        # flag words carry epoch * n_buckets + bucket: the numbering never changes over the life of
        # the engine (a whole-arena launch counts as the round's last bucket), so sequence numbers
        # stay monotone when overlapped and single-launch rounds are mixed
        nb = self.n_buckets
        if whole or nb == 1:
            off, ln, k = 0, self.d_pad, nb - 1
        else:
            off, ln = self.bucket_range(k)
        s_off, s_len = self._shard(off, ln)
        self.ext.fused_ps_cw(
            self._rows, self._scales, plan.mode, plan.f, nv, nh, va, vb, self.d_pad,
            s_off, s_len, self.rank,
            [self.sym.peer_ptr(r, self._off_agg) for r in range(self.world)],
            [self.sym.peer_ptr(r, self._off_pad) for r in range(self.world)],
            0, ctl + 8, ctl + 0, ctl + 4,
            self._upd_params, self._upd_moms, self.lr, self.momentum, self.weight_decay,
            self.sm, stream, grid_limit, off, ln, nb, k, self._agg_mc,
            self.live_mask, self.spin_seconds,
            0 if self._trace is None else self._trace.data_ptr() + 8 * (32 + 8 * (0 if whole else k)),
        )
        self._round_launches += 1

    def launch_aggregate(self) -> None:
        """Enqueue the fused gather+aggregate+broadcast+update on the current stream (one launch
        over the whole arena for the coordinate-wise family: the non-overlapped form of the round)."""
        stream = torch.cuda.current_stream(self.device).cuda_stream
        ctl = self.ctl.data_ptr()
        self.ext.bump_u32(ctl + 8, stream)
        plan = self.plan
        if isinstance(plan, CwPlan):
            self._round_launches = 1
            self._launch_cw_bucket(0, whole=True)
            self.launches_per_step = self._round_launches
        elif isinstance(plan, GramPlan):
            self._launch_gram_round(stream, ctl)
        elif isinstance(plan, MapCwPlan):
            self._launch_mapcw_round(stream, ctl)
        else:
            raise TypeError(f"unsupported plan {plan!r}")

    # ------------------------------------------------------------- Gram-family round
    def _setup_gram_plan(self) -> None:
        """Static buffers of the Gram-family round (addresses must not change under graph replay)."""
        lay = self.layout
        dev = self.device
        self._n_real = lay.n_workers + lay.n_virtual
        self._n_aux = len(self.plan.aux)
        nt = self._n_real + self._n_aux
        if nt > self.nt_max or nt > ops.MAXN:
            raise ValueError("too many rows for the fused Gram round")
        self._nt = nt
        self._virt_buf = (torch.zeros(self.d_pad, dtype=torch.float32, device=dev)
                          if lay.n_virtual else None)
        self._aux_bufs = [torch.zeros(self.d_pad, dtype=torch.float32, device=dev)
                          for _ in range(self._n_aux)]
        for kind, buf in zip(self.plan.aux, self._aux_bufs):
            if isinstance(kind, tuple) and kind[0] == "const":
                buf[: self.d].copy_(kind[1](self.d, buf).to(torch.float32))
            elif kind != "median":
                raise ValueError(f"unknown aux row {kind!r}")
        self._g_local64 = torch.zeros((nt, nt), dtype=torch.float64, device=dev)
        self._g_local32 = torch.zeros((nt, nt), dtype=torch.float32, device=dev)
        self._g_tail64 = torch.zeros((nt, nt), dtype=torch.float64, device=dev)
        self._g_tail32 = torch.zeros((nt, nt), dtype=torch.float32, device=dev)
        self._g_total64 = torch.zeros((nt, nt), dtype=torch.float64, device=dev)
        self._w_dev = torch.zeros(nt, dtype=torch.float32, device=dev)
        need = self.ext.gram_partials_needed(nt, self.sm)
        self._gram_scratch = torch.empty(need, dtype=torch.float32, device=dev)
        self._umma_scratch = torch.empty(self.sm * 8 * 2 * nt * nt, dtype=torch.float32, device=dev)
        # row tables: workers (peer pointers) + virtual rows + aux rows (both local, shard-valid)
        rows = list(self._rows)
        scales = list(self._scales)
        if lay.n_virtual:
            rows += [self._virt_buf.data_ptr()] * lay.n_virtual
            scales += [1.0] * lay.n_virtual
        self._real_rows, self._real_scales = list(rows), list(scales)
        for b in self._aux_bufs:
            rows.append(b.data_ptr())
            scales.append(1.0)
        self._all_rows, self._all_scales = rows, scales

    def _launch_gram_round(self, stream: int, ctl: int) -> None:
        ext, lay, plan = self.ext, self.layout, self.plan
        pads = [self.sym.peer_ptr(r, self._off_pad) for r in range(self.world)]
        # this rank's share of the coordinates among the LIVE ranks (equal to the static shard while every
        # rank takes part; after recover() the survivors cover the dropped rank's range as well)
        off, ln = self._shard(0, self.d_pad)
        nt = self._nt
        launches = 1
        # every rank's gradient rows must be complete before anybody reads them
        ext.flag_barrier(pads, self.rank, ext.PAD_READY, ctl + 8, ctl + 4, stream, 0, 0, self.live_mask,
                         self.spin_seconds)
        launches += 1
        if lay.n_virtual:
            nh = lay.n_honest
            ext.colstat(self._rows[:nh], self._scales[:nh], float(self.virtual_fold.a),
                        float(self.virtual_fold.b), off, ln, self._virt_buf.data_ptr(), self.sm, stream)
            launches += 1
        for kind, buf in zip(plan.aux, self._aux_bufs):
            if kind != "median":
                continue                    # constant rows were filled once at set-up
            ext.cw_select(self._real_rows, self._real_scales, ops.MODE_MEDIAN, 0, 0, 0, 0.0, 0.0, off, ln,
                          buf.data_ptr(), [], [], 0.0, 0.0, 0.0, self.sm, stream)
            launches += 1
        # pass 1: partial Gram over my coordinate shard (tcgen05 for n > 16, exact fp32 otherwise)
        tc = ext.gram_umma_tile_cols(nt)
        main = (ln // tc) * tc if nt > 16 else 0
        if main > 0:
            tail_ptr = 0
            if main < ln:
                ext.gram(self._all_rows, self._all_scales, off + main, ln - main, self._gram_scratch.data_ptr(),
                         self._gram_scratch.numel() // (nt * nt), self._g_tail32.data_ptr(),
                         self._g_tail64.data_ptr(), self.sm, stream)
                tail_ptr = self._g_tail64.data_ptr()
                launches += 2
            # TMA-fed when the table is a few matrix segments (one (L, d_pad) gradient matrix per peer GPU
            # + the local virtual / auxiliary rows), per-thread cp.async otherwise
            from ..ops import umma

            self.gram_path = umma.launch(ext, self._all_rows, self._all_scales, off, main, self.d_pad,
                                         self._umma_scratch, nt, tail_ptr, self._g_local32.data_ptr(),
                                         self._g_local64.data_ptr(), self.sm, stream)
        else:
            ext.gram(self._all_rows, self._all_scales, off, ln, self._gram_scratch.data_ptr(),
                     self._gram_scratch.numel() // (nt * nt), self._g_local32.data_ptr(),
                     self._g_local64.data_ptr(), self.sm, stream)
        launches += 2
        # all-reduce the (nt, nt) partials through peer stores
        # (NVLS multicast: summed inside the switch by multimem.ld_reduce, broadcast by multimem.st)
        ext.gram_exchange(self._g_local64.data_ptr(),
                          [self.sym.peer_ptr(r, self._off_gslots) for r in range(self.world)], pads,
                          self.rank, nt, ctl + 8, ctl + 4, self._g_total64.data_ptr(), 0, stream,
                          self.live_mask, self.spin_seconds,
                          self.sym.mc_ptr(self._off_gslots) if self._agg_mc else 0)
        launches += 1
        # n-space solve (device side) -> weights
        w = plan.solver(self._g_total64)
        self._w_dev.copy_(w.reshape(-1).to(torch.float32), non_blocking=True)
        launches += 2
        # pass 2: weighted sum on my shard + broadcast + SGD
        ext.fused_ps_wsum(self._all_rows, self._all_scales, self._w_dev.data_ptr(), self.d_pad, off, ln,
                          self.rank, [self.sym.peer_ptr(r, self._off_agg) for r in range(self.world)], pads,
                          ctl + 8, ctl + 0, ctl + 4, self._upd_params, self._upd_moms, self.lr,
                          self.momentum, self.weight_decay, self.sm, stream, 0, 0, 0, 0, 0, self._agg_mc,
                          self.live_mask, self.spin_seconds)
        self.launches_per_step = launches + 1

    # ------------------------------------------- pre-aggregator -> coordinate-wise round
    def _setup_mapcw_plan(self) -> None:
        """Static buffers of the map -> coordinate-wise round."""
        lay, dev, plan = self.layout, self.device, self.plan
        n = lay.n_workers + lay.n_virtual
        m = int(plan.m)
        if n > ops.MAXN or m > ops.MAXN or n > self.nt_max:
            raise ValueError("too many rows for the fused map + coordinate-wise round")
        self._map_n, self._map_m = n, m
        self._virt_buf = (torch.zeros(self.d_pad, dtype=torch.float32, device=dev) if lay.n_virtual else None)
        rows, scales = list(self._rows), list(self._scales)
        if lay.n_virtual:
            rows += [self._virt_buf.data_ptr()] * lay.n_virtual
            scales += [1.0] * lay.n_virtual
        self._map_rows, self._map_scales = rows, scales
        # the m mixed rows exist only over THIS rank's coordinate shard (recover() re-runs this set-up
        # with the new live set, so the buffer follows the shard)
        self._map_cap = padded_size(self._shard(0, self.d_pad)[1], 1024)
        self._map_Y = torch.zeros((m, self._map_cap), dtype=torch.float32, device=dev)
        self._map_W = torch.zeros((m, n), dtype=torch.float32, device=dev)
        if plan.needs_gram:
            self._g_local64 = torch.zeros((n, n), dtype=torch.float64, device=dev)
            self._g_local32 = torch.zeros((n, n), dtype=torch.float32, device=dev)
            self._g_tail64 = torch.zeros((n, n), dtype=torch.float64, device=dev)
            self._g_tail32 = torch.zeros((n, n), dtype=torch.float32, device=dev)
            self._g_total64 = torch.zeros((n, n), dtype=torch.float64, device=dev)
            self._gram_scratch = torch.empty(self.ext.gram_partials_needed(n, self.sm), dtype=torch.float32, device=dev)
            self._umma_scratch = torch.empty(self.sm * 8 * 2 * n * n, dtype=torch.float32, device=dev)

    def _launch_mapcw_round(self, stream: int, ctl: int) -> None:
        """flag barrier -> (virtual rows) -> (Gram pass + in-switch all-reduce -> map kernel) ->
        ``Y = W_p X`` on my shard -> fused select / deliver / SGD over the m local rows."""
        ext, lay, plan = self.ext, self.layout, self.plan
        pads = [self.sym.peer_ptr(r, self._off_pad) for r in range(self.world)]
        off, ln = self._shard(0, self.d_pad)
        if ln > self._map_cap:
            raise RuntimeError("coordinate shard exceeds the mixed-row buffer")
        n, m = self._map_n, self._map_m
        rows, scales = self._map_rows, self._map_scales
        launches = 1
        # every rank's gradient rows must be complete before anybody reads them (its own kernel: the passes
        # below read the rows with non-coherent loads)
        ext.flag_barrier(pads, self.rank, ext.PAD_READY, ctl + 8, ctl + 4, stream, 0, 0, self.live_mask,
                         self.spin_seconds)
        launches += 1
        if lay.n_virtual:
            nh = lay.n_honest
            ext.colstat(self._rows[:nh], self._scales[:nh], float(self.virtual_fold.a), float(self.virtual_fold.b),
                        off, ln, self._virt_buf.data_ptr(), self.sm, stream)
            launches += 1
        G = None
        if plan.needs_gram:
            tc = ext.gram_umma_tile_cols(n)
            main = (ln // tc) * tc if n > 16 else 0
            if main > 0:
                tail_ptr = 0
                if main < ln:
                    ext.gram(rows, scales, off + main, ln - main, self._gram_scratch.data_ptr(),
                             self._gram_scratch.numel() // (n * n), self._g_tail32.data_ptr(),
                             self._g_tail64.data_ptr(), self.sm, stream)
                    tail_ptr = self._g_tail64.data_ptr()
                    launches += 2
                from ..ops import umma

                self.gram_path = umma.launch(ext, rows, scales, off, main, self.d_pad, self._umma_scratch, n, tail_ptr,
                                             self._g_local32.data_ptr(), self._g_local64.data_ptr(), self.sm, stream)
            else:
                ext.gram(rows, scales, off, ln, self._gram_scratch.data_ptr(), self._gram_scratch.numel() // (n * n),
                         self._g_local32.data_ptr(), self._g_local64.data_ptr(), self.sm, stream)
            launches += 2
            ext.gram_exchange(self._g_local64.data_ptr(),
                              [self.sym.peer_ptr(r, self._off_gslots) for r in range(self.world)], pads,
                              self.rank, n, ctl + 8, ctl + 4, self._g_total64.data_ptr(), 0, stream,
                              self.live_mask, self.spin_seconds,
                              self.sym.mc_ptr(self._off_gslots) if self._agg_mc else 0)
            launches += 1
            G = self._g_total64
        W = plan.weights(G)
        self._map_W.copy_(W.reshape(m, n).to(torch.float32), non_blocking=True)
        launches += 2
        # Y[:, 0:ln] = W X[:, off:off+ln]: the kernels index outputs with the global coordinate, so the
        # output pointers are shifted back by the shard offset (a multiple of 4 elements: stays 16-byte aligned)
        outs = [self._map_Y[r].data_ptr() - 4 * off for r in range(m)]
        wptr = self._map_W.data_ptr()
        main = 0
        if m > 8 and hasattr(ext, "wsum_multi"):
            tile = ext.WSUM_MULTI_TILE
            main = (ln // tile) * tile
            if main:
                ext.wsum_multi(rows, scales, wptr, m, off, main, outs, self.sm, stream)
                launches += 1
        if main < ln:
            for r0 in range(0, m, 8):
                mb = min(8, m - r0)
                ext.wsum(rows, scales, wptr + 4 * r0 * n, mb, off + main, ln - main, outs[r0:r0 + mb],
                         [], [], 0.0, 0.0, 0.0, self.sm, stream)
                launches += 1
        # the fused kernel over the m LOCAL mixed rows: its ready wait is already satisfied (same sequence
        # number as the barrier above), delivery and SGD are those of the plain coordinate-wise round
        cw = plan.cw
        self.ext.fused_ps_cw(
            outs, [1.0] * m, cw.mode, cw.f, 0, 0, 0.0, 0.0, self.d_pad, off, ln, self.rank,
            [self.sym.peer_ptr(r, self._off_agg) for r in range(self.world)], pads,
            0, ctl + 8, ctl + 0, ctl + 4,
            self._upd_params, self._upd_moms, self.lr, self.momentum, self.weight_decay,
            self.sm, stream, 0, 0, self.d_pad, 1, 0, self._agg_mc, self.live_mask, self.spin_seconds, 0)
        self.launches_per_step = launches + 1

    def _upd_index(self) -> List[int]:
        return [i for i, w in enumerate(self.workers) if w.role == "honest" or self.update_byzantines]

    def _body(self) -> None:
        main = torch.cuda.current_stream(self.device)
        l0 = ops.launches()
        bucketed = isinstance(self.plan, CwPlan) and self._use_buckets
        self._stamp(0)
        if bucketed:
            # the round's epoch is bumped up front: bucket launches are enqueued from inside backward
            self.ext.bump_u32(self.ctl.data_ptr() + 8, main.cuda_stream)
            self._round_launches = 1
            self._agg_stream.wait_stream(main)
            self._bk_events = [[] for _ in range(self.n_buckets)]
            self._bk_count = [0] * self.n_buckets
            self._next_bucket = 0
            self._in_round = True
        try:
            self._run_replicas(main)
        finally:
            self._in_round = False
        self.model_launches_per_step = ops.launches() - l0   # this library's BN / pooling kernels
        self._stamp(1)
        if not bucketed:
            self.launch_aggregate()
            self._stamp(2)
            return
        # buckets whose mark never fired (or ranks without a replica) and the head bucket: after the
        # whole backward pass, on the aggregation stream, full grid
        agg = self._agg_stream
        agg.wait_stream(main)
        with torch.cuda.stream(agg):
            while self._next_bucket < self.n_buckets:
                last = self._next_bucket == self.n_buckets - 1
                self._launch_cw_bucket(self._next_bucket, 0 if last else self.overlap_grid)
                self._next_bucket += 1
        main.wait_stream(agg)
        self._stamp(2)
        self.launches_per_step = self._round_launches

    def _run_replicas(self, main) -> None:
        if self._side_streams:
            streams = [main] + self._side_streams
            for s in self._side_streams:
                s.wait_stream(main)
            for i, w in enumerate(self.workers):
                with torch.cuda.stream(streams[i % len(streams)]):
                    w.forward_backward(self.amp_dtype)
            for s in self._side_streams:
                main.wait_stream(s)
        else:
            for w in self.workers:
                w.forward_backward(self.amp_dtype)

    # ---------------------------------------------------------------- bucket validation
    def _snapshot(self):
        return (self.params.clone(), None if self.moms is None else self.moms.clone(),
                [[b.clone() for b in w.model.buffers()] for w in self.workers])

    def _restore(self, snap) -> None:
        snap_p, snap_m, snap_b = snap
        with torch.no_grad():
            self.params.copy_(snap_p)
            if snap_m is not None:
                self.moms.copy_(snap_m)
            for w, bufs in zip(self.workers, snap_b):
                for b, saved in zip(w.model.buffers(), bufs):
                    b.copy_(saved)

    def _validate_buckets(self) -> None:
        """Run the same round twice from the same state -- one launch over the whole arena after
        backward, then the overlapped bucket sequence -- and compare the aggregates.  A bucket
        launched before all of its gradients exist (a model whose execution order differs from its
        registration order) reads the zeroed arena and shows up here; the round then falls back to
        the single launch.  Both rounds are rolled back."""
        self._buckets_validated = True
        if not (isinstance(self.plan, CwPlan) and self._use_buckets):
            return
        snap = self._snapshot()
        s = torch.cuda.Stream(self.device)
        s.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(s):
            self._use_buckets = False
            self._body()
            ref = self.agg.clone()
            s.synchronize()
            self._restore(snap)
            self._use_buckets = True
            self._body()
            got = self.agg.clone()
        torch.cuda.current_stream(self.device).wait_stream(s)
        torch.cuda.synchronize(self.device)
        self._restore(snap)
        scale = ref.abs().max().clamp_min(1e-30)
        bad = torch.tensor([float(not torch.allclose(got, ref, rtol=5e-2, atol=float(scale) * 1e-3))],
                           device=self.device)
        if self.world > 1:
            dist.all_reduce(bad, op=dist.ReduceOp.MAX, group=self.group)
        if bad.item() != 0.0:
            import warnings

            warnings.warn("gradient-bucket overlap disabled: the bucketed round did not reproduce the "
                          "single-launch aggregate (the model's backward order does not follow its "
                          "parameter order)")
            self._use_buckets = False
            for w in self.workers:
                w.remove_marks()
        torch.cuda.synchronize(self.device)

    def capture(self, warmup: int = 2) -> None:
        """Warm up eagerly on a side stream, then capture the whole round in a CUDA graph."""
        for w in self.workers:
            if w.static_x is None:
                raise RuntimeError("stage a batch on every worker before capture()")
        # Warm-up rounds are real rounds: snapshot the training state and restore it afterwards
        # so that capture() is invisible to the optimisation trajectory.
        if not self._buckets_validated:
            self._validate_buckets()
        snap = self._snapshot()
        s = self._capture_stream
        s.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(s):
            for _ in range(warmup):
                self._body()
        torch.cuda.current_stream(self.device).wait_stream(s)
        torch.cuda.synchronize(self.device)
        if self.world > 1 and self._recoveries == 0:      # (after a recovery the process group has a dead member)
            dist.barrier(group=self.group)
        g = torch.cuda.CUDAGraph()
        other = self._graphs[1 - self._buf]
        with torch.cuda.graph(g, stream=self._capture_stream, **({"pool": other.pool()} if other is not None else {})):
            self._body()
        self._graphs[self._buf] = g
        self._restore(snap)
        torch.cuda.synchronize(self.device)

    # --------------------------------------------------------------------- step
    @property
    def _graph(self) -> Optional[torch.cuda.CUDAGraph]:
        """The captured round reading the ACTIVE input buffer set."""
        return self._graphs[self._buf]

    def step(self, batches: Optional[Sequence[Tuple[torch.Tensor, torch.Tensor]]] = None, *,
             prefetch: bool = True) -> torch.Tensor:
        """One training round.  Returns the device tensor of per-worker losses.

        ``batches[i]`` = this step's (x, y) for local worker i: pinned host tensors are copied H2D
        asynchronously on the launching stream, then the round runs.

        ``batches=None``: every worker pulls from its ``data`` source.  The inputs are then
        **double buffered**: as soon as round k is launched, batch k+1 is fetched and copied H2D on a
        copy stream into the other buffer set (its own captured graph), so the copy overlaps the
        round instead of preceding it; round k+1 only waits for that copy's event."""
        if self._fault == "silent":
            return self.losses              # test hook: this rank has stopped taking part (inject_fault)
        main = torch.cuda.current_stream(self.device)
        explicit = batches is not None
        b = self._buf
        for w in self.workers:
            w.buf = b
        if explicit or not self._prefetched[b]:
            if batches is None:
                batches = [w.data() for w in self.workers]
            stale = False
            for w, (x, y) in zip(self.workers, batches):
                stale |= w.stage_batch(x, y, b)
            if stale:
                self._graphs[b] = None
        else:
            main.wait_event(self._h2d_done[b])
        self._prefetched[b] = False
        refresh = getattr(self.plan, "refresh", None)
        if refresh is not None:
            refresh()
        if self.use_cuda_graph:
            if self._graphs[b] is None:
                self.capture()
            self._graphs[b].replay()
        else:
            if not self._buckets_validated:
                self._validate_buckets()
            self._body()
        self._consumed[b].record(main)
        if (not explicit) and prefetch and all(w.data is not None for w in self.workers):
            nb = 1 - b
            nxt = [w.data() for w in self.workers]
            cs = self._copy_stream
            cs.wait_event(self._consumed[nb])       # the last round that read set nb has finished
            stale = False
            with torch.cuda.stream(cs):
                for w, (x, y) in zip(self.workers, nxt):
                    stale |= w.stage_batch(x, y, nb)
            if stale:
                self._graphs[nb] = None
            self._h2d_done[nb].record(cs)
            self._prefetched[nb] = True
            self._buf = nb
        return self.losses

    def replay(self) -> torch.Tensor:
        """Run one more round on the batches already resident in the active input buffers (no H2D
        copy): the device-only form of :meth:`step`, for kernel-level timing."""
        refresh = getattr(self.plan, "refresh", None)
        if refresh is not None:
            refresh()
        if self.use_cuda_graph:
            if self._graphs[self._buf] is None:
                self.capture()
            self._graphs[self._buf].replay()
        else:
            if not self._buckets_validated:
                self._validate_buckets()
            self._body()
        return self.losses

    def read_losses(self) -> torch.Tensor:
        """Device->host read of the round's losses (synchronises the stream).  The fused kernels'
        status word rides along in the same copy: a flag-wait timeout raises here instead of
        leaving a training loop running on frozen replicas."""
        self._status_host.copy_(self.ctl[1:2], non_blocking=True)
        self.losses_host.copy_(self.losses, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        self._raise_on_status(int(self._status_host.item()))
        return self.losses_host

    _STATUS_BITS = {1: "gradient-ready wait", 2: "delivery wait", 4: "flag barrier", 8: "Gram exchange"}

    @staticmethod
    def decode_status(st: int) -> Tuple[List[str], List[int]]:
        """(names of the waits that timed out, ranks that did not arrive)."""
        kinds = [name for bit, name in DeviceRound._STATUS_BITS.items() if st & bit]
        ranks = [r for r in range(8) if (st >> (8 + r)) & 1]
        return kinds, ranks

    def _raise_on_status(self, st: int) -> None:
        if st != 0:
            kinds, ranks = self.decode_status(st)
            raise RuntimeError(f"fused PS round failed on rank {self.rank}: status {st:#x} "
                               f"({', '.join(kinds) or 'wait'} timed out; silent ranks {ranks}). The status "
                               f"is sticky: later rounds skip aggregation until reset_status() or recover().")

    def status(self) -> int:
        return int(self.ctl[1].item())

    def check_status(self) -> None:
        self._raise_on_status(self.status())

    def reset_status(self) -> None:
        """Clear the sticky error word (after the cause of a timeout has been dealt with)."""
        torch.cuda.synchronize(self.device)
        self.ctl[0:2].zero_()
        torch.cuda.synchronize(self.device)

    # ------------------------------------------------------------------ failure handling
    def inject_fault(self, kind: Optional[str] = "silent") -> None:
        """Test hook: ``"silent"`` makes this rank stop taking part in rounds -- its
        ``step()`` returns without launching anything, exactly what its peers see when the process
        hangs; ``None`` clears it.  The peers' fused kernels then time out (``spin_seconds``),
        report this rank in their status word and can :meth:`recover`."""
        self._fault = kind

    def silent_ranks(self) -> List[int]:
        """Ranks named by the sticky status word of the last failed wait."""
        return self.decode_status(self.status())[1]

    def recover(self, plan_factory: Optional[Callable[[int], object]] = None) -> List[int]:
        """Drop the ranks that did not arrive and continue with the remaining rows.

        Every surviving rank calls this after a flag wait timed out (``read_losses`` /
        ``check_status`` raised).  The silent ranks are read from the status word; their gradient
        rows leave the row table, their flags are no longer awaited and nothing is delivered to
        them (``live_mask``); the aggregation plan is rebuilt for the smaller row count by
        ``plan_factory(n_rows)`` (``ParameterServer`` passes ``aggregator.fused_plan``).  Because a
        failure in the middle of a round can leave the survivors with different sets of applied
        buckets, the replicas (parameters and momentum) are re-synchronised from the lowest live
        rank through the symmetric ``agg`` buffer.  Captured graphs are dropped (row tables and
        masks are baked into them).  Returns the list of dropped ranks.

        The reference has no failure handling on this path: a hung node actor hangs
        ``ParameterServer.round()`` (reference engine/parameter_server/ps.py:121-144)."""
        torch.cuda.synchronize(self.device)
        st = self.status()
        _, silent = self.decode_status(st)
        silent = [r for r in silent if r < self.world]
        if not silent:
            self.reset_status()
            return []
        if self.rank in silent:
            raise RuntimeError(f"rank {self.rank} was reported silent by itself; cannot recover")
        live_before = [r for r in range(self.world) if self.live_mask == 0 or (self.live_mask >> r) & 1]
        live = [r for r in live_before if r not in silent]
        if not live:
            raise RuntimeError("no live rank left")
        self.live_mask = sum(1 << r for r in live)
        lay = self.layout
        keep = [g for g in range(lay.n_workers) if lay.rank_of[g] in live]
        dropped_honest = sum(1 for g in range(lay.n_honest) if lay.rank_of[g] not in live)
        new_layout = RowLayout(lay.n_honest - dropped_honest,
                               lay.n_byz_workers - (lay.n_workers - len(keep) - dropped_honest),
                               lay.n_virtual, lay.world, [lay.rank_of[g] for g in keep],
                               [lay.slot_of[g] for g in keep])
        old_rows, old_scales = self._rows, self._scales
        self._rows = [old_rows[g] for g in keep]
        self._scales = [old_scales[g] for g in keep]
        self.layout = new_layout
        n_rows = new_layout.n_workers + new_layout.n_virtual
        if plan_factory is not None:
            plan = plan_factory(n_rows)
            if plan is None or type(plan) is not type(self.plan):
                raise RuntimeError(f"no fused plan of the same family for the remaining {n_rows} rows")
            self.plan = plan
        if isinstance(self.plan, GramPlan):
            self._setup_gram_plan()
        if isinstance(self.plan, MapCwPlan):
            self._setup_mapcw_plan()
        self._graphs = [None, None]
        self._prefetched = [False, False]
        # fresh flag numbering: the survivors agree on a new epoch base above anything published so far
        self._recoveries += 1
        self.ctl[0:2].zero_()
        torch.cuda.synchronize(self.device)
        self._resync_from_lowest_live(live)
        return silent

    def _resync_from_lowest_live(self, live: List[int]) -> None:
        """Copy replica 0 of the lowest live rank (parameters, then momentum) into every replica of
        every live rank: source -> its symmetric agg buffer, device flag barrier among the live
        ranks, peers read it over NVLink."""
        src = live[0]
        stream = torch.cuda.current_stream(self.device).cuda_stream
        pads = [self.sym.peer_ptr(r, self._off_pad) for r in range(self.world)]
        ctl = self.ctl.data_ptr()
        view = tensor_from_ptr(self.sym.peer_ptr(src, self._off_agg), self.d_pad * 4, self.device,
                               owner=self.sym).view(torch.float32)
        for which in ("params", "moms"):
            bank = self.params if which == "params" else self.moms
            if bank is None:
                continue
            has = torch.tensor([1 if self.L else 0], device=self.device)
            if self.rank == src:
                if not self.L:
                    raise RuntimeError("the lowest live rank hosts no replica to resynchronise from")
                self.agg.copy_(bank[0])
            # two barriers per bank on their own flag words and sequence numbers (PAD_SYNC)
            for phase in range(2):
                self._resync_seq += 1
                self.ext.flag_barrier(pads, self.rank, PAD_SYNC, self._resync_ctr.data_ptr(), ctl + 4, stream,
                                      1, self._resync_seq, self.live_mask, max(self.spin_seconds, 5.0))
                if phase == 0 and self.L:
                    for i in range(self.L):
                        bank[i].copy_(view)
            del has
        torch.cuda.synchronize(self.device)
        self.check_status()

    def aggregated(self) -> torch.Tensor:
        return self.agg[: self.d]

    def close(self) -> None:
        self._graphs = [None, None]
        self.sym.close()


__all__ = ["CwPlan", "GramPlan", "MapCwPlan", "RowFold", "DeviceWorker", "RowLayout", "DeviceRound", "pick_bucket_offsets",
           "bucket_bounds"]
