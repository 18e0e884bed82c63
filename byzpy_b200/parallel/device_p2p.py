"""B200-native peer-to-peer (gossip) round: the CUDA runtime behind
:class:`byzpy_b200.engine.peer_to_peer.train.PeerToPeer` when its nodes are device nodes.

Reference semantics (reference engine/node/mixin.py:59-105, examples/p2p/*): every honest node i
takes a local half step ``theta_i <- theta_i - lr * grad_i``, publishes ``theta_i^{t+1/2}``, and
replaces its parameters by a robust aggregate of its own vector and its in-neighbours' vectors;
Byzantine nodes publish an attack vector built from the honest vectors they can see.

Here the published vectors live in CUDA-IPC symmetric memory (one row per node), so "broadcast to
the out-neighbours" costs nothing: a topology only selects WHICH peer rows a node's aggregation
kernels load over NVLink.  A round on a rank is

    fwd/bwd + flat SGD half step + publish (one copy into the symmetric row)       per local honest node
    flag barrier (device side)                                                      all rows published
    attack kernels (column statistics / alias copy over peer rows) + flag barrier   local Byzantine nodes
    robust aggregation straight from peer HBM into the local parameter arena        per local honest node
    flag barrier                                                                     rows may be overwritten

All synchronisation is by release/acquire flags in device memory; the whole round is captured in
a CUDA graph when the aggregator's n-space solve is device-resident.
"""
from __future__ import annotations

import contextlib
from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist
import torch.nn as nn

from .. import ops
from .arena import ParamArena, padded_size
from .device_ps import CwPlan, GramPlan, RowFold
from .symmetric import SymmetricBuffer, _dist_on

PAD_BYZ = 48  # spare region of the signal pad used for the "attack vectors published" barrier


class DevicePeer:
    """One P2P node resident on the local GPU."""

    def __init__(self, *, role: str, model: Optional[nn.Module] = None, loss_fn: Optional[Callable] = None,
                 plan=None, fold: Optional[RowFold] = None, preprocess: Optional[Callable] = None,
                 data: Optional[Callable] = None, name: str = "peer"):
        self.role, self.model, self.loss_fn = role, model, loss_fn
        self.plan, self.fold, self.preprocess, self.data, self.name = plan, fold, preprocess, data, name
        self.arena: Optional[ParamArena] = None
        self.static_x = self.static_y = None
        self.loss_slot: Optional[torch.Tensor] = None
        self.sink = None        # ops.fused_layers.GradSink once direct gradients are enabled

    def stage_batch(self, x: torch.Tensor, y: torch.Tensor, dev: torch.device) -> None:
        if self.static_x is None or self.static_x.shape != x.shape or self.static_x.dtype != x.dtype:
            self.static_x = torch.empty(x.shape, dtype=x.dtype, device=dev)
            self.static_y = torch.empty(y.shape, dtype=y.dtype, device=dev)
        self.static_x.copy_(x, non_blocking=True)
        self.static_y.copy_(y, non_blocking=True)


@dataclass
class PeerLayout:
    """Global node order (honest first, then Byzantine), block-distributed over ranks."""

    n_honest: int
    n_byz: int
    world: int = 1

    @property
    def n(self) -> int:
        return self.n_honest + self.n_byz

    def rank_of(self, g: int) -> int:
        per = self.n // self.world
        return g // per

    def slot_of(self, g: int) -> int:
        return g % (self.n // self.world)

    def local_ids(self, rank: int) -> List[int]:
        if self.n % self.world:
            raise ValueError(f"{self.n} nodes do not divide evenly over {self.world} ranks")
        return [g for g in range(self.n) if self.rank_of(g) == rank]


class DeviceP2PRound:
    """Fused gossip round on GPUs: the device path of :class:`~byzpy_b200.engine.peer_to_peer.train.PeerToPeer`.

    Every peer's published vector lives in a row of a symmetric-memory arena.  A round on a rank: the local half steps
    of its honest peers (forward, backward, flat SGD, publish); a device-side flag barrier; every remote in-neighbour's
    row is pulled over NVLink ONCE into local staging rows (a multi-pass aggregator run for several local peers would
    otherwise fetch each remote byte many times); the attack kernels of local Byzantine peers (column statistics or an
    alias copy over the honest rows) and a second barrier; then, per local honest peer, the robust aggregate of its
    own and its in-neighbours' rows is written into its parameter arena -- the topology only selects which rows a
    kernel loads.  Synchronisation is by release/acquire flags in device memory and the whole round is one CUDA
    graph when every aggregator's small solve runs on the device.

    Parameters
    ----------
    peers : sequence of DevicePeer
        This rank's peers (honest first).
    layout : PeerLayout
        Which global peer ids each rank hosts.
    topology : Topology
    lr : float
    device, group, amp_dtype, use_cuda_graph :
        As for :class:`byzpy_b200.parallel.device_ps.DeviceRound`.

    Notes
    -----
    ``step()`` enqueues one round; ``read_losses()`` waits for it and raises when a rank stopped answering
    (``check_status``); ``close()`` releases the symmetric memory.
    """
    def __init__(self, peers: Sequence[DevicePeer], layout: PeerLayout, topology, *, lr: float,
                 device: Optional[torch.device] = None, group=None,
                 amp_dtype: Optional[torch.dtype] = torch.bfloat16, use_cuda_graph: bool = True):
        if not torch.cuda.is_available():
            raise RuntimeError("DeviceP2PRound needs a CUDA device")
        self.ext = ops.require_ext()
        self.peers = list(peers)
        self.layout, self.topology, self.lr = layout, topology, float(lr)
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self.group, self.amp_dtype = group, amp_dtype
        self.rank = dist.get_rank(group) if _dist_on() else 0
        self.world = dist.get_world_size(group) if _dist_on() else 1
        if layout.world != self.world:
            raise ValueError("layout.world does not match the process group")
        self.local_ids = layout.local_ids(self.rank)
        if len(self.local_ids) != len(self.peers):
            raise ValueError("number of local peers does not match the layout")
        L = self.L = len(self.peers)
        model0 = next(p.model for p in self.peers if p.model is not None) if any(
            p.model is not None for p in self.peers) else None
        if model0 is None:
            d_local = 0
        else:
            d_local = sum(p.numel() for p in model0.parameters())
        if self.world > 1:
            t = torch.tensor([d_local], device=self.device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
            d_local = int(t.item())
        self.d = d_local
        self.d_pad = padded_size(self.d, 1024)
        self.sm = ops.sm_count(self.device)
        f4 = 4
        self._off_theta = 0
        self._off_pad = L * self.d_pad * f4
        self._off_ctl = self._off_pad + 256
        self.sym = SymmetricBuffer(self._off_ctl + 256, self.device, group)
        self.theta = self.sym.view(torch.float32, L * self.d_pad, 0).view(L, self.d_pad)
        self.ctl = self.sym.view(torch.int32, 64, self._off_ctl)
        self.params = torch.zeros((L, self.d_pad), dtype=torch.float32, device=self.device)
        self.grads = torch.zeros((L, self.d_pad), dtype=torch.float32, device=self.device)
        self.losses = torch.zeros(L, dtype=torch.float32, device=self.device)
        self.losses_host = torch.zeros(L, dtype=torch.float32).pin_memory()
        for i, p in enumerate(self.peers):
            p.loss_slot = self.losses[i]
            if p.model is not None:
                p.model.to(self.device)
                p.arena = ParamArena(p.model, flat_params=self.params[i], flat_grads=self.grads[i])
                if amp_dtype == torch.bfloat16:
                    # in-place parameter gradients + weight-gradient GEMMs on a side stream
                    from ..ops.fused_layers import enable_direct_grads

                    if not hasattr(self, "_wgrad_stream"):
                        self._wgrad_stream = torch.cuda.Stream(self.device)
                    p.sink = enable_direct_grads(p.model, side_stream=self._wgrad_stream)
        self._pads = [self.sym.peer_ptr(r, self._off_pad) for r in range(self.world)]
        self._has_byz = layout.n_byz > 0
        # Remote vectors are staged ONCE per round into local HBM (one pass over NVLink): with L local
        # peers and a 2-3 pass aggregator every remote byte would otherwise cross the link 2-3 x L
        # times (measured: BERT-base GM at 2 GPUs was NVLink bound, profiles/training_configs.md).
        needed = []
        for g in self.local_ids:
            for j in dict.fromkeys(self.topology.in_.get(g, [])):
                if self.layout.rank_of(j) != self.rank and j not in needed:
                    needed.append(j)
        self._staged_ids = needed
        self._stage = (torch.empty((len(needed), self.d_pad), dtype=torch.float32, device=self.device)
                       if needed else None)
        self._stage_slot = {j: k for k, j in enumerate(needed)}
        self._tables = [self._neighbour_table(g) for g in self.local_ids]
        self._work = [self._workspace(i) for i in range(L)]
        self.use_cuda_graph = use_cuda_graph and all(
            not isinstance(p.plan, GramPlan) or p.plan.capturable for p in self.peers if p.role == "honest")
        self._graph: Optional[torch.cuda.CUDAGraph] = None
        self.launches_per_step = 0
        if self.world > 1:
            dist.barrier(group=group)
        torch.cuda.synchronize(self.device)

    # ---------------------------------------------------------------------------------
    def _remote_ptr(self, g: int) -> int:
        return self.sym.peer_ptr(self.layout.rank_of(g), self.layout.slot_of(g) * self.d_pad * 4)

    def _row_ptr(self, g: int) -> int:
        """Where peer g's published vector is read from: its symmetric slot when it lives on this
        rank, otherwise this rank's staged copy."""
        k = self._stage_slot.get(g)
        if k is not None:
            return self._stage[k].data_ptr()
        return self._remote_ptr(g)

    def _stage_remote(self, honest: bool, stream: int) -> None:
        """Pull the remote peers' vectors over NVLink into the local staging rows (P2P loads)."""
        for j in self._staged_ids:
            if (j < self.layout.n_honest) == honest:
                self.ext.scale_copy(self._remote_ptr(j), self._stage[self._stage_slot[j]].data_ptr(), 1.0,
                                    self.d_pad, self.sm, stream)

    def _neighbour_table(self, g: int) -> List[int]:
        ins = list(dict.fromkeys(self.topology.in_.get(g, [])))
        if g < self.layout.n_honest:
            return [self._row_ptr(g)] + [self._row_ptr(j) for j in ins if j != g]
        return [self._row_ptr(j) for j in ins if j < self.layout.n_honest]  # Byzantine: honest vectors only

    def _workspace(self, i: int) -> dict:
        p = self.peers[i]
        ws: dict = {}
        if p.role == "honest" and isinstance(p.plan, GramPlan):
            nt = len(self._tables[i]) + len(p.plan.aux)
            dev = self.device
            ws["aux"] = [torch.zeros(self.d_pad, dtype=torch.float32, device=dev) for _ in p.plan.aux]
            for kind, buf in zip(p.plan.aux, ws["aux"]):       # constant rows (CAF's start direction): filled once
                if isinstance(kind, tuple) and kind[0] == "const":
                    buf[: self.d].copy_(kind[1](self.d, buf).to(torch.float32))
            ws["G32"] = torch.zeros((nt, nt), dtype=torch.float32, device=dev)
            ws["G64"] = torch.zeros((nt, nt), dtype=torch.float64, device=dev)
            ws["T32"] = torch.zeros((nt, nt), dtype=torch.float32, device=dev)
            ws["T64"] = torch.zeros((nt, nt), dtype=torch.float64, device=dev)
            ws["W"] = torch.zeros(nt, dtype=torch.float32, device=dev)
            ws["scratch"] = torch.empty(self.ext.gram_partials_needed(nt, self.sm), dtype=torch.float32, device=dev)
            ws["umma"] = torch.empty(self.sm * 8 * 2 * nt * nt, dtype=torch.float32, device=dev)
            ws["nt"] = nt
        return ws

    # ----------------------------------------------------------------------------- body
    def _half_step(self, i: int) -> None:
        p = self.peers[i]
        p.arena.flat_grads.zero_()
        x = p.preprocess(p.static_x) if p.preprocess is not None else p.static_x
        ctx = (torch.autocast("cuda", dtype=self.amp_dtype) if self.amp_dtype is not None
               else contextlib.nullcontext())
        if p.sink is not None:
            p.sink.refresh_shadows()
        with ctx:
            loss = p.loss_fn(p.model(x), p.static_y)
        loss.backward()
        if p.sink is not None:
            p.sink.join()
        p.loss_slot.copy_(loss.detach().float())
        ops.sgd_step(self.grads[i], [self.params[i]], None, lr=self.lr)   # theta <- theta - lr * grad
        self.theta[i].copy_(self.params[i])                                # publish

    def _attack(self, i: int, stream: int) -> None:
        p, rows = self.peers[i], self._tables[i]
        fold = p.fold
        if not rows or fold is None:
            return
        out = self.theta[i].data_ptr()
        if fold.kind == "virtual":
            self.ext.colstat(rows, [], float(fold.a), float(fold.b), 0, self.d_pad, out, self.sm, stream)
        elif fold.kind == "alias":
            self.ext.scale_copy(rows[min(fold.index, len(rows) - 1)], out, 1.0, self.d_pad, self.sm, stream)
        else:
            raise ValueError(f"attack fold {fold.kind!r} is not usable in P2P")

    def _aggregate(self, i: int, stream: int) -> None:
        p, rows, ws = self.peers[i], self._tables[i], self._work[i]
        ext, out, d = self.ext, self.params[i].data_ptr(), self.d_pad
        plan = p.plan
        if isinstance(plan, CwPlan):
            ext.cw_select(rows, [], plan.mode, plan.f, 0, 0, 0.0, 0.0, 0, d, out, [], [], 0.0, 0.0, 0.0,
                          self.sm, stream)
            return
        all_rows = list(rows)
        nt = ws["nt"]
        fused_median = tuple(plan.aux) == ("median",) and len(rows) <= 16
        if not fused_median:
            for kind, buf in zip(plan.aux, ws["aux"]):
                if kind == "median":
                    ext.cw_select(rows, [], ops.MODE_MEDIAN, 0, 0, 0, 0.0, 0.0, 0, d, buf.data_ptr(), [], [], 0.0,
                                  0.0, 0.0, self.sm, stream)
                all_rows.append(buf.data_ptr())
        tc = ext.gram_umma_tile_cols(nt)
        main = (d // tc) * tc if nt > 16 else 0
        if fused_median:
            # median start row + Gram of [rows..., median] in ONE pass (csrc/gram.cu, AUX variant)
            med = ws["aux"][0]
            ext.gram(rows, [], 0, d, ws["scratch"].data_ptr(), ws["scratch"].numel() // (nt * nt),
                     ws["G32"].data_ptr(), ws["G64"].data_ptr(), self.sm, stream, med.data_ptr())
            all_rows.append(med.data_ptr())
        elif main > 0:
            tail = 0
            if main < d:
                ext.gram(all_rows, [], main, d - main, ws["scratch"].data_ptr(), ws["scratch"].numel() // (nt * nt),
                         ws["T32"].data_ptr(), ws["T64"].data_ptr(), self.sm, stream)
                tail = ws["T64"].data_ptr()
            ext.gram_umma(all_rows, [], 0, main, ws["umma"].data_ptr(), ws["umma"].numel() // (2 * nt * nt), tail,
                          ws["G32"].data_ptr(), ws["G64"].data_ptr(), self.sm, stream)
        else:
            ext.gram(all_rows, [], 0, d, ws["scratch"].data_ptr(), ws["scratch"].numel() // (nt * nt),
                     ws["G32"].data_ptr(), ws["G64"].data_ptr(), self.sm, stream)
        w = plan.solver(ws["G64"])
        ws["W"].copy_(w.reshape(-1).to(torch.float32), non_blocking=True)
        ext.wsum(all_rows, [], ws["W"].data_ptr(), 1, 0, d, [out], [], [], 0.0, 0.0, 0.0, self.sm, stream)

    def _body(self) -> None:
        stream = torch.cuda.current_stream(self.device).cuda_stream
        ctl = self.ctl.data_ptr()
        ext = self.ext
        ext.bump_u32(ctl + 8, stream)
        for i, p in enumerate(self.peers):
            if p.role == "honest":
                self._half_step(i)
        ext.flag_barrier(self._pads, self.rank, ext.PAD_READY, ctl + 8, ctl + 4, stream)
        self._stage_remote(True, stream)
        if self._has_byz:
            for i, p in enumerate(self.peers):
                if p.role != "honest":
                    self._attack(i, stream)
            ext.flag_barrier(self._pads, self.rank, PAD_BYZ, ctl + 8, ctl + 4, stream)
            self._stage_remote(False, stream)
        for i, p in enumerate(self.peers):
            if p.role == "honest":
                self._aggregate(i, stream)
        ext.flag_barrier(self._pads, self.rank, ext.PAD_DONE, ctl + 8, ctl + 4, stream)

    def capture(self, warmup: int = 2) -> None:
        snap = self.params.clone()
        bufs = [[b.clone() for b in p.model.buffers()] if p.model is not None else [] for p in self.peers]
        s = torch.cuda.Stream(self.device)
        s.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(s):
            for _ in range(warmup):
                self._body()
        torch.cuda.current_stream(self.device).wait_stream(s)
        torch.cuda.synchronize(self.device)
        if self.world > 1:
            dist.barrier(group=self.group)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._body()
        self._graph = g
        with torch.no_grad():
            self.params.copy_(snap)
            for p, saved in zip(self.peers, bufs):
                if p.model is not None:
                    for b, sv in zip(p.model.buffers(), saved):
                        b.copy_(sv)
        torch.cuda.synchronize(self.device)

    def step(self, batches: Optional[Sequence[Optional[Tuple[torch.Tensor, torch.Tensor]]]] = None) -> torch.Tensor:
        if batches is None:
            batches = [p.data() if (p.role == "honest" and p.data is not None) else None for p in self.peers]
        for p, b in zip(self.peers, batches):
            if p.role == "honest" and b is not None:
                p.stage_batch(b[0], b[1], self.device)
        if self.use_cuda_graph:
            if self._graph is None:
                self.capture()
            self._graph.replay()
        else:
            self._body()
        return self.losses

    def replay(self) -> torch.Tensor:
        """One more round on the batches already staged on the device (no H2D copy)."""
        if self.use_cuda_graph:
            if self._graph is None:
                self.capture()
            self._graph.replay()
        else:
            self._body()
        return self.losses

    def read_losses(self) -> torch.Tensor:
        """Device->host read of the round's losses (synchronises the stream).  A flag barrier that timed out
        (a peer that stopped taking part) raises here rather than letting a training loop run on stale rows."""
        self.losses_host.copy_(self.losses, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        self.check_status()
        return self.losses_host

    def check_status(self) -> None:
        st = int(self.ctl[1].item())
        if st != 0:
            raise RuntimeError(f"device P2P round reported flag-barrier error {st}")

    def param_vector(self, i: int) -> torch.Tensor:
        return self.params[i][: self.d]

    def close(self) -> None:
        self._graph = None
        self.sym.close()


__all__ = ["DevicePeer", "PeerLayout", "DeviceP2PRound", "PAD_BYZ"]
