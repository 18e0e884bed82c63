"""Synchronous Byzantine-robust parameter server.

Same constructor and ``round()/shutdown()`` contract as the reference
(reference engine/parameter_server/ps.py:18-158): collect honest gradients, let every
Byzantine node see them and emit its vector (appended after the honest ones), optional
pre-aggregation, robust aggregation, fan the aggregate out to the honest nodes (and the
Byzantine ones if ``update_byzantines``).

Two execution paths:

* **device path** -- when every node is a device node (``DeviceHonestNode`` /
  ``DeviceByzantineNode``) on CUDA and the aggregator has a fused plan, the whole round is
  :class:`byzpy_b200.parallel.device_ps.DeviceRound`: fwd/bwd per replica, then ONE fused
  sm_100a kernel doing gather (P2P loads over NVLink) + aggregation + broadcast + SGD, replayed
  as a CUDA graph.  Multi-GPU: one process per GPU, each constructs a ParameterServer over its
  LOCAL nodes and passes ``layout=RowLayout.block(...)``; ``round()`` is then collective.
* **generic path** -- arbitrary node actors (thread / process / tcp backends): gradients are
  gathered through actor RPCs like the reference, in *submission* order (deterministic row
  order; the reference uses completion order).  Aggregators accept
  tensors and shm handles alike, so -- unlike the reference at this commit --
  every aggregator works here, with or without an ``actor_pool``.
"""
from __future__ import annotations

import asyncio
import inspect
import logging
import os
import time
from typing import Any, List, Optional, Sequence

import torch

from ...aggregators.base import Aggregator
from ...pre_aggregators.base import PreAggregator
from ...utils import metrics


async def _call(obj: Any, method: str, *args, **kwargs) -> Any:
    res = getattr(obj, method)(*args, **kwargs)
    if inspect.isawaitable(res):
        res = await res
    return res


_log = logging.getLogger(__name__)


def _is_device_node(node: Any) -> bool:
    try:
        from ..node.device import DeviceByzantineNode, DeviceHonestNode
    except Exception:  # pragma: no cover
        return False
    return isinstance(node, (DeviceHonestNode, DeviceByzantineNode))


class ParameterServer:
    """Synchronous robust training with a central aggregator.

    One :meth:`round`: every honest node computes a gradient on its next batch; every Byzantine node produces its vector
    after seeing the honest ones (omniscient adversary); the optional pre-aggregator and then the aggregator reduce the
    ``n`` vectors to one; the result is sent back to the honest nodes (and to the Byzantine ones when
    ``update_byzantines``), which apply it with their own optimizer.

    Parameters
    ----------
    honest_nodes, byzantine_nodes : list
        Node objects or node actors (:class:`~byzpy_b200.engine.node.actors.HonestNodeActor` /
        ``ByzantineNodeActor``); on the multi-GPU device path each process passes the nodes it hosts.
    aggregator : Aggregator
    pre_aggregator : PreAggregator, optional
    update_byzantines : bool, default False
    actor_pool : ActorPool, optional
        Run the aggregation through a :class:`~byzpy_b200.engine.graph.scheduler.NodeScheduler` on this pool
        (generic path only).
    scheduler_metadata : dict, optional
    node_timeout : float, optional
        Seconds a node may take to answer before it counts as failed for the round.
    tolerate_failures : bool, default False
        Skip nodes that raise or time out (recorded in ``failed`` as ``(round, "honest:<i>", reason)``) instead of
        propagating the error.
    fused : bool, optional
        ``None``: use the fused device round when every node is a device node on CUDA; ``True``: require it;
        ``False``: always take the generic, actor-driven path.
    layout, process_group, lr, momentum, weight_decay, amp_dtype, use_cuda_graph, worker_streams, direct_grads,
    overlap_wgrad, branch_streams, buckets, multicast, device_options :
        Device-path configuration, see :class:`byzpy_b200.parallel.device_ps.DeviceRound` and
        ``docs/source/device_path.md``.  ``layout=RowLayout(...)`` says which worker rows each GPU hosts;
        ``buckets`` cuts the gradient into buckets that are aggregated while backward is still running;
        ``multicast`` selects NVLS multicast delivery of the aggregate.

    Notes
    -----
    Device path: all replicas of a GPU train in one captured graph, gradients land in a symmetric-memory arena, and
    aggregation, the attack's row folds, the optimizer step and the delivery of the update happen in one kernel per
    bucket reading peer memory over NVLink (``csrc/fused_ps.cu``).  :meth:`step` enqueues a round without waiting,
    :meth:`round` additionally returns the aggregate, :meth:`recover` drops a rank that stopped answering and carries
    on with the rest.  Generic path: gradients are collected in node order (the reference: completion order), so
    order-sensitive operators see the same worker in the same row every round.

    Examples
    --------
    >>> import asyncio, torch
    >>> from byzpy_b200.aggregators.coordinate_wise import CoordinateWiseMedian
    >>> from byzpy_b200.engine.parameter_server.ps import ParameterServer
    >>> class Honest:
    ...     def __init__(self, g):
    ...         self.g, self.seen = torch.tensor(g), None
    ...     def honest_gradient_for_next_batch(self):
    ...         return self.g
    ...     def apply_server_gradient(self, g):
    ...         self.seen = g
    >>> class Liar:
    ...     def byzantine_gradient_for_next_batch(self, honest_grads=None):
    ...         return -100.0 * torch.stack(list(honest_grads)).mean(0)
    ...     def apply_server_gradient(self, g):
    ...         pass
    >>> honest = [Honest([1.0, 1.0]), Honest([2.0, 0.0]), Honest([3.0, 2.0])]
    >>> ps = ParameterServer(honest, [Liar()], CoordinateWiseMedian())
    >>> asyncio.run(ps.round())
    tensor([1., 0.])
    >>> honest[0].seen
    tensor([1., 0.])
    """

    def __init__(self, honest_nodes: List[Any], byzantine_nodes: List[Any], aggregator: Aggregator,
                 pre_aggregator: Optional[PreAggregator] = None, update_byzantines: bool = False, *,
                 actor_pool=None, scheduler_metadata: Optional[dict] = None, layout=None,
                 process_group=None, lr: Optional[float] = None, momentum: Optional[float] = None,
                 weight_decay: Optional[float] = None, amp_dtype: Optional[torch.dtype] = torch.bfloat16,
                 use_cuda_graph: bool = True, worker_streams: int = 1, fused: Optional[bool] = None,
                 node_timeout: Optional[float] = None, tolerate_failures: bool = False,
                 direct_grads: bool = True, overlap_wgrad: bool = True, branch_streams: bool = True,
                 buckets: Optional[int] = None, multicast: Optional[bool] = None,
                 device_options: Optional[dict] = None):
        # ``buckets``: gradient buckets of the fused round (None = automatic, 1 = one launch after
        # backward); ``multicast``: NVLS multicast broadcast (None = when the heap supports it);
        # ``device_options``: further DeviceRound keyword arguments (bucket_cuts, spin_seconds, ...)
        self._device_opts = dict(direct_grads=direct_grads, overlap_wgrad=overlap_wgrad,
                                 branch_streams=branch_streams, buckets=buckets, multicast=multicast,
                                 **(device_options or {}))
        self.hon = list(honest_nodes)
        self.byz = list(byzantine_nodes)
        self.agg = aggregator
        self.pre = pre_aggregator
        self.update_byz = update_byzantines
        self.pool = actor_pool
        self.scheduler = None
        self.rounds = 0
        # failure detection (the reference has none: a hung actor hangs the round):
        # a node that raises or exceeds ``node_timeout`` is treated as silent for that round when
        # ``tolerate_failures`` is set; ``failed`` records (round, node index, reason).
        self.node_timeout = node_timeout
        self.tolerate_failures = tolerate_failures
        self.failed: List[tuple] = []
        if actor_pool is not None:
            from ..graph.ops import make_single_operator_graph
            from ..graph.scheduler import NodeScheduler

            graph = make_single_operator_graph(node_name="agg", operator=self.agg,
                                               input_keys=("gradients",))
            self.scheduler = NodeScheduler(graph, pool=actor_pool, metadata=scheduler_metadata)
        self.device_round = None
        want_fused = fused if fused is not None else True
        # the pre-aggregator -> coordinate-wise fused round has not been timed on hardware yet: automatic selection
        # (fused=None) keeps to the measured plans, fused=True or BYZPY_FUSED_MAPCW=1 asks for it explicitly
        self._allow_mapcw = bool(fused) or os.environ.get("BYZPY_FUSED_MAPCW", "0") not in ("", "0")
        if want_fused and actor_pool is None:
            self.device_round = self._try_build_device_round(
                layout, process_group, lr, momentum, weight_decay, amp_dtype, use_cuda_graph,
                worker_streams)
        if fused and self.device_round is None:
            raise RuntimeError("fused=True requested but the configuration has no fused device path")

    # ------------------------------------------------------------------ device path
    def _try_build_device_round(self, layout, group, lr, momentum, weight_decay, amp_dtype,
                                use_cuda_graph, worker_streams):
        nodes = self.hon + self.byz
        if not nodes:
            # a rank that hosts no replica (RowLayout.spread) still joins the fused aggregation
            if layout is None or not torch.cuda.is_available() or lr is None:
                return None
        elif not all(_is_device_node(n) for n in nodes):
            return None
        if not torch.cuda.is_available() or any(n.device.type != "cuda" for n in nodes):
            return None
        from ...parallel.device_ps import DeviceRound, RowFold, RowLayout

        import torch.distributed as dist

        world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        workers = [n.worker for n in self.hon]
        virtual_fold = None
        n_virtual_local = 0
        for b in self.byz:
            fold = b.fold(0)
            if b.worker is not None:
                if fold is None or fold.kind != "scale":
                    return None
                b.worker.fold = fold
                workers.append(b.worker)
            else:
                if fold is None or fold.kind != "virtual":
                    return None
                n_virtual_local += 1
        if layout is None:
            if world != 1:
                raise ValueError("multi-rank ParameterServer needs an explicit layout=RowLayout...")
            layout = RowLayout.block(len(self.hon), len(workers) - len(self.hon), 1, n_virtual_local)
        if layout.n_virtual:
            virt = [b for b in self.byz if b.worker is None]
            if not virt:
                return None
            f0 = virt[0].attack.fold(layout.n_honest)
            virtual_fold = RowFold("virtual", a=f0.a, b=f0.b)
        n_rows = layout.n_workers + layout.n_virtual
        plan = self._fused_plan(n_rows)
        if plan is None:
            return None
        first = self.hon[0] if self.hon else (self.byz[0] if self.byz else None)
        return DeviceRound(
            workers, layout, plan,
            lr=first.lr if lr is None else lr,
            momentum=(first.momentum if first is not None else 0.0) if momentum is None else momentum,
            weight_decay=(first.weight_decay if first is not None else 0.0) if weight_decay is None else weight_decay,
            update_byzantines=self.update_byz,
            device=first.device if first is not None else torch.device("cuda", torch.cuda.current_device()),
            group=group,
            amp_dtype=amp_dtype, use_cuda_graph=use_cuda_graph, worker_streams=worker_streams,
            virtual_fold=virtual_fold, **self._device_opts)

    def _fused_plan(self, n_rows: int):
        """Aggregator plan, composed with a linear pre-aggregator when present.  Gram-family aggregator:
        in n-space, ``X' = W_p X``  =>  ``G' = W_p G W_p^T`` and final weights ``w_agg @ W_p`` -- the
        pre-aggregated vectors are never materialised.  Coordinate-wise aggregator: a
        :class:`MapCwPlan` (the mixed rows exist per coordinate shard only)."""
        if self.pre is None:
            return self.agg.fused_plan(n_rows)
        from ...parallel.device_ps import CwPlan, GramPlan, MapCwPlan
        from ...pre_aggregators.base import LinearPreAggregator

        if not isinstance(self.pre, LinearPreAggregator):
            return None
        import numpy as np
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1 \
                and getattr(self.pre, "perm", 0) is None and hasattr(self.pre, "rng"):
            # every rank solves the n-space problem itself, so a randomised map (Bucketing without a
            # fixed ``perm``) must draw the SAME permutation on every rank: share one seed
            import random

            box = [random.SystemRandom().getrandbits(63)]
            dist.broadcast_object_list(box, src=0)
            self.pre.rng = random.Random(box[0])

        pre = self.pre
        pre._validate(n_rows)
        probe = pre.row_map(np.eye(n_rows) if pre.needs_gram else None, n_rows)
        m = probe.shape[0]
        inner = self.agg.fused_plan(m)
        state = {"W": None}

        def refresh():
            if not pre.needs_gram:
                Wn = np.asarray(pre.row_map(None, n_rows), dtype=np.float64)
                if state["W"] is None:
                    state["W"] = torch.from_numpy(Wn).to(self.hon[0].device if self.hon else "cuda")
                else:
                    state["W"].copy_(torch.from_numpy(Wn), non_blocking=True)

        device_map = type(pre).row_map_device is not LinearPreAggregator.row_map_device

        if isinstance(inner, CwPlan):
            if not getattr(self, "_allow_mapcw", True):
                return None
            # map -> coordinate-wise selection is not linear in n-space: the m mixed rows are produced on
            # every rank's coordinate shard and the fused selection kernel runs over them (MapCwPlan)
            def weights(G: Optional[torch.Tensor]) -> torch.Tensor:
                if pre.needs_gram:
                    Wp = pre.row_map_device(G.double(), n_rows) if (device_map and G.is_cuda) else None
                    if Wp is None:
                        Wp = torch.from_numpy(np.asarray(pre.row_map(G.detach().double().cpu().numpy(), n_rows),
                                                         dtype=np.float64)).to(G.device)
                    return Wp
                if state["W"] is None:
                    refresh()
                return state["W"]

            return MapCwPlan(inner, m, weights, needs_gram=bool(pre.needs_gram), name=f"{pre.name}+{self.agg.name}",
                             capturable=device_map or not pre.needs_gram,
                             refresh=None if pre.needs_gram else refresh)
        if not isinstance(inner, GramPlan) or inner.aux:
            return None

        def solver(G: torch.Tensor) -> torch.Tensor:
            if pre.needs_gram:
                # the map of Clipping / ARC / NNM is a single-CTA kernel on the device Gram
                # (csrc/nspace_maps.cu): no host round trip, the round stays CUDA-graph capturable
                Wp = pre.row_map_device(G.double(), n_rows) if (device_map and G.is_cuda) else None
                if Wp is None:
                    Wp = torch.from_numpy(np.asarray(pre.row_map(G.detach().double().cpu().numpy(), n_rows),
                                                     dtype=np.float64)).to(G.device)
            else:
                if state["W"] is None:
                    refresh()
                Wp = state["W"]
            G2 = Wp @ G.double() @ Wp.T
            w2 = inner.solver(G2)
            return (w2.double() @ Wp).float()

        return GramPlan(solver, f"{pre.name}+{inner.name}", aux=(),
                        capturable=inner.capturable and (device_map or not pre.needs_gram),
                        refresh=None if pre.needs_gram else refresh)

    def step(self, batches=None) -> torch.Tensor:
        """Device path: one fused round, asynchronous on the current CUDA stream.
        Returns the device tensor of local per-worker losses."""
        if self.device_round is None:
            raise RuntimeError("step() is only available on the fused device path; use round()")
        self.rounds += 1
        metrics.inc("byzpy_ps_rounds_total", labels={"path": "device"})
        return self.device_round.step(batches)

    def recover(self) -> List[int]:
        """Device path: after a fused round timed out on a silent rank (``read_losses`` raised), drop
        that rank's rows, rebuild the aggregation plan for the remaining ones and resynchronise the
        replicas (:meth:`DeviceRound.recover`).  Returns the dropped ranks.  The generic path handles
        failures per call (``node_timeout`` / ``tolerate_failures``)."""
        if self.device_round is None:
            raise RuntimeError("recover() applies to the fused device path")
        return self.device_round.recover(self._fused_plan)

    # ------------------------------------------------------------------- generic path
    async def _guarded(self, kind: str, idx: int, node: Any, method: str, *args):
        try:
            coro = _call(node, method, *args)
            if self.node_timeout is not None:
                return await asyncio.wait_for(coro, timeout=self.node_timeout)
            return await coro
        except Exception as exc:  # noqa: BLE001
            if not self.tolerate_failures:
                raise
            self.failed.append((self.rounds, f"{kind}:{idx}", repr(exc)))
            metrics.inc("byzpy_ps_node_failures_total", labels={"kind": kind})
            _log.warning("round %d: %s node %d skipped (%r)", self.rounds, kind, idx, exc)
            return None

    async def _gather_honest(self) -> List[torch.Tensor]:
        got = await asyncio.gather(*[self._guarded("honest", i, h, "honest_gradient_for_next_batch")
                                     for i, h in enumerate(self.hon)])
        return [g for g in got if g is not None]

    async def _gather_byzantine(self, honest: Sequence[torch.Tensor]) -> List[torch.Tensor]:
        if not self.byz:
            return []
        got = await asyncio.gather(*[self._guarded("byzantine", i, b, "byzantine_gradient_for_next_batch", honest)
                                     for i, b in enumerate(self.byz)])
        return [g for g in got if g is not None]

    async def round(self) -> torch.Tensor:
        """One training round; returns the aggregated gradient (device path: the device tensor of the aggregate after
        enqueuing the round).
        """
        if self.device_round is not None:
            self.step()
            return self.device_round.aggregated()
        t_round = time.perf_counter()
        grads = await self._gather_honest()
        grads += await self._gather_byzantine(tuple(grads))
        if not grads:
            raise RuntimeError("no gradients were collected this round (all nodes failed)")
        if self.pre is not None:
            grads = list(self.pre.pre_aggregate(grads))
        if self.scheduler is not None:
            g = (await self.scheduler.run({"gradients": grads}))["agg"]
        else:
            g = self.agg.aggregate(grads)
        if not grads:
            raise RuntimeError("no gradients were collected this round (all nodes failed)")
        targets = self.hon + (self.byz if self.update_byz else [])
        await asyncio.gather(*[self._guarded("apply", i, n, "apply_server_gradient", g)
                               for i, n in enumerate(targets)])
        self.rounds += 1
        metrics.inc("byzpy_ps_rounds_total", labels={"path": "generic"})
        metrics.observe("byzpy_ps_round_seconds", time.perf_counter() - t_round)
        return g

    def round_sync(self) -> torch.Tensor:
        """:meth:`round` for synchronous callers (spins an event loop)."""
        return asyncio.run(self.round())

    async def shutdown(self) -> None:
        """Release the device round's symmetric memory and close the node actors' backends."""
        if self.device_round is not None:
            self.device_round.close()
            self.device_round = None
        closers = []
        for n in self.hon + self.byz:
            ref = getattr(n, "_ref", None)
            backend = getattr(ref, "_backend", None)
            if backend is not None:
                closers.append(backend.close())
        if closers:
            await asyncio.gather(*closers)


__all__ = ["ParameterServer"]
