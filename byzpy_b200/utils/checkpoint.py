"""Checkpoint / resume.

The reference has no checkpoint subsystem; the only persisted artefact is the example nodes'
``dump_state_dict()`` -- ``{name: cpu_tensor}`` from ``model.state_dict()`` -- loaded with
``load_state_dict(strict=True)`` (reference examples/ps/nodes.py:127-128,
examples/ps/thread/mnist.py:117-118).  That mapping is kept verbatim as the ``"state_dict"`` entry
of every node record, so a reference checkpoint loads here and vice versa; on top of it a
checkpoint carries what the reference never saves: optimizer (momentum) state, the round counter and
the RNG state.

File = ``torch.save`` of::

    {"format": "byzpy_b200.ckpt.v1", "round": int, "rng": {...},
     "nodes": [{"name", "role", "state_dict": {k: cpu tensor}, "momentum": flat cpu tensor | None,
                "optimizer": optimizer.state_dict() | None}, ...]}

One file per rank (each rank owns its local replicas).
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional

import torch

FORMAT = "byzpy_b200.ckpt.v1"


def _rng_state() -> Dict[str, Any]:
    st = {"torch": torch.get_rng_state()}
    if torch.cuda.is_available():
        st["cuda"] = torch.cuda.get_rng_state_all()
    return st


def _set_rng_state(st: Dict[str, Any]) -> None:
    if "torch" in st:
        torch.set_rng_state(st["torch"])
    if "cuda" in st and torch.cuda.is_available():
        try:
            torch.cuda.set_rng_state_all(st["cuda"])
        except Exception:
            pass


def _node_records(ps) -> List[Dict[str, Any]]:
    recs: List[Dict[str, Any]] = []
    rnd = getattr(ps, "device_round", None)
    if rnd is not None:
        for i, w in enumerate(rnd.workers):
            recs.append({"name": w.name, "role": w.role, "state_dict": w.state_dict_cpu(),
                         "momentum": None if rnd.moms is None else rnd.moms[i].detach().cpu().clone(),
                         "optimizer": None})
        return recs
    for node in list(getattr(ps, "hon", [])) + list(getattr(ps, "byz", [])):
        model = getattr(node, "model", None)
        if model is None:
            continue
        opt = getattr(node, "_opt", None) or getattr(node, "optimizer", None)
        recs.append({"name": getattr(node, "name", type(node).__name__),
                     "role": "honest" if node in getattr(ps, "hon", []) else "byzantine",
                     "state_dict": {k: v.detach().cpu() for k, v in model.state_dict().items()},
                     "momentum": None, "optimizer": None if opt is None else opt.state_dict()})
    return recs


def save_checkpoint(path: str, ps, *, round_index: Optional[int] = None, extra: Optional[dict] = None) -> None:
    """Write models, momentum / optimizer state and RNG state of every node of the parameter server ``ps`` to ``path``
    (one ``torch.save`` blob that :func:`load_checkpoint` reads with the restricted unpickler).
    """
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    blob = {"format": FORMAT, "round": int(round_index if round_index is not None else getattr(ps, "rounds", 0)),
            "rng": _rng_state(), "nodes": _node_records(ps), "extra": dict(extra or {})}
    torch.save(blob, path)


def load_checkpoint(path: str, ps, *, strict: bool = True, restore_rng: bool = True) -> int:
    """Restores models, momentum/optimizer state and RNG; returns the saved round index."""
    # the blob holds only tensors, dicts, lists and scalars: the restricted unpickler is enough, and a
    # checkpoint file obtained from elsewhere cannot run code when loaded
    blob = torch.load(path, map_location="cpu", weights_only=True)
    if blob.get("format") != FORMAT:
        raise ValueError(f"not a {FORMAT} checkpoint: {blob.get('format')!r}")
    nodes = blob["nodes"]
    rnd = getattr(ps, "device_round", None)
    if rnd is not None:
        if len(nodes) != len(rnd.workers):
            raise ValueError("checkpoint has a different number of local replicas")
        with torch.no_grad():
            for i, (w, rec) in enumerate(zip(rnd.workers, nodes)):
                # parameters are views into the flat arena: load_state_dict copies in place
                w.model.load_state_dict(rec["state_dict"], strict=strict)
                if rnd.moms is not None and rec.get("momentum") is not None:
                    rnd.moms[i].copy_(rec["momentum"].to(rnd.moms.device))
                if not w.arena.check_bound():
                    raise RuntimeError("parameters detached from the flat arena while loading")
    else:
        targets = [n for n in list(getattr(ps, "hon", [])) + list(getattr(ps, "byz", []))
                   if getattr(n, "model", None) is not None]
        if len(nodes) != len(targets):
            raise ValueError("checkpoint has a different number of nodes")
        for node, rec in zip(targets, nodes):
            node.model.load_state_dict(rec["state_dict"], strict=strict)
            opt = getattr(node, "_opt", None) or getattr(node, "optimizer", None)
            if opt is None and rec.get("optimizer") is not None and hasattr(node, "ensure_optimizer"):
                opt = node.ensure_optimizer()      # lazily-built optimizers: momentum must not be dropped on resume
            if opt is not None and rec.get("optimizer") is not None:
                opt.load_state_dict(rec["optimizer"])
    if restore_rng:
        _set_rng_state(blob.get("rng", {}))
    if hasattr(ps, "rounds"):
        ps.rounds = int(blob["round"])
    return int(blob["round"])


def load_reference_state_dict(model: torch.nn.Module, state_dict: Dict[str, torch.Tensor], strict: bool = True):
    """Load a reference-format snapshot (``dump_state_dict()`` mapping) into ``model``."""
    return model.load_state_dict(state_dict, strict=strict)


__all__ = ["save_checkpoint", "load_checkpoint", "load_reference_state_dict", "FORMAT"]
