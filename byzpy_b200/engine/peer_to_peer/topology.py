"""Directed communication graph over nodes ``0..n-1`` (reference
engine/peer_to_peer/topology.py:13-38).  ``out[i]`` / ``in_[i]`` are the out-/in-neighbour
lists.  ``ring(n, k)`` adds both directions for every offset ``1..k`` (so duplicates appear when
``2k >= n``, as in the reference; routers de-duplicate on broadcast).

On one NVSwitch box every peer is reachable at full NVLink bandwidth, so a topology costs nothing
to route: it only selects WHICH peer buffers a rank's aggregation kernel loads.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Iterable, List, Tuple


@dataclass(frozen=True)
class Edge:
    u: int
    v: int  # directed u -> v


class Topology:
    def __init__(self, n_nodes: int, edges: Iterable[Tuple[int, int]]):
        self.n = int(n_nodes)
        self.out: Dict[int, List[int]] = {i: [] for i in range(self.n)}
        self.in_: Dict[int, List[int]] = {i: [] for i in range(self.n)}
        for u, v in edges:
            if not (0 <= u < self.n and 0 <= v < self.n):
                raise ValueError(f"edge ({u}, {v}) outside 0..{self.n - 1}")
            self.out[u].append(v)
            self.in_[v].append(u)

    @classmethod
    def complete(cls, n: int) -> "Topology":
        return cls(n, ((i, j) for i in range(n) for j in range(n) if i != j))

    @classmethod
    def ring(cls, n: int, k: int = 1) -> "Topology":
        edges = []
        for i in range(n):
            for d in range(1, k + 1):
                edges.append((i, (i + d) % n))
                edges.append((i, (i - d) % n))
        return cls(n, edges)

    def edges(self) -> List[Edge]:
        return [Edge(u, v) for u in range(self.n) for v in self.out[u]]

    def in_neighbors(self, i: int, unique: bool = True) -> List[int]:
        return list(dict.fromkeys(self.in_[i])) if unique else list(self.in_[i])

    def out_neighbors(self, i: int, unique: bool = True) -> List[int]:
        return list(dict.fromkeys(self.out[i])) if unique else list(self.out[i])


__all__ = ["Topology", "Edge"]
