"""Directed communication graph over nodes ``0 .. n-1`` (reference
engine/peer_to_peer/topology.py:13-38).

``out[i]`` / ``in_[i]`` list node i's out- / in-neighbours in insertion order.  ``ring(n, k)`` links
every node to its ``k`` successors and ``k`` predecessors, so an edge appears twice once ``2k >= n``
-- the reference keeps those duplicates and lets the routers de-duplicate; ``in_neighbors`` /
``out_neighbors`` do that here.

On one NVSwitch box every peer is reachable at full NVLink bandwidth, so a topology costs nothing
to route: it only selects WHICH peer buffers a rank's aggregation kernel loads.
"""
from __future__ import annotations

from typing import Dict, Iterable, List, NamedTuple, Tuple


class Edge(NamedTuple):
    """Directed edge ``u -> v``."""

    u: int
    v: int


def _unique(seq: List[int]) -> List[int]:
    return list(dict.fromkeys(seq))


class Topology:
    """Directed communication graph over nodes ``0 .. n-1``: who sends its model to whom in a gossip round.

    Parameters
    ----------
    n_nodes : int
    edges : iterable of (int, int)
        Directed edges ``(source, destination)``; an undirected link is two edges.

    Notes
    -----
    ``Topology.complete(n)`` links every ordered pair, ``Topology.ring(n, k=1)`` every node with its ``k`` nearest
    neighbours on both sides.  ``out_neighbors(i)`` / ``in_neighbors(i)`` list a node's peers (duplicates removed unless
    ``unique=False``), ``edges()`` all edges, ``out`` / ``in_`` the raw adjacency lists.

    Examples
    --------
    >>> from byzpy_b200.engine.peer_to_peer.topology import Topology
    >>> Topology.ring(5, 1).out_neighbors(0)
    [1, 4]
    >>> Topology.complete(3).in_neighbors(2)
    [0, 1]
    >>> Topology(3, [(0, 1), (1, 2)]).edges()
    [Edge(u=0, v=1), Edge(u=1, v=2)]
    """

    def __init__(self, n_nodes: int, edges: Iterable[Tuple[int, int]]):
        self.n = n = int(n_nodes)
        self.out: Dict[int, List[int]] = {i: [] for i in range(n)}
        self.in_: Dict[int, List[int]] = {i: [] for i in range(n)}
        for src, dst in edges:
            if min(src, dst) < 0 or max(src, dst) >= n:
                raise ValueError(f"edge ({src}, {dst}) outside 0..{n - 1}")
            self.out[src].append(dst)
            self.in_[dst].append(src)

    # -- constructors -----------------------------------------------------------------------
    @classmethod
    def complete(cls, n: int) -> "Topology":
        """Every ordered pair ``i != j``."""
        return cls(n, [(i, j) for i in range(n) for j in range(n) if j != i])

    @classmethod
    def ring(cls, n: int, k: int = 1) -> "Topology":
        """Each node talks to its ``k`` nearest neighbours on both sides of a cycle."""
        pairs: List[Tuple[int, int]] = []
        for i in range(n):
            for step in range(1, k + 1):
                pairs += [(i, (i + step) % n), (i, (i - step) % n)]
        return cls(n, pairs)

    # -- queries ----------------------------------------------------------------------------
    def edges(self) -> List[Edge]:
        """All directed edges, grouped by source in insertion order."""
        return [Edge(u, v) for u, targets in self.out.items() for v in targets]

    def in_neighbors(self, i: int, unique: bool = True) -> List[int]:
        """Nodes that send to ``i`` (duplicates removed unless ``unique=False``)."""
        return _unique(self.in_[i]) if unique else list(self.in_[i])

    def out_neighbors(self, i: int, unique: bool = True) -> List[int]:
        """Nodes ``i`` sends to (duplicates removed unless ``unique=False``)."""
        return _unique(self.out[i]) if unique else list(self.out[i])


__all__ = ["Topology", "Edge"]
