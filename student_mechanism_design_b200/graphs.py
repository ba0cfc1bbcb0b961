"""Host-side graph pool: generation, CSR packing and the static observation tensors.

`generate_connected_graph` draws from the SAME DISTRIBUTION as the reference's
ConnectedGraph.sample (/root/reference/src/environment/graph_layout.py:9-80) -- a random
recursive tree over a uniformly random insertion order, then uniformly shuffled extra edges
under the degree cap (which the tree ignores), weights U{1..4} -- but in O(N^2) numpy instead
of the reference's O(N^3) Python, and from its own `numpy.random.Generator` stream.  Parity
runs hand over the reference's own graphs instead (see tests).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Sequence, Tuple

import numpy as np


@dataclass
class GraphSpec:
    """Same fields as the reference's GraphInstance (graph_layout.py:53)."""

    num_nodes: int
    edge_links: np.ndarray  # int32 [E, 2]
    edges: np.ndarray  # int64 [E]  weights

    def __post_init__(self):
        self.edge_links = np.ascontiguousarray(np.asarray(self.edge_links, dtype=np.int32).reshape(-1, 2))
        self.edges = np.ascontiguousarray(np.asarray(self.edges, dtype=np.int64).reshape(-1))
        if len(self.edges) != len(self.edge_links):
            raise ValueError("edges and edge_links disagree")

    @property
    def nodes(self) -> np.ndarray:
        return np.arange(self.num_nodes, dtype=np.int64)


def generate_connected_graph(num_nodes: int, num_edges: int | None, rng: np.random.Generator,
                             max_edges_per_node: int = 4, max_weight: int = 5) -> GraphSpec:
    n = int(num_nodes)
    order = rng.permutation(n)
    # graph_layout.py:55-80: every new node attaches to a uniformly random already-visited node
    parents = (rng.random(n - 1) * np.arange(1, n)).astype(np.int64)
    links = np.stack([order[parents], order[1:]], axis=1).astype(np.int64)
    if num_edges is None:
        num_edges = n - 1
    extra = int(num_edges) - (n - 1)
    if extra > 0:
        deg = np.bincount(links.ravel(), minlength=n)
        iu, ju = np.triu_indices(n, k=1)
        is_tree = np.zeros((n, n), dtype=bool)
        is_tree[links[:, 0], links[:, 1]] = True
        is_tree[links[:, 1], links[:, 0]] = True
        keep = ~is_tree[iu, ju]
        cand = np.stack([iu[keep], ju[keep]], axis=1)
        cand = cand[rng.permutation(len(cand))]  # graph_layout.py:31
        added = []
        for a, b in cand.tolist():  # graph_layout.py:34-48: first-fit under the degree cap
            if extra <= 0:
                break
            if deg[a] < max_edges_per_node and deg[b] < max_edges_per_node:
                added.append((a, b))
                deg[a] += 1
                deg[b] += 1
                extra -= 1
        if added:
            links = np.concatenate([links, np.asarray(added, dtype=np.int64)], axis=0)
    weights = rng.integers(1, max_weight, size=len(links))  # graph_layout.py:17,47 -> {1..4}
    return GraphSpec(n, links.astype(np.int32), weights.astype(np.int64))


def generate_graph_pool(num_graphs: int, num_nodes: int, num_edges: int | None, seed: int = 0) -> List[GraphSpec]:
    """`num_graphs` graphs with one common edge count (the reference freezes the achievable
    count at construction and resamples until it is met, yard.py:67-101)."""
    rng = np.random.default_rng(seed)
    first = generate_connected_graph(num_nodes, num_edges, rng)
    want = len(first.edges)
    pool = [first]
    attempts = 0
    while len(pool) < num_graphs:
        g = generate_connected_graph(num_nodes, num_edges, rng)
        attempts += 1
        if len(g.edges) == want:
            pool.append(g)
            attempts = 0
        elif attempts >= 100:
            raise RuntimeError(f"Failed to generate graph with {want} edges after 100 attempts.")  # yard.py:96-101
    return pool


def pack_csr(graphs: Sequence[GraphSpec]) -> Tuple[np.ndarray, np.ndarray, np.ndarray, int]:
    """(row_ptr int32[G,N+1], col int32[G,S], w int32[G,S], S): both directions of every edge,
    neighbours ascending, parallel edges collapsed to their minimum weight (yard.py:454-465)."""
    n = graphs[0].num_nodes
    rows = []
    for g in graphs:
        if g.num_nodes != n:
            raise ValueError("all graphs in a pool must have the same number of nodes")
        u = np.concatenate([g.edge_links[:, 0], g.edge_links[:, 1]]).astype(np.int64)
        v = np.concatenate([g.edge_links[:, 1], g.edge_links[:, 0]]).astype(np.int64)
        w = np.concatenate([g.edges, g.edges]).astype(np.int64)
        if len(u) and (u.min() < 0 or u.max() >= n or v.min() < 0 or v.max() >= n or np.any(u == v)):
            raise ValueError("edge endpoint out of range or self-loop")
        order = np.lexsort((w, v, u))
        u, v, w = u[order], v[order], w[order]
        first = np.ones(len(u), dtype=bool)
        first[1:] = (u[1:] != u[:-1]) | (v[1:] != v[:-1])
        u, v, w = u[first], v[first], w[first]
        rp = np.zeros(n + 1, dtype=np.int32)
        np.cumsum(np.bincount(u, minlength=n), out=rp[1:])
        rows.append((rp, v.astype(np.int32), w.astype(np.int32)))
    stride = max(1, max(len(r[1]) for r in rows))
    G = len(rows)
    row_ptr = np.zeros((G, n + 1), dtype=np.int32)
    col = np.zeros((G, stride), dtype=np.int32)
    wgt = np.zeros((G, stride), dtype=np.int32)
    for i, (rp, c, w) in enumerate(rows):
        row_ptr[i] = rp
        col[i, : len(c)] = c
        wgt[i, : len(w)] = w
    return row_ptr, col, wgt, stride
