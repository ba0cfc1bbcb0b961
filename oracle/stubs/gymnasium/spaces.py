"""Minimal `gymnasium.spaces` stand-in used only to import the unmodified reference env.

The reference relies on exactly one behaviour: `Graph.from_jsonable` turning the
json-able sample into a `GraphInstance` of numpy arrays whose dtypes come from the
node/edge spaces (reference: src/environment/graph_layout.py:53, yard.py:34-36).
"""
from typing import Any, NamedTuple

import numpy as np


class GraphInstance(NamedTuple):
    nodes: Any
    edges: Any
    edge_links: Any


class _Space:
    def __init__(self, *args, **kwargs):
        self.args, self.kwargs = args, kwargs


class Discrete(_Space):
    dtype = np.int64

    def __init__(self, n, start=0):
        self.n, self.start = n, start


class MultiDiscrete(_Space):
    pass


class Box(_Space):
    pass


class Dict(_Space):
    pass


class MultiBinary(_Space):
    pass


class Graph(_Space):
    def __init__(self, node_space, edge_space, seed=None):
        self.node_space, self.edge_space = node_space, edge_space

    def from_jsonable(self, sample_n):
        return [
            GraphInstance(
                np.asarray(s["nodes"], dtype=self.node_space.dtype),
                np.asarray(s["edges"], dtype=self.edge_space.dtype),
                np.asarray(s["edge_links"], dtype=np.int32),
            )
            for s in sample_n
        ]
