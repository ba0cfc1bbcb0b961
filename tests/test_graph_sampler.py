"""the counter-based (Philox) graph sampler the device generator implements draws from the same
distribution as the reference's ConnectedGraph.sample (graph_layout.py:9-80).

`sample_connected_graph` in the oracle is the draw-for-draw restatement of the reference sampler (pinned against the
unmodified reference in test_oracle_golden.py); `philox_sample_graph_once` is the restatement of the device kernel
(bit-exact against it in the GPU tests).  Here the two are compared as distributions: exactly (labelled-graph
frequencies on tiny graphs, chi-square) and through summary statistics on larger ones.  Seeds are fixed, so the
outcome is deterministic."""
import random
from collections import Counter

import numpy as np
import pytest
from scipy import stats

from oracle import sy_oracle as so


def _ref_samples(n, N, E, cap, seed):
    pr, nr = random.Random(seed), np.random.RandomState(seed)
    return [so.sample_connected_graph(N, E, pr, nr, cap) for _ in range(n)]


def _new_samples(n, N, E, cap, seed):
    return [so.philox_sample_graph_once(seed, g, 0, 0, N, E, cap) for g in range(n)]


def _label(g):
    W = g.weight_matrix() > 0
    iu = np.triu_indices(g.num_nodes, 1)
    return int(np.packbits(W[iu]).tobytes().hex(), 16)


def _chi2_two_sample(ca: Counter, cb: Counter, min_expected=8):
    keys = sorted(set(ca) | set(cb))
    a = np.asarray([ca.get(k, 0) for k in keys], dtype=np.float64)
    b = np.asarray([cb.get(k, 0) for k in keys], dtype=np.float64)
    big = (a + b) >= 2 * min_expected  # pool the rare categories
    a = np.append(a[big], a[~big].sum())
    b = np.append(b[big], b[~big].sum())
    keep = (a + b) > 0
    return stats.chi2_contingency(np.stack([a[keep], b[keep]]))[1]


@pytest.mark.parametrize("N,E,cap", [(4, 5, 3), (5, 7, 3), (5, 6, 2), (6, 9, 4)])
def test_labelled_graph_frequencies_match(N, E, cap):
    n = 12000
    ca = Counter(_label(g) for g in _ref_samples(n, N, E, cap, 11))
    cb = Counter(_label(g) for g in _new_samples(n, N, E, cap, 12))
    assert set(cb) <= set(ca) or len(ca) > 500  # never a graph the reference cannot produce (small supports only)
    p = _chi2_two_sample(ca, cb)
    assert p > 1e-3, p


@pytest.mark.parametrize("N,E,cap", [(15, 20, 4), (30, 58, 4), (40, 70, 3)])
def test_summary_statistics_match(N, E, cap):
    n = 1500
    ref, new = _ref_samples(n, N, E, cap, 3), _new_samples(n, N, E, cap, 4)

    def summarise(gs):
        out = dict(edges=Counter(), deg=Counter(), tree_deg=Counter(), weight=Counter(), maxdeg=Counter(), diam=Counter())
        for g in gs:
            W = g.weight_matrix()
            d = (W > 0).sum(1)
            out["edges"][len(g.edges)] += 1
            out["deg"].update(d.tolist())
            out["tree_deg"].update(np.bincount(g.edge_links[: g.num_nodes - 1].ravel(), minlength=g.num_nodes).tolist())
            out["weight"].update(g.edges.tolist())
            out["maxdeg"][int(d.max())] += 1
            out["diam"][int(g.apsp().max())] += 1
        return out

    a, b = summarise(ref), summarise(new)
    for k in a:
        p = _chi2_two_sample(a[k], b[k])
        assert p > 1e-3, (k, p)
    assert set(b["weight"]) == {1, 2, 3, 4}


def test_structure_and_retry_rule():
    for g in range(20):
        gr = so.philox_sample_graph_once(9, g, 1, 0, 50, 110)
        W = gr.weight_matrix()
        assert (W == W.T).all() and (np.diag(W) == 0).all()
        assert (gr.apsp() < so.INF_U16).all(), "not connected"
        assert len(gr.edges) <= 110 and len({tuple(sorted(e)) for e in gr.edge_links.tolist()}) == len(gr.edges)
        extra_deg = np.bincount(gr.edge_links[49:].ravel(), minlength=50)
        tree_deg = np.bincount(gr.edge_links[:49].ravel(), minlength=50)
        assert ((tree_deg + extra_deg)[extra_deg > 0] <= 4).all()  # the cap binds the extra edges only
    pool = so.philox_graph_pool(9, 12, 40, 80, generation=2)
    assert len({len(g.edges) for g in pool}) == 1  # yard.py:87-101: one edge count per pool
    again = so.philox_graph_pool(9, 12, 40, 80, generation=2)
    assert all((a.edge_links == b.edge_links).all() and (a.edges == b.edges).all() for a, b in zip(pool, again))
    shard = so.philox_graph_pool(9, 6, 40, 80, generation=2, graph_offset=6)  # second half of the same pool
    assert all((a.edge_links == b.edge_links).all() for a, b in zip(pool[6:], shard))
    other = so.philox_graph_pool(9, 12, 40, 80, generation=3)
    assert any((a.edge_links.shape != b.edge_links.shape) or (a.edge_links != b.edge_links).any() for a, b in zip(pool, other))
    with pytest.raises(RuntimeError):
        so.philox_sample_graph(1, 0, 0, 10, 30, want_edges=31)  # unreachable count -> the reference's RuntimeError
