"""CPU: host-side graph pool logic (generator distribution facts, CSR packing)."""
import numpy as np

import sy_oracle as so
from student_mechanism_design_b200.graphs import GraphSpec, generate_connected_graph, generate_graph_pool, pack_csr


def test_generator_matches_reference_distribution_facts():
    """graph_layout.py:9-80: connected, no self loops / parallel edges, weights in 1..4, the
    degree cap binds only the extra edges, edge count <= requested."""
    rng = np.random.default_rng(0)
    for N, E in [(10, 10), (15, 20), (50, 110), (200, 400)]:
        g = generate_connected_graph(N, E, rng)
        assert len(g.edges) <= E and len(g.edges) >= N - 1
        assert g.edges.min() >= 1 and g.edges.max() <= 4
        pairs = {tuple(sorted(p)) for p in g.edge_links.tolist()}
        assert len(pairs) == len(g.edge_links) and all(a != b for a, b in pairs)
        og = so.Graph(N, g.edge_links, g.edges)
        assert og.apsp().max() < so.INF_U16  # connected
        deg = np.bincount(g.edge_links.ravel(), minlength=N)
        extra = g.edge_links[N - 1:]
        tree_deg = np.bincount(g.edge_links[: N - 1].ravel(), minlength=N)
        for a, b in extra.tolist():
            assert max(tree_deg[a], tree_deg[b]) < 4
        assert deg.sum() == 2 * len(g.edges)


def test_generator_tree_depth_statistics_match_reference_sampler():
    """Same distribution as the oracle's draw-for-draw restatement: compare mean degree-1 node
    count of the tree over many samples (random recursive trees: ~N/2 leaves)."""
    import random

    N = 12
    rng = np.random.default_rng(1)
    mine = [np.sum(np.bincount(generate_connected_graph(N, None, rng).edge_links.ravel(), minlength=N) == 1) for _ in range(600)]
    py, npr = random.Random(5), np.random.RandomState(5)
    ref = [np.sum(np.bincount(so.sample_connected_graph(N, None, py, npr).edge_links.ravel(), minlength=N) == 1) for _ in range(600)]
    assert abs(np.mean(mine) - np.mean(ref)) < 0.25


def test_pool_has_common_edge_count():
    pool = generate_graph_pool(5, 50, 110, seed=3)
    assert len({len(g.edges) for g in pool}) == 1


def test_pack_csr_matches_oracle_csr_and_collapses_parallel_edges():
    g = GraphSpec(4, [[0, 1], [1, 0], [1, 2], [3, 2]], [3, 2, 4, 1])
    row_ptr, col, w, stride = pack_csr([g])
    og = so.Graph(4, g.edge_links, g.edges)
    rp, c, ww = og.csr()
    assert np.array_equal(row_ptr[0], rp) and np.array_equal(col[0, : len(c)], c) and np.array_equal(w[0, : len(c)], ww)
    assert og.weight_matrix()[0, 1] == 2  # min weight of the duplicated edge (yard.py:454-465)
    pool = generate_graph_pool(3, 15, 20, seed=1)
    row_ptr, col, w, stride = pack_csr(pool)
    for i, gg in enumerate(pool):
        rp, c, ww = so.Graph(15, gg.edge_links, gg.edges).csr()
        assert np.array_equal(row_ptr[i], rp) and np.array_equal(col[i, : len(c)], c) and np.array_equal(w[i, : len(c)], ww)
