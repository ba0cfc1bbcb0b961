"""Pins the plain-C oracle (oracle/sy_oracle.c) against the vectors recorded from the UNMODIFIED
reference (tests/golden/traces.npz) and against the Python oracle on seeded rollouts.  CPU only."""
import os

import numpy as np
import pytest

import sy_oracle as so
import sy_oracle_c as oc
from conftest import GOLDEN


@pytest.fixture(scope="module")
def tables():
    t = np.load(os.path.join(GOLDEN, "tables.npz"))
    return t["exp_neg"], t["coverage"]


def _trace(gt, name):
    keys = ["edge_links", "edges", "start", "obs0_mask", "actions", "pos", "money", "masks", "reward", "reward32",
            "terminated", "truncated", "winner", "visits_at_police", "final_visits", "weights", "config"]
    return {k: gt[f"{name}/{k}"] for k in keys}


def test_c_philox_known_answers():
    assert oc.philox4x32((0, 0, 0, 0), (0, 0)) == (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)
    assert oc.philox4x32((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2) == (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)
    assert oc.philox4x32((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0)) == (
        0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)


def test_c_apsp_matches_scipy_and_python(golden_traces):
    """pathfinding.py:34-137"""
    from scipy.sparse import csr_matrix
    from scipy.sparse.csgraph import dijkstra

    gt = golden_traces
    names = [str(n) for n in gt["names"]]
    for n in names[::6] + [names[-1]]:
        t = _trace(gt, n)
        N = int(t["config"][1])
        g = so.Graph(N, t["edge_links"], t["edges"])
        D = oc.apsp(N, *g.csr())
        ref = dijkstra(csr_matrix(g.weight_matrix().astype(float)), directed=False)
        assert np.array_equal(D.astype(float), ref), n
    g = so.Graph(6, [[0, 1], [1, 2], [3, 4]], [2, 3, 1])
    D = oc.apsp(6, *g.csr())
    assert D[0, 2] == 5 and D[0, 3] == 0xFFFF and D[5, 5] == 0 and D[5, 1] == 0xFFFF


@pytest.mark.parametrize("mode", ["fp64", "fp32"])
def test_c_oracle_replays_reference_traces(golden_traces, tables, mode):
    """yard.py:144-269, reward_calculator.py:26-266, action_mask.py:54-83 -- bit-exact replay."""
    gt = golden_traces
    total = 0
    for n in [str(x) for x in gt["names"]]:
        t = _trace(gt, n)
        seed, N, E, P, money = [int(x) for x in t["config"]]
        cfg = so.OracleConfig(num_police=P, agent_money=money, reward_mode=mode, exp_table=tables[0], cov_table=tables[1],
                              reward_weights=dict(zip(so.REWARD_WEIGHT_NAMES, t["weights"].tolist())))
        cb = oc.CBatch(cfg, [so.Graph(N, t["edge_links"], t["edges"])], 1, auto_reset=False,
                       start_positions=t["start"][None], graph_id=[0])
        assert np.array_equal(cb.masks()[0], t["obs0_mask"]), n
        masks = np.unpackbits(t["masks"], axis=-1)[..., :N].astype(bool)
        want_r = t["reward"] if mode == "fp64" else t["reward32"]
        for s, act in enumerate(t["actions"]):
            out = cb.step(act[None])
            tag = (n, s)
            assert cb.pos()[0].tolist() == t["pos"][s].tolist(), tag
            assert cb.money()[0].tolist() == t["money"][s].tolist(), tag
            assert (bool(out["terminated"][0]), bool(out["truncated"][0]), int(out["winner"][0])) == (
                bool(t["terminated"][s]), bool(t["truncated"][s]), int(t["winner"][s])), tag
            assert np.array_equal(cb.masks()[0], masks[s]), tag
            assert out["reward"][0].astype(want_r.dtype).tobytes() == want_r[s].tobytes(), (tag, out["reward"][0], want_r[s])
            v = cb.visits()[0]
            assert [int(v[p]) for p in cb.pos()[0, 1:]] == t["visits_at_police"][s].tolist(), tag
            total += 1
        assert np.array_equal(cb.visits()[0].astype(np.int64), t["final_visits"]), n
    assert total > 800


CASES = [
    dict(N=15, E=20, P=2, money=10, G=3, B=60, kw={}, mode="fp64", steps=40, max_t=15),
    dict(N=50, E=110, P=3, money=10, G=2, B=40, kw=dict(reveal_interval=5), mode="fp32", steps=30, max_t=250),
    dict(N=30, E=55, P=6, money=12, G=4, B=50, kw=dict(reveal_interval=3, toll=1, belief=True), mode="fp64", steps=40, max_t=250),
    dict(N=24, E=40, P=15, money=9, G=2, B=20, kw=dict(toll=2, belief=True, reveal_interval=4), mode="fp32", steps=25, max_t=250),
]


@pytest.mark.parametrize("ci", range(len(CASES)))
@pytest.mark.parametrize("auto_reset", [True, False])
def test_c_oracle_matches_python_oracle(tables, ci, auto_reset):
    """same seeds, same Philox policy: C and Python restatements agree on everything (belief to 1e-12:
    NumPy's pairwise sum vs a sequential sum in the normalisation)."""
    from student_mechanism_design_b200.graphs import generate_graph_pool

    c = CASES[ci]
    pool = [so.Graph(g.num_nodes, g.edge_links, g.edges) for g in generate_graph_pool(c["G"], c["N"], c["E"], seed=4)]
    cfg = so.OracleConfig(num_police=c["P"], agent_money=c["money"], reward_mode=c["mode"], max_timestep=c["max_t"],
                          exp_table=tables[0], cov_table=tables[1], **c["kw"])
    py = so.OracleBatch.from_seed(cfg, pool, c["B"], seed=21, env_offset=7, auto_reset=auto_reset, resample_graph=True)
    cb = oc.CBatch(cfg, pool, c["B"], seed=21, env_offset=7, auto_reset=auto_reset, resample_graph=True, threads=2)
    rng = np.random.default_rng(ci)
    for s in range(c["steps"]):
        assert np.array_equal(cb.pos(), py.pos()), s
        acts = py.sample_actions(s)
        assert np.array_equal(cb.sample_actions(s), acts), s
        junk = rng.integers(-3, c["N"] + 3, size=acts.shape)
        acts = np.where(rng.random(acts.shape) < 0.85, acts, junk)
        a, b = cb.step(acts), py.step(acts)
        for k in ("reward", "terminated", "truncated", "winner"):
            assert a[k].tobytes() == b[k].tobytes(), (k, s)
        for k in ("pos", "money", "timestep", "visits", "masks", "node_features", "revealed"):
            assert np.array_equal(getattr(cb, k)(), getattr(py, k)()), (k, s)
        assert cb.graph_id == py.graph_id and cb.episode == py.episode and cb.done == list(py.done), s
        if cfg.belief:
            assert np.abs(cb.belief() - py.belief()).max() < 1e-12, s
