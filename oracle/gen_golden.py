"""Generate tests/golden/*.npz by running the UNMODIFIED reference env (build container only).

Usage:  python oracle/gen_golden.py            (needs /root/reference; takes a few minutes)

What is recorded (the reference's tests pin only the mask rule, so
the oracle is pinned by trace replay of the reference implementation itself):

* traces.npz   -- per trace: the seed, the sampled graph and start nodes (pins the graph
                  sampler restatement), the replayed actions, and after every step the
                  reference's positions, budgets, all agents' action masks, terminated /
                  truncated / winner, visit counts at the police nodes, and the rewards
                  from TWO runs of the same trace: python-float weights (fp64 arithmetic,
                  evaluator.py:59-71) and 0-dim fp32 torch weights (gnn_trainer.py:98-110).
* masks.npz    -- compute_action_mask (action_mask.py:30-83) on random dense inputs with
                  None / scalar / vector / matrix tolls and float budgets.
* belief.npz   -- ParticleBeliefTracker (belief_module.py:41-111) with 200k particles on
                  small graphs: empirical distributions over a few updates incl. a reveal
                  and a hint, to pin the exact-expectation definition statistically.

This script is test infrastructure; it is the only place that touches /root/reference.
"""
from __future__ import annotations

import os
import random
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

import ref_loader  # noqa: E402
from sy_oracle import DEFAULT_REWARD_WEIGHTS, REWARD_WEIGHT_NAMES  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def _weights_py(rng, kind):
    if kind == "default":
        return dict(DEFAULT_REWARD_WEIGHTS)
    # sigmoid-range weights like RewardWeightNet's output (reward_net.py:34-39)
    return {k: float(rng.uniform(0.02, 0.98)) for k in REWARD_WEIGHT_NAMES}


def _as_torch32(weights):
    import torch

    return {k: torch.tensor(np.float32(v)) for k, v in weights.items()}


def _pick_action(rng, env, agent_idx, n_nodes, policy):
    moves = env.get_possible_moves(agent_idx)
    u = rng.random()
    if policy == "mixed":
        if u < 0.75 and len(moves):
            return int(rng.choice(moves))
        if u < 0.90:
            return int(rng.integers(-3, n_nodes + 3))
        return -1
    if policy == "valid":
        return int(rng.choice(moves)) if len(moves) else -1
    if policy == "police_idle_invalid":  # police send their own node: not skipped, never move
        if agent_idx == 0:
            return int(rng.choice(moves)) if len(moves) else -1
        return int(env.police_positions[agent_idx - 1])
    if policy == "lazy":  # police mostly send their own node (invalid, not skipped) -> long episodes
        if agent_idx == 0 or u < 0.3:
            return int(rng.choice(moves)) if len(moves) else -1
        if u < 0.9:
            return int(env.police_positions[agent_idx - 1])
        return int(rng.integers(-3, n_nodes + 3))
    if policy == "all_default":  # test/env_test.py:53-77
        return -1
    raise ValueError(policy)


def run_trace(seed, N, E, P, money, weights, policy, max_steps, actions=None):
    """Run the reference once.  If `actions` is given they are replayed, else drawn."""
    env = ref_loader.make_reference_env(P, money, weights, N, E, seed=seed)
    A = P + 1
    names = ["MrX"] + [f"Police{i}" for i in range(P)]
    rec = dict(
        edge_links=np.asarray(env.board.edge_links, dtype=np.int32),
        edges=np.asarray(env.board.edges, dtype=np.int64),
        start=np.asarray([env.MrX_pos[0]] + list(env.police_positions), dtype=np.int32),
        obs0_mask=np.stack([env._get_graph_observations()[n]["action_mask"] for n in names]),
    )
    rng = np.random.default_rng(seed + 1000003)
    acts, pos, mon, masks, rew, term, trunc, win, vis = [], [], [], [], [], [], [], [], []
    for s in range(max_steps):
        if actions is None:
            a = [_pick_action(rng, env, i, N, policy) for i in range(A)]
        else:
            a = [int(x) for x in actions[s]]
        obs, r, te, tr, _ = env.step({n: a[i] for i, n in enumerate(names)})
        acts.append(a)
        pos.append([int(env.MrX_pos[0])] + [int(p) for p in env.police_positions])
        mon.append([int(m) for m in env.agents_money])
        masks.append(np.stack([obs[n]["action_mask"] for n in names]))
        rew.append([float(r[n]) for n in names])
        term.append(bool(te["MrX"]))
        trunc.append(bool(tr["MrX"]))
        win.append({None: 0, "MrX": 1, "Police": 2}[env.current_winner])
        vis.append([int(env.node_visit_counts[p]) for p in env.police_positions])
        # obs sanity the oracle also reproduces: one-hot node features, budgets
        nf = obs["MrX"]["node_features"]
        assert nf.sum() == A and all(nf[pos[-1][i], i] == 1 for i in range(A))
        if term[-1] or trunc[-1] or (actions is not None and s + 1 == len(actions)):
            break
    rec.update(
        actions=np.asarray(acts, dtype=np.int64),
        pos=np.asarray(pos, dtype=np.int32),
        money=np.asarray(mon, dtype=np.int32),
        masks=np.packbits(np.asarray(masks, dtype=bool), axis=-1),
        reward=np.asarray(rew, dtype=np.float64),
        terminated=np.asarray(term),
        truncated=np.asarray(trunc),
        winner=np.asarray(win, dtype=np.int8),
        visits_at_police=np.asarray(vis, dtype=np.int32),
        final_visits=np.asarray([env.node_visit_counts[n] for n in range(N)], dtype=np.int32),
    )
    return rec


# (name, N, E, P, money, weight kind, policy, max_steps, seeds)
_FAMILIES = [
    ("c1_mixed", 15, 20, 2, 10, "default", "mixed", 60, range(100, 108)),
    ("c1_valid", 15, 20, 2, 10, "random", "valid", 60, range(110, 116)),
    ("c1_lazy", 15, 20, 2, 10, "random", "lazy", 80, range(120, 126)),
    ("c1_all_default", 15, 20, 2, 10, "default", "all_default", 3, range(130, 131)),
    ("n10_p1", 10, 12, 1, 8, "random", "mixed", 60, range(140, 146)),
    ("n10_p3_poor", 10, 10, 3, 4, "random", "lazy", 60, range(150, 156)),
    ("n10_p5", 10, 14, 5, 8, "random", "lazy", 60, range(160, 164)),
    ("n30_p3", 30, 55, 3, 20, "random", "mixed", 60, range(170, 176)),
    ("n30_p5_poor", 30, 50, 5, 4, "default", "valid", 60, range(180, 186)),
    ("n30_p2_rich", 30, 60, 2, 20, "random", "lazy", 80, range(190, 194)),
    ("n20_p6", 20, 36, 6, 8, "random", "lazy", 60, range(200, 206)),
    ("n25_p4", 25, 45, 4, 12, "random", "mixed", 60, range(210, 216)),
    ("c2_mixed", 50, 110, 3, 10, "default", "mixed", 40, range(220, 223)),
    ("c2_lazy", 50, 110, 3, 10, "random", "lazy", 40, range(230, 233)),
    ("n12_timeout", 12, 16, 2, 1000, "random", "police_idle_invalid", 260, range(240, 241)),
    ("c3_short", 200, 400, 6, 20, "default", "valid", 6, range(250, 251)),
]
TRACES = [
    (f"{name}_s{seed}", seed, N, E, P, money, wk, pol, ms)
    for name, N, E, P, money, wk, pol, ms, seeds in _FAMILIES
    for seed in seeds
]


def gen_traces():
    out = {}
    names = []
    t0 = time.time()
    for name, seed, N, E, P, money, wkind, policy, max_steps in TRACES:
        wrng = np.random.default_rng(seed + 77)
        w = _weights_py(wrng, wkind)
        r64 = run_trace(seed, N, E, P, money, w, policy, max_steps)
        r32 = run_trace(seed, N, E, P, money, _as_torch32(w), policy, max_steps, actions=r64["actions"])
        for k in ("pos", "money", "masks", "terminated", "truncated", "winner", "edge_links", "edges", "start"):
            assert np.array_equal(r64[k], r32[k]), (name, k)
        for k, v in r64.items():
            out[f"{name}/{k}"] = v
        out[f"{name}/reward32"] = r32["reward"].astype(np.float32)
        assert np.array_equal(out[f"{name}/reward32"].astype(np.float64), r32["reward"]), "fp32 rewards must be fp32-exact"
        out[f"{name}/weights"] = np.asarray([w[k] for k in REWARD_WEIGHT_NAMES], dtype=np.float64)
        out[f"{name}/config"] = np.asarray([seed, N, E, P, money], dtype=np.int64)
        names.append(name)
        print(f"  {name}: {len(r64['actions'])} steps, end winner={r64['winner'][-1]} "
              f"term={r64['terminated'][-1]} trunc={r64['truncated'][-1]}  [{time.time()-t0:.0f}s]", flush=True)
    out["names"] = np.asarray(names)
    np.savez_compressed(os.path.join(OUT, "traces.npz"), **out)


def gen_masks():
    ref = ref_loader.load_reference()
    rng = np.random.default_rng(123)
    out = {}
    n_cases = 40
    for c in range(n_cases):
        n = int(rng.integers(2, 24))
        adj = np.triu((rng.random((n, n)) < 0.3).astype(np.int64), 1)
        adj = adj + adj.T
        w = rng.integers(1, 6, size=(n, n)).astype(np.float64)
        w = np.triu(w, 1) + np.triu(w, 1).T
        kind = c % 5
        tolls = [None, float(rng.choice([0.25, 0.5, 1.0, 2.0])), rng.random(n) * 2, rng.random((n, n)) * 2, None][kind]
        weights = None if kind == 4 else w
        cur = int(rng.integers(0, n))
        budget = float(rng.choice([0.5, 1.0, 1.5, 2.0, 3.0, 4.25, 6.0, 100.0]))
        res = ref.compute_action_mask(adj, cur, budget, tolls=tolls, edge_weights=weights)
        assert res.valid_actions == np.nonzero(res.mask)[0].tolist()
        out[f"{c}/adj"] = adj
        out[f"{c}/w"] = w if weights is not None else np.zeros((0, 0))
        out[f"{c}/toll_kind"] = np.asarray(kind)
        out[f"{c}/tolls"] = np.zeros(0) if tolls is None else np.asarray(tolls, dtype=np.float64)
        out[f"{c}/cur"] = np.asarray(cur)
        out[f"{c}/budget"] = np.asarray(budget)
        out[f"{c}/mask"] = res.mask
    out["n_cases"] = np.asarray(n_cases)
    np.savez_compressed(os.path.join(OUT, "masks.npz"), **out)


def gen_belief():
    ref = ref_loader.load_reference()
    import sy_oracle

    out = {}
    cases = [(0, 8, 11), (1, 12, 18)]
    for ci, (seed, N, E) in enumerate(cases):
        g = sy_oracle.sample_connected_graph(N, E, random.Random(seed), np.random.RandomState(seed))
        adj = g.adjacency()
        tr = ref.ParticleBeliefTracker(N, num_particles=200_000, rng=np.random.default_rng(seed))
        dists = []
        # script: 2 plain updates, reveal at node 3, 2 plain updates, 1 update with a hint
        script = [("plain", None)] * 2 + [("reveal", 3)] + [("plain", None)] * 2 + [("hint", [1, 2, 5])]
        for kind, arg in script:
            if kind == "plain":
                dists.append(tr.update(adj))
            elif kind == "reveal":
                dists.append(tr.update(adj, reveal=arg))
            else:
                dists.append(tr.update(adj, observation_hint=arg))
        out[f"{ci}/edge_links"] = g.edge_links
        out[f"{ci}/edges"] = g.edges
        out[f"{ci}/N"] = np.asarray(N)
        out[f"{ci}/empirical"] = np.asarray(dists)
    out["n_cases"] = np.asarray(len(cases))
    out["script"] = np.asarray(["plain", "plain", "reveal:3", "plain", "plain", "hint:1,2,5"])
    np.savez_compressed(os.path.join(OUT, "belief.npz"), **out)


def gen_tables():
    """The float64 tables NumPy produced on the host that generated the golden rewards.
    NumPy's SIMD exp is not correctly rounded (e.g. exp(-26) differs from libm by 1 ulp on
    AVX512 hosts), so `bit-exact with the reference` is relative to the host's NumPy; golden
    replays therefore pass these recorded tables instead of recomputing them."""
    np.savez_compressed(
        os.path.join(OUT, "tables.npz"),
        exp_neg=np.exp(-np.arange(1100, dtype=np.float64)),
        coverage=np.exp(-np.log1p(np.arange(1024, dtype=np.float64))),
    )


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    which = sys.argv[1:] or ["tables", "traces", "masks", "belief"]
    if "tables" in which:
        gen_tables()
        print("tables.npz written")
    if "masks" in which:
        gen_masks()
        print("masks.npz written")
    if "belief" in which:
        gen_belief()
        print("belief.npz written")
    if "traces" in which:
        gen_traces()
        print("traces.npz written")
