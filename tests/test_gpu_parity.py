"""GPU parity tests proper: the CUDA path (through the C ABI, via the host mirror) against
(i) the vectors recorded from the UNMODIFIED reference (tests/golden/*.npz) and (ii) the CPU
oracle on the same seeded inputs.  Bar: bit-exact for positions, budgets, masks, flags, visit
counts and rewards (fp64 and fp32 modes); belief_map within 1e-6 abs (north_star)."""
import os

import numpy as np
import pytest

import sy_oracle as so
from conftest import GOLDEN

pytestmark = pytest.mark.gpu

BELIEF_TOL = 1e-6  # north_star: "belief_map must match within 1e-6 abs"


@pytest.fixture(scope="module")
def torch_cuda():
    import torch

    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    return torch


@pytest.fixture(scope="module")
def tables():
    t = np.load(os.path.join(GOLDEN, "tables.npz"))
    return t["exp_neg"], t["coverage"]


def _pkg():
    import student_mechanism_design_b200 as pkg

    return pkg


def _trace(gt, name):
    keys = ["edge_links", "edges", "start", "obs0_mask", "actions", "pos", "money", "masks", "reward", "reward32",
            "terminated", "truncated", "winner", "visits_at_police", "final_visits", "weights", "config"]
    return {k: gt[f"{name}/{k}"] for k in keys}


def _groups(gt):
    """golden traces grouped into batches that can share one env handle"""
    groups = {}
    for n in [str(x) for x in gt["names"]]:
        t = _trace(gt, n)
        seed, N, E, P, money = [int(x) for x in t["config"]]
        key = (N, P, money, t["weights"].tobytes(), t["edge_links"].shape[0])
        groups.setdefault(key, []).append(t)
    return list(groups.values())


def test_library_loaded_is_in_tree(torch_cuda):
    pkg = _pkg()
    lib = pkg.load_library()
    assert os.path.dirname(pkg.LIB_PATH).endswith("student_mechanism_design_b200")
    assert lib.sy_abi_version() == pkg._cabi.SY_ABI_VERSION == 4


def test_graph_tables_match_oracle(torch_cuda, golden_traces):
    """sy_load_graphs: dense weights (yard.py:404-418) and the all-pairs table that replaces
    Pathfinder.get_distance (pathfinding.py:34-137), incl. the 200-node golden graph."""
    pkg = _pkg()
    gt = golden_traces
    names = [str(x) for x in gt["names"]]
    for n in names[::9] + [names[-1]]:
        t = _trace(gt, n)
        N, P = int(t["config"][1]), int(t["config"][3])
        g = so.Graph(N, t["edge_links"], t["edges"])
        env = pkg.BatchedScotlandYardEnv(1, P, 5, graphs=[pkg.GraphSpec(N, t["edge_links"], t["edges"])])
        W, D = env.graph_tables(0)
        assert np.array_equal(W.astype(np.int64), g.weight_matrix()), n
        assert np.array_equal(D.astype(np.int64), g.apsp()), n
        assert env.get_distance(0, 0) == 0.0
        env.close()


def test_apsp_disconnected_graph(torch_cuda):
    """pathfinding.py:133-137: unreachable -> inf (0xFFFF in the table)."""
    pkg = _pkg()
    g = pkg.GraphSpec(6, [[0, 1], [1, 2], [3, 4]], [2, 3, 1])
    env = pkg.BatchedScotlandYardEnv(1, 1, 5, graphs=[g])
    _, D = env.graph_tables(0)
    assert D[0, 2] == 5 and D[3, 4] == 1 and D[0, 3] == 0xFFFF and D[5, 0] == 0xFFFF and D[5, 5] == 0
    assert env.get_distance(0, 5) == float("inf")
    env.close()


@pytest.mark.parametrize("mode", ["fp64", "fp32"])
def test_golden_trace_replay(torch_cuda, golden_traces, tables, mode):
    """Replay of the reference's recorded episodes (yard.py:144-269, reward_calculator.py:26-266,
    action_mask.py:54-83): every step bit-exact, batched with one graph per env."""
    torch = torch_cuda
    pkg = _pkg()
    total = 0
    for group in _groups(golden_traces):
        t0 = group[0]
        seed, N, E, P, money = [int(x) for x in t0["config"]]
        A, B = P + 1, len(group)
        weights = dict(zip(so.REWARD_WEIGHT_NAMES, t0["weights"].tolist()))
        graphs = [pkg.GraphSpec(N, t["edge_links"], t["edges"]) for t in group]
        env = pkg.BatchedScotlandYardEnv(B, P, money, weights, graphs=graphs, reward_mode=mode, reward_tables=tables,
                                         keep_reward64=True)
        start = np.stack([t["start"] for t in group])
        obs = env.reset(init_pos=start, graph_id=np.arange(B))
        assert np.array_equal(obs["action_mask"].cpu().numpy(), np.stack([t["obs0_mask"] for t in group]))
        assert np.array_equal(obs["agent_position"].cpu().numpy(), start)
        lens = [len(t["actions"]) for t in group]
        for s in range(max(lens)):
            acts = np.full((B, A), -1, dtype=np.int64)
            for b, t in enumerate(group):
                if s < lens[b]:
                    acts[b] = t["actions"][s]
            obs, rew, term, trunc, info = env.step(torch.from_numpy(acts).cuda())
            pos, mon = env.pos.cpu().numpy(), env.money.cpu().numpy()
            mask = obs["action_mask"].cpu().numpy()
            r32, r64 = rew.cpu().numpy(), env.reward64.cpu().numpy()
            te, tr, win = term.cpu().numpy(), trunc.cpu().numpy(), info["winner"].cpu().numpy()
            nf = obs["node_features"].cpu().numpy()
            vis = env.visits.cpu().numpy()
            for b, t in enumerate(group):
                if s >= lens[b]:
                    continue
                tag = (int(t["config"][0]), s)
                assert pos[b].tolist() == t["pos"][s].tolist(), tag
                assert mon[b].tolist() == t["money"][s].tolist(), tag
                want_mask = np.unpackbits(t["masks"][s], axis=-1)[..., :N].astype(bool)
                assert np.array_equal(mask[b], want_mask), tag
                assert bool(te[b, 0]) == bool(t["terminated"][s]) and bool(tr[b, 0]) == bool(t["truncated"][s]), tag
                assert te[b].all() == te[b].any() and tr[b].all() == tr[b].any(), tag
                assert int(win[b]) == int(t["winner"][s]), tag
                if mode == "fp64":
                    assert r64[b].tobytes() == t["reward"][s].tobytes(), (tag, r64[b], t["reward"][s])
                    assert r32[b].tobytes() == t["reward"][s].astype(np.float32).tobytes(), tag
                else:
                    assert r32[b].tobytes() == t["reward32"][s].tobytes(), (tag, r32[b], t["reward32"][s])
                assert [int(vis[b, p]) for p in pos[b, 1:]] == t["visits_at_police"][s].tolist(), tag
                assert nf[b].sum() == A and all(nf[b, pos[b, a], a] == 1 for a in range(A)), tag
                if s == lens[b] - 1:
                    assert np.array_equal(vis[b].astype(np.int64), t["final_visits"]), tag
                total += 1
        env.close()
    assert total > 800


CASES = [
    # N, E, P, money, G, B, kw, mode, steps
    dict(N=15, E=20, P=2, money=10, G=3, B=200, kw={}, mode="fp64", steps=40),
    dict(N=50, E=110, P=3, money=10, G=2, B=97, kw=dict(reveal_interval=5), mode="fp32", steps=40),
    dict(N=30, E=55, P=6, money=12, G=4, B=130, kw=dict(reveal_interval=3, tolls=1, belief=True), mode="fp64", steps=45),
    dict(N=40, E=70, P=1, money=6, G=1, B=64, kw=dict(belief=True), mode="fp32", steps=30),
    dict(N=24, E=40, P=15, money=9, G=2, B=33, kw=dict(tolls=2, belief=True, reveal_interval=4), mode="fp64", steps=25),
    dict(N=200, E=400, P=6, money=20, G=1, B=96, kw=dict(tolls=1, belief=True, reveal_interval=5), mode="fp64", steps=12),
]


def _make_pair(pkg, c, tables, auto_reset=True, seed=17, env_offset=0, max_timestep=250, resample_graph=False, B=None):
    pool = pkg.generate_graph_pool(c["G"], c["N"], c["E"], seed=3)
    B = B or c["B"]
    env = pkg.BatchedScotlandYardEnv(B, c["P"], c["money"], graphs=pool, seed=seed, auto_reset=auto_reset,
                                     reward_mode=c["mode"], keep_reward64=True, reward_tables=tables,
                                     env_offset=env_offset, max_timestep=max_timestep, resample_graph=resample_graph,
                                     belief_ce=True, **c["kw"])
    ocfg = so.OracleConfig(num_police=c["P"], agent_money=c["money"], reward_mode=c["mode"],
                           reveal_interval=c["kw"].get("reveal_interval", 0), toll=c["kw"].get("tolls", 0),
                           belief=c["kw"].get("belief", False), max_timestep=max_timestep,
                           exp_table=tables[0], cov_table=tables[1])
    ograph = [so.Graph(g.num_nodes, g.edge_links, g.edges) for g in pool]
    ob = so.OracleBatch.from_seed(ocfg, ograph, B, seed=seed, env_offset=env_offset, auto_reset=auto_reset,
                                  resample_graph=resample_graph)
    return env, ob


def _compare_state(env, ob, c, tag):
    assert np.array_equal(env.pos.cpu().numpy(), ob.pos()), tag
    assert np.array_equal(env.money.cpu().numpy(), ob.money()), tag
    assert np.array_equal(env.timestep.cpu().numpy(), ob.timestep()), tag
    assert np.array_equal(env.graph_id.cpu().numpy(), np.asarray(ob.graph_id, dtype=np.int32)), tag
    assert np.array_equal(env.episode.cpu().numpy(), np.asarray(ob.episode, dtype=np.int32)), tag
    assert np.array_equal(env.visits.cpu().numpy(), ob.visits()), tag
    assert np.array_equal(env.action_mask.cpu().numpy(), ob.masks()), tag
    assert np.array_equal(env.node_features.cpu().numpy(), ob.node_features()), tag
    assert np.array_equal(env.agent_budget.cpu().numpy(), ob.money().astype(np.float32)), tag
    assert np.array_equal(env.mrx_revealed.cpu().numpy(), ob.revealed()), tag
    if c["kw"].get("belief"):
        got = env.belief_map.cpu().numpy().astype(np.float64)
        err = np.abs(got - ob.belief()).max()
        assert err <= BELIEF_TOL, (tag, err)
        assert np.abs(got.sum(axis=1) - 1.0).max() < 1e-5, tag


def _compare_out(env, want, c, tag):
    r64 = env.reward64.cpu().numpy()
    r32 = env.reward.cpu().numpy()
    if c["mode"] == "fp64":
        assert r64.tobytes() == want["reward"].tobytes(), tag
        assert r32.tobytes() == want["reward"].astype(np.float32).tobytes(), tag
    else:
        assert r32.tobytes() == want["reward"].tobytes(), tag
    te, tr = env.terminated.cpu().numpy(), env.truncated.cpu().numpy()
    assert np.array_equal(te, np.repeat(want["terminated"][:, None], te.shape[1], 1)), tag
    assert np.array_equal(tr, np.repeat(want["truncated"][:, None], tr.shape[1], 1)), tag
    assert np.array_equal(env.winner.cpu().numpy(), want["winner"]), tag


@pytest.mark.parametrize("ci", range(len(CASES)))
def test_random_policy_rollout_matches_oracle(torch_cuda, tables, ci):
    """Batched rollout with the on-device Philox random-valid policy and same-step auto-reset:
    sampler, dynamics, budgets (with tolls), rewards, flags, visit counts, masks, node features,
    reveal schedule and belief map against the CPU oracle.  B is not a multiple of the tile."""
    c = CASES[ci]
    pkg = _pkg()
    env, ob = _make_pair(pkg, c, tables, max_timestep=20 if ci == 0 else 250)
    env.reset()
    _compare_state(env, ob, c, "reset")
    n_done = 0
    for s in range(c["steps"]):
        acts = env.sample_actions(step_counter=s)
        a_h = acts.cpu().numpy()
        assert np.array_equal(a_h, ob.sample_actions(s)), ("sampler", s)
        env.step(acts)
        want = ob.step(a_h)
        _compare_out(env, want, c, ("out", s))
        _compare_state(env, ob, c, ("state", s))
        n_done += int(want["terminated"].sum() + want["truncated"].sum())
    st = env.stats()
    assert st["env_steps"] == c["B"] * c["steps"] and st["episodes"] == n_done
    assert st["mrx_wins"] + st["police_wins"] == n_done
    assert n_done > 0
    # the aggregates of the reference's MetricsTracker (src/eval/metrics.py:168-232) from the device statistics
    fin = np.asarray(ob.finished, dtype=np.int64)  # (length, winner, budget spent) per finished episode
    lengths, winners, spent = fin[:, 0], fin[:, 1], fin[:, 2]
    moves = sum(e.moves for e in ob.envs)
    m = env.metrics()
    assert m["num_episodes"] == len(fin) and m["mrx_wins"] == int((winners == so.WINNER_MRX).sum())
    assert m["police_wins"] == int((winners == so.WINNER_POLICE).sum()) and st["police_moves"] == moves
    np.testing.assert_allclose(m["win_rate"], np.mean(winners == so.WINNER_MRX), rtol=1e-12)
    np.testing.assert_allclose(m["win_rate_std"], np.std((winners == so.WINNER_MRX).astype(float)), rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(m["mean_episode_length"], lengths.mean(), rtol=1e-12)
    np.testing.assert_allclose(m["episode_length_std"], lengths.std(), rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(m["mean_budget_spent"], spent.mean(), rtol=1e-12)
    np.testing.assert_allclose(m["mean_budget_efficiency"], spent.mean() / max(c["P"] * c["money"], 1), rtol=1e-12)
    np.testing.assert_allclose(m["mean_tolls_paid"], c["kw"].get("tolls", 0) * moves / len(fin), rtol=1e-12)
    pol, mrx = lengths[winners == so.WINNER_POLICE], lengths[winners == so.WINNER_MRX]
    np.testing.assert_allclose(m["mean_time_to_catch"], pol.mean() if len(pol) else 0.0, rtol=1e-12)
    np.testing.assert_allclose(m["mean_survival_time"], mrx.mean() if len(mrx) else 0.0, rtol=1e-12)
    # belief quality at reveal steps (belief_quality.py:8-11): device fp32 prediction vs the float64 oracle
    ces = np.asarray(ob.belief_ces)
    assert st["reveals"] == len(ces)
    if c["kw"].get("belief") and c["kw"].get("reveal_interval", 0) > 0:
        assert len(ces) > 0
    if len(ces):
        np.testing.assert_allclose(m["mean_belief_ce"], ces.mean(), rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(m["belief_ce_std"], ces.std(), rtol=1e-4, atol=1e-5)
    else:
        assert m["mean_belief_ce"] == 0.0 and m["belief_ce_std"] == 0.0
    env.close()


def test_arbitrary_actions_and_freeze_without_auto_reset(torch_cuda, tables):
    """Invalid actions are data, not errors (yard.py:171-178,223-229): out-of-range, negative,
    huge int64, non-adjacent and unaffordable targets all mean `stay`; -1 means `skipped` for
    police.  Without auto-reset finished envs freeze until reset(reset_mask)."""
    torch = torch_cuda
    pkg = _pkg()
    c = dict(N=20, E=34, P=4, money=7, G=3, B=150, kw=dict(belief=True, reveal_interval=2), mode="fp64")
    env, ob = _make_pair(pkg, c, tables, auto_reset=False, max_timestep=12)
    env.reset()
    rng = np.random.default_rng(5)
    B, A, N = c["B"], c["P"] + 1, c["N"]
    for s in range(30):
        valid = ob.sample_actions(s)
        junk = rng.integers(-4, N + 4, size=(B, A))
        huge = rng.choice(np.asarray([2**40, -2**40, 2**63 - 1, -2**63, N, -2], dtype=np.int64), size=(B, A))
        u = rng.random((B, A))
        acts = np.where(u < 0.6, valid, np.where(u < 0.85, junk, np.where(u < 0.93, huge, -1))).astype(np.int64)
        env.step(torch.from_numpy(acts).cuda())
        want = ob.step(acts)
        _compare_out(env, want, c, ("out", s))
        _compare_state(env, ob, c, ("state", s))
        assert np.array_equal(env.done.cpu().numpy().astype(bool), np.asarray(ob.done)), s
        if s in (9, 19):  # partial reset of the finished envs (torchrl "_reset" semantics)
            m = np.asarray(ob.done)
            assert m.any()
            env.reset(reset_mask=m)
            for b in np.nonzero(m)[0]:
                ob.episode[b] += 1
                e = ob.envs[b]
                e.reset(ob.graphs[ob.graph_id[b]], so.philox_start_positions(ob.seed, b, ob.episode[b], N, A))
                ob.done[b] = False
            _compare_state(env, ob, c, ("after partial reset", s))
    env.close()


def test_timeout_branch_batched(torch_cuda, tables):
    """reward_calculator.py:68-74: `timestep > max_timestep` truncates; the pre-increment
    timestep is compared (yard.py:345,355), so with max_timestep=3 step index 4 truncates."""
    torch = torch_cuda
    pkg = _pkg()
    c = dict(N=15, E=20, P=2, money=1000, G=1, B=40, kw={}, mode="fp64")
    env, ob = _make_pair(pkg, c, tables, auto_reset=False, max_timestep=3)
    env.reset()
    seen_trunc = False
    for s in range(6):
        # police send their own node (not skipped, never move), MrX stays: only timeouts can end it
        acts = ob.pos().astype(np.int64)
        env.step(torch.from_numpy(acts).cuda())
        want = ob.step(acts)
        _compare_out(env, want, c, s)
        if want["truncated"].any():
            assert s == 4
            seen_trunc = True
    assert seen_trunc
    env.close()


def test_shard_invariance(torch_cuda, tables):
    """Sharding by batch: a shard created with env_offset=k reproduces envs
    [k, k+b) of the full batch bit for bit (Philox streams are keyed by the global env index)."""
    pkg = _pkg()
    c = dict(N=30, E=55, P=3, money=10, G=4, B=160, kw=dict(belief=True, reveal_interval=4, tolls=1), mode="fp64")
    pool = pkg.generate_graph_pool(c["G"], c["N"], c["E"], seed=3)
    mk = lambda B, off: pkg.BatchedScotlandYardEnv(B, c["P"], c["money"], graphs=pool, seed=9, auto_reset=True,  # noqa: E731
                                                   resample_graph=True, env_offset=off, keep_reward64=True,
                                                   reward_tables=tables, **c["kw"])
    full, shard = mk(160, 0), mk(50, 100)
    full.reset()
    shard.reset()
    for s in range(30):
        full.step(full.sample_actions(step_counter=s))
        shard.step(shard.sample_actions(step_counter=s))
        for k in ("pos", "money", "timestep", "graph_id", "visits", "belief_map", "reward64", "terminated",
                  "action_mask", "node_features"):
            a, b = getattr(full, k)[100:150].cpu().numpy(), getattr(shard, k).cpu().numpy()
            assert a.tobytes() == b.tobytes(), (k, s)
    sf, ss = full.stats(), shard.stats()
    assert sf["env_steps"] == 160 * 30 and ss["env_steps"] == 50 * 30
    full.close()
    shard.close()


def test_dense_action_mask_matches_reference_vectors(torch_cuda):
    """sy_action_mask_dense == compute_action_mask (action_mask.py:30-113) on the vectors recorded
    from the reference function and on the reference's known-answer cases
    (test/test_action_mask.py:9-124)."""
    pkg = _pkg()
    gm = np.load(os.path.join(GOLDEN, "masks.npz"))
    for ci in range(int(gm["n_cases"])):
        kind = int(gm[f"{ci}/toll_kind"])
        tolls = None if kind in (0, 4) else (float(gm[f"{ci}/tolls"]) if kind == 1 else gm[f"{ci}/tolls"])
        w = None if kind == 4 else gm[f"{ci}/w"]
        got = pkg.dense_action_mask(gm[f"{ci}/adj"], int(gm[f"{ci}/cur"]), float(gm[f"{ci}/budget"]), tolls, w)
        assert np.array_equal(got[0].cpu().numpy(), gm[f"{ci}/mask"]), ci
    adj = np.array([[0, 1, 1], [1, 0, 0], [1, 0, 0]])
    w = np.array([[0, 2, 4], [2, 0, 0], [4, 0, 0]])
    assert pkg.dense_action_mask(adj, 0, 3, edge_weights=w)[0].tolist() == [False, True, False]
    adj2 = np.ones((2, 2)) - np.eye(2)
    assert pkg.dense_action_mask(adj2, 0, 0.5, tolls=0.25).sum().item() == 0
    assert pkg.dense_action_mask(adj2, 0, 1.5, tolls=0.25)[0].tolist() == [False, True]
    star = np.array([[0, 1, 1, 1], [1, 0, 0, 0], [1, 0, 0, 0], [1, 0, 0, 0]])
    assert pkg.dense_action_mask(star, 0, 100)[0].tolist() == [False, True, True, True]
    assert pkg.dense_action_mask(np.array([[0, 1], [1, 0]]), 0, 1, edge_weights=np.array([[0, 100], [100, 0]])).sum().item() == 0
    assert pkg.dense_action_mask(np.array([[0, 0, 0], [0, 0, 1], [0, 1, 0]]), 0, 100).sum().item() == 0
    # batched queries: every node of a graph at two budgets
    q = pkg.dense_action_mask(adj, [0, 1, 2, 0], [3, 3, 3, 10], edge_weights=w).cpu().numpy()
    for i, (cur, bud) in enumerate([(0, 3), (1, 3), (2, 3), (0, 10)]):
        assert np.array_equal(q[i], so.action_mask_dense(adj, cur, bud, edge_weights=w))


def test_full_size_properties(torch_cuda):
    """BASELINE config 3 at full size (N=200, P=6, B=65536, tolls + belief + reveal): properties
    that do not need the oracle -- one-hot node features, masks == affordable neighbours recomputed
    from the state with torch, police on distinct nodes, budgets never increase inside an episode,
    belief rows sum to 1, flags consistent, statistics add up."""
    torch = torch_cuda
    pkg = _pkg()
    N, P, B = 200, 6, 65536
    A = P + 1
    env = pkg.BatchedScotlandYardEnv(B, P, 20, graph_nodes=N, graph_edges=400, seed=1, auto_reset=True, tolls=1,
                                     belief=True, reveal_interval=5)
    env.reset()
    W = torch.from_numpy(env.graph_tables(0)[0].astype(np.int32)).cuda()  # [N, N]
    prev_money, prev_ep = env.money.clone(), env.episode.clone()
    for s in range(25):
        env.step(env.sample_actions())
        pos, money = env.pos.long(), env.money
        assert int(pos.min()) >= 0 and int(pos.max()) < N
        pol = pos[:, 1:].sort(dim=1).values
        assert bool((pol[:, 1:] != pol[:, :-1]).all()), "police share a node"
        rows = W[pos]  # [B, A, N]
        want_mask = (rows > 0) & (rows + 1 <= money.unsqueeze(-1))
        assert bool((want_mask == env.action_mask).all())
        nf = env.node_features
        hidden = env.mrx_revealed < 0
        assert bool((nf.sum(dim=1)[:, 1:] == 1).all())
        assert bool((nf.sum(dim=1)[:, 0] == (~hidden).float()).all())
        assert bool((nf[:, :, 1:].gather(1, pos[:, None, 1:]) == 1).all())
        same_ep = env.episode == prev_ep
        assert bool((money[same_ep] <= prev_money[same_ep]).all())
        assert bool((money[~same_ep][:, 1:] == 20).all()) and bool((env.timestep[~same_ep] == 0).all())
        bsum = env.belief_map.sum(dim=1)
        assert float((bsum - 1).abs().max()) < 1e-5 and float(env.belief_map.min()) >= 0
        rev = env.mrx_revealed >= 0
        assert bool((env.timestep[rev] % 5 == 0).all()) and bool((env.timestep[rev] > 0).all())
        assert bool((env.belief_map[rev].max(dim=1).values == 1).all())
        assert bool(((env.terminated | env.truncated) == env.done_flags).all())
        assert bool(torch.isfinite(env.reward).all())
        prev_money, prev_ep = money.clone(), env.episode.clone()
    st = env.stats()
    assert st["env_steps"] == B * 25
    assert st["episodes"] == int(env.episode.sum()) == st["mrx_wins"] + st["police_wins"]
    assert st["reveals"] == 0 and env.metrics()["mean_belief_ce"] == 0.0  # belief scoring is opt-in (belief_ce=True)
    env.close()


def test_error_conventions(torch_cuda):
    """C-ABI status codes surface as SyError with the library's message; nothing crashes."""
    torch = torch_cuda
    pkg = _pkg()
    with pytest.raises(pkg.SyError):
        pkg.BatchedScotlandYardEnv(4, 0, 10, graph_nodes=10, graph_edges=12)  # no police
    with pytest.raises(pkg.SyError):
        pkg.BatchedScotlandYardEnv(4, 9, 10, graph_nodes=8, graph_edges=10)  # more agents than nodes
    env = pkg.BatchedScotlandYardEnv(4, 2, 10, graph_nodes=10, graph_edges=12)
    with pytest.raises(pkg.SyError):
        env.step(torch.zeros(4, 3, dtype=torch.int64, device="cuda"))  # step before reset
    env.reset()
    with pytest.raises(ValueError):
        env.step(torch.zeros(4, 2, dtype=torch.int64, device="cuda"))  # wrong shape
    with pytest.raises(ValueError):
        env.reset(init_pos=np.full((4, 3), 99))
    env.close()


def test_step_host_equals_step(torch_cuda, tables):
    """sy_step_host (host buffers in/out, the reference's call shape) == sy_step on device buffers."""
    pkg = _pkg()
    c = dict(N=30, E=55, P=3, money=10, G=2, B=70, kw=dict(belief=True, reveal_interval=4), mode="fp64")
    a, _ = _make_pair(pkg, c, tables)
    b, _ = _make_pair(pkg, c, tables)
    a.reset()
    b.reset()
    for s in range(20):
        acts = a.sample_actions(step_counter=s)
        host_acts = b.sample_actions_host(step_counter=s)
        assert not host_acts.is_cuda and np.array_equal(host_acts.numpy(), acts.cpu().numpy())
        _, rew, term, trunc, info = a.step(acts)
        res = b.step_host(host_acts if s % 2 == 0 else host_acts.numpy().copy())
        assert not res["reward"].is_cuda
        assert res["reward"].numpy().tobytes() == rew.cpu().numpy().tobytes()
        assert np.array_equal(res["terminated"].numpy(), term.cpu().numpy())
        assert np.array_equal(res["truncated"].numpy(), trunc.cpu().numpy())
        assert np.array_equal(res["done"].numpy(), info["done"].cpu().numpy())
        assert np.array_equal(res["winner"].numpy(), info["winner"].cpu().numpy())
        assert np.array_equal(a.pos.cpu().numpy(), b.pos.cpu().numpy())
        assert np.array_equal(a.belief_map.cpu().numpy(), b.belief_map.cpu().numpy())
    a.close()
    b.close()


def test_compat_env_replays_reference_traces(torch_cuda, golden_traces, tables):
    """The single-episode view with the reference's call shape (dicts keyed by agent name, numpy observations with
    the reference's dtypes) replays recorded reference episodes bit for bit (yard.py:80-269,319-332)."""
    pkg = _pkg()
    gt = golden_traces
    names = [str(x) for x in gt["names"]]
    total = 0
    for n in names[::8]:
        t = _trace(gt, n)
        seed, N, E, P, money = [int(x) for x in t["config"]]
        weights = dict(zip(so.REWARD_WEIGHT_NAMES, t["weights"].tolist()))
        g = pkg.GraphSpec(N, t["edge_links"], t["edges"])
        env = pkg.CustomEnvironment(P, money, weights, None, 0, N, E, None, graph=g, reward_tables=tables)
        obs, infos = env.reset(graph=g, start_positions=t["start"])
        agents = ["MrX"] + [f"Police{i}" for i in range(P)]
        assert list(obs) == agents and env.agents == agents and env.action_space("MrX").n == N
        assert np.array_equal(np.stack([obs[a]["action_mask"] for a in agents]), t["obs0_mask"])
        o = obs["Police0"]
        assert o["adjacency_matrix"].dtype == np.float64 and o["adjacency_matrix"].shape == (N, N)
        assert o["node_features"].dtype == np.float64 and o["node_features"].shape == (N, P + 1)
        assert o["edge_index"].shape == (2, len(t["edges"])) and o["edge_index"].dtype == np.int32
        assert o["agent_budget"].dtype == np.float32 and o["agent_budget"].shape == (1,)
        assert o["action_mask"].dtype == bool and o["agent_position"] == int(t["start"][1])
        for s, act in enumerate(t["actions"]):
            obs, rew, term, trunc, infos = env.step({a: int(act[i]) for i, a in enumerate(agents)})
            assert [env.MrX_pos[0]] + env.police_positions == t["pos"][s].tolist(), (n, s)
            assert env.agents_money == t["money"][s].tolist(), (n, s)
            assert np.asarray([rew[a] for a in agents]).tobytes() == t["reward"][s].tobytes(), (n, s)
            assert all(term[a] == bool(t["terminated"][s]) and trunc[a] == bool(t["truncated"][s]) for a in agents)
            assert {None: 0, "MrX": 1, "Police": 2}[env.current_winner] == int(t["winner"][s])
            want_mask = np.unpackbits(t["masks"][s], axis=-1)[..., :N].astype(bool)
            assert np.array_equal(np.stack([obs[a]["action_mask"] for a in agents]), want_mask), (n, s)
            for i in range(P + 1):
                assert env.get_possible_moves(i).tolist() == np.nonzero(want_mask[i])[0].tolist()
            total += 1
        assert env.agents == [] or not (t["terminated"][-1] or t["truncated"][-1])
        env.close()
    assert total > 60


def test_torchrl_adapter_on_device(torch_cuda, tables):
    """The torchrl EnvBase adapter over a real device env (stand-in base classes; torchrl is not installed):
    zero-copy keys, `_reset` masks, per-agent action entries."""
    torch = torch_cuda
    import fake_torchrl
    from student_mechanism_design_b200 import torchrl_env

    pkg = _pkg()
    c = dict(N=20, E=34, P=2, money=8, G=2, B=40, kw=dict(belief=True, reveal_interval=2), mode="fp64")
    env, ob = _make_pair(pkg, c, tables, auto_reset=False)
    tenv = torchrl_env.make_env_class(fake_torchrl.EnvBase, fake_torchrl.TensorDict, fake_torchrl.SPECS)(env)
    # the declared specs against what a real device env emits (stand-in for torchrl's check_env_specs): every key,
    # shape, dtype and bound, on reset and on random steps drawn from the action spec
    probe, _ = _make_pair(pkg, c, tables, auto_reset=False)
    assert fake_torchrl.check_env_specs(torchrl_env.make_env_class(fake_torchrl.EnvBase, fake_torchrl.TensorDict, fake_torchrl.SPECS)(probe))
    probe.close()
    td = tenv.reset()
    assert td.get(("agents", "observation", "action_mask")).data_ptr() == env.action_mask.data_ptr()
    hidden = (env.mrx_revealed < 0)
    assert bool((td.get(("Police0", "observation", "MrX_pos"))[:, 0][hidden] == -1).all())  # no leak while MrX is hidden
    assert np.array_equal(td.get(("MrX", "observation", "agent_position"))[:, 0].cpu().numpy(), ob.pos()[:, 0])
    for s in range(15):
        acts = ob.sample_actions(s)
        tin = fake_torchrl.TensorDict({}, batch_size=[c["B"]])
        for i, name in enumerate(env.possible_agents):
            tin.set((name, "action"), torch.from_numpy(acts[:, i:i + 1]).cuda())
        out = tenv.step(tin)
        want = ob.step(acts)
        assert out.get(("agents", "reward")).squeeze(-1).cpu().numpy().tobytes() == want["reward"].astype(np.float32).tobytes()
        assert np.array_equal(out.get("terminated")[:, 0].cpu().numpy(), want["terminated"])
        assert np.array_equal(out.get(("Police1", "observation", "action_mask"))[:, 0].cpu().numpy(), ob.masks()[:, 2])
    done = np.asarray(ob.done)
    if done.any():
        m = torch.from_numpy(done).cuda().reshape(-1, 1)
        td = tenv.reset(fake_torchrl.TensorDict({"_reset": m}, batch_size=[c["B"]]))
        assert not env.done[torch.from_numpy(done).cuda()].any()
        assert bool((env.timestep[torch.from_numpy(done).cuda()] == 0).all())
    env.close()


def test_large_graph_generic_paths_vs_c_oracle(torch_cuda, tables):
    """BASELINE config 4 shape (N=1000, 6 police): too large for the transposed-tile belief path, so the generic
    warp-per-env belief, the staged-CSR writers and the u16 tables are exercised; checked against the C oracle."""
    import sy_oracle_c as oc

    pkg = _pkg()
    N, E, P, B = 1000, 2000, 6, 70
    pool = pkg.generate_graph_pool(2, N, E, seed=11)
    kw = dict(tolls=1, belief=True, reveal_interval=5)
    env = pkg.BatchedScotlandYardEnv(B, P, 20, graphs=pool, seed=3, auto_reset=True, keep_reward64=True,
                                     reward_tables=tables, resample_graph=True, max_timestep=9, **kw)
    cfg = so.OracleConfig(num_police=P, agent_money=20, toll=1, belief=True, reveal_interval=5, max_timestep=9,
                          exp_table=tables[0], cov_table=tables[1])
    ob = oc.CBatch(cfg, pool, B, seed=3, auto_reset=True, resample_graph=True)
    env.reset()
    W, D = env.graph_tables(1)
    assert np.array_equal(D.astype(np.int32), ob._D[1]) and np.array_equal(W.astype(np.int32), ob._W[1])
    for s in range(14):
        acts = env.sample_actions(step_counter=s)
        assert np.array_equal(acts.cpu().numpy(), ob.sample_actions(s)), s
        env.step(acts)
        want = ob.step(acts.cpu().numpy())
        assert env.reward64.cpu().numpy().tobytes() == want["reward"].tobytes(), s
        assert np.array_equal(env.terminated[:, 0].cpu().numpy(), want["terminated"])
        assert np.array_equal(env.truncated[:, 0].cpu().numpy(), want["truncated"])
        assert np.array_equal(env.pos.cpu().numpy(), ob.pos()) and np.array_equal(env.money.cpu().numpy(), ob.money())
        assert np.array_equal(env.graph_id.cpu().numpy(), np.asarray(ob.graph_id, dtype=np.int32))
        assert np.array_equal(env.visits.cpu().numpy(), ob.visits())
        assert np.array_equal(env.action_mask.cpu().numpy(), ob.masks())
        assert np.array_equal(env.node_features.cpu().numpy(), ob.node_features())
        assert np.array_equal(env.mrx_revealed.cpu().numpy(), ob.revealed())
        assert np.abs(env.belief_map.cpu().numpy().astype(np.float64) - ob.belief()).max() <= BELIEF_TOL
    assert env.stats()["truncations"] > 0
    env.close()


def test_rollout_random_and_cuda_graph_replay(torch_cuda, tables):
    """sy_rollout_random (T steps issued from C) == T python-level sample + step calls == the same steps replayed
    from a captured CUDA graph (every launch of the library is stream-ordered and capturable)."""
    torch = torch_cuda
    pkg = _pkg()
    c = dict(N=30, E=55, P=3, money=10, G=2, B=100, kw=dict(belief=True, reveal_interval=4, tolls=1), mode="fp64")
    a, _ = _make_pair(pkg, c, tables)
    b, _ = _make_pair(pkg, c, tables)
    g, _ = _make_pair(pkg, c, tables)
    for e in (a, b, g):
        e.reset()
    T = 24
    for s in range(T):
        a.step(a.sample_actions(step_counter=s))
    b.rollout_random(T, step_counter=0)
    # CUDA graph: capture 4 steps, replay 6 times (step counters 0..3 are baked in, so compare against the same)
    ref, _ = _make_pair(pkg, c, tables)
    ref.reset()
    side = torch.cuda.Stream()
    acts = torch.empty(c["B"], c["P"] + 1, dtype=torch.int64, device="cuda")
    graph = torch.cuda.CUDAGraph()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        g.rollout_random(1, actions=acts, step_counter=0)  # warm-up outside the capture
        ref.rollout_random(1, step_counter=0)
    torch.cuda.current_stream().wait_stream(side)
    with torch.cuda.graph(graph, stream=side):
        g.rollout_random(4, actions=acts, step_counter=1)
    for _ in range(3):
        graph.replay()
        ref.rollout_random(4, step_counter=1)
    torch.cuda.synchronize()
    # device-resident step counter (sy_rollout_random_dev): the replayed graph keeps drawing fresh actions, so
    # 4 warm-up steps + 5 replays of a 4-step graph are the same 24 steps as the python loop
    d, _ = _make_pair(pkg, c, tables)
    d.reset()
    dgraph, counter = d.capture_rollout(4)
    for _ in range(5):
        dgraph.replay()
    torch.cuda.synchronize()
    assert int(counter.item()) == T
    for e in (a, b, g, ref, d):
        e.stats()  # folds the library's accumulators into stats_vec
    for k in ("pos", "money", "timestep", "episode", "visits", "belief_map", "reward64", "terminated", "action_mask",
              "node_features", "stats_vec"):
        assert getattr(a, k).cpu().numpy().tobytes() == getattr(b, k).cpu().numpy().tobytes(), k
        assert getattr(g, k).cpu().numpy().tobytes() == getattr(ref, k).cpu().numpy().tobytes(), k
        assert getattr(a, k).cpu().numpy().tobytes() == getattr(d, k).cpu().numpy().tobytes(), ("graph + device counter", k)
    for e in (a, b, g, ref, d):
        e.close()


def test_int32_wire_format_equals_int64(torch_cuda, tables):
    """sy_step_i32 / sy_step_host_i32 / sy_sample_actions_i32: the narrow action format gives identical results."""
    torch = torch_cuda
    pkg = _pkg()
    c = dict(N=30, E=55, P=3, money=10, G=2, B=70, kw=dict(belief=True, reveal_interval=4), mode="fp64")
    a, _ = _make_pair(pkg, c, tables)
    b, _ = _make_pair(pkg, c, tables)
    h, _ = _make_pair(pkg, c, tables)
    for e in (a, b, h):
        e.reset()
    a32 = torch.empty(c["B"], c["P"] + 1, dtype=torch.int32, device="cuda")
    for s in range(15):
        acts = a.sample_actions(step_counter=s)
        b.sample_actions(out=a32, step_counter=s)
        assert torch.equal(acts.to(torch.int32), a32)
        junk = acts.clone()
        junk[::7] = 10_000 + s  # out of range in both formats
        a.step(junk)
        b.step(junk.to(torch.int32))
        hacts = h.sample_actions_host(step_counter=s, dtype=torch.int32).clone()
        assert hacts.dtype == torch.int32
        hacts[::7] = 10_000 + s
        res = h.step_host(hacts)
        for k in ("pos", "money", "reward64", "terminated", "action_mask", "belief_map"):
            assert getattr(a, k).cpu().numpy().tobytes() == getattr(b, k).cpu().numpy().tobytes(), (k, s)
            assert getattr(a, k).cpu().numpy().tobytes() == getattr(h, k).cpu().numpy().tobytes(), (k, s)
        assert res["reward"].numpy().tobytes() == a.reward.cpu().numpy().tobytes()
    for e in (a, b, h):
        e.close()


def test_int16_wire_and_compact_status(torch_cuda, tables):
    """sy_step_i16 / sy_step_host_i16 / sy_sample_actions_i16 and the one-byte-per-env status output: identical
    dynamics, and the status bits expand to exactly the three per-agent flag arrays (also for frozen envs)."""
    torch = torch_cuda
    pkg = _pkg()
    c = dict(N=30, E=55, P=3, money=6, G=2, B=70, kw=dict(belief=True, reveal_interval=4), mode="fp64")
    a, _ = _make_pair(pkg, c, tables, auto_reset=False, max_timestep=12)  # no auto-reset: finished envs freeze
    b, _ = _make_pair(pkg, c, tables, auto_reset=False, max_timestep=12)
    h, _ = _make_pair(pkg, c, tables, auto_reset=False, max_timestep=12)
    for e in (a, b, h):
        e.reset()
    a16 = torch.empty(c["B"], c["P"] + 1, dtype=torch.int16, device="cuda")
    seen = set()
    for s in range(20):
        acts = a.sample_actions(step_counter=s)
        b.sample_actions(out=a16, step_counter=s)
        assert torch.equal(acts.to(torch.int16), a16)
        junk = acts.clone()
        junk[::5] = 20_000 + s  # out of range, representable in int16
        a.step(junk)
        b.step(junk.to(torch.int16))
        hacts = h.sample_actions_host(step_counter=s, dtype=torch.int16).clone()
        hacts[::5] = 20_000 + s
        res = h.step_host(hacts, flags="compact")
        assert set(res) == {"reward", "winner", "status"}
        for k in ("pos", "money", "reward64", "terminated", "truncated", "done_flags", "action_mask", "belief_map", "status"):
            assert getattr(a, k).cpu().numpy().tobytes() == getattr(b, k).cpu().numpy().tobytes(), (k, s)
            assert getattr(a, k).cpu().numpy().tobytes() == getattr(h, k).cpu().numpy().tobytes(), (k, s)
        assert res["reward"].numpy().tobytes() == a.reward.cpu().numpy().tobytes()
        assert torch.equal(res["status"], a.status.cpu()) and torch.equal(res["winner"], a.winner.cpu())
        ex = h.expand_status(res["status"])
        assert torch.equal(ex["terminated"], a.terminated.cpu()) and torch.equal(ex["truncated"], a.truncated.cpu())
        assert torch.equal(ex["done"], a.done_flags.cpu())
        seen |= set(res["status"].tolist())
    assert {0, 1, 2, 4} <= seen  # running, terminated, truncated and frozen envs all occurred
    for e in (a, b, h):
        e.close()


def _custom_pair(pkg, tables, graph_spec, P, money, B, mode="fp64", **kw):
    env = pkg.BatchedScotlandYardEnv(B, P, money, graphs=[graph_spec], seed=23, auto_reset=True, reward_mode=mode,
                                     keep_reward64=True, reward_tables=tables, max_timestep=kw.pop("max_timestep", 250), **kw)
    ocfg = so.OracleConfig(num_police=P, agent_money=money, reward_mode=mode, reveal_interval=kw.get("reveal_interval", 0),
                           toll=kw.get("tolls", 0), belief=kw.get("belief", False), max_timestep=env.config.max_timestep,
                           exp_table=tables[0], cov_table=tables[1])
    ob = so.OracleBatch.from_seed(ocfg, [so.Graph(graph_spec.num_nodes, graph_spec.edge_links, graph_spec.edges)], B, seed=23,
                                  auto_reset=True)
    return env, ob


@pytest.mark.parametrize("mode", ["fp64", "fp32"])
def test_disconnected_graph_with_isolated_nodes(torch_cuda, tables, mode):
    """Graphs the reference generator never produces but its code handles: unreachable pairs give distance inf
    (pathfinding.py:133-137) -> -1/(inf+1) = -0.0, exp(-inf) = 0, mean(inf) = inf; isolated nodes have no moves and
    keep their belief mass (belief_module.py:93-97)."""
    pkg = _pkg()
    # two components {0..4}, {5..8}, isolated nodes 9, 10
    links = [[0, 1], [1, 2], [2, 3], [3, 4], [0, 4], [5, 6], [6, 7], [7, 8], [5, 8]]
    g = pkg.GraphSpec(11, links, [1, 2, 3, 4, 2, 1, 1, 3, 2])
    c = dict(kw=dict(belief=True, reveal_interval=3, tolls=1), mode=mode)
    env, ob = _custom_pair(pkg, tables, g, P=3, money=9, B=96, mode=mode, belief=True, reveal_interval=3, tolls=1)
    env.reset()
    _compare_state(env, ob, c, "reset")
    saw_inf = False
    for s in range(40):
        acts = env.sample_actions(step_counter=s)
        a_h = acts.cpu().numpy()
        assert np.array_equal(a_h, ob.sample_actions(s)), s
        env.step(acts)
        want = ob.step(a_h)
        _compare_out(env, want, c, s)
        _compare_state(env, ob, c, s)
        D = ob.graphs[0].apsp()
        pos = ob.pos()
        saw_inf |= bool((D[pos[:, 0], pos[:, 1]] >= so.INF_U16).any())
    assert saw_inf, "the test never exercised an unreachable pair"
    env.close()


def test_crowded_and_broke_edge_cases(torch_cuda, tables):
    """N == A (every node occupied: nobody can ever move onto a free node except by capture rules), police that start
    broke (money 0 -> all skipped -> out-of-money ending at the first step, reward_calculator.py:75-79), toll larger
    than any budget, single police."""
    pkg = _pkg()
    ring = pkg.GraphSpec(5, [[0, 1], [1, 2], [2, 3], [3, 4], [0, 4]], [1, 2, 1, 3, 2])
    for P, money, kw in [(4, 6, {}), (2, 0, {}), (3, 5, dict(tolls=9)), (1, 4, dict(belief=True))]:
        c = dict(kw=kw, mode="fp64")
        env, ob = _custom_pair(pkg, tables, ring, P=P, money=money, B=64, **dict(kw))
        env.reset()
        for s in range(25):
            acts = env.sample_actions(step_counter=s)
            a_h = acts.cpu().numpy()
            assert np.array_equal(a_h, ob.sample_actions(s)), (P, money, s)
            env.step(acts)
            want = ob.step(a_h)
            _compare_out(env, want, c, (P, money, s))
            _compare_state(env, ob, c, (P, money, s))
        st = env.stats()
        if money == 0 or kw.get("tolls", 0) > money:
            assert st["out_of_money"] == st["episodes"] == 64 * 25  # every step ends an episode
        env.close()


def test_rollout_collector_with_a_policy(torch_cuda, tables):
    """the batched rollout loop with a policy in it (a hand-written `chase MrX` heuristic over the
    batched observation: logits = -distance to the last known MrX node) -- actions are legal or DEFAULT_ACTION, the
    stored trajectory is consistent with the env, and the heuristic police catch MrX more often than random play."""
    torch = torch_cuda
    pkg = _pkg()
    N, P, B, T = 40, 4, 512, 60
    pool = pkg.generate_graph_pool(1, N, 75, seed=8)
    mk = lambda: pkg.BatchedScotlandYardEnv(B, P, 30, graphs=pool, seed=2, auto_reset=True, reward_tables=tables)  # noqa: E731
    env = mk()
    env.reset()
    D = torch.from_numpy(env.graph_tables(0)[1].astype(np.int64)).cuda().float()  # [N, N]
    gd = pkg.batched_graph_data(env)
    assert gd["x"].data_ptr() == env.node_features.data_ptr() and gd["edge_index"].shape == (1, 2, len(pool[0].edges))

    def chase(obs):
        mrx = obs["MrX_pos"].long()  # [B]
        to_mrx = -D[mrx]  # [B, N]: police prefer nodes close to MrX
        away = D[obs["Polices_pos"].long()].min(dim=1).values  # [B, N]: MrX prefers nodes far from the nearest police
        return torch.cat([away.unsqueeze(1), to_mrx.unsqueeze(1).expand(-1, P, -1)], dim=1)

    col = pkg.RolloutCollector(env, chase, T, greedy=True)
    traj = col.collect()
    acts, pos0, money0 = traj["actions"], traj["pos"], traj["money"]
    W = torch.from_numpy(env.graph_tables(0)[0].astype(np.int64)).cuda()
    w = W[pos0.long(), acts.clamp_min(0)]  # weight of the chosen edge
    legal = (w > 0) & (w <= money0)
    assert bool((legal | (acts == -1)).all()), "the collector produced an illegal action"
    assert traj["reward"].shape == (T, B, P + 1) and bool(torch.isfinite(traj["reward"]).all())
    chase_stats = env.metrics()
    rnd = mk()
    rnd.reset()
    rnd.rollout_random(T)
    rnd_stats = rnd.metrics()
    assert chase_stats["num_episodes"] > 0 and rnd_stats["num_episodes"] > 0
    assert chase_stats["police_wins"] / chase_stats["num_episodes"] > rnd_stats["police_wins"] / rnd_stats["num_episodes"]
    env.close()
    rnd.close()


def test_million_envs_config5_scale(torch_cuda):
    """BASELINE config 5's batch (1 048 576 envs of the 200-node / 6-police config) on one GPU: index arithmetic beyond
    2^31 bytes per tensor, statistics add up, observations stay consistent with the state."""
    torch = torch_cuda
    pkg = _pkg()
    N, P, B = 200, 6, 1 << 20
    env = pkg.BatchedScotlandYardEnv(B, P, 20, graph_nodes=N, graph_edges=400, seed=4, auto_reset=True, tolls=1, belief=True,
                                     reveal_interval=5)
    env.reset()
    assert env.node_features.numel() * 4 > 2**31  # 5.9 GB tensor
    env.rollout_random(6)
    st = env.stats()
    assert st["env_steps"] == 6 * B and st["episodes"] == int(env.episode.sum())
    # the last tile is as healthy as the first: one-hot features, masks consistent with budgets, belief rows sum to 1
    W = torch.from_numpy(env.graph_tables(0)[0].astype(np.int32)).cuda()
    for sl in (slice(0, 4096), slice(B - 4096, B)):
        pos, money = env.pos[sl].long(), env.money[sl]
        rows = W[pos]
        assert bool((((rows > 0) & (rows + 1 <= money.unsqueeze(-1))) == env.action_mask[sl]).all())
        nf = env.node_features[sl]
        assert bool((nf.sum(dim=1)[:, 1:] == 1).all()) and bool((nf[:, :, 1:].gather(1, pos[:, None, 1:]) == 1).all())
        assert float((env.belief_map[sl].sum(dim=1) - 1).abs().max()) < 1e-5
    assert float(env.node_features.sum()) == float((P + (env.mrx_revealed >= 0).float()).sum())
    env.close()


@pytest.mark.gpu
@pytest.mark.parametrize("N,E,G,P", [(50, 110, 24, 3), (200, 400, 6, 6), (12, 11, 5, 2), (33, 80, 40, 4)])
def test_device_graph_sampler_matches_oracle(torch_cuda, tables, N, E, G, P):
    """the pool sampled ON THE DEVICE (sy_generate_graphs: ConnectedGraph.sample + the edge-count
    resample loop, graph_layout.py:9-80 / yard.py:67-101) is, draw for draw, the oracle's counter-based sampler
    (whose distribution is checked against the reference sampler in tests/test_graph_sampler.py): same edge lists in
    the reference's order, same attempts, same dense weights / APSP; a rollout on the sampled graphs (belief on, so the
    device-built neighbour lists are exercised) matches the oracle; pool refresh and pool sharding reproduce."""
    pkg = _pkg()
    seed, B = 21, 70
    kw = dict(belief=True, reveal_interval=4, tolls=1)
    env = pkg.BatchedScotlandYardEnv(B, P, 12, graph_nodes=N, graph_edges=E, graphs="device", num_graphs=G, seed=seed,
                                     auto_reset=True, reward_mode="fp64", keep_reward64=True, reward_tables=tables, **kw)

    def check_pool(e, generation, offset=0):
        want_pool = so.philox_graph_pool(seed, len(e.graphs), N, E, generation=generation, graph_offset=offset)
        for g, (got, want) in enumerate(zip(e.graphs, want_pool)):
            assert np.array_equal(got.edge_links, want.edge_links), (generation, g)
            assert np.array_equal(got.edges, want.edges), (generation, g)
            W, D = e.graph_tables(g)
            assert np.array_equal(W.astype(np.int64), want.weight_matrix()), (generation, g)
            assert np.array_equal(D.astype(np.int64), want.apsp()), (generation, g)
        assert len({len(g.edges) for g in e.graphs}) == 1
        return want_pool

    pool = check_pool(env, 0)
    first_edges = len(pool[0].edges)
    for g in range(1, G):  # attempts used per slot (yard.py:89-101)
        assert env.generation_attempts[g] == so.philox_sample_graph(seed, g, 0, N, E, first_edges)[1]
    for generation in (0, 1):
        if generation:
            env.regenerate_graphs()
            pool = check_pool(env, generation)
            with pytest.raises(pkg.SyError):
                env.step(env.sample_actions())  # a refreshed pool needs a reset
        ocfg = so.OracleConfig(num_police=P, agent_money=12, reward_mode="fp64", reveal_interval=4, toll=1, belief=True,
                               exp_table=tables[0], cov_table=tables[1])
        ob = so.OracleBatch.from_seed(ocfg, pool, B, seed=seed, auto_reset=True)
        env.reset()
        c = dict(kw=kw, mode="fp64")
        _compare_state(env, ob, c, ("reset", generation))
        for s in range(6):
            acts = env.sample_actions(step_counter=s)
            a_h = acts.cpu().numpy()
            assert np.array_equal(a_h, ob.sample_actions(s))
            env.step(acts)
            want = ob.step(a_h)
            _compare_out(env, want, c, ("out", generation, s))
            _compare_state(env, ob, c, ("state", generation, s))
    env.close()
    if G >= 6:  # the second half of the pool as its own shard
        half = pkg.BatchedScotlandYardEnv(8, P, 12, graph_nodes=N, graph_edges=E, graphs="device", num_graphs=G - G // 2,
                                          seed=seed, graph_offset=G // 2)
        check_pool(half, 0, offset=G // 2)
        half.close()


@pytest.mark.gpu
def test_device_graph_sampler_large_graphs(torch_cuda):
    """BASELINE config 4 sized graphs (1000 nodes / 2000 edges) sampled on the device: edge lists equal the oracle's
    counter-based sampler draw for draw, the all-pairs table equals scipy's Dijkstra on the sampled graph."""
    from scipy.sparse import csr_matrix
    from scipy.sparse.csgraph import dijkstra

    pkg = _pkg()
    N, E, G, seed = 1000, 2000, 3, 77
    env = pkg.BatchedScotlandYardEnv(64, 6, 20, graph_nodes=N, graph_edges=E, graphs="device", num_graphs=G, seed=seed,
                                     belief=True)
    want_pool = so.philox_graph_pool(seed, G, N, E)
    for g in range(G):
        assert np.array_equal(env.graphs[g].edge_links, want_pool[g].edge_links), g
        assert np.array_equal(env.graphs[g].edges, want_pool[g].edges), g
        W, D = env.graph_tables(g)
        assert np.array_equal(W.astype(np.int64), want_pool[g].weight_matrix()), g
        ref = dijkstra(csr_matrix(W.astype(np.float64)), directed=False)
        assert np.isfinite(ref).all() and np.array_equal(D.astype(np.float64), ref), g
    env.reset()
    for s in range(5):  # the generic (large-N) belief path runs on the device-built neighbour lists
        env.step(env.sample_actions(step_counter=s))
    assert float((env.belief_map.sum(dim=1) - 1).abs().max()) < 1e-5
    env.close()


@pytest.mark.gpu
def test_interleaved_graph_pool_matches_oracle(torch_cuda, tables):
    """Every 32-env tile holds a mix of graphs (explicit graph ids e % G instead of the default blocks of 32 envs):
    the writers' per-env CSR walk and the generic warp-per-env belief path against the oracle, incl. belief scoring."""
    pkg = _pkg()
    c = CASES[2]
    pool = pkg.generate_graph_pool(c["G"], c["N"], c["E"], seed=3)
    B, seed = 130, 29
    env = pkg.BatchedScotlandYardEnv(B, c["P"], c["money"], graphs=pool, seed=seed, auto_reset=True, reward_mode=c["mode"],
                                     keep_reward64=True, reward_tables=tables, belief_ce=True, **c["kw"])
    gid = np.arange(B) % c["G"]
    env.reset(graph_id=gid)
    ocfg = so.OracleConfig(num_police=c["P"], agent_money=c["money"], reward_mode=c["mode"], reveal_interval=c["kw"]["reveal_interval"],
                           toll=c["kw"]["tolls"], belief=True, exp_table=tables[0], cov_table=tables[1])
    ograph = [so.Graph(g.num_nodes, g.edge_links, g.edges) for g in pool]
    starts = [so.philox_start_positions(seed, e, 0, c["N"], c["P"] + 1) for e in range(B)]
    ob = so.OracleBatch(ocfg, ograph, gid.tolist(), starts, seed=seed, auto_reset=True)
    _compare_state(env, ob, c, "reset")
    for s in range(20):
        acts = env.sample_actions(step_counter=s)
        a_h = acts.cpu().numpy()
        assert np.array_equal(a_h, ob.sample_actions(s))
        env.step(acts)
        want = ob.step(a_h)
        _compare_out(env, want, c, ("out", s))
        _compare_state(env, ob, c, ("state", s))
    ces = np.asarray(ob.belief_ces)
    st = env.stats()
    assert st["reveals"] == len(ces) > 0
    np.testing.assert_allclose(env.metrics()["mean_belief_ce"], ces.mean(), rtol=1e-5, atol=1e-6)
    env.close()


@pytest.mark.gpu
def test_host_overlap_mode_is_equivalent(torch_cuda, tables):
    """sy_set_host_overlap: step_host returns when the results are on the host (the observation kernel may still be
    running) and sample_actions_host runs next to it on the library stream -- same trajectory, results and
    observations as the fully synchronous calls, also when resets / device steps are mixed in."""
    torch = torch_cuda
    pkg = _pkg()
    c = dict(N=30, E=55, P=3, money=8, G=2, B=300, kw=dict(belief=True, reveal_interval=4, tolls=1), mode="fp64")
    a, _ = _make_pair(pkg, c, tables)
    b, _ = _make_pair(pkg, c, tables)
    a.reset()
    b.reset()
    b.set_host_overlap(True)
    for s in range(40):
        ha = a.sample_actions_host(step_counter=s, dtype=torch.int16)
        hb = b.sample_actions_host(step_counter=s, dtype=torch.int16)
        assert torch.equal(ha, hb), s
        ra = a.step_host(ha, flags="compact")
        rb = b.step_host(hb, flags="compact")
        for k in ("reward", "winner", "status"):
            assert torch.equal(ra[k], rb[k]), (k, s)
        if s % 7 == 3:  # observations read on the stream are ordered behind the still-running kernel
            for k in ("pos", "money", "action_mask", "node_features", "belief_map"):
                assert torch.equal(getattr(a, k), getattr(b, k)), (k, s)
        if s == 20:  # a device-side step and a partial reset in between: the library re-orders its stream behind them
            acts = a.sample_actions(step_counter=1000)
            a.step(acts)
            b.step(b.sample_actions(step_counter=1000))
            m = torch.zeros(c["B"], dtype=torch.bool, device="cuda")
            m[::3] = True
            a.reset(reset_mask=m)
            b.reset(reset_mask=m)
    torch.cuda.synchronize()
    for k in ("pos", "money", "timestep", "episode", "visits", "action_mask", "node_features", "belief_map", "reward64"):
        assert getattr(a, k).cpu().numpy().tobytes() == getattr(b, k).cpu().numpy().tobytes(), k
    a.close()
    b.close()


@pytest.mark.gpu
def test_uint8_node_features_option(torch_cuda, tables):
    """node_features_dtype=torch.uint8 (SyObs.node_features_u8): the same one-hot as bytes, everything else untouched."""
    torch = torch_cuda
    pkg = _pkg()
    for c in (CASES[1], CASES[2], CASES[5]):
        pool = pkg.generate_graph_pool(c["G"], c["N"], c["E"], seed=3)
        mk = lambda dt: pkg.BatchedScotlandYardEnv(c["B"], c["P"], c["money"], graphs=pool, seed=5, auto_reset=True,  # noqa: E731
                                                   reward_mode=c["mode"], keep_reward64=True, reward_tables=tables,
                                                   node_features_dtype=dt, **c["kw"])
        a, b = mk(torch.float32), mk(torch.uint8)
        a.reset()
        b.reset()
        assert b.node_features.dtype == torch.uint8 and b.observation()["node_features"].dtype == torch.uint8
        for s in range(12):
            acts = a.sample_actions(step_counter=s)
            a.step(acts)
            b.step(acts)
            assert torch.equal(a.node_features, b.node_features.float()), s
            for k in ("pos", "money", "action_mask", "reward64", "terminated", "mrx_revealed") + (("belief_map",) if c["kw"].get("belief") else ()):
                assert getattr(a, k).cpu().numpy().tobytes() == getattr(b, k).cpu().numpy().tobytes(), (k, s)
        a.close()
        b.close()


@pytest.mark.gpu
@pytest.mark.parametrize("step_kernel", ["two_kernels", "fused"])
@pytest.mark.parametrize("mixed_graphs", [False, True])
def test_belief_hint_and_reveal_skip_match_oracle(torch_cuda, tables, step_kernel, mixed_graphs):
    """The two remaining parts of the belief definition: the observation-hint re-weighting of
    ParticleBeliefTracker.update (belief_module.py:102-106: likelihood 0.1 + 0.9 * [j in hint], then normalise) and the
    reveal-skip robustness hook (ood_eval.py:227-229: a scheduled reveal is skipped with probability p).  Device fp32 vs
    the float64 oracle within 1e-6; the skip decisions (one Philox draw per env and reveal step) bit-exact.  Mixed graphs
    per tile take the generic belief path, a single graph per tile the lane = env fast path."""
    torch = torch_cuda
    pkg = _pkg()
    N, E, P, B, G = 30, 55, 3, 96, (3 if mixed_graphs else 1)
    pool = pkg.generate_graph_pool(G, N, E, seed=3)
    kw = dict(reveal_interval=3, tolls=1, belief=True)
    env = pkg.BatchedScotlandYardEnv(B, P, 10, graphs=pool, seed=23, auto_reset=True, keep_reward64=True, reward_tables=tables,
                                     resample_graph=mixed_graphs, reveal_skip_prob=0.4, **kw)
    env.set_option("step_kernel", step_kernel)
    cfg = so.OracleConfig(num_police=P, agent_money=10, reveal_interval=3, toll=1, belief=True, reveal_skip_prob=0.4,
                          exp_table=tables[0], cov_table=tables[1])
    ob = so.OracleBatch.from_seed(cfg, [so.Graph(g.num_nodes, g.edge_links, g.edges) for g in pool], B, seed=23, auto_reset=True,
                                  resample_graph=mixed_graphs)
    env.reset()
    rng = np.random.default_rng(4)
    skipped = revealed = 0
    for s in range(24):
        hints = rng.random((B, N)) < 0.15
        hints[::5] = False  # envs without any candidate: no re-weighting (`if observation_hint:` is False)
        use_hint = s % 3 != 2
        env.set_belief_hint(torch.from_numpy(hints).cuda() if use_hint else None)
        acts = env.sample_actions(step_counter=s)
        env.step(acts)
        want = ob.step(acts.cpu().numpy(), hints=hints if use_hint else None)
        assert env.reward64.cpu().numpy().tobytes() == want["reward"].tobytes(), s
        assert np.array_equal(env.pos.cpu().numpy(), ob.pos()) and np.array_equal(env.timestep.cpu().numpy(), ob.timestep())
        assert np.array_equal(env.mrx_revealed.cpu().numpy(), ob.revealed()), ("reveal / skip decisions", s)
        assert np.array_equal(env.node_features.cpu().numpy(), ob.node_features()), s
        err = np.abs(env.belief_map.cpu().numpy().astype(np.float64) - ob.belief()).max()
        assert err <= BELIEF_TOL, (s, err)
        t = ob.timestep()
        due = (t > 0) & (t % 3 == 0)
        revealed += int((ob.revealed()[due] >= 0).sum())
        skipped += int((ob.revealed()[due] < 0).sum())
    assert revealed > 50 and skipped > 50  # both outcomes of the Bernoulli(0.4) draw occurred
    assert abs(skipped / (skipped + revealed) - 0.4) < 0.1
    env.close()


@pytest.mark.gpu
def test_pointer_structs_of_another_abi_revision_are_rejected(torch_cuda):
    """every pointer struct carries struct_bytes = sizeof(struct); a binding built against another header revision (e.g.
    from a stale copy of INTEGRATION.md) is refused with SY_ERR_INVALID_ARGUMENT before any pointer reaches a kernel"""
    import ctypes as C

    pkg = _pkg()
    from student_mechanism_design_b200 import _cabi

    env = pkg.BatchedScotlandYardEnv(64, 2, 10, graph_nodes=15, graph_edges=20, seed=0)
    env.reset()
    acts = env.sample_actions()
    lib = _cabi.load_library()
    stream = env._stream()
    for name in ("_state", "_obs", "_out"):
        good = getattr(env, name)
        bad = type(good).from_buffer_copy(good)
        bad.struct_bytes = good.struct_bytes - 8  # the previous revision's size
        args = {"_state": (bad, env._obs, env._out), "_obs": (env._state, bad, env._out), "_out": (env._state, env._obs, bad)}[name]
        rc = lib.sy_step(env._handle, acts.data_ptr(), C.byref(args[0]), C.byref(args[1]), C.byref(args[2]), stream)
        assert rc == 1 and b"size mismatch" in lib.sy_last_error(), name
    host = env._host_buffers()
    bad = type(env._host_out).from_buffer_copy(env._host_out)
    bad.struct_bytes = 0
    rc = lib.sy_step_host(env._handle, host["actions"].data_ptr(), env._actions_dev.data_ptr(), C.byref(env._state), C.byref(env._obs),
                          C.byref(env._out), C.byref(bad), stream)
    assert rc == 1 and b"SyHostOut size mismatch" in lib.sy_last_error()
    before = env.pos.clone()
    env.step(acts)  # the handle is still usable and the rejected calls changed nothing
    assert env.timestep.max().item() == 1 and before.shape == env.pos.shape
    env.close()


@pytest.mark.gpu
def test_allreduce_stats_through_the_c_abi(torch_cuda):
    """sy_allreduce_stats: fold + one ncclAllReduce of the statistics vector on a raw ncclComm_t, here a
    1-rank communicator created through NCCL's C API (a non-PyTorch host would hand over its own); the global vector equals
    the local one, repeated calls do not double count, and the torch.distributed helper agrees."""
    import ctypes as C
    import glob

    torch = torch_cuda
    pkg = _pkg()
    from student_mechanism_design_b200 import _cabi

    cands = glob.glob(os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "nccl", "lib", "libnccl.so.2")) + ["libnccl.so.2"]
    nccl = C.CDLL(cands[0])

    class UniqueId(C.Structure):
        _fields_ = [("internal", C.c_char * 128)]

    uid, comm = UniqueId(), C.c_void_p()
    nccl.ncclGetUniqueId.argtypes = [C.POINTER(UniqueId)]
    nccl.ncclCommInitRank.argtypes = [C.POINTER(C.c_void_p), C.c_int, UniqueId, C.c_int]
    nccl.ncclCommDestroy.argtypes = [C.c_void_p]
    assert nccl.ncclGetUniqueId(C.byref(uid)) == 0
    assert nccl.ncclCommInitRank(C.byref(comm), 1, uid, 0) == 0
    env = pkg.BatchedScotlandYardEnv(512, 3, 8, graph_nodes=30, graph_edges=55, seed=2, auto_reset=True)
    env.reset()
    lib = _cabi.load_library()
    glob_vec = torch.zeros(_cabi.SY_NUM_STATS, dtype=torch.int64, device="cuda")
    for rounds in range(2):
        env.rollout_random(20)
        _cabi.check(lib.sy_allreduce_stats(env._handle, comm, env.stats_vec.data_ptr(), glob_vec.data_ptr(), env._stream()))
        torch.cuda.synchronize()
        assert torch.equal(glob_vec, env.stats_vec)
        assert int(glob_vec[0]) == 512 * 20 * (rounds + 1)  # cumulative env-steps, not double counted
    assert env.stats()["env_steps"] == 512 * 40 and env.stats()["episodes"] == int(glob_vec[1])
    assert lib.sy_allreduce_stats(env._handle, None, env.stats_vec.data_ptr(), glob_vec.data_ptr(), env._stream()) == 1
    nccl.ncclCommDestroy(comm)
    env.close()


@pytest.mark.gpu
def test_host_rollout_random_equals_the_python_host_loop(torch_cuda, tables):
    """sy_host_rollout_random (the host-buffer loop issued from C, what bench.py's e2e leg times) == the same steps driven
    from Python with sample_actions_host + step_host, in both overlap modes and for two concurrent half-batch handles on
    their own streams and threads (the ping-pong of the e2e leg)."""
    import threading

    torch = torch_cuda
    pkg = _pkg()
    c = dict(N=30, E=55, P=3, money=10, G=2, B=128, kw=dict(belief=True, reveal_interval=4, tolls=1), mode="fp64")
    K = 17
    ref, _ = _make_pair(pkg, c, tables)
    ref.reset()
    for s in range(K):
        out_ref = ref.step_host(ref.sample_actions_host(step_counter=s, dtype=torch.int16), flags="compact")
    for overlap in (False, True):
        env, _ = _make_pair(pkg, c, tables)
        env.reset()
        env.set_host_overlap(overlap)
        out = env.host_rollout_random(K, step_counter=0, dtype=torch.int16, flags="compact")
        torch.cuda.synchronize()
        for k in ("reward", "winner", "status"):
            assert out[k].numpy().tobytes() == out_ref[k].numpy().tobytes(), (overlap, k)
        for k in ("pos", "money", "timestep", "visits", "belief_map", "action_mask", "node_features"):
            assert getattr(env, k).cpu().numpy().tobytes() == getattr(ref, k).cpu().numpy().tobytes(), (overlap, k)
        env.close()
    # two half-batch handles driven concurrently reproduce the halves of the full batch (env_offset-keyed streams)
    halves = [_make_pair(pkg, c, tables, B=64, env_offset=o)[0] for o in (0, 64)]
    streams = [torch.cuda.Stream() for _ in halves]
    for h, st in zip(halves, streams):
        with torch.cuda.stream(st):
            h.reset()
            h.set_host_overlap(True)

    def drive(h, st):
        with torch.cuda.stream(st):
            h.host_rollout_random(K, step_counter=0, dtype=torch.int16, flags="compact")
            torch.cuda.current_stream().synchronize()

    ts = [threading.Thread(target=drive, args=(h, st)) for h, st in zip(halves, streams)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    for k in ("pos", "money", "timestep", "visits", "belief_map", "action_mask", "node_features", "reward"):
        got = np.concatenate([getattr(h, k).cpu().numpy() for h in halves])
        assert got.tobytes() == getattr(ref, k).cpu().numpy().tobytes(), ("halves", k)
    for e in halves + [ref]:
        e.close()
