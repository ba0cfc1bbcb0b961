"""GPU parity of the policy forward through the C ABI of include/sy_policy.h against
oracle/policy_oracle.py: the GNN Q values and epsilon-greedy actions of both agents, the MAPPO distribution / sampled
action / log-prob / critic incl. the golden vectors recorded from the unmodified reference modules."""
import os

import numpy as np
import pytest

from conftest import ROOT
from oracle import policy_oracle as po

pytestmark = pytest.mark.gpu
Q_TOL = 2e-5  # fp32 kernel vs float64 oracle


@pytest.fixture(scope="module")
def torch_cuda():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    return torch


def _pkg():
    import student_mechanism_design_b200 as pkg

    return pkg


def _roll(env, steps):
    for s in range(steps):
        env.step(env.sample_actions(step_counter=s))


@pytest.mark.parametrize("N,E,P,G,features", [(30, 55, 3, 3, "env"), (50, 110, 6, 2, "reference"), (200, 400, 6, 1, "env"),
                                               (24, 40, 15, 2, "env"), (20, 30, 2, 1, "reference")])
def test_gnn_q_values_and_actions_match_oracle(torch_cuda, N, E, P, G, features):
    torch = torch_cuda
    pkg = _pkg()
    B, A, seed = 70, P + 1, 13
    env = pkg.BatchedScotlandYardEnv(B, P, 8, graph_nodes=N, graph_edges=E, num_graphs=G, seed=seed, auto_reset=True,
                                     reveal_interval=3, tolls=1, env_offset=1000)
    env.reset()
    _roll(env, 7)  # spread the agents, spend budgets, hide MrX in most envs
    pol = pkg.GNNPolicy(env, features=features, seed=5)
    # give the biases some weight so the "untouched node" constant is not trivially zero
    gen = torch.Generator().manual_seed(1)
    for sd in pol.state:
        sd["conv1.bias"] = torch.randn(A, generator=gen) * 0.3
        sd["conv2.bias"] = torch.randn(A, generator=gen) * 0.3
    pol._pack()
    q = pol.q_values().cpu().numpy()
    pos, money = env.pos.cpu().numpy(), env.money.cpu().numpy()
    rev, gid = env.mrx_revealed.cpu().numpy(), env.graph_id.cpu().numpy()
    mode = po.FEATURES_ENV if features == "env" else po.FEATURES_REFERENCE
    want_q = np.zeros_like(q, dtype=np.float64)
    for b in range(B):
        x = po.graph_features(mode, pos[b], rev[b], N, A)
        for m in range(2):
            want_q[b, m] = po.gnn_forward(x, env.graphs[gid[b]].edge_links, {k: v.numpy() for k, v in pol.state[m].items()})
    assert np.abs(q - want_q).max() <= Q_TOL, np.abs(q - want_q).max()
    assert (rev < 0).any() and (rev >= 0).any()
    mask = env.action_mask.cpu().numpy()
    for eps_m, eps_p, step in ((0.0, 0.0, 3), (1.0, 1.0, 4), (0.3, 0.6, 5)):
        acts, qt = pol.act(eps_m, eps_p, step_counter=step, return_q=True)
        acts, qt = acts.cpu().numpy(), qt.cpu().numpy()
        n_explored = 0
        for b in range(B):
            for a in range(A):
                moves = np.nonzero(mask[b, a])[0]
                want, explored, ok = po.gnn_select(want_q[b, 0 if a == 0 else 1], moves, eps_m if a == 0 else eps_p, seed,
                                                   1000 + b, step, a, tie_tol=2 * Q_TOL)
                assert acts[b, a] in ok, (b, a, acts[b, a], want)
                n_explored += explored
                if explored or want < 0:
                    assert np.isnan(qt[b, a])
                else:
                    assert abs(qt[b, a] - want_q[b, 0 if a == 0 else 1][acts[b, a]]) <= Q_TOL
        assert (n_explored == 0) if eps_m == eps_p == 0.0 else n_explored > 0
    # the chosen actions are legal moves for the env: stepping with them never leaves an agent off the graph
    env.step(pol.act(0.1, 0.1))
    env.close()


def test_gnn_policy_rollout_and_state_dict(torch_cuda):
    """a few hundred policy-driven steps with auto-reset: actions always valid (or -1 when broke), episodes finish"""
    torch = torch_cuda
    pkg = _pkg()
    env = pkg.BatchedScotlandYardEnv(4096, 6, 20, graph_nodes=200, graph_edges=400, seed=3, auto_reset=True, tolls=1,
                                     belief=True, reveal_interval=5)
    env.reset()
    pol = pkg.GNNPolicy(env, seed=9)
    pol.load_state_dicts(pol.state[0], pol.state[1])
    for s in range(60):
        acts = pol.act(0.05, 0.05)
        valid = env.action_mask.gather(2, acts.clamp_min(0).unsqueeze(-1)).squeeze(-1)
        has_move = env.action_mask.any(dim=-1)
        assert bool(((acts >= 0) == has_move).all()) and bool(valid[acts >= 0].all())
        env.step(acts)
    assert env.stats()["episodes"] > 0
    env.close()


def test_gnn_policy_in_the_rollout_collector(torch_cuda):
    """row f1 + f2 together: RolloutCollector drives the env with the GNN agents' own kernel (no dense logits)."""
    torch = torch_cuda
    pkg = _pkg()
    env = pkg.BatchedScotlandYardEnv(512, 3, 12, graph_nodes=40, graph_edges=75, seed=4, auto_reset=True)
    env.reset()
    pol = pkg.GNNPolicy(env, seed=1)
    pol.epsilon = (0.2, 0.1)
    traj = pkg.RolloutCollector(env, pol, 30).collect()
    acts, pos0, money0 = traj["actions"], traj["pos"], traj["money"]
    W = torch.from_numpy(env.graph_tables(0)[0].astype(np.int64)).cuda()
    w = W[pos0.long(), acts.clamp_min(0)]
    assert bool((((w > 0) & (w <= money0)) | (acts == -1)).all())
    assert bool((acts >= 0).any()) and env.stats()["episodes"] > 0
    # the collector stepped deferred (GNNPolicy reads only the compact state) and flushed: a twin env stepped plainly with
    # the recorded actions ends in the same state AND the same dense observations
    assert not env.observations_pending
    twin = pkg.BatchedScotlandYardEnv(512, 3, 12, graph_nodes=40, graph_edges=75, seed=4, auto_reset=True)
    twin.reset()
    for t in range(30):
        assert torch.equal(twin.pos, pos0[t]) and torch.equal(twin.money, money0[t])
        twin.step(acts[t].contiguous())
        assert torch.equal(twin.reward, traj["reward"][t])
    for k in ("pos", "money", "timestep", "visits", "action_mask", "node_features", "agent_budget", "mrx_revealed"):
        assert torch.equal(getattr(twin, k), getattr(env, k)), k
    twin.close()
    env.close()


def test_mappo_matches_oracle_and_reference_goldens(torch_cuda):
    torch = torch_cuda
    pkg = _pkg()
    gold = np.load(os.path.join(ROOT, "tests", "golden", "mappo.npz"))
    # (1) live env: distribution, sampled action, log-prob vs the oracle
    for N, E, P, H, D in ((50, 110, 3, 64, 7), (200, 400, 6, 64, 9), (33, 60, 2, 40, 5)):
        B, A, seed = 150, P + 1, 4
        env = pkg.BatchedScotlandYardEnv(B, P, 6, graph_nodes=N, graph_edges=E, num_graphs=2, seed=seed, auto_reset=True, tolls=1)
        env.reset()
        _roll(env, 9)
        mp = pkg.MappoPolicy(env, obs_size=D, hidden_size=H, global_obs_size=D * A, seed=2)
        with torch.no_grad():  # push one policy into the "mass underflows" branch now and then
            mp.policies[1]["actor.2.bias"][:4] += 70.0
        mp2 = pkg.MappoPolicy(env, obs_size=D, hidden_size=H, policies=mp.policies, policy_of_agent=list(range(A)),
                              critic=mp.critic, global_obs_size=D * A)
        obs = torch.randn(B, A, D, generator=torch.Generator().manual_seed(7)).cuda()
        mask = env.action_mask.cpu().numpy()
        obs_h = obs.cpu().numpy()
        for tensor_cores in (True, False):  # tcgen05 3xTF32 path and the CUDA-core path
            mp2.tensor_cores = tensor_cores
            acts, lp, probs = mp2.act(obs, step_counter=11, return_probs=True)
            mp2.check()
            acts, lp, probs = acts.cpu().numpy(), lp.cpu().numpy(), probs.cpu().numpy()
            for b in range(B):
                for a in range(A):
                    sd = {k: v.numpy() for k, v in mp.policies[a].items()}
                    want = po.mappo_probs(obs_h[b, a], sd, mask[b, a].astype(np.float64))
                    np.testing.assert_allclose(probs[b, a], want, rtol=3e-5, atol=1e-9, err_msg=f"tc={tensor_cores} {b} {a}")
                    assert acts[b, a] == po.mappo_sample(probs[b, a], seed, b, 11, a), (tensor_cores, b, a)
                    np.testing.assert_allclose(lp[b, a], po.categorical_log_prob(probs[b, a], acts[b, a]), rtol=1e-5, atol=2e-6)
        gobs = torch.randn(37, D * A, generator=torch.Generator().manual_seed(8)).cuda()
        v = mp2.values(gobs).cpu().numpy()
        want_v = [po.critic_value(x, {k: t.numpy() for k, t in mp.critic.items()}) for x in gobs.cpu().numpy()]
        np.testing.assert_allclose(v, want_v, rtol=1e-5, atol=1e-6)
        env.close()
    # (2) the reference's own outputs (golden): same distribution and log-prob for the reference's sampled action.
    # The env only supplies masks here, so each golden mask is re-created as "valid moves" of a star graph:
    # node 0 is adjacent to exactly the masked nodes and the agent stands on node 0 (node 0 itself is never masked in).
    for ci, (n_agents, D, H, N) in enumerate(gold["cases"].tolist()):
        obs_all, mask_all = gold[f"c{ci}_obs"], gold[f"c{ci}_mask"]
        for t in range(len(obs_all)):
            mask = mask_all[t].copy()
            if mask[0] == 1 or mask.sum() == 0:
                continue  # cannot be expressed as neighbours of node 0 / covered by the live test
            nbrs = np.nonzero(mask)[0]
            rest = [int(n) for n in range(1, N) if mask[n] == 0]  # hang the other nodes off as a chain (small degrees)
            links = [(0, int(n)) for n in nbrs] + list(zip([int(nbrs[0])] + rest[:-1], rest))
            g = pkg.GraphSpec(N, np.asarray(links), np.ones(len(links), dtype=np.int64))
            env = pkg.BatchedScotlandYardEnv(1, 1, 50, graphs=[g], seed=1)
            env.reset(init_pos=np.asarray([[int(nbrs[0]), 0]], dtype=np.int32))  # MrX on a neighbour, the officer on node 0
            pid = int(gold[f"c{ci}_pid"][t])
            sd = {k: torch.from_numpy(gold[f"c{ci}_p{pid}_{k}"]) for k in ("actor.0.weight", "actor.0.bias", "actor.2.weight", "actor.2.bias")}
            mp = pkg.MappoPolicy(env, obs_size=D, hidden_size=H, policies=[sd], policy_of_agent=[0, 0])
            obs = torch.from_numpy(np.stack([obs_all[t], obs_all[t]])[None]).cuda()
            _, lp, probs = mp.act(obs, return_probs=True)
            mp.check()
            got = probs[0, 1].cpu().numpy()
            # the officer may not move onto MrX's node?  the mask rule only looks at budgets (yard.py:420-472), so it can
            np.testing.assert_allclose(got, gold[f"c{ci}_probs"][t], rtol=3e-5, atol=1e-9)
            a_ref = int(gold[f"c{ci}_action"][t])
            np.testing.assert_allclose(po.categorical_log_prob(got, a_ref, np.float32), gold[f"c{ci}_logp"][t], rtol=2e-5, atol=2e-6)
            env.close()


@pytest.mark.gpu
def test_masked_sample_kernel(torch_cuda):
    """sy_masked_sample: greedy == torch's first argmax over the legal nodes; sampling == Gumbel-max with
    the documented Philox noise (recomputed on the CPU), rows without a legal node -> -1, frequencies follow
    softmax(logits | mask)."""
    import sy_oracle as so

    torch = torch_cuda
    pkg = _pkg()
    from student_mechanism_design_b200.rollout import masked_sample, masked_sample_device

    g = torch.Generator().manual_seed(1)
    B, A, N = 16, 3, 37
    logits = torch.randn(B, A, N, generator=g)
    logits[1, 0, 5] = logits[1, 0, 9] = 7.0  # a tie: the first maximum wins
    mask = torch.rand(B, A, N, generator=g) < 0.3
    mask[1, 0, 5] = mask[1, 0, 9] = True
    mask[0, 1] = False
    lg, mk = logits.cuda(), mask.cuda()
    greedy = masked_sample_device(lg, mk, greedy=True).cpu()
    assert torch.equal(greedy, masked_sample(logits, mask, greedy=True)) and greedy[0, 1] == -1 and greedy[1, 0] == 5
    seed, step, off = 1234567, 9, 100
    got = masked_sample_device(lg, mk, seed=seed, step_counter=step, env_offset=off).cpu().numpy()
    key = (seed & 0xFFFFFFFF, seed >> 32)
    for b in range(B):
        for a in range(A):
            legal = np.nonzero(mask[b, a].numpy())[0]
            if len(legal) == 0:
                assert got[b, a] == -1
                continue
            keys = np.full(N, -np.inf)
            for j in legal:
                r = so.philox4x32((off + b, step, 5 + 16 * a, j >> 2), key)[j & 3]
                u = ((r >> 8) + 0.5) / 16777216.0
                keys[j] = float(logits[b, a, j]) - np.log(-np.log(u))
            assert got[b, a] in legal and keys[got[b, a]] >= keys.max() - 1e-4, (b, a)
    # distribution: 1 row replicated over many envs and steps
    l1 = torch.tensor([0.0, 1.0, 2.0, 9.0, -1.0]).cuda().expand(4096, 1, 5).contiguous()
    m1 = torch.tensor([True, True, True, False, True]).cuda().expand(4096, 1, 5).contiguous()
    draws = torch.cat([masked_sample_device(l1, m1, seed=3, step_counter=s).flatten() for s in range(8)])
    freq = torch.bincount(draws, minlength=5).float().cpu() / draws.numel()
    want = torch.softmax(torch.tensor([0.0, 1.0, 2.0, -1e30, -1.0]), -1)
    assert freq[3] == 0 and torch.allclose(freq, want, atol=0.015)


@pytest.mark.gpu
def test_mappo_trainer_observation_built_in_the_kernel(torch_cuda):
    """sy_mappo_act with obs == NULL builds MappoTrainer's observations from the env state (mappo_trainer.py:171-199: MrX
    <- MrX_pos, officer i <- Polices_pos.sum(dim=1)): bit-identical to passing that tensor explicitly, on both kernels
    (tcgen05 and CUDA cores)."""
    torch = torch_cuda
    pkg = _pkg()
    B, P, N = 300, 4, 60
    env = pkg.BatchedScotlandYardEnv(B, P, 12, graph_nodes=N, graph_edges=110, seed=5, auto_reset=True)
    env.reset()
    env.rollout_random(3)
    A = P + 1
    pol = pkg.MappoPolicy(env, obs_size=A, hidden_size=64, seed=2)
    obs = torch.zeros(B, A, A, device="cuda")
    obs[:, 0, 0] = env.pos[:, 0].float()
    obs[:, 1:, :P] = env.pos[:, 1:].float().unsqueeze(1).expand(B, P, P)
    for tc in (True, False):
        pol.tensor_cores = tc
        a_ref, lp_ref, pr_ref = pol.act(obs, step_counter=4, return_probs=True)
        a_new, lp_new, pr_new = pol.act(None, step_counter=4, return_probs=True)
        assert torch.equal(a_ref, a_new) and torch.equal(lp_ref, lp_new) and torch.equal(pr_ref, pr_new), tc
    env.close()


@pytest.mark.gpu
def test_copy_segments_records_a_transition_in_one_launch(torch_cuda):
    torch = torch_cuda
    _pkg()
    from student_mechanism_design_b200.rollout import copy_segments

    g = torch.Generator().manual_seed(0)
    srcs = [torch.randint(0, 100, (1000, 7), generator=g, dtype=torch.int32).cuda(), torch.randn(333, generator=g).cuda(),
            torch.randint(0, 2, (77,), generator=g, dtype=torch.uint8).cuda(),
            torch.randint(0, 255, (121,), generator=g, dtype=torch.uint8).cuda()]
    base = torch.zeros(4096, dtype=torch.uint8, device="cuda")
    dsts = [torch.zeros_like(s) for s in srcs[:3]] + [base[3:3 + 121]]  # an unaligned destination: byte path
    copy_segments(dsts, srcs)
    torch.cuda.synchronize()
    for d, s in zip(dsts, srcs):
        assert torch.equal(d, s)
