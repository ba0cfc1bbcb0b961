"""CPU tests of the host-side logic around the C ABI: torchrl adapter key layout (against the stand-in base
classes of tests/fake_torchrl.py and a stub env holding CPU tensors), batch sharding, and the statistics
all-reduce on a world_size-2 gloo group."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

import fake_torchrl
from student_mechanism_design_b200 import shard_range, torchrl_env


class StubEnv:
    """the attributes the adapter reads from BatchedScotlandYardEnv, on the CPU"""

    DEFAULT_ACTION = -1

    def __init__(self, B=6, P=2, N=9, belief=True, reveal_interval=0, auto_reset=False):
        A = P + 1
        self.num_envs, self.num_agents, self.number_of_agents = B, A, P
        self.graph_nodes, self.agent_money, self.num_graphs = N, 10, 1
        self.reveal_interval, self.auto_reset, self._is_reset = reveal_interval, auto_reset, False
        self.possible_agents = ["MrX"] + [f"Police{i}" for i in range(P)]
        self.device = torch.device("cpu")
        self.belief_on = belief
        g = torch.Generator().manual_seed(0)
        self.pos = torch.randint(0, N, (B, A), dtype=torch.int32, generator=g)
        self.money = torch.randint(0, 10, (B, A), dtype=torch.int32, generator=g)
        self.node_features = torch.zeros(B, N, A)
        self.action_mask = torch.zeros(B, A, N, dtype=torch.bool)
        self.agent_budget = self.money.float()
        self.mrx_revealed = self.pos[:, 0].clone()
        if reveal_interval:
            self.mrx_revealed[::2] = -1  # hidden in every other env
        self.graph_id = torch.zeros(B, dtype=torch.int32)
        self.belief_map = torch.full((B, N), 1.0 / N)
        self.reward = torch.zeros(B, A)
        self.terminated = torch.zeros(B, A, dtype=torch.bool)
        self.truncated = torch.zeros(B, A, dtype=torch.bool)
        self.done_flags = torch.zeros(B, A, dtype=torch.bool)
        self.winner = torch.zeros(B, dtype=torch.int8)
        self.calls = []

    def reset(self, reset_mask=None, **kw):
        self.calls.append(("reset", None if reset_mask is None else reset_mask.clone(), kw))
        self._is_reset = True

    def step(self, actions):
        self.calls.append(("step", actions.clone()))
        self.reward += 1
        self.terminated[0] = True
        self.done_flags[0] = True

    def set_seed(self, seed):
        self.calls.append(("seed", seed))

    def get_possible_moves(self, agent_idx, env_index=0):
        return np.arange(2)


def _adapter(stub):
    cls = torchrl_env.make_env_class(fake_torchrl.EnvBase, fake_torchrl.TensorDict, fake_torchrl.SPECS)
    return cls(stub)


def test_adapter_declares_specs_and_passes_check_env_specs():
    """torchrl's EnvBase.rollout / check_env_specs / SyncDataCollector need observation / action / reward / done specs:
    every key the adapter emits is declared with its shape, dtype and bounds, and random actions drawn from the action
    spec step the env (against the stand-in's check_env_specs; a real torchrl is not installed here)."""
    for kw in (dict(), dict(belief=False), dict(reveal_interval=3), dict(P=5, N=20)):
        stub = StubEnv(**kw)
        env = _adapter(stub)
        B, A, N = stub.num_envs, stub.num_agents, stub.graph_nodes
        act = env.action_spec[("agents", "action")]
        assert tuple(act.shape) == (B, A) and act.dtype == torch.int64 and act.n == N  # Categorical(N) per agent
        assert tuple(env.reward_spec[("agents", "reward")].shape) == (B, A, 1)
        for k in ("done", "terminated", "truncated"):
            assert tuple(env.done_spec[k].shape) == (B, 1) and env.done_spec[k].dtype == torch.bool
        assert fake_torchrl.check_env_specs(env, steps=3)
        out = env.rollout(4)  # EnvBase.rollout with random actions from the spec
        assert len(out) == 4 and out[-1].get(("agents", "reward")).shape == (B, A, 1)
        assert stub.calls[-1][0] == "step" and stub.calls[-1][1].dtype == torch.int64


def test_adapter_hides_mrx_from_the_police_groups():
    """with a reveal schedule the police groups see MrX only at reveal steps: MrX_pos = -1 while hidden, never his
    position or mask row; MrX's own group and the privileged state keep the true node"""
    stub = StubEnv(reveal_interval=3)
    env = _adapter(stub)
    td = env.reset()
    hidden = stub.mrx_revealed < 0
    assert hidden.any() and (~hidden).any()
    for i, name in enumerate(stub.possible_agents):
        seen = td.get((name, "observation", "MrX_pos"))[:, 0]
        if i == 0:
            assert torch.equal(seen, stub.pos[:, 0])
        else:
            assert bool((seen[hidden] == -1).all()) and torch.equal(seen[~hidden], stub.pos[~hidden, 0])
            assert td.get((name, "observation", "action_mask")).shape[1] == 1  # own row only
    shared = td.get(("agents", "observation", "agent_position"))
    assert bool((shared[hidden, 0] == -1).all()) and torch.equal(shared[:, 1:], stub.pos[:, 1:])
    assert torch.equal(td.get(("agents", "state", "agent_position")), stub.pos)
    # without a schedule nothing is masked (reference behaviour, yard.py:282,324)
    td0 = _adapter(StubEnv()).reset()
    assert td0.get(("agents", "observation", "agent_position")).data_ptr() == td0.get(("agents", "state", "agent_position")).data_ptr()


def test_adapter_partial_reset_is_a_noop_under_auto_reset():
    stub = StubEnv(auto_reset=True)
    env = _adapter(stub)
    env.reset()
    n = len(stub.calls)
    B = stub.num_envs
    m = torch.tensor([1, 0, 0, 1, 0, 0], dtype=torch.bool).reshape(B, 1)
    td = env.reset(fake_torchrl.TensorDict({"_reset": m}, batch_size=[B]))
    assert len(stub.calls) == n, "finished envs were already reset inside the step: no second reset"
    assert td.get("done").shape == (B, 1)
    env.reset()  # a full reset is still a reset
    assert len(stub.calls) == n + 1


def test_adapter_reset_keys_and_partial_reset():
    stub = StubEnv()
    env = _adapter(stub)
    td = env.reset()
    B, A, N = stub.num_envs, stub.num_agents, 9
    assert td.get(("agents", "observation", "action_mask")).shape == (B, A, N)
    assert td.get(("agents", "observation", "node_features")).shape == (B, N, A)
    assert td.get(("agents", "observation", "belief_map")).shape == (B, N)
    assert td.get(("agents", "observation", "agent_budget")).shape == (B, A, 1)
    assert td.get("done").shape == (B, 1) and not td.get("done").any()
    for i, name in enumerate(stub.possible_agents):  # PettingZooWrapper layout: one group of size 1 per agent
        assert td.get((name, "observation", "action_mask")).shape == (B, 1, N)
        assert td.get((name, "observation", "Polices_pos")).shape == (B, 1, A - 1)
        assert td.get((name, "observation", "Polices_pos")).sum(dim=1).shape == (B, A - 1)  # mappo_trainer.py:197-199
        assert torch.equal(td.get((name, "observation", "agent_position"))[:, 0], stub.pos[:, i])
        # zero-copy: the per-agent mask is a view of the batched buffer
        assert td.get((name, "action_mask")).data_ptr() == stub.action_mask[:, i:i + 1].data_ptr()
    m = torch.tensor([1, 0, 0, 1, 0, 0], dtype=torch.bool).reshape(B, 1)
    env.reset(fake_torchrl.TensorDict({"_reset": m}, batch_size=[B]))
    assert stub.calls[-1][0] == "reset" and torch.equal(stub.calls[-1][1], m.reshape(B))
    env.set_seed(5)
    assert stub.calls[-1] == ("seed", 5)


def test_adapter_step_accepts_batched_and_per_agent_actions():
    stub = StubEnv()
    env = _adapter(stub)
    B, A = stub.num_envs, stub.num_agents
    acts = torch.arange(B * A, dtype=torch.int64).reshape(B, A)
    out = env.step(fake_torchrl.TensorDict({("agents", "action"): acts}, batch_size=[B]))
    assert torch.equal(stub.calls[-1][1], acts)
    assert out.get(("agents", "reward")).shape == (B, A, 1) and out.get(("agents", "reward")).dtype == torch.float32
    assert out.get("terminated").shape == (B, 1) and bool(out.get("terminated")[0]) and not bool(out.get("terminated")[1])
    assert out.get(("Police0", "reward")).shape == (B, 1, 1)
    # the reference's loops write td[agent]["action"] = tensor([a]) per agent (gnn_trainer.py:234-241)
    td = fake_torchrl.TensorDict({}, batch_size=[B])
    td.set(("MrX", "action"), acts[:, 0:1])
    td.set(("Police1", "action"), acts[:, 2:3])
    env.step(td)
    want = acts.clone()
    want[:, 1] = -1  # missing agent -> DEFAULT_ACTION (yard.py:210-215)
    assert torch.equal(stub.calls[-1][1], want)


def test_shard_range_partitions_the_batch():
    for B, W in [(65536, 8), (10, 3), (7, 8), (0, 2), (262144, 8)]:
        spans = [shard_range(B, W, r) for r in range(W)]
        assert spans[0][0] == 0 and sum(c for _, c in spans) == B
        for (o0, c0), (o1, _) in zip(spans, spans[1:]):
            assert o0 + c0 == o1
        assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
    with pytest.raises(ValueError):
        shard_range(8, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _stats_worker(rank, world, port, q):
    import torch.distributed as dist

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from student_mechanism_design_b200 import allreduce_stats, shard_range
    from student_mechanism_design_b200._cabi import SY_NUM_STATS

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    off, cnt = shard_range(1000, world, rank)
    vec = torch.zeros(SY_NUM_STATS, dtype=torch.int64)
    vec[0] = cnt * 3  # env_steps of this shard
    vec[1] = rank + 1  # episodes
    vec[2] = off
    out = allreduce_stats(vec)
    assert vec[0] == cnt * 3, "the local vector must not be modified"
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


def test_stats_allreduce_world_size_2_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_stats_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in range(world):
        assert res[r]["env_steps"] == 3000 and res[r]["episodes"] == 3 and res[r]["mrx_wins"] == 500
    assert res[0] == res[1]


def test_oracle_shards_reproduce_the_global_batch():
    """The Philox streams are keyed by the GLOBAL env index: two shards with env_offset reproduce the slices of
    the unsharded batch (CPU statement of the multi-GPU invariant the gpu tests check on the device)."""
    import sy_oracle as so
    import sy_oracle_c as oc
    from student_mechanism_design_b200.graphs import generate_graph_pool

    pool = [so.Graph(g.num_nodes, g.edge_links, g.edges) for g in generate_graph_pool(3, 20, 34, seed=2)]
    cfg = so.OracleConfig(num_police=3, agent_money=9, belief=True, reveal_interval=3, toll=1)
    full = oc.CBatch(cfg, pool, 40, seed=5, resample_graph=True)
    shards = [oc.CBatch(cfg, pool, c, seed=5, env_offset=o, resample_graph=True) for o, c in
              (shard_range(40, 2, r) for r in range(2))]
    for s in range(25):
        full.step(full.sample_actions(s))
        for sh in shards:
            sh.step(sh.sample_actions(s))
        for k in ("pos", "money", "timestep", "visits", "belief", "revealed"):
            got = np.concatenate([getattr(sh, k)() for sh in shards])
            assert np.array_equal(got, getattr(full, k)()), (k, s)


def test_masked_sample_respects_mask_and_default_action():
    from student_mechanism_design_b200 import masked_sample

    g = torch.Generator().manual_seed(0)
    B, A, N = 64, 3, 11
    mask = torch.rand(B, A, N, generator=g) < 0.3
    mask[0, 1] = False  # an agent without legal moves
    logits = torch.randn(B, A, N, generator=g)
    for greedy in (False, True):
        a = masked_sample(logits, mask, generator=g, greedy=greedy)
        assert a.shape == (B, A) and a.dtype == torch.int64 and a[0, 1] == -1
        ok = mask.any(-1)
        assert bool((a[~ok] == -1).all()) and bool(mask.gather(-1, a.clamp_min(0).unsqueeze(-1)).squeeze(-1)[ok].all())
    # greedy picks the best legal logit
    best = torch.where(mask, logits, torch.full_like(logits, -1e30)).argmax(-1)
    assert torch.equal(masked_sample(logits, mask, greedy=True)[mask.any(-1)], best[mask.any(-1)])
    # sampling frequencies follow softmax over the legal nodes
    m1 = torch.zeros(1, 1, 4, dtype=torch.bool)
    m1[..., :3] = True
    lg = torch.tensor([[[0.0, 1.0, 2.0, 9.0]]])
    draws = torch.stack([masked_sample(lg.expand(4096, 1, 4), m1.expand(4096, 1, 4), generator=g) for _ in range(4)]).flatten()
    freq = torch.bincount(draws, minlength=4).float() / draws.numel()
    want = torch.softmax(lg[0, 0, :3], -1)
    assert freq[3] == 0 and torch.allclose(freq[:3], want, atol=0.02)
