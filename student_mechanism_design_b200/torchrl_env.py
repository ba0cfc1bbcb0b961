"""torchrl `EnvBase` adapter of the batched env (the reference wraps its PettingZoo env
in `torchrl.envs.libs.pettingzoo.PettingZooWrapper`, gnn_trainer.py:129; agent names without `_` give one
group per agent).

torchrl / tensordict are NOT installed in the build image, so they are imported lazily and the adapter is
exercised in the CPU tests against a stand-in (tests/fake_torchrl.py) that mimics the parts of torchrl the adapter
touches: `EnvBase.__init__/reset/step/rollout`, the spec classes (Composite / Bounded / Unbounded / Categorical /
Binary) and a `check_env_specs` approximation that verifies every emitted key, shape and dtype against the declared
specs -- its behaviour against a real torchrl is still unverified.  The adapter itself is thin: all state lives in the
tensors owned by `BatchedScotlandYardEnv`, keys are zero-copy views of them.

Specs (what torchrl's `EnvBase.rollout`, `check_env_specs` and `SyncDataCollector` need; built in `__init__` from
whichever spec classes the installed torchrl exports): `observation_spec` = Composite of the keys below,
`action_spec` = Categorical(n=N) int64 [B, A] under ("agents", "action"), `reward_spec` = Unbounded float32
[B, A, 1] under ("agents", "reward"), `done_spec` = done / terminated / truncated bool [B, 1].

Partial observability: while MrX is hidden (`reveal_interval > 0` and no reveal this step) the per-agent groups of
the POLICE carry `MrX_pos = -1` and never MrX's own position or mask row; `node_features` has his column blank
(written that way by the kernel).  The shared ("agents", ...) entries are the controller's view: `agent_position`
has column 0 replaced by `MrX_revealed`, the full positions live under ("agents", "state", "agent_position") for
a centralised critic, and ("agents", "action_mask") holds all rows because a central controller samples all agents.

Auto-reset: with `auto_reset=True` the env resets finished envs inside the same step, so a torchrl collector's
follow-up `reset` with a `"_reset"` mask must not reset them again: the adapter then only returns the current
observation.  With `auto_reset=False` the mask is honoured (partial reset).

Key layout (batch_size = [B]):
  ("agents", "observation", k)  k in node_features [B,N,A] | action_mask [B,A,N] | agent_position [B,A] |
                                agent_budget [B,A,1] | belief_map [B,N] | MrX_revealed [B] | graph_id [B]
  ("agents", "action")          int64 [B, A]          (input of step)
  ("agents", "reward")          float32 [B, A, 1]
  "done" / "terminated" / "truncated"   bool [B, 1]
  (name, "observation", k), (name, "action_mask"), (name, "reward"), ... per agent name ("MrX", "Police0", ...):
      the same data sliced per agent with a leading group dim of 1, the shape PettingZooWrapper produced
      (mappo_trainer.py:197-199 sums over it); `(name, "action")` int64 [B, 1] is accepted as input too.
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Optional

import torch


def _import_torchrl():
    try:
        from tensordict import TensorDict
        from torchrl.envs import EnvBase
    except ImportError as e:  # pragma: no cover - depends on the image
        raise ImportError("torchrl / tensordict are required for ScotlandYardTorchRLEnv (pip install torchrl)") from e
    return EnvBase, TensorDict


def _import_torchrl_specs():
    """(Composite, Bounded, Unbounded, Categorical, Binary) under either generation of torchrl's names"""
    import torchrl.data as D  # pragma: no cover - depends on the image

    pick = lambda *names: next(getattr(D, n) for n in names if hasattr(D, n))  # noqa: E731
    return SimpleNamespace(Composite=pick("Composite", "CompositeSpec"), Bounded=pick("Bounded", "BoundedTensorSpec"),
                           Unbounded=pick("Unbounded", "UnboundedContinuousTensorSpec"),
                           Categorical=pick("Categorical", "DiscreteTensorSpec"),
                           Binary=pick("Binary", "BinaryDiscreteTensorSpec"))


def _masked_positions(env):
    """[B, A] positions with MrX's column replaced by MrX_revealed (-1 while hidden) when a reveal schedule is on"""
    if not getattr(env, "reveal_interval", 0):
        return env.pos
    return torch.cat([env.mrx_revealed.unsqueeze(1).to(env.pos.dtype), env.pos[:, 1:]], dim=1)


def collect_keys(env, per_agent: bool = True):
    """Flat {key tuple: tensor} view of the env's current observation (zero-copy except the masked positions)."""
    hidden_aware = bool(getattr(env, "reveal_interval", 0))
    out = {
        ("agents", "observation", "node_features"): env.node_features,
        ("agents", "observation", "action_mask"): env.action_mask,
        ("agents", "observation", "agent_position"): _masked_positions(env),
        ("agents", "observation", "agent_budget"): env.agent_budget.unsqueeze(-1),
        ("agents", "observation", "MrX_revealed"): env.mrx_revealed,
        ("agents", "observation", "graph_id"): env.graph_id,
        ("agents", "action_mask"): env.action_mask,
        ("agents", "state", "agent_position"): env.pos,  # privileged: centralised critic / logging
    }
    if env.belief_on:
        out[("agents", "observation", "belief_map")] = env.belief_map
    if per_agent:
        mrx_seen = env.mrx_revealed.unsqueeze(1).to(env.pos.dtype) if hidden_aware else env.pos[:, 0:1]
        for i, name in enumerate(env.possible_agents):
            out[(name, "observation", "node_features")] = env.node_features.unsqueeze(1)
            out[(name, "observation", "action_mask")] = env.action_mask[:, i:i + 1]
            out[(name, "action_mask")] = env.action_mask[:, i:i + 1]
            out[(name, "observation", "agent_position")] = env.pos[:, i:i + 1]
            out[(name, "observation", "agent_budget")] = env.agent_budget[:, i:i + 1].unsqueeze(-1)
            # MrX knows where he is; the police see him only at reveal steps (-1 otherwise)
            out[(name, "observation", "MrX_pos")] = env.pos[:, 0:1] if i == 0 else mrx_seen
            out[(name, "observation", "Polices_pos")] = env.pos[:, 1:].unsqueeze(1)
            out[(name, "observation", "Currency")] = env.money[:, 1:].unsqueeze(1)
            if env.belief_on:
                out[(name, "observation", "belief_map")] = env.belief_map.unsqueeze(1)
    return out


def build_specs(env, S, per_agent: bool = True):
    """(observation_spec, action_spec, reward_spec, done_spec) for `env` from the spec classes in `S`
    (torchrl's, or the stand-in's).  Every leaf carries the full shape incl. the batch dimension."""
    B, A, N, P = env.num_envs, env.num_agents, env.graph_nodes, env.number_of_agents
    G = max(int(getattr(env, "num_graphs", 1)), 1)
    dev, bs = env.device, torch.Size([B])
    i32, f32, i64 = torch.int32, torch.float32, torch.int64
    money_hi = max(int(getattr(env, "agent_money", 0)), 1000)  # MAX_MONEY_LIMIT, yard.py:11

    def comp(d):
        return S.Composite(d, shape=bs, device=dev)

    def obs_leaves(lead, mrx_lo=-1):
        """observation leaves of a group whose tensors carry `lead` after the batch dim ([] shared, [1] per agent)"""
        L = list(lead)
        d = {
            "node_features": S.Bounded(low=0, high=1, shape=[B, *L, N, A], dtype=env.node_features.dtype, device=dev),
            "agent_budget": S.Bounded(low=0, high=money_hi, shape=[B, *(L or [A]), 1], dtype=f32, device=dev),
        }
        if env.belief_on:
            d["belief_map"] = S.Bounded(low=0, high=1, shape=[B, *L, N], dtype=f32, device=dev)
        return d

    shared_obs = obs_leaves([])
    shared_obs.update({
        "action_mask": S.Binary(n=N, shape=[B, A, N], dtype=torch.bool, device=dev),
        "agent_position": S.Bounded(low=-1, high=N - 1, shape=[B, A], dtype=i32, device=dev),
        "MrX_revealed": S.Bounded(low=-1, high=N - 1, shape=[B], dtype=i32, device=dev),
        "graph_id": S.Bounded(low=0, high=G - 1, shape=[B], dtype=i32, device=dev),
    })
    obs = {"agents": comp({
        "observation": comp(shared_obs),
        "action_mask": S.Binary(n=N, shape=[B, A, N], dtype=torch.bool, device=dev),
        "state": comp({"agent_position": S.Bounded(low=0, high=N - 1, shape=[B, A], dtype=i32, device=dev)}),
    })}
    reward = {"agents": comp({"reward": S.Unbounded(shape=[B, A, 1], dtype=f32, device=dev)})}
    flag = lambda *shape: S.Binary(n=1, shape=list(shape), dtype=torch.bool, device=dev)  # noqa: E731
    done = {"done": flag(B, 1), "terminated": flag(B, 1), "truncated": flag(B, 1)}
    if per_agent:
        for i, name in enumerate(env.possible_agents):
            leaves = obs_leaves([1])
            leaves.update({
                "action_mask": S.Binary(n=N, shape=[B, 1, N], dtype=torch.bool, device=dev),
                "agent_position": S.Bounded(low=0, high=N - 1, shape=[B, 1], dtype=i32, device=dev),
                "MrX_pos": S.Bounded(low=-1, high=N - 1, shape=[B, 1], dtype=i32, device=dev),
                "Polices_pos": S.Bounded(low=0, high=N - 1, shape=[B, 1, P], dtype=i32, device=dev),
                "Currency": S.Bounded(low=0, high=money_hi, shape=[B, 1, P], dtype=i32, device=dev),
            })
            obs[name] = comp({"observation": comp(leaves), "action_mask": S.Binary(n=N, shape=[B, 1, N], dtype=torch.bool, device=dev)})
            reward[name] = comp({"reward": S.Unbounded(shape=[B, 1, 1], dtype=f32, device=dev)})
            done[name] = comp({"done": flag(B, 1, 1), "terminated": flag(B, 1, 1), "truncated": flag(B, 1, 1)})
    action = {"agents": comp({"action": S.Categorical(n=N, shape=[B, A], dtype=i64, device=dev)})}
    return comp(obs), comp(action), comp(reward), comp(done)


def collect_results(env, per_agent: bool = True):
    done = env.done_flags[:, :1]
    out = {
        ("agents", "reward"): env.reward.unsqueeze(-1),
        ("done",): done,
        ("terminated",): env.terminated[:, :1],
        ("truncated",): env.truncated[:, :1],
    }
    if per_agent:
        for i, name in enumerate(env.possible_agents):
            out[(name, "reward")] = env.reward[:, i:i + 1].unsqueeze(-1)
            out[(name, "done")] = env.done_flags[:, i:i + 1].unsqueeze(-1)
            out[(name, "terminated")] = env.terminated[:, i:i + 1].unsqueeze(-1)
            out[(name, "truncated")] = env.truncated[:, i:i + 1].unsqueeze(-1)
    return out


def gather_actions(env, get) -> torch.Tensor:
    """int64 [B, A] from either ("agents", "action") or the per-agent (name, "action") entries;
    `get(key)` returns a tensor or None."""
    a = get(("agents", "action"))
    if a is not None:
        return a.reshape(env.num_envs, env.num_agents).to(torch.int64)
    cols = []
    for name in env.possible_agents:
        x = get((name, "action"))
        if x is None:
            cols.append(torch.full((env.num_envs,), env.DEFAULT_ACTION, dtype=torch.int64, device=env.device))
        else:
            cols.append(x.reshape(env.num_envs).to(torch.int64))
    return torch.stack(cols, dim=1).contiguous()


def make_env_class(EnvBase, TensorDict, specs=None):
    """Build the adapter class over the given torchrl-like base classes (real or the test stand-in); `specs` is the
    namespace of spec classes (`_import_torchrl_specs()` for a real torchrl)."""

    class ScotlandYardTorchRLEnv(EnvBase):
        def __init__(self, env, per_agent_groups: bool = True):
            super().__init__(device=env.device, batch_size=torch.Size([env.num_envs]))
            self.sy = env
            self.per_agent_groups = per_agent_groups
            self.possible_agents = env.possible_agents
            self.group_map = {n: [n] for n in env.possible_agents}  # PettingZooWrapper default for names without `_`
            S = specs if specs is not None else _import_torchrl_specs()
            obs, act, rew, done = build_specs(env, S, per_agent_groups)
            self.observation_spec = obs
            self.action_spec = act
            self.reward_spec = rew
            self.done_spec = done

        def _td(self, flat):
            td = TensorDict({}, batch_size=[self.sy.num_envs], device=self.sy.device)
            for k, v in flat.items():
                td.set(k if len(k) > 1 else k[0], v)
            return td

        def _reset(self, tensordict=None, **kwargs):
            mask = None
            if tensordict is not None:
                m = tensordict.get("_reset", None)
                if m is not None:
                    mask = m.reshape(self.sy.num_envs)
            # same-step auto-reset already restarted the finished envs: a collector's partial reset is then a no-op
            if not (mask is not None and getattr(self.sy, "auto_reset", False) and getattr(self.sy, "_is_reset", True)):
                self.sy.reset(reset_mask=mask, **{k: v for k, v in kwargs.items() if k in ("init_pos", "graph_id", "seed")})
            flat = collect_keys(self.sy, self.per_agent_groups)
            B = self.sy.num_envs
            z = torch.zeros(B, 1, dtype=torch.bool, device=self.sy.device)
            flat.update({("done",): z, ("terminated",): z.clone(), ("truncated",): z.clone()})
            if self.per_agent_groups:
                for name in self.possible_agents:
                    for k in ("done", "terminated", "truncated"):
                        flat[(name, k)] = torch.zeros(B, 1, 1, dtype=torch.bool, device=self.sy.device)
            return self._td(flat)

        def _step(self, tensordict):
            acts = gather_actions(self.sy, lambda k: tensordict.get(k, None))
            self.sy.step(acts)
            flat = collect_keys(self.sy, self.per_agent_groups)
            flat.update(collect_results(self.sy, self.per_agent_groups))
            return self._td(flat)

        def _set_seed(self, seed: Optional[int]):
            if seed is not None:
                self.sy.set_seed(int(seed))
            return seed

        # attribute forwarding the reference's loops rely on (gnn_trainer.py:133-135,207,221,294,303)
        def get_possible_moves(self, agent_idx, env_index=0):
            return self.sy.get_possible_moves(agent_idx, env_index)

        @property
        def number_of_agents(self):
            return self.sy.number_of_agents

        @property
        def current_winner(self):
            return self.sy.winner

    return ScotlandYardTorchRLEnv


def make_torchrl_env(env, per_agent_groups: bool = True):
    """Wrap a BatchedScotlandYardEnv as a torchrl EnvBase (needs torchrl + tensordict)."""
    EnvBase, TensorDict = _import_torchrl()
    return make_env_class(EnvBase, TensorDict, _import_torchrl_specs())(env, per_agent_groups)
