"""torchrl `EnvBase` adapter of the batched env (the reference wraps its PettingZoo env
in `torchrl.envs.libs.pettingzoo.PettingZooWrapper`, gnn_trainer.py:129; agent names without `_` give one
group per agent).

torchrl / tensordict are NOT installed in the build image, so they are imported lazily and the adapter is
exercised in the CPU tests against a minimal stand-in (tests/fake_torchrl.py) -- its behaviour against a real
torchrl is unverified.  The adapter itself is thin: all state lives in the tensors owned by
`BatchedScotlandYardEnv`, keys are zero-copy views of them.

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

from typing import Optional

import torch


def _import_torchrl():
    try:
        from tensordict import TensorDict
        from torchrl.envs import EnvBase
    except ImportError as e:  # pragma: no cover - depends on the image
        raise ImportError("torchrl / tensordict are required for ScotlandYardTorchRLEnv (pip install torchrl)") from e
    return EnvBase, TensorDict


def collect_keys(env, per_agent: bool = True):
    """Flat {key tuple: tensor} view of the env's current observation (zero-copy)."""
    out = {
        ("agents", "observation", "node_features"): env.node_features,
        ("agents", "observation", "action_mask"): env.action_mask,
        ("agents", "observation", "agent_position"): env.pos,
        ("agents", "observation", "agent_budget"): env.agent_budget.unsqueeze(-1),
        ("agents", "observation", "MrX_revealed"): env.mrx_revealed,
        ("agents", "observation", "graph_id"): env.graph_id,
        ("agents", "action_mask"): env.action_mask,
    }
    if env.belief_on:
        out[("agents", "observation", "belief_map")] = env.belief_map
    if per_agent:
        for i, name in enumerate(env.possible_agents):
            out[(name, "observation", "node_features")] = env.node_features.unsqueeze(1)
            out[(name, "observation", "action_mask")] = env.action_mask[:, i:i + 1]
            out[(name, "action_mask")] = env.action_mask[:, i:i + 1]
            out[(name, "observation", "agent_position")] = env.pos[:, i:i + 1]
            out[(name, "observation", "agent_budget")] = env.agent_budget[:, i:i + 1].unsqueeze(-1)
            out[(name, "observation", "MrX_pos")] = env.pos[:, 0:1]
            out[(name, "observation", "Polices_pos")] = env.pos[:, 1:].unsqueeze(1)
            out[(name, "observation", "Currency")] = env.money[:, 1:].unsqueeze(1)
            if env.belief_on:
                out[(name, "observation", "belief_map")] = env.belief_map.unsqueeze(1)
    return out


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


def make_env_class(EnvBase, TensorDict):
    """Build the adapter class over the given torchrl-like base classes (real or the test stand-in)."""

    class ScotlandYardTorchRLEnv(EnvBase):
        def __init__(self, env, per_agent_groups: bool = True):
            super().__init__(device=env.device, batch_size=torch.Size([env.num_envs]))
            self.sy = env
            self.per_agent_groups = per_agent_groups
            self.possible_agents = env.possible_agents
            self.group_map = {n: [n] for n in env.possible_agents}  # PettingZooWrapper default for names without `_`

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
            self.sy.reset(reset_mask=mask, **{k: v for k, v in kwargs.items() if k in ("init_pos", "graph_id", "seed")})
            flat = collect_keys(self.sy, self.per_agent_groups)
            B = self.sy.num_envs
            z = torch.zeros(B, 1, dtype=torch.bool, device=self.sy.device)
            flat.update({("done",): z, ("terminated",): z.clone(), ("truncated",): z.clone()})
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
    return make_env_class(EnvBase, TensorDict)(env, per_agent_groups)
