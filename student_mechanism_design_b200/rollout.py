"""Batched rollout loop: the caller side of the hot path.

The reference's trainers step ONE episode in Python, build one PyG `Data` object per agent per step
(`create_graph_data`, src/training/utils.py:151-211) and pick actions agent by agent
(src/training/gnn_trainer.py:201-250, mappo_trainer.py:167-232).  With 65 536 envs on the device the same loop
is: observation tensors (already batched, written by the kernels) -> one policy call for all envs and agents ->
masked sampling on the device -> `env.step`.  Nothing here copies an observation; the stored trajectory keeps the
compact state (nodes, budgets, reveal flags) from which every dense observation can be re-assembled.
"""
from __future__ import annotations

from typing import Callable, Dict, Optional

import torch


def batched_graph_data(env) -> Dict[str, torch.Tensor]:
    """The tensors `create_graph_data` (src/training/utils.py:151-211) packs into a PyG `Data`, for all envs at once:
    `x` [B, N, A] one-hot agent nodes (a view of the kernel-written node_features), `edge_index` [G, 2, E] and
    `edge_attr` [G, E] of the graph pool plus `graph_id` [B] (shared graphs are not replicated per env)."""
    adj, ei, ef = env._static_tensors()
    return {"x": env.node_features, "edge_index": ei, "edge_attr": None if ef is None else ef.float(),
            "graph_id": env.graph_id, "adjacency": adj}


def masked_sample(logits: torch.Tensor, mask: torch.Tensor, generator: Optional[torch.Generator] = None,
                  greedy: bool = False, default_action: int = -1) -> torch.Tensor:
    """Sample one node per (env, agent) from `logits` [B, A, N] restricted to `mask` [B, A, N] (the env's
    action_mask); agents without a legal move get DEFAULT_ACTION, as the trainers do (gnn_trainer.py:227-229)."""
    neg = torch.finfo(logits.dtype).min
    masked = torch.where(mask, logits, torch.full_like(logits, neg))
    has_move = mask.any(dim=-1)
    if greedy:
        choice = masked.argmax(dim=-1)
    else:  # Gumbel-max: one pass, no normalisation, works for any batch shape
        u = torch.rand(masked.shape, device=masked.device, dtype=masked.dtype, generator=generator).clamp_min_(1e-20)
        g = -torch.log(-torch.log(u))
        choice = torch.where(mask, masked + g, torch.full_like(masked, neg)).argmax(dim=-1)
    return torch.where(has_move, choice, torch.full_like(choice, default_action)).to(torch.int64)


class RolloutCollector:
    """Collect `horizon` steps of all envs with a policy `policy(obs) -> logits [B, A, N]` (obs = env.observation());
    a policy object with `returns_actions = True` (e.g. `GNNPolicy`, whose kernel does the masked selection itself)
    is called as `policy(obs) -> actions int64 [B, A]` instead.

    Stored per step (leading dim T): `actions` int64 [T,B,A], `reward` f32 [T,B,A], `terminated` / `truncated`
    bool [T,B], `pos` / `money` int32 [T,B,A] (state BEFORE the step) and `mrx_revealed` int32 [T,B]."""

    def __init__(self, env, policy: Callable[[Dict[str, torch.Tensor]], torch.Tensor], horizon: int,
                 greedy: bool = False, seed: Optional[int] = None):
        self.env, self.policy, self.T, self.greedy = env, policy, int(horizon), greedy
        B, A, dev = env.num_envs, env.num_agents, env.device
        self.gen = None
        if seed is not None:
            self.gen = torch.Generator(device=dev)
            self.gen.manual_seed(int(seed))
        z = lambda *s, dtype: torch.zeros(*s, dtype=dtype, device=dev)  # noqa: E731
        self.buf = dict(actions=z(self.T, B, A, dtype=torch.int64), reward=z(self.T, B, A, dtype=torch.float32),
                        terminated=z(self.T, B, dtype=torch.bool), truncated=z(self.T, B, dtype=torch.bool),
                        pos=z(self.T, B, A, dtype=torch.int32), money=z(self.T, B, A, dtype=torch.int32),
                        mrx_revealed=z(self.T, B, dtype=torch.int32))

    @torch.no_grad()
    def collect(self) -> Dict[str, torch.Tensor]:
        env, buf = self.env, self.buf
        obs = env.observation()
        for t in range(self.T):
            buf["pos"][t].copy_(env.pos)
            buf["money"][t].copy_(env.money)
            buf["mrx_revealed"][t].copy_(env.mrx_revealed)
            if getattr(self.policy, "returns_actions", False):
                actions = self.policy(obs)
            else:
                actions = masked_sample(self.policy(obs), env.action_mask, self.gen, self.greedy, env.DEFAULT_ACTION)
            obs, reward, terminated, truncated, _ = env.step(actions)
            buf["actions"][t].copy_(actions)
            buf["reward"][t].copy_(reward)
            buf["terminated"][t].copy_(terminated[:, 0])
            buf["truncated"][t].copy_(truncated[:, 0])
        return buf
