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
    action_mask); agents without a legal move get DEFAULT_ACTION, as the trainers do (gnn_trainer.py:227-229).
    Plain-torch form (CPU tensors, tests); on the device use `masked_sample_device` (one kernel, no [B, A, N] temporaries)."""
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


def masked_sample_device(logits: torch.Tensor, mask: torch.Tensor, seed: int = 0, step_counter: int = 0, greedy: bool = False,
                         env_offset: int = 0, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """`masked_sample` as ONE kernel of libsy_policy.so (sy_masked_sample): a warp per (env, agent) row, Gumbel-max with
    counter-based Philox noise (reproducible per (seed, env, step, agent, node); no [B, A, N] temporaries -- the eager
    form allocates five of them, 92 M floats each at BASELINE config 3) or greedy first-argmax over the legal nodes."""
    import ctypes as C

    from . import _policy_cabi as pc

    if not logits.is_cuda:
        raise ValueError("masked_sample_device needs device tensors (use masked_sample on the CPU)")
    B, A, N = logits.shape
    lg = logits.contiguous().float()
    mk = mask.contiguous()
    mk = mk.view(torch.uint8) if mk.dtype == torch.bool else mk.to(torch.uint8)
    if out is None:
        out = torch.empty(B, A, dtype=torch.int64, device=logits.device)
    lib = pc.load_library()
    with torch.cuda.device(logits.device):
        pc.check(lib.sy_masked_sample(lg.data_ptr(), mk.data_ptr(), B, A, N, int(env_offset), int(seed) & 0xFFFFFFFFFFFFFFFF,
                                      int(step_counter) & 0xFFFFFFFF, int(bool(greedy)), out.data_ptr(),
                                      torch.cuda.current_stream(logits.device).cuda_stream))
    return out


def copy_segments(dsts, srcs):
    """dst[i].copy_(src[i]) for up to 8 pairs of contiguous same-size device tensors in ONE launch (sy_copy_segments)"""
    import ctypes as C

    from . import _cabi

    n = len(dsts)
    dst = (C.c_void_p * n)(*[d.data_ptr() for d in dsts])
    src = (C.c_void_p * n)(*[x.data_ptr() for x in srcs])
    nbytes = (C.c_uint64 * n)(*[d.numel() * d.element_size() for d in dsts])
    for d, x in zip(dsts, srcs):
        if not (d.is_contiguous() and x.is_contiguous()) or d.numel() * d.element_size() != x.numel() * x.element_size():
            raise ValueError("copy_segments needs contiguous tensors of equal byte size")
    with torch.cuda.device(dsts[0].device):
        _cabi.check(_cabi.load_library().sy_copy_segments(n, dst, src, nbytes, torch.cuda.current_stream(dsts[0].device).cuda_stream))


class RolloutCollector:
    """Collect `horizon` steps of all envs with a policy `policy(obs) -> logits [B, A, N]` (obs = env.observation());
    a policy object with `returns_actions = True` (e.g. `GNNPolicy`, whose kernel does the masked selection itself)
    is called as `policy(obs) -> actions int64 [B, A]` instead.

    Stored per step (leading dim T): `actions` int64 [T,B,A], `reward` f32 [T,B,A], `status` uint8 [T,B] (bit 0
    terminated, bit 1 truncated; `terminated` / `truncated` bool [T,B] are derived from it after the rollout), `pos` /
    `money` int32 [T,B,A] (state BEFORE the step) and `mrx_revealed` int32 [T,B].  Two kernel launches per step record
    the transition (sy_copy_segments), and a policy that returns logits is sampled by one kernel (sy_masked_sample)."""

    def __init__(self, env, policy: Callable[[Dict[str, torch.Tensor]], torch.Tensor], horizon: int,
                 greedy: bool = False, seed: Optional[int] = None):
        self.env, self.policy, self.T, self.greedy = env, policy, int(horizon), greedy
        B, A, dev = env.num_envs, env.num_agents, env.device
        self.gen = None
        self._step = 0
        if seed is not None:
            self.gen = torch.Generator(device=dev)
            self.gen.manual_seed(int(seed))
        z = lambda *s, dtype: torch.zeros(*s, dtype=dtype, device=dev)  # noqa: E731
        self.buf = dict(actions=z(self.T, B, A, dtype=torch.int64), reward=z(self.T, B, A, dtype=torch.float32),
                        status=z(self.T, B, dtype=torch.uint8),
                        pos=z(self.T, B, A, dtype=torch.int32), money=z(self.T, B, A, dtype=torch.int32),
                        mrx_revealed=z(self.T, B, dtype=torch.int32))

    @torch.no_grad()
    def collect(self) -> Dict[str, torch.Tensor]:
        env, buf = self.env, self.buf
        obs = env.observation()
        on_device = env.device.type == "cuda"
        # a policy that reads only the compact state (GNNPolicy: positions, budgets, reveal flags, the graph pool) lets
        # the env step software-pipelined: the dense observations trail by one step and are flushed once at the end
        deferred = bool(getattr(self.policy, "reads_state_only", False)) and hasattr(env, "step_deferred")
        for t in range(self.T):
            # the state BEFORE the step: one launch for the three tensors
            self._copy([buf["pos"][t], buf["money"][t], buf["mrx_revealed"][t]], [env.pos, env.money, env.mrx_revealed])
            if getattr(self.policy, "returns_actions", False):
                actions = self.policy(obs)
            elif on_device and self.gen is None:
                actions = masked_sample_device(self.policy(obs), env.action_mask, seed=env.seed, step_counter=self._step,
                                               greedy=self.greedy, env_offset=env.env_offset, out=buf["actions"][t])
            else:
                actions = masked_sample(self.policy(obs), env.action_mask, self.gen, self.greedy, env.DEFAULT_ACTION)
            self._step += 1
            if deferred:
                reward, terminated, truncated, _ = env.step_deferred(actions)
            else:
                obs, reward, terminated, truncated, _ = env.step(actions)
            # the step's results: one launch (the per-env flags come from the status byte: bit 0 terminated, 1 truncated)
            dst, src = [buf["reward"][t], buf["status"][t]], [reward, env.status]
            if actions.data_ptr() != buf["actions"][t].data_ptr():
                dst.append(buf["actions"][t])
                src.append(actions.contiguous())
            self._copy(dst, src)
        if deferred:
            env.flush_observations()
        buf["terminated"] = (buf["status"] & 1).bool()
        buf["truncated"] = (buf["status"] & 2).bool()
        return buf

    def _copy(self, dsts, srcs):
        if self.env.device.type == "cuda":
            copy_segments(dsts, srcs)
        else:
            for d, s in zip(dsts, srcs):
                d.copy_(s)
