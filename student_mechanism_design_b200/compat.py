"""Single-episode view with the reference's exact call shape, for the reference's own loops.

`CustomEnvironment` here takes the constructor arguments of the reference class
(/root/reference/src/environment/yard.py:18-28) and returns what it returns: `reset()` ->
`(observations, infos)`, `step({agent: int | None})` -> `(observations, rewards, terminations,
truncations, infos)`, all dicts keyed by "MrX", "Police0", ... holding numpy values with the
dtypes the reference produces (yard.py:319-332).  Underneath it is a
`BatchedScotlandYardEnv` with B = 1: every step is one sy_step on the GPU plus a few small
device->host copies, i.e. this class exists for drop-in compatibility (the trainers of
src/training/*.py run unmodified against it), not for throughput -- use the batched env for that.

Like the reference it draws a new random graph on every reset (yard.py:87-101), from this
package's generator (same distribution, own numpy stream) unless `graph=` / `start_positions=`
hand over explicit ones (parity harness).
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Dict, Optional, Sequence

import numpy as np
import torch

from .env import BatchedScotlandYardEnv, DEFAULT_REWARD_WEIGHTS
from .graphs import GraphSpec, generate_connected_graph

WINNER_NAMES = {0: None, 1: "MrX", 2: "Police"}  # env.current_winner, yard.py:250


class _Discrete:
    """stand-in for gymnasium.spaces.Discrete (only `.n` is used by the trainers, gnn_trainer.py:133-135)"""

    def __init__(self, n):
        self.n, self.start, self.dtype = int(n), 0, np.int64


class CustomEnvironment:
    DEFAULT_ACTION = -1  # yard.py:16
    metadata = {"name": "scotland_yard_env_b200"}

    def __init__(self, number_of_agents, agent_money, reward_weights=None, logger=None, epoch=0, graph_nodes=50,
                 graph_edges=110, vis_configs=None, *, reveal_interval=0, tolls=0, belief=False, reward_mode=None,
                 seed=0, device="cuda:0", graph: Optional[GraphSpec] = None, reward_tables=None):
        self.number_of_agents, self.agent_money = int(number_of_agents), int(agent_money)
        self.reward_weights = dict(DEFAULT_REWARD_WEIGHTS) if reward_weights is None else dict(reward_weights)
        self.logger, self.epoch, self.vis_config = logger, epoch, vis_configs
        self.graph_nodes, self.graph_edges = int(graph_nodes), graph_edges
        self.possible_agents = ["MrX"] + [f"Police{i}" for i in range(self.number_of_agents)]
        self.agents = list(self.possible_agents)
        self._kw = dict(reveal_interval=reveal_interval, tolls=tolls, belief=belief, reward_mode=reward_mode,
                        device=device, reward_tables=reward_tables)
        self._rng = np.random.default_rng(seed)
        self._seed = int(seed)
        self._env: Optional[BatchedScotlandYardEnv] = None
        self.current_winner = None
        self.timestep = 0
        # yard.py:67-76: one probe sample fixes the achievable edge count
        probe = graph if graph is not None else generate_connected_graph(self.graph_nodes, graph_edges, self._rng)
        self.actual_num_edges = len(probe.edges)
        self.reset(graph=graph)

    # ------------------------------------------------------------------ reference API
    def reset(self, episode=0, seed=None, options=None, *, graph: Optional[GraphSpec] = None,
              start_positions: Optional[Sequence[int]] = None):
        """yard.py:80-142.  Returns (observations, infos)."""
        if graph is None:
            for _ in range(100):  # yard.py:89-101
                graph = generate_connected_graph(self.graph_nodes, self.graph_edges, self._rng)
                if len(graph.edges) == self.actual_num_edges:
                    break
            else:
                raise RuntimeError(f"Failed to generate graph with {self.actual_num_edges} edges after 100 attempts.")
        if self._env is not None:
            self._env.close()
        self._env = BatchedScotlandYardEnv(1, self.number_of_agents, self.agent_money, self.reward_weights, graphs=[graph],
                                           seed=int(self._rng.integers(0, 2**62)), auto_reset=False, keep_reward64=True,
                                           **self._kw)
        self.board = SimpleNamespace(nodes=graph.nodes, edges=graph.edges, edge_links=graph.edge_links)
        init = None if start_positions is None else np.asarray(start_positions, dtype=np.int32).reshape(1, -1)
        self._env.reset(init_pos=init, graph_id=np.zeros(1, dtype=np.int32))
        self.agents = list(self.possible_agents)
        self.current_winner, self.timestep = None, 0
        obs = self._observations()
        return obs, {a: {} for a in self.possible_agents}

    def step(self, actions: Dict[str, Optional[int]]):
        """yard.py:144-269.  `actions`: {agent: node or None}; invalid targets mean `stay`."""
        A = self.number_of_agents + 1
        acts = np.full((1, A), -1, dtype=np.int64)
        for i, name in enumerate(self.possible_agents):
            a = actions.get(name, None) if hasattr(actions, "get") else actions[name]
            if a is None:
                # MrX `None` skips his move entirely (yard.py:155-160): any non-adjacent target does the same
                acts[0, i] = -1
            else:
                acts[0, i] = int(np.asarray(a.cpu() if isinstance(a, torch.Tensor) else a).reshape(-1)[0])
        env = self._env
        env.step(torch.from_numpy(acts).to(env.device))
        f64 = env.reward_mode == "fp64"
        rew = (env.reward64 if f64 else env.reward)[0].cpu().numpy()
        term, trunc = bool(env.terminated[0, 0]), bool(env.truncated[0, 0])
        self.current_winner = WINNER_NAMES[int(env.winner[0])]
        self.timestep = int(env.timestep[0])
        names = self.possible_agents
        rewards = {n: (float(rew[i]) if f64 else np.float32(rew[i])) for i, n in enumerate(names)}
        terminations = {n: term for n in names}
        truncations = {n: trunc for n in names}
        obs = self._observations()
        if term or trunc:
            self.agents = []  # yard.py:260-266
        return obs, rewards, terminations, truncations, {n: {} for n in names}

    def get_possible_moves(self, agent_idx):
        """yard.py:474-480: sorted affordable neighbour ids, int32."""
        return self._env.get_possible_moves(agent_idx, 0)

    def action_space(self, agent):
        return _Discrete(self.graph_nodes)  # yard.py:482-498

    def get_distance(self, node1, node2):
        return self._env.get_distance(int(node1), int(node2), 0)  # yard.py:375-388

    @property
    def MrX_pos(self):
        return [int(self._env.pos[0, 0])]

    @property
    def police_positions(self):
        return [int(x) for x in self._env.pos[0, 1:].cpu().tolist()]

    @property
    def agents_money(self):
        return [int(x) for x in self._env.money[0].cpu().tolist()]

    def render(self):  # visualisation is out of scope
        return None

    def close(self):
        if self._env is not None:
            self._env.close()
            self._env = None

    def save_visualizations(self):
        return None

    # ------------------------------------------------------------------ observation dict (yard.py:319-332)
    def _observations(self):
        env = self._env
        N = self.graph_nodes
        adj = np.zeros((N, N), dtype=np.float64)
        el = self.board.edge_links
        adj[el[:, 0], el[:, 1]] = 1
        adj[el[:, 1], el[:, 0]] = 1
        nf = env.node_features[0].cpu().numpy().astype(np.float64)
        pos = env.pos[0].cpu().numpy()
        money = env.money[0].cpu().numpy()
        mask = env.action_mask[0].cpu().numpy()
        extra = {}
        if env.belief_on:
            extra["belief_map"] = env.belief_map[0].cpu().numpy()
        if env.reveal_interval:
            extra["MrX_revealed"] = int(env.mrx_revealed[0])
        out = {}
        for i, name in enumerate(self.possible_agents):
            out[name] = {
                "adjacency_matrix": adj,
                "node_features": nf,
                "edge_index": self.board.edge_links.T,
                "edge_features": self.board.edges,
                "MrX_pos": int(pos[0]) if int(env.mrx_revealed[0]) >= 0 or not env.reveal_interval else -1,
                "Polices_pos": [int(x) for x in pos[1:]],
                "Currency": [int(x) for x in money[1:]],
                "action_mask": mask[i].copy(),
                "agent_position": int(pos[i]),
                "agent_budget": np.array([money[i]], dtype=np.float32),
                **extra,
            }
        return out
