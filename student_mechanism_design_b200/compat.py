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


class _Space:
    """stand-in for a gymnasium space when gymnasium is not installed: same attribute names"""

    def __repr__(self):
        return f"{type(self).__name__}({', '.join(f'{k}={v!r}' for k, v in vars(self).items())})"


class _Discrete(_Space):
    def __init__(self, n, start=0):
        self.n, self.start, self.dtype, self.shape = int(n), int(start), np.int64, ()


class _MultiDiscrete(_Space):
    def __init__(self, nvec):
        self.nvec = np.asarray(nvec, dtype=np.int64)
        self.dtype, self.shape = np.int64, self.nvec.shape


class _MultiBinary(_Space):
    def __init__(self, n):
        self.n, self.dtype, self.shape = int(n), np.int8, (int(n),)


class _Box(_Space):
    def __init__(self, low, high, shape, dtype):
        self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), dtype


class _Dict(_Space):
    def __init__(self, spaces):
        self.spaces = dict(spaces)

    def __getitem__(self, k):
        return self.spaces[k]

    def keys(self):
        return self.spaces.keys()


def _space_classes():
    """gymnasium's classes when it is installed (what the reference builds, yard.py:5), else the stand-ins above"""
    try:
        import gymnasium
        from gymnasium.spaces import Box, Dict, Discrete, MultiBinary, MultiDiscrete

        if not hasattr(gymnasium, "__version__"):  # a test stub on sys.path, not the library
            raise ImportError
        return Discrete, MultiDiscrete, MultiBinary, Box, Dict
    except Exception:
        return _Discrete, _MultiDiscrete, _MultiBinary, _Box, _Dict


MAX_WEIGHT = 5  # ConnectedGraph.MAX_WEIGHT, graph_layout.py:12
MAX_MONEY_LIMIT = 1000  # yard.py:11


def reference_action_space(num_nodes: int):
    """yard.py:482-498: Discrete(num_nodes), action i = move to node i (invalid ones are masked)"""
    return _space_classes()[0](int(num_nodes))


def reference_observation_space(num_nodes: int, num_police: int, num_edges: int, agent_money: int, *, belief: bool = False,
                                reveal: bool = False):
    """yard.py:500-554, key for key and with the DECLARED dtypes (the reference's produced arrays differ: adjacency and
    node_features are float64, edge_features int64 -- SURVEY.md 8(a) a5; the declaration is reproduced as written).
    `belief` / `reveal` add the keys this env's extensions emit (belief_map f32 [N]; MrX_revealed in {-1..N-1})."""
    Discrete, MultiDiscrete, MultiBinary, Box, Dict = _space_classes()
    N, P, E = int(num_nodes), int(num_police), int(num_edges)
    spaces = {
        "adjacency_matrix": Box(low=0.0, high=1.0, shape=(N, N), dtype=np.int64),
        "node_features": Box(low=0.0, high=1.0, shape=(N, P + 1), dtype=np.int64),
        "edge_index": Box(low=0, high=N, shape=(2, E), dtype=np.int32),
        "edge_features": Box(low=0, high=MAX_WEIGHT, shape=(E,), dtype=np.int32),
        "MrX_pos": Discrete(N),
        "Polices_pos": MultiDiscrete([N] * P),
        "Currency": MultiDiscrete([int(agent_money) + 1] * P),
        "action_mask": MultiBinary(N),
        "agent_position": Discrete(N),
        "agent_budget": Box(low=0.0, high=MAX_MONEY_LIMIT, shape=(1,), dtype=np.float32),
    }
    if belief:
        spaces["belief_map"] = Box(low=0.0, high=1.0, shape=(N,), dtype=np.float32)
    if reveal:
        spaces["MrX_revealed"] = Discrete(N + 1, start=-1)
    return Dict(spaces)


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
        return reference_action_space(self.graph_nodes)  # yard.py:482-498

    def observation_space(self, agent):
        """yard.py:500-554 (same keys and declared dtypes), plus the keys of this env's extensions when they are on"""
        return reference_observation_space(self.graph_nodes, self.number_of_agents, self.actual_num_edges, self.agent_money,
                                           belief=bool(self._kw["belief"]), reveal=bool(self._kw["reveal_interval"]))

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
