"""Write tests/golden/mappo.npz from the UNMODIFIED reference modules (src/agent/mappo_agent.py only needs torch):
AgentPolicy / CentralCritic parameters, observations, masks and what MappoAgent.select_action returns for them
(sampled action, its log-prob, the masked + renormalised distribution), incl. the fall-back branches
(mappo_agent.py:121-133).  Run in the build container: python oracle/gen_policy_golden.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
import importlib.util  # noqa: E402

# load the file itself: the package __init__ also imports the GNN agent, which needs torch_geometric (absent here)
_spec = importlib.util.spec_from_file_location("ref_mappo_agent", "/root/reference/src/agent/mappo_agent.py")
_mod = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_mod)
MappoAgent = _mod.MappoAgent

torch.manual_seed(0)
rng = np.random.default_rng(0)
out = {}
cases = []
for ci, (n_agents, obs_size, hidden, n_nodes) in enumerate([(1, 4, 64, 15), (3, 7, 64, 50), (6, 7, 32, 200), (2, 12, 48, 33)]):
    agent = MappoAgent(n_agents=n_agents, obs_size=obs_size, global_obs_size=obs_size * (n_agents + 1), action_size=n_nodes,
                       hidden_size=hidden, device="cpu")
    if ci == 2:  # make some logits extreme so the `probs_sum <= 1e-8` branch is reachable
        with torch.no_grad():
            agent.policies[0].actor[2].bias[:5] += 60.0
    for i, pol in enumerate(agent.policies):
        for k, v in pol.state_dict().items():
            out[f"c{ci}_p{i}_{k}"] = v.numpy().copy()
    for k, v in agent.critic.state_dict().items():
        out[f"c{ci}_{k}"] = v.numpy().copy()
    obs_l, mask_l, act_l, lp_l, pr_l, pid_l = [], [], [], [], [], []
    for t in range(24):
        pid = t % n_agents
        obs = torch.from_numpy(rng.normal(size=obs_size).astype(np.float32) * (3.0 if t % 5 == 0 else 1.0))
        k = int(rng.integers(0, 6))
        mask = np.zeros(n_nodes, dtype=np.float32)
        if t % 7 == 3:
            pass  # empty mask -> uniform over all nodes
        elif ci == 2 and t % 4 == 1 and pid == 0:
            mask[rng.choice(np.arange(5, n_nodes), size=3, replace=False)] = 1  # mass underflows -> uniform over the mask
        else:
            mask[rng.choice(n_nodes, size=max(k, 1), replace=False)] = 1
        a, lp, probs = agent.select_action(pid, obs, torch.from_numpy(mask))
        obs_l.append(obs.numpy()); mask_l.append(mask); act_l.append(a); lp_l.append(float(lp)); pr_l.append(probs.numpy()); pid_l.append(pid)
    gobs = rng.normal(size=(8, obs_size * (n_agents + 1))).astype(np.float32)
    with torch.no_grad():
        out[f"c{ci}_values"] = agent.critic(torch.from_numpy(gobs)).squeeze(-1).numpy()
    out[f"c{ci}_gobs"] = gobs
    out[f"c{ci}_obs"], out[f"c{ci}_mask"] = np.stack(obs_l), np.stack(mask_l)
    out[f"c{ci}_action"], out[f"c{ci}_logp"] = np.asarray(act_l), np.asarray(lp_l, dtype=np.float32)
    out[f"c{ci}_probs"], out[f"c{ci}_pid"] = np.stack(pr_l), np.asarray(pid_l)
    cases.append((n_agents, obs_size, hidden, n_nodes))
out["cases"] = np.asarray(cases)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "mappo.npz"), **out)
print("wrote tests/golden/mappo.npz", len(out), "arrays")
