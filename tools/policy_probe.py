"""Run the policy kernels alone on a c3-sized batch (for ncu / quick timing): python tools/policy_probe.py [gnn|mappo|both] [B]"""
import sys

import torch

sys.path.insert(0, ".")
import student_mechanism_design_b200 as pkg  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "both"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
N, P = 200, 6
A = P + 1
env = pkg.BatchedScotlandYardEnv(B, P, 20, graph_nodes=N, graph_edges=400, seed=0, auto_reset=True, tolls=1, belief=True,
                                 reveal_interval=5)
env.reset()
for s in range(20):
    env.step(env.sample_actions(step_counter=s))
acts = torch.empty(B, A, dtype=torch.int64, device="cuda")


def timed(fn, n=20):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


if which in ("gnn", "both"):
    gnn = pkg.GNNPolicy(env, seed=0)
    for eps in (0.05, 1.0):
        print(f"gnn act eps={eps}: {timed(lambda: gnn.act(eps, eps, out=acts)) * 1e3:.1f} us", flush=True)
    print(f"gnn dense q: {timed(lambda: gnn.q_values(), 5) * 1e3:.1f} us", flush=True)
if which in ("mappo", "both"):
    mp = pkg.MappoPolicy(env, obs_size=A, hidden_size=64, seed=0)
    obs = (env.pos.float() / N).unsqueeze(1).expand(B, A, A).contiguous()
    for tc in (True, False):
        mp.tensor_cores = tc
        print(f"mappo act (tensor cores {tc}, max degree {mp._graphs() and mp._tables.max_degree}): "
              f"{timed(lambda: mp.act(obs), 5) * 1e3:.1f} us", flush=True)
        mp.check()
