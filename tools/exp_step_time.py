#!/usr/bin/env python
"""Time the step of one workload for one library build / option set (experiments; prints one JSON line).
    SY_LIB_PATH=variants/libsy_env_x.so python tools/exp_step_time.py --workload c3 --writer bulk
Reports: CUDA-event time of sy_step alone (python loop), and per-step time of the replayed rollout graph (the timed path
of bench.py), median of 5 passes."""
import argparse
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from bench import WORKLOADS, algorithmic_bytes_per_env_step  # noqa: E402
from student_mechanism_design_b200 import BatchedScotlandYardEnv  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c3")
    ap.add_argument("--writer", default=None)
    ap.add_argument("--opt", action="append", default=[], help="name=value for env.set_option")
    ap.add_argument("--envs", type=int, default=0)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--tag", default="")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    B = args.envs or wl["B"]
    env = BatchedScotlandYardEnv(B, wl["P"], wl["money"], graph_nodes=wl["N"], graph_edges=wl["E"], seed=0, tolls=wl["toll"],
                                 belief=wl["belief"], reveal_interval=wl["reveal"], auto_reset=True)
    if args.writer:
        env.set_option("writer_path", args.writer)
    for o in args.opt:
        k, v = o.split("=")
        env.set_option(k, int(v) if v.lstrip("-").isdigit() else v)
    env.reset()
    A = wl["P"] + 1
    actions = torch.empty(B, A, dtype=torch.int64, device=env.device)
    for s in range(30):
        env.sample_actions(out=actions, step_counter=s)
        env.step(actions)
    K = args.steps
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    torch.cuda.synchronize()
    for k in range(K):
        env.sample_actions(out=actions, step_counter=30 + k)
        ev[k][0].record()
        env.step(actions)
        ev[k][1].record()
    torch.cuda.synchronize()
    step_ms = statistics.median(a.elapsed_time(b) for a, b in ev)
    seg = 50
    env._sample_counter = 30 + K
    graph, _ = env.capture_rollout(seg, actions=actions)
    passes = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(5):
        torch.cuda.synchronize()
        e0.record()
        for _ in range(max(1, K // seg)):
            graph.replay()
        e1.record()
        torch.cuda.synchronize()
        passes.append(e0.elapsed_time(e1) / (max(1, K // seg) * seg))
    bstep = algorithmic_bytes_per_env_step(wl["N"], wl["P"], wl["belief"])
    gms = statistics.median(passes)
    print(json.dumps({"tag": args.tag, "lib": os.environ.get("SY_LIB_PATH", "in-tree"), "workload": args.workload, "envs": B,
                      "writer": args.writer, "opts": args.opt, "sy_step_ms": round(step_ms, 5), "sy_step_frac": round(bstep * B / (step_ms * 1e-3) / 6549.4e9, 4),
                      "graph_ms_per_step": round(gms, 5), "graph_env_steps_per_s": round(B / (gms * 1e-3)),
                      "graph_frac": round(bstep * B / (gms * 1e-3) / 6549.4e9, 4), "passes": [round(x, 5) for x in passes]}), flush=True)
    env.close()


if __name__ == "__main__":
    main()
