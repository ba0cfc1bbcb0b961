#!/usr/bin/env python
"""Time the UNMODIFIED reference environment (/root/reference/src/environment/yard.py, imported through the stubs of
oracle/ref_loader.py: no-op logger and visualiser -- this flatters the reference, whose shipped logger writes ~8 + 8P
TensorBoard scalars per step) on all cores of THIS host: one process per core, each stepping its own env with uniformly
random valid actions and resetting on termination.  BUILD CONTAINER ONLY (/root/reference does not exist on the GPU box);
writes profiles/r02_reference_python.json, the "reference CPU env timed on the same host's cores" record that BASELINE.md
section 4 cites next to the C-oracle CPU arm of bench.py.

    python tools/time_reference_python.py [--seconds 10] [--c3-steps 6]
"""
import argparse
import json
import multiprocessing as mp
import os
import platform
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

WEIGHTS = {"Police_distance": 0.1, "Police_group": 0.1, "Police_position": 0.1, "Police_time": 0.0, "Mrx_closest": 0.3,
           "Mrx_average": 0.2, "Mrx_position": 0.1, "Mrx_time": 0.0, "Police_coverage": 0.05, "Police_proximity": 0.05,
           "Police_overlap_penalty": 0.0}  # src/training/evaluator.py:59-71
CONFIGS = {  # BASELINE.json configs 1-3 as far as the reference env implements them (no tolls / belief / reveal in its step)
    "c1": dict(N=15, E=20, P=2, money=10),
    "c2": dict(N=50, E=110, P=3, money=10),
    "c3": dict(N=200, E=400, P=6, money=20),
}


def worker(args):
    name, cfg, seconds, max_steps, seed = args
    import numpy as np

    import ref_loader

    t0 = time.perf_counter()
    env = ref_loader.make_reference_env(cfg["P"], cfg["money"], dict(WEIGHTS), cfg["N"], cfg["E"], seed=seed)
    build_s = time.perf_counter() - t0
    rng = np.random.default_rng(seed)
    env.reset()
    n, t0 = 0, time.perf_counter()
    while (time.perf_counter() - t0 < seconds) and (max_steps is None or n < max_steps):
        acts = {}
        for i, a in enumerate(env.possible_agents):  # the trainers' random valid move (gnn_trainer.py:221-229)
            moves = env.get_possible_moves(i)
            acts[a] = int(rng.choice(moves)) if len(moves) else env.DEFAULT_ACTION
        _, _, term, trunc, _ = env.step(acts)
        n += 1
        if any(term.values()) or any(trunc.values()):
            env.reset()
    return n, time.perf_counter() - t0, build_s


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=10.0)
    ap.add_argument("--c3-steps", type=int, default=6)
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r02_reference_python.json"))
    args = ap.parse_args()
    try:
        cores = len(os.sched_getaffinity(0))
    except Exception:
        cores = os.cpu_count() or 1
    out = {"what": "unmodified reference env (yard.py + reward_calculator.py + pathfinding.py + action_mask.py), stub gymnasium / "
                   "pettingzoo, no-op logger and visualiser, random valid actions, one process per core",
           "host": platform.processor() or platform.machine(), "cores": cores, "python": platform.python_version(), "configs": {}}
    ctx = mp.get_context("spawn")
    for name, cfg in CONFIGS.items():
        max_steps = args.c3_steps if name == "c3" else None
        seconds = 600.0 if name == "c3" else args.seconds
        with ctx.Pool(cores) as pool:
            res = pool.map(worker, [(name, cfg, seconds, max_steps, 100 + r) for r in range(cores)])
        steps = sum(r[0] for r in res)
        per_core = [r[0] / r[1] for r in res]
        out["configs"][name] = {**cfg, "env_steps": steps, "env_steps_per_s_per_core": sum(per_core) / len(per_core),
                                "env_steps_per_s_all_cores": sum(per_core), "seconds_per_process": max(r[1] for r in res),
                                "constructor_seconds": sum(r[2] for r in res) / len(res)}
        print(name, json.dumps(out["configs"][name]), flush=True)
    json.dump(out, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
