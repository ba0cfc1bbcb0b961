#!/usr/bin/env python
"""A handful of steps of one workload (for an ncu launch list): python tools/exp_two_steps.py c3 [opt=value ...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from bench import WORKLOADS  # noqa: E402
from student_mechanism_design_b200 import BatchedScotlandYardEnv  # noqa: E402

wl = WORKLOADS[sys.argv[1]]
env = BatchedScotlandYardEnv(wl["B"], wl["P"], wl["money"], graph_nodes=wl["N"], graph_edges=wl["E"], seed=0, tolls=wl["toll"],
                             belief=wl["belief"], reveal_interval=wl["reveal"], auto_reset=True)
for o in sys.argv[2:]:
    k, v = o.split("=")
    env.set_option(k, v)
env.reset()
for s in range(12):
    env.step(env.sample_actions(step_counter=s))
torch.cuda.synchronize()
print("ok")
