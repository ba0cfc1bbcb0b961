#!/usr/bin/env python
"""Per-role clock accounting inside the lagged step kernel (needs a -DSY_FUSED_CLOCKS variant via SY_LIB_PATH):
average time a CTA's writer warp 0, belief warp 0 and dynamics warp 0 spend in their role."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from bench import WORKLOADS  # noqa: E402
from student_mechanism_design_b200 import BatchedScotlandYardEnv, _cabi  # noqa: E402

wl = WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c3"]
env = BatchedScotlandYardEnv(wl["B"], wl["P"], wl["money"], graph_nodes=wl["N"], graph_edges=wl["E"], seed=0, tolls=wl["toll"],
                             belief=wl["belief"], reveal_interval=wl["reveal"], auto_reset=True)
env.set_option("step_kernel", "two_kernels")
plain = len(sys.argv) > 2 and sys.argv[2] == "plain"  # the two-launch sy_step for comparison
step = env.step if plain else env.step_deferred
env.reset()
lib = _cabi.load_library()
out = (C.c_ulonglong * 16)()
a = torch.empty(wl["B"], wl["P"] + 1, dtype=torch.int64, device="cuda")
for s in range(20):
    env.sample_actions(out=a, step_counter=s)
    step(a)
lib.sy_debug_fused_clocks(out, 1)
K = 50
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record()
for s in range(K):
    step(a)
e1.record()
torch.cuda.synchronize()
lib.sy_debug_fused_clocks(out, 1)
v = [int(x) for x in out]
ctas = v[6] if plain else v[5]
us = lambda c: c / ctas / 1965.0  # noqa: E731
print(json.dumps({"ms_per_step": e0.elapsed_time(e1) / K, "ctas_per_step": ctas / K, "writer_us": us(v[1]), "belief_us": us(v[3]), "logic_us": us(v[4])}))
