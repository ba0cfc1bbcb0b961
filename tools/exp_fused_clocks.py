#!/usr/bin/env python
"""Per-role clock accounting inside the fused step kernel (needs a -DSY_FUSED_CLOCKS variant via SY_LIB_PATH)."""
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
env.set_option("step_kernel", "fused")
env.reset()
lib = _cabi.load_library()
out = (C.c_ulonglong * 16)()
for s in range(20):
    env.step(env.sample_actions(step_counter=s))
lib.sy_debug_fused_clocks(out, 1)
K = 50
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
acts = [env.sample_actions(step_counter=100 + s).clone() for s in range(1)]
torch.cuda.synchronize()
e0.record()
for s in range(K):
    env.step(acts[0])
e1.record()
torch.cuda.synchronize()
lib.sy_debug_fused_clocks(out, 1)
v = [int(x) for x in out]
ctas = v[5] / K
mhz = 1965.0
us = lambda c: c / ctas / K / mhz  # noqa: E731  per CTA per step, microseconds at max clock
print(json.dumps({"ms_per_step": e0.elapsed_time(e1) / K, "ctas": ctas,
                  "writer_wait_logic_us": us(v[0]), "writer_total_us": us(v[1]), "writer_wait_read_us": us(v[6]), "writer_wait_fill_us": us(v[7]),
                  "w_fill_issue_us": us(v[8]), "w_img_wait_zero_us": us(v[9]), "w_img_ones_us": us(v[10]), "w_fence_us": us(v[11]), "w_img_issue_us": us(v[12]), "w_ones_stage_us": us(v[13]),
                  "belief_wait_logic_us": us(v[2]), "belief_total_us": us(v[3]), "logic_total_us": us(v[4])}))
