"""profiling experiment: average cycles per logic phase (library built with -DSY_PHASE_CLOCKS), for the two-kernel path and
the fused persistent kernel"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from student_mechanism_design_b200 import BatchedScotlandYardEnv, _cabi
lib = _cabi.load_library()
names = ["P0 stage+barrier", "P1 moves (warp0)", "P1 barrier wait", "P2 rewards (warp0)", "P2 barrier wait", "P3 advance (warp0)", "P3 barrier wait", "P4 store", "w0 moves", "w1 philox", "w0 P2a", "w3 P2a", "-", "-", "-", "-"]
for mode in ("two_kernels", "fused", "lagged"):
    env = BatchedScotlandYardEnv(65536, 6, 20, graph_nodes=200, graph_edges=400, seed=0, tolls=1, belief=True, reveal_interval=5, auto_reset=True)
    env.set_option("step_kernel", "two_kernels" if mode == "lagged" else mode)
    step = env.step_deferred if mode == "lagged" else env.step
    env.reset()
    a = torch.empty(65536, 7, dtype=torch.int64, device="cuda")
    for s in range(20):
        env.sample_actions(out=a, step_counter=s); step(a)
    buf = (ctypes.c_ulonglong * 16)()
    lib.sy_debug_phase_clocks(buf, 1)
    K = 50
    for s in range(K):
        env.sample_actions(out=a, step_counter=20 + s); step(a)
    lib.sy_debug_phase_clocks(buf, 0)
    tiles = 2048 * K
    tot = 0
    print("==", mode)
    for n, v in zip(names, buf):
        print(f"{n:24s} {v / tiles:9.0f} cycles"); tot += v / tiles if n.startswith("P") else 0
    print("sum", tot)
    env.close()
