import sys, torch, statistics
sys.path.insert(0, '/root/repo')
from student_mechanism_design_b200 import BatchedScotlandYardEnv
for cs in (True, False):
    env = BatchedScotlandYardEnv(65536, 6, 20, graph_nodes=200, graph_edges=400, seed=0, tolls=1, belief=True, reveal_interval=5, auto_reset=True, collect_stats=cs)
    env.reset()
    a = torch.empty(65536, 7, dtype=torch.int64, device='cuda')
    for s in range(30):
        env.sample_actions(out=a, step_counter=s); env.step(a)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(200)]
    torch.cuda.synchronize()
    for k in range(200):
        env.sample_actions(out=a, step_counter=30 + k)
        ev[k][0].record(); env.step(a); ev[k][1].record()
    torch.cuda.synchronize()
    print('collect_stats', cs, 'sy_step ms', statistics.mean(x.elapsed_time(y) for x, y in ev))
    env.close()
