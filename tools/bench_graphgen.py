"""Row f3 measurement: graphs/s of the device pool sampler (sy_generate_graphs, all tables incl. APSP built) against
the host numpy generator (graphs.py) and the reference's own sampler as restated draw for draw in the oracle
(graph_layout.py:9-80 is O(N^3) Python).  Usage: python tools/bench_graphgen.py [out.json]"""
import json
import random
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import student_mechanism_design_b200 as pkg  # noqa: E402
from oracle import sy_oracle as so  # noqa: E402

out = {}
for name, N, E, G in (("c3_200n_400e", 200, 400, 4096), ("c1_50n_110e", 50, 110, 16384), ("c4_1000n_2000e", 1000, 2000, 256)):
    env = pkg.BatchedScotlandYardEnv(1024, 6, 20, graph_nodes=N, graph_edges=E, graphs="device", num_graphs=G, seed=1,
                                     belief=True)
    lib, h = env._lib, env._handle
    ts = []
    for gen in range(1, 4):
        torch.cuda.synchronize()
        t = time.perf_counter()
        pkg._cabi.check(lib.sy_generate_graphs(h, G, E, 4, 5, 1, gen, 0, None, env._stream()))
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t)
    dev = G / min(ts)
    attempts = float(np.mean(env.generation_attempts))
    env.close()
    rng = np.random.default_rng(0)
    t = time.perf_counter()
    n_host = 20 if N >= 1000 else 100
    for _ in range(n_host):
        pkg.generate_connected_graph(N, E, rng)
    host = n_host / (time.perf_counter() - t)
    n_ref = 1 if N >= 1000 else (3 if N >= 200 else 20)
    pr, nr = random.Random(0), np.random.RandomState(0)
    t = time.perf_counter()
    for _ in range(n_ref):
        so.sample_connected_graph(N, E, pr, nr)
    ref = n_ref / (time.perf_counter() - t)
    out[name] = dict(nodes=N, edges=E, pool=G, device_graphs_per_s=dev, device_ms_per_pool=min(ts) * 1e3,
                     mean_attempts=attempts, host_numpy_graphs_per_s=host, reference_sampler_graphs_per_s=ref,
                     note="device time includes CSR, dense weights, move counts, neighbour lists and all-pairs distances")
    print(name, out[name], flush=True)
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"), indent=1)
