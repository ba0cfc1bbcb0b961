"""Write tests/golden/gnn.npz from the UNMODIFIED reference GNN (src/agent/gnn_agent.py:230-257: two torch_geometric
AntiSymmetricConv layers + Linear) -- the golden vectors that would PIN the GNN restatement of oracle/policy_oracle.py
and the sy_gnn_act / sy_gnn_q_values kernels.

torch_geometric is a requirements.txt dependency of the reference that is not installed in the build image and cannot be
fetched (no network), so this script cannot run there: the GNN's parity stays UNPINNED until somebody runs it where PyG
exists (`pip install torch_geometric`; CPU is enough) and commits the file.  tests/test_policy_oracle.py and
tests/test_gpu_policy.py pick the file up automatically when it is present.

    python oracle/gen_gnn_golden.py [/path/to/reference]
"""
import importlib.util
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"


def main():
    try:
        import torch
        from torch_geometric.data import Data
    except ImportError as e:
        raise SystemExit(f"gen_gnn_golden.py needs torch_geometric ({e}); the GNN parity stays unpinned")
    spec = importlib.util.spec_from_file_location("ref_gnn_agent", os.path.join(REF, "src", "agent", "gnn_agent.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import sy_oracle as so

    torch.manual_seed(0)
    rng = np.random.default_rng(0)
    out, cases = {}, []
    for ci, (N, E, K) in enumerate([(15, 20, 3), (50, 110, 4), (200, 400, 7), (33, 60, 16)]):
        g = so.philox_sample_graph_once(1, ci, 0, 0, N, E)
        model = mod.GNNModel(K)
        with torch.no_grad():
            for p in model.parameters():  # the default init leaves W = 0: make every term matter
                p.copy_(torch.from_numpy(rng.normal(size=tuple(p.shape)).astype(np.float32) * 0.5))
        for k, v in model.state_dict().items():
            out[f"c{ci}_{k}"] = v.numpy().copy()
        xs, qs = [], []
        for t in range(6):
            x = np.zeros((N, K), dtype=np.float32)
            pos = rng.choice(N, size=K, replace=False)
            for k in range(K):
                if not (k == 0 and t % 3 == 2):  # MrX hidden in a third of the cases
                    x[pos[k], k] = 1
            data = Data(x=torch.from_numpy(x), edge_index=torch.from_numpy(np.asarray(g.edge_links, dtype=np.int64).T.copy()))
            with torch.no_grad():
                qs.append(model(data).numpy().copy())
            xs.append(x)
        out[f"c{ci}_edge_links"] = np.asarray(g.edge_links, dtype=np.int32)
        out[f"c{ci}_x"], out[f"c{ci}_q"] = np.stack(xs), np.stack(qs)
        cases.append((N, E, K))
    out["cases"] = np.asarray(cases)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "gnn.npz"), **out)
    print("wrote tests/golden/gnn.npz", len(out), "arrays")


if __name__ == "__main__":
    main()
