"""BASELINE.json configs[4] style run: saturated tree (min-modularity -1: every cell ends as its own leaf)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import too_many_cells_b200 as T
from too_many_cells_b200 import synth
name = sys.argv[1] if len(sys.argv) > 1 else "c2"
n, g, k, d, s, binary = synth.CONFIGS[name]
dev = torch.device("cuda", 0)
rp, col, val = synth.synth_counts_torch(n, g, k, d, s, dev, binary)
b = T.DeviceB.adopt_device(n, g, rp, col, val, normalize=True, device=0)
for mm in (0.0, -1.0):
    t0 = time.perf_counter()
    tr = T.hierarchical_spectral_cluster(b, min_modularity=mm)
    dt = time.perf_counter() - t0
    depth = np.zeros(tr.n_nodes, dtype=np.int32)
    for i in range(1, tr.n_nodes):
        depth[i] = depth[tr.parent[i]] + 1
    print(f"{name} min_modularity={mm}: {dt*1e3:.0f} ms, nodes {tr.n_nodes}, leaves {len(tr.leaves())}, depth {int(depth.max())}, "
          f"sweeps {tr.stats['n_sweeps']}, lanczos steps {int(tr.iters.sum())}, perm ok {bool(np.array_equal(np.sort(tr.perm), np.arange(n)))}", flush=True)
