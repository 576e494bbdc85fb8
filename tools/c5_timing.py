import os, sys, time
sys.path.insert(0, "/root/repo")
import torch
import too_many_cells_b200 as T
from too_many_cells_b200 import synth
n, g, k, d, s, binary = synth.CONFIGS["c3"]
dev = torch.device("cuda", 0)
rp, col, val = synth.synth_counts_torch(n, g, k, d, s, dev, binary)
b = T.DeviceB.adopt_device(n, g, rp, col, val, normalize=True, device=0)
for it in range(2):
    t0 = time.perf_counter()
    tr = T.hierarchical_spectral_cluster(b, min_modularity=-1.0, max_active=16384)
    t1 = time.perf_counter()
    h = tr.tree_hash()
    t2 = time.perf_counter()
    print(f"it {it}: call {1e3*(t1-t0):.1f} ms (engine {tr.stats['engine_ms']:.1f}), tree_hash {1e3*(t2-t1):.1f} ms, nodes {tr.n_nodes}", flush=True)
