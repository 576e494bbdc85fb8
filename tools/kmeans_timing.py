"""KMeansGroup (--eigen-group KMeansGroup, one eigenvector) inside the batched sweeps, at scale.   python tools/kmeans_timing.py [c2]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import too_many_cells_b200 as T
from too_many_cells_b200 import synth
name = sys.argv[1] if len(sys.argv) > 1 else "c2"
n, g, k, d, s, binary = synth.CONFIGS[name]
dev = torch.device("cuda", 0)
rp, col, val = synth.synth_counts_torch(n, g, k, d, s, dev, binary)
b = T.DeviceB.adopt_device(n, g, rp, col, val, normalize=True, device=0)
ts = T.hierarchical_spectral_cluster(b)
for it in range(2):
    t0 = time.perf_counter(); tk = T.hierarchical_spectral_cluster(b, eigen_group="KMeansGroup"); dt = time.perf_counter() - t0
print(f"{name}: SignGroup {ts.n_nodes} nodes in {ts.stats['engine_ms']:.0f} ms ({int(ts.iters.sum())} steps, {ts.stats['n_sweeps']} sweeps); "
      f"KMeansGroup {tk.n_nodes} nodes in {1e3 * dt:.0f} ms (engine {tk.stats['engine_ms']:.0f} ms, {int(tk.iters.sum())} steps, {tk.stats['n_sweeps']} sweeps)", flush=True)
b.free()
