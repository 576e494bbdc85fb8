"""--numRuns at scale: the batched permutation test on a synthetic config.   python tools/perm_timing.py [c2] [runs]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import too_many_cells_b200 as T
from too_many_cells_b200 import synth
name = sys.argv[1] if len(sys.argv) > 1 else "c2"
runs = int(sys.argv[2]) if len(sys.argv) > 2 else 10
n, g, k, d, s, binary = synth.CONFIGS[name]
dev = torch.device("cuda", 0)
rp, col, val = synth.synth_counts_torch(n, g, k, d, s, dev, binary)
b = T.DeviceB.adopt_device(n, g, rp, col, val, normalize=True, device=0)
t = T.hierarchical_spectral_cluster(b)
t0 = time.perf_counter()
p = T.permutation_test(b, t, runs)
dt = time.perf_counter() - t0
inner = t.left >= 0
print(f"{name}: {t.n_nodes} nodes, {int(inner.sum())} splits x {runs} runs: {1e3 * dt:.0f} ms (tree itself {t.stats['engine_ms']:.0f} ms); "
      f"p == 0 for {int((p[inner] == 0).sum())} splits, max p {np.nanmax(p):.2f}", flush=True)
b.free()
