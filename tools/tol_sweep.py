"""Effect of the Lanczos residual tolerance on time and on the tree (vs the strict default 1e-10)."""
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
ref = None
for tol in (1e-10, 1e-8, 1e-7, 1e-6):
    T.hierarchical_spectral_cluster(b, tol=tol)
    t0 = time.perf_counter(); tr = T.hierarchical_spectral_cluster(b, tol=tol); dt = time.perf_counter() - t0
    ls = tr.leaf_sets()
    if ref is None: ref = ls
    moved = 0
    if ls != ref:
        a = {c: i for i, l in enumerate(ref) for c in l}; bb = {c: i for i, l in enumerate(ls) for c in l}
        # cells whose leaf (as a set) differs
        refsets = {frozenset(l) for l in ref}
        moved = sum(len(l) for l in ls if frozenset(l) not in refsets)
    print(f"{name} tol={tol:g}: {dt*1e3:.0f} ms, lanczos steps {int(tr.iters.sum())}, nodes {tr.n_nodes}, identical leaves {ls == ref}, cells in differing leaves {moved}", flush=True)
