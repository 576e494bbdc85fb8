"""More than 2^31 non-zeros on one GPU: 1.35 M cells x 30 k genes at the c3 density (~2.26e9 nnz, int64 row_ptr throughout).
Checks that the tree is a permutation of the cells with the usual invariants (split iff Q > 0, children tile the parent, leaves
ascending) and prints the time.      python tools/big_nnz_check.py [cells]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import too_many_cells_b200 as T
from too_many_cells_b200 import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1350000
_, g, k, d, s, binary = synth.CONFIGS["c3"]
dev = torch.device("cuda", 0)
t0 = time.time()
rp, col, val = synth.synth_counts_torch(n, g, k, d, s, dev, binary)
torch.cuda.synchronize()
nnz = int(val.numel())
print(f"generated {n} x {g}, nnz {nnz} ({'>' if nnz > 2**31 else '<='} 2^31) in {time.time() - t0:.1f} s", flush=True)
b = T.DeviceB.adopt_device(n, g, rp, col, val, normalize=True, device=0)
print("storage", b.info(), flush=True)
t0 = time.perf_counter()
t = T.hierarchical_spectral_cluster(b)
ms = 1e3 * (time.perf_counter() - t0)
assert np.array_equal(np.sort(t.perm), np.arange(n))
inner = np.nonzero(t.left >= 0)[0]; leaves = t.leaves(); sizes = t.item_end - t.item_begin
assert len(leaves) == len(inner) + 1
assert np.all(t.q[inner] > 0) and np.all(t.q[inner] <= 0.5) and np.all(t.q[leaves[sizes[leaves] >= 2]] <= 0)
assert np.array_equal(t.item_end[t.left[inner]], t.item_begin[t.right[inner]])
assert np.array_equal(sizes[inner], sizes[t.left[inner]] + sizes[t.right[inner]])
for i in leaves[:: max(1, len(leaves) // 100)]:
    assert np.all(np.diff(t.items(i)) > 0)
print(f"ok: {t.n_nodes} nodes, {len(leaves)} leaves, {int(t.iters.sum())} Lanczos steps, {t.stats['n_sweeps']} sweeps, {ms:.0f} ms (engine {t.stats['engine_ms']:.0f} ms), hash {t.tree_hash()[:12]}", flush=True)
b.free()
