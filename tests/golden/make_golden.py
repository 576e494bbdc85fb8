"""Regenerates tests/golden/synth_*.npz: synthetic inputs + the oracle's outputs for them.

The reference holds no numeric fixtures for this path (no tests; the
workshop outputs have no input matrix), and its Haskell engine cannot be built here, so these
vectors pin the ORACLE (oracle/tmc_oracle.py, "parity unpinned" against the reference) and give
the GPU tests a fixed, committed target.  Run:  python tests/golden/make_golden.py
"""
import os, sys
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, scipy.sparse as sp
from too_many_cells_b200 import synth
import tmc_oracle as O

for name in ("tiny", "small"):
    ip, ix, dv, leaf = synth.synth_config(name)
    n, g = synth.CONFIGS[name][:2]
    a = sp.csr_matrix((dv, ix, ip), shape=(n, g))
    b2 = O.b1_to_b2(a)
    b = O.b2_to_b(b2)
    tr = O.hierarchical_spectral_cluster(b2, keep_u2=True)
    leaf_of_row = np.full(n, -1, dtype=np.int32)
    for i in tr.leaves():
        leaf_of_row[tr.items[i]] = i
    lab, u2, th, gap, d = O.spectral_cluster(b)
    np.savez_compressed(
        os.path.join(HERE, f"synth_{name}.npz"),
        indptr=ip.astype(np.int64), indices=ix.astype(np.int32), counts=dv.astype(np.uint16), shape=np.array([n, g]),
        tfidf_head=b2.data[:4096], b_head=b.data[:4096],
        left=np.array(tr.left, dtype=np.int64), right=np.array(tr.right, dtype=np.int64), q=np.array(tr.q),
        theta=np.array(tr.theta), gap=np.array(tr.gap), margin=np.array(tr.margin),
        size=np.array([x.size for x in tr.items], dtype=np.int64), leaf_of_row=leaf_of_row,
        root_u2=u2, root_d=d, root_q=np.array(O.get_b_modularity(lab, b)))
    print(name, "nodes", tr.n_nodes, "leaves", len(tr.leaves()))

# ---- the non-default variants on `tiny`: KMeansGroup trees and permutation p-values
ip, ix, dv, _ = synth.synth_config("tiny")
n, g = synth.CONFIGS["tiny"][:2]
a = sp.csr_matrix((dv, ix, ip), shape=(n, g))
out = {}
for tag, kw in (("km1", dict(eigen_group="KMeansGroup", num_eigen=1)), ("km3", dict(eigen_group="KMeansGroup", num_eigen=3)),
                ("perm", dict(num_runs=20))):
    tr = O.hierarchical_spectral_cluster(a, normalize=True, **kw)
    leaf_of_row = np.full(n, -1, dtype=np.int32)
    for i in tr.leaves():
        leaf_of_row[tr.items[i]] = i
    out[f"{tag}_left"] = np.array(tr.left, dtype=np.int64); out[f"{tag}_right"] = np.array(tr.right, dtype=np.int64)
    out[f"{tag}_q"] = np.array(tr.q); out[f"{tag}_leaf_of_row"] = leaf_of_row
    out[f"{tag}_p_val"] = np.array([np.nan if p is None else p for p in tr.p_val])
    print("variants", tag, "nodes", tr.n_nodes)
sat = O.hierarchical_spectral_cluster(a[:60], normalize=True, num_runs=10, min_modularity=-1.0, seed=7)
out["sat_left"] = np.array(sat.left, dtype=np.int64); out["sat_p_val"] = np.array([np.nan if p is None else p for p in sat.p_val])
out["splitmix_1234567"] = np.array([O.SplitMix64(1234567).next() for _ in range(1)], dtype=np.uint64)
np.savez_compressed(os.path.join(HERE, "variants_tiny.npz"), **out)
