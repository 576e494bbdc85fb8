"""First-contact GPU check: every fine-grained export against the oracle, then a tree."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))      # checker scripts live under tests/: only tests may import the oracle
import numpy as np, scipy.sparse as sp
import too_many_cells_b200 as T
from too_many_cells_b200 import synth
import tmc_oracle as O

def relerr(a, b):
    return float(np.max(np.abs(a - b)) / max(1e-300, np.max(np.abs(b))))

name = sys.argv[1] if len(sys.argv) > 1 else "tiny"
ip, ix, dv, leaf = synth.synth_config(name)
n, g = synth.CONFIGS[name][:2]
A = sp.csr_matrix((dv, ix, ip), shape=(n, g))
print("matrix", A.shape, A.nnz, flush=True)
# A1
v, idf = T.b1_to_b2(A)
B2 = O.b1_to_b2(A)
print("tfidf relerr", relerr(v, B2.data), flush=True)
# A2
v2, nrm = T.b2_to_b(B2)
B = O.b2_to_b(B2)
print("row_l2 relerr", relerr(v2, B.data), flush=True)
dB = T.DeviceB.from_host(B2)
# A3
d = T.b_to_d(dB)
print("degree relerr", relerr(d, O.b_to_d(B)), flush=True)
rows = np.arange(0, n, 3)
d = T.b_to_d(dB, rows)
print("degree(subset) relerr", relerr(d, O.b_to_d(B[rows])), flush=True)
# pair
x = np.random.default_rng(0).standard_normal(n)
Cm = O.bd_to_c(B, O.b_to_d(B))
print("spmv_pair relerr", relerr(T.spmv_pair(dB, x), Cm @ (Cm.T @ x)), flush=True)
# A4-A6
t = time.time()
u2, th, it = T.second_left(dB)
print("second_left: theta", th, "iters", it, "time", time.time() - t, flush=True)
lab_o, u2o, tho, gap, _ = O.spectral_cluster(B)
s = 1.0 if u2 @ u2o >= 0 else -1.0
print("  oracle theta", tho, "gap", gap, "u2 err", float(np.linalg.norm(s * u2 - u2o)), "label mismatches",
      int(np.sum((s * u2 >= 0) != (u2o >= 0))), flush=True)
# A7
q = T.get_b_modularity(lab_o, dB)
print("modularity", q, "oracle", O.get_b_modularity(lab_o, B), flush=True)
# A8
ro, nl = T.partition_rows(np.arange(n), lab_o)
print("partition ok", bool(np.array_equal(ro, np.concatenate([np.nonzero(lab_o == 0)[0], np.nonzero(lab_o == 1)[0]]))), nl, flush=True)
# A0
t = time.time()
tree = T.hierarchical_spectral_cluster(dB, profile=True)
print("tree: nodes", tree.n_nodes, "leaves", len(tree.leaves()), "wall", time.time() - t, tree.stats, flush=True)
t = time.time()
ot = O.hierarchical_spectral_cluster(B2)
print("oracle: nodes", ot.n_nodes, "leaves", len(ot.leaves()), "wall", time.time() - t, flush=True)
print("leaf sets equal:", tree.leaf_sets() == ot.leaf_sets(), flush=True)
ca = O.canonical_tree(tree.left, tree.right, tree.items)
cb = O.canonical_tree(ot.left, ot.right, lambda i: ot.items[i])
print("topology equal (mod child swap):", ca == cb, flush=True)
