"""Small end-to-end target for compute-sanitizer: trees in both storage modes, both transposed-product paths, API calls, ingest."""
import os, sys, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, scipy.sparse as sp, scipy.io
import too_many_cells_b200 as T
from too_many_cells_b200 import synth
ip, ix, dv, _ = synth.synth_counts(700, 1500, 90, 3, 5)
a = sp.csr_matrix((dv, ix, ip), shape=(700, 1500))
t = T.hierarchical_spectral_cluster(a, normalize=True)                       # factored uint16
b2 = T.tfidf_scale_sparse_mat(a)
t2 = T.hierarchical_spectral_cluster(b2)                                      # fp64 B
t3 = T.hierarchical_spectral_cluster(a, normalize=True, scatter_mode=1)       # CSR scatter with atomics
t4 = T.hierarchical_spectral_cluster(a[:200], normalize=True, min_modularity=-1.0)   # saturated: 1- and 2-cell nodes
ab = a.copy(); ab.data[:] = 1.0
t5 = T.hierarchical_spectral_cluster(ab, normalize=True)                      # all-ones storage mode
db = T.DeviceB.from_host(b2)
rows = np.arange(3, 700, 7)
d = T.b_to_d(db, rows); u2, th, it = T.second_left(db, rows); q = T.get_b_modularity((u2 >= 0).astype(np.uint8), db, rows)
x = T.spmv_pair(db, np.ones(rows.size), rows)
buf = io.BytesIO(); scipy.io.mmwrite(buf, sp.coo_matrix(a.T), field="integer")
m = T.parse_matrix_market(buf.getvalue()); f, kr, kc = T.filter_num_sparse_mat(m, 50, 3)
print("ok", t.n_nodes, t2.n_nodes, t3.n_nodes, t4.n_nodes, t5.n_nodes, d.shape, th, q, f.shape, t.leaf_sets() == t2.leaf_sets() == t3.leaf_sets())
