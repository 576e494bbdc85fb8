"""Robustness sweep on odd inputs: must finish (no hang / crash) and return a valid partition of the cells."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))      # checker scripts live under tests/: only tests may import the oracle
import numpy as np, scipy.sparse as sp
import too_many_cells_b200 as T
import tmc_oracle as O

def check(name, a, **kw):
    a = sp.csr_matrix(a, dtype=np.float64)
    try:
        t = T.hierarchical_spectral_cluster(a, **kw)
    except T.TmcError as e:
        print(f"{name:40s} -> error {e.code}: {e.msg[:60]}"); return
    ok = np.array_equal(np.sort(t.perm), np.arange(a.shape[0]))
    extra = ""
    if a.shape[0] <= 400:
        try:
            ot = O.hierarchical_spectral_cluster(a, **{k: v for k, v in kw.items() if k in ("normalize", "min_modularity", "min_size")})
            extra = f" oracle nodes {ot.n_nodes} same leaves {t.leaf_sets() == ot.leaf_sets()}"
        except Exception as e:
            extra = f" oracle raised {type(e).__name__}"
    print(f"{name:40s} -> nodes {t.n_nodes:5d} leaves {len(t.leaves()):5d} perm ok {ok}{extra}", flush=True)

rng = np.random.default_rng(0)
base = sp.random(3, 50, 0.5, random_state=1, data_rvs=lambda k: rng.integers(1, 5, k).astype(float)).toarray() + np.eye(3, 50)
check("200 copies of 3 distinct rows", np.repeat(base, 67, axis=0))
check("all rows identical", np.tile(base[0], (120, 1)))
check("single column", np.ones((50, 1)))
check("identity (orthogonal rows)", sp.identity(64))
check("one nonzero per row, 4 columns", sp.csr_matrix((np.ones(100), (np.arange(100), np.arange(100) % 4)), shape=(100, 4)))
check("two disconnected blocks", sp.block_diag([np.ones((30, 5)) + np.eye(30, 5), np.ones((40, 6)) + np.eye(40, 6)]))
check("very wide, sparse", sp.random(300, 1000000, 2e-5, random_state=2, data_rvs=lambda k: np.ones(k)) + sp.csr_matrix((np.ones(300), (np.arange(300), np.arange(300))), shape=(300, 1000000)))
for m in (2, 3, 4, 5):
    for s in range(3):
        a = sp.random(m, 8, 0.6, random_state=10 * m + s, data_rvs=lambda k: rng.integers(1, 9, k).astype(float)).toarray() + np.eye(m, 8)
        check(f"random m={m} seed={s}", a, min_modularity=-1.0)
check("huge values", sp.random(200, 300, 0.1, random_state=3, data_rvs=lambda k: rng.uniform(1e150, 1e152, k)) + sp.identity(200, format="csr").dot(sp.eye(200, 300)) * 1e150)
check("tiny values", sp.random(200, 300, 0.1, random_state=4, data_rvs=lambda k: rng.uniform(1e-150, 1e-148, k)) + sp.eye(200, 300) * 1e-150)
check("normalize=True with a gene in every cell (idf=0 column)", np.hstack([np.ones((80, 1)), (rng.random((80, 30)) < 0.3).astype(float) + np.eye(80, 30)]), normalize=True)
check("normalize=True making a zero row", np.vstack([np.hstack([np.ones((20, 1)), np.zeros((20, 5))]), np.hstack([np.ones((1, 1)), np.ones((1, 5))])]), normalize=True)
check("counts > 65535 (falls back to fp64 storage)", sp.random(150, 200, 0.2, random_state=5, data_rvs=lambda k: rng.integers(1, 200000, k).astype(float)) + sp.eye(150, 200), normalize=True)
print("done")
