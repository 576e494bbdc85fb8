"""the non-default variants of the engine call — KMeansGroup / --num-eigen, the --numRuns
permutation test and --dense (Options.hs:172-174, Cluster.hs:157-167) — through the C ABI against the oracle's restatement
of the same conventions (oracle/tmc_oracle.py: kmeans2_labels, permutation_p_value).  Upstream bodies are un-vendored, so
like the rest of the numeric path this is parity with the restated algorithm ("parity unpinned")."""
import numpy as np
import scipy.sparse as sp
import pytest

import tmc_oracle as O
import too_many_cells_b200 as T
from too_many_cells_b200 import _lib

pytestmark = pytest.mark.gpu


def same_tree(tree, ot):
    ca = O.canonical_tree(tree.left, tree.right, tree.items)
    cb = O.canonical_tree(ot.left, ot.right, lambda i: ot.items[i])
    return ca == cb


def by_items(tree_items, n_nodes, values):
    return {frozenset(int(x) for x in tree_items(i)): values[i] for i in range(n_nodes)}


def assert_q_match(tree, ot, tol=1e-9):
    qa = by_items(tree.items, tree.n_nodes, tree.q)
    qb = by_items(lambda i: ot.items[i], ot.n_nodes, ot.q)
    assert qa.keys() == qb.keys()
    for k, v in qb.items():
        if len(k) > 1:
            assert abs(qa[k] - v) <= tol * max(1.0, abs(v)), (len(k), qa[k], v)


def test_dense_call_builds_the_sparse_tree(golden_tiny):
    a, _ = golden_tiny
    a = a[:200]
    ts = T.hierarchical_spectral_cluster(a, normalize=True)
    td = T.hierarchical_spectral_cluster(np.asarray(a.todense()), normalize=True)        # tmc_make_tree_dense
    assert np.array_equal(ts.left, td.left) and np.array_equal(ts.perm, td.perm)
    assert np.allclose(ts.q, td.q, rtol=0, atol=1e-12)       # two runs of the engine agree to rounding, not bit for bit
    assert same_tree(td, O.hierarchical_spectral_cluster(a, normalize=True))
    cl, th = T.h_spec_clust(T.tfidf_scale_sparse_mat(a), [f"c{i}" for i in range(a.shape[0])], dense=True)   # hSpecCommand True
    assert th.leaf_sets() == ts.leaf_sets() and len(cl) == a.shape[0]
    assert T.hierarchical_spectral_cluster(np.array([[1.0, -2.0], [1.0, 1.0]])).n_nodes >= 1      # signed dense input is accepted


@pytest.mark.parametrize("num_eigen", [1, 3])
def test_kmeans_group_vs_oracle(golden_tiny, num_eigen):
    a, _ = golden_tiny
    t = T.hierarchical_spectral_cluster(a, eigen_group="KMeansGroup", num_eigen=num_eigen, normalize=True)
    ot = O.hierarchical_spectral_cluster(a, normalize=True, eigen_group="KMeansGroup", num_eigen=num_eigen)
    assert t.n_nodes == ot.n_nodes and t.leaf_sets() == ot.leaf_sets()
    assert same_tree(t, ot)
    assert_q_match(t, ot)
    for i in t.leaves():                                      # stable splits: leaves ascend in original row id
        assert np.all(np.diff(t.items(i)) > 0)
    assert np.all(np.isnan(t.p_val))                          # no test asked for: every _pVal is Nothing


def test_kmeans_group_on_device_matrix_and_flags(golden_tiny):
    a, _ = golden_tiny
    a = a[:150]
    db = T.DeviceB.from_host(a, normalize=True)
    t = T.hierarchical_spectral_cluster(db, eigen_group="KMeansGroup")
    ot = O.hierarchical_spectral_cluster(a, normalize=True, eigen_group="KMeansGroup")
    assert same_tree(t, ot)
    t5 = T.hierarchical_spectral_cluster(db, eigen_group="KMeansGroup", min_size=5, min_modularity=0.01)
    o5 = O.hierarchical_spectral_cluster(a, normalize=True, eigen_group="KMeansGroup", min_size=5, min_modularity=0.01)
    assert same_tree(t5, o5)
    with pytest.raises(T.TmcError) as ei:                     # several eigenvectors need the host matrix (C_S is assembled per node)
        T.hierarchical_spectral_cluster(db, eigen_group="KMeansGroup", num_eigen=2)
    assert ei.value.code == _lib.ERR_UNSUPPORTED
    db.free()
    two = sp.csr_matrix(np.array([[1.0, 2.0, 0.0], [0.0, 1.0, 1.0]]))
    t = T.hierarchical_spectral_cluster(two, eigen_group="KMeansGroup")
    assert t.n_nodes == 1 and abs(t.q[0] + 0.5) < 1e-12
    one = sp.csr_matrix(np.array([[1.0, 2.0, 0.0]]))
    assert T.hierarchical_spectral_cluster(one, eigen_group="KMeansGroup", num_runs=3).n_nodes == 1


def check_p_values(t, ot):
    pa = by_items(t.items, t.n_nodes, t.p_val)
    pb = by_items(lambda i: ot.items[i], ot.n_nodes, ot.p_val)
    assert pa.keys() == pb.keys()
    n_inner = 0
    for k, v in pb.items():
        if v is None:
            assert np.isnan(pa[k])
        else:
            n_inner += 1
            assert pa[k] == v, (len(k), pa[k], v)
    return n_inner


def test_permutation_test_vs_oracle(golden_tiny):
    a, _ = golden_tiny
    t = T.hierarchical_spectral_cluster(a, normalize=True, num_runs=20)
    ot = O.hierarchical_spectral_cluster(a, normalize=True, num_runs=20)
    assert same_tree(t, ot)
    assert check_p_values(t, ot) == (t.n_nodes - 1) // 2
    # a saturated tree splits noise too: p-values away from 0, other seed
    s = a[:60]
    ts = T.hierarchical_spectral_cluster(s, normalize=True, num_runs=10, min_modularity=-1.0, seed=7)
    os_ = O.hierarchical_spectral_cluster(s, normalize=True, num_runs=10, min_modularity=-1.0, seed=7)
    assert same_tree(ts, os_) and len(ts.leaves()) == 60
    check_p_values(ts, os_)
    assert np.nanmax(ts.p_val) > 0.0
    # the export on a finished tree gives the same numbers as the in-call test
    db = T.DeviceB.from_host(a, normalize=True)
    t0 = T.hierarchical_spectral_cluster(db)
    assert np.all(np.isnan(t0.p_val))
    p = T.permutation_test(db, t0, 20)
    assert np.array_equal(np.isnan(p), np.isnan(t.p_val)) and np.array_equal(p[~np.isnan(p)], t.p_val[~np.isnan(t.p_val)])
    db.free()


def test_kmeans_group_in_the_sweep_engine_equals_the_per_node_composition(golden_small):
    """KMeansGroup on one eigenvector runs inside the batched sweeps (ST_KM: one Lloyd round per sweep); TMC_KMEANS_COMPOSED=1 is
    the per-node host recursion it replaced.  Same tree and Q, also from device-resident input and with a permutation test."""
    import os
    import time
    a, _ = golden_small
    db = T.DeviceB.from_host(a, normalize=True)
    tb = time.perf_counter(); t_sweeps = T.hierarchical_spectral_cluster(db, eigen_group="KMeansGroup", num_runs=4); tb = time.perf_counter() - tb
    os.environ["TMC_KMEANS_COMPOSED"] = "1"
    try:
        tc = time.perf_counter(); t_comp = T.hierarchical_spectral_cluster(db, eigen_group="KMeansGroup", num_runs=4); tc = time.perf_counter() - tc
    finally:
        del os.environ["TMC_KMEANS_COMPOSED"]
    db.free()
    assert t_sweeps.tree_hash_strict() == t_comp.tree_hash_strict() and t_sweeps.n_nodes > 10
    qa = {tuple(sorted(int(x) for x in t_sweeps.items(i))): (t_sweeps.q[i], t_sweeps.p_val[i]) for i in range(t_sweeps.n_nodes)}
    for i in range(t_comp.n_nodes):
        k = tuple(sorted(int(x) for x in t_comp.items(i)))
        if len(k) >= 2:
            assert abs(qa[k][0] - t_comp.q[i]) <= 1e-9
            assert (np.isnan(qa[k][1]) and np.isnan(t_comp.p_val[i])) or qa[k][1] == t_comp.p_val[i]
    ot = O.hierarchical_spectral_cluster(a, normalize=True, eigen_group="KMeansGroup", num_eigen=1)
    assert same_tree(t_sweeps, ot)
    print(f"KMeansGroup, {t_sweeps.n_nodes} nodes: in the sweeps {1e3 * tb:.0f} ms, per node {1e3 * tc:.0f} ms")


def test_batched_permutation_test_equals_the_per_node_version(golden_small):
    """The permutation test runs batched through the sweep engine (one slot per split of a tree depth, one modularity sweep per
    run); TMC_PERM_COMPOSED=1 is the per-node composition it replaced: identical p-values, also on a signed matrix."""
    import os
    import time
    a, _ = golden_small
    for mat, kw in ((a, dict(normalize=True)), (_signed_matrix(seed=8), dict())):
        db = T.DeviceB.from_host(mat, **kw)
        t0 = T.hierarchical_spectral_cluster(db)
        tb = time.perf_counter(); p_batched = T.permutation_test(db, t0, 12, seed=3); tb = time.perf_counter() - tb
        os.environ["TMC_PERM_COMPOSED"] = "1"
        try:
            tc = time.perf_counter(); p_composed = T.permutation_test(db, t0, 12, seed=3); tc = time.perf_counter() - tc
        finally:
            del os.environ["TMC_PERM_COMPOSED"]
        db.free()
        assert np.array_equal(np.isnan(p_batched), np.isnan(p_composed))
        assert np.array_equal(p_batched[~np.isnan(p_batched)], p_composed[~np.isnan(p_composed)])
        assert np.array_equal(np.isnan(p_batched), t0.left < 0)
        print(f"permutation test, {t0.n_nodes} nodes x 12 runs: batched {1e3 * tb:.0f} ms, per node {1e3 * tc:.0f} ms")


def test_p_values_reach_the_output_files(golden_tiny):
    a, _ = golden_tiny
    a = a[:120]
    names = [f"CELL{i:04d}-1" for i in range(a.shape[0])]
    t = T.hierarchical_spectral_cluster(a, normalize=True, num_runs=5)
    lines = T.tree_io.node_info_csv(t).splitlines()
    assert lines[0] == "node,size,proportion,modularity,significance,composition,subtree"
    for i, ln in enumerate(lines[1:]):
        sig = ln.split(",")[4]
        assert (sig == "") == bool(t.left[i] < 0)
        if sig:
            assert sig == T.tree_io.haskell_show_double(float(t.p_val[i]))
    back, _ = T.tree_io.parse_cluster_tree_json(T.tree_io.cluster_tree_json(t, names, significance=True))
    inner = t.left >= 0
    assert np.array_equal(back.p_val[inner], t.p_val[inner]) and np.all(np.isnan(back.p_val[~inner]))


# ----------------------------------------------------------------------------- signed matrices, --svd / --pca
def _signed_matrix(seed=3, n=260, g=900):
    """A tf-idf matrix with planted structure shifted so that ~40 % of its stored entries are negative (sparsity kept): what a
    centred / batch-corrected input looks like.  The oracle builds a 60-70 node tree on it (smallest gap ~4e-3)."""
    import scipy.sparse as sp
    from too_many_cells_b200 import synth
    ip, ix, dv, _ = synth.synth_counts(n, g, 90, 3, seed)
    b2 = O.b1_to_b2(sp.csr_matrix((dv, ix, ip), shape=(n, g)))
    sg = b2.copy()
    sg.data = sg.data - 0.9 * np.median(sg.data)
    sg.eliminate_zeros()
    return sg


def test_signed_matrix_per_node_exports_vs_oracle():
    """Negative entries: d = |B| (|B|^T 1) (bToD), C keeps the signs, nothing is deflated and u2 is the second singular vector."""
    import scipy.sparse as sp
    a = _signed_matrix()
    assert a.data.min() < 0
    b = O.b2_to_b(a)
    db = T.DeviceB.from_host(a)
    assert db.info()["value_type"] == "f64"
    d = T.b_to_d(db)
    assert np.max(np.abs(d - O.b_to_d(b))) <= 1e-12 * np.max(np.abs(d))
    lab, uo, tho, gap, _ = O.spectral_cluster(b)
    u2, th, it = T.second_left(db)
    assert abs(th - tho) <= 1e-10
    sgn = 1.0 if u2 @ uo >= 0 else -1.0
    assert np.linalg.norm(sgn * u2 - uo) <= 1e-7 / max(gap, 1e-3)
    glab = (u2 >= 0).astype(np.uint8)
    assert abs(T.get_b_modularity(glab, db) - O.get_b_modularity(glab, b)) <= 1e-12
    rows = np.arange(3, a.shape[0], 4)
    assert np.max(np.abs(T.b_to_d(db, rows) - O.b_to_d(b[rows]))) <= 1e-12 * np.max(np.abs(d))
    db.free()


def test_signed_matrix_full_tree_vs_oracle():
    a = _signed_matrix(seed=8)
    tree = T.hierarchical_spectral_cluster(a)
    ot = O.hierarchical_spectral_cluster(a)
    assert ot.n_nodes > 20
    assert tree.n_nodes == ot.n_nodes and tree.leaf_sets() == ot.leaf_sets()
    qo = {tuple(sorted(int(x) for x in ot.items[i])): ot.q[i] for i in range(ot.n_nodes)}
    for i in range(tree.n_nodes):
        key = tuple(sorted(int(x) for x in tree.items(i)))
        if len(key) >= 2:
            assert abs(tree.q[i] - qo[key]) <= 1e-9
    # the dense call (typical after PCA / batch correction) builds the same tree
    td = T.hierarchical_spectral_cluster(np.asarray(a.todense()))
    assert td.tree_hash_strict() == tree.tree_hash_strict()


def test_svd_and_pca_pre_reductions_vs_lapack():
    """`--svd N` / `--pca N` (Preprocess.hs:543-556, 573-585) through tmc_left_singular on the standardised (signed) matrix."""
    import scipy.sparse as sp
    from too_many_cells_b200 import synth
    ip, ix, dv, _ = synth.synth_counts(220, 700, 80, 3, 17)
    a = sp.csr_matrix((dv, ix, ip), shape=(220, 700))
    u, sig, conv = T.svd_sparse_mat(a, 6)
    uo, so = O.svd_sparse_mat(a, 6)
    assert conv and np.allclose(sig, so, rtol=1e-9)
    for i in range(6):
        assert abs(abs(u[:, i] @ uo[:, i]) - 1.0) < 1e-7
    sc, sig, conv = T.pca_sparse_mat(a, 5)
    so_, sgo = O.pca_sparse_mat(a, 5)
    assert conv and np.allclose(sig, sgo, rtol=1e-9)
    for i in range(5):      # scores are defined up to the sign of the principal direction
        s = 1.0 if sc[:, i] @ so_[:, i] >= 0 else -1.0
        assert np.linalg.norm(s * sc[:, i] - so_[:, i]) <= 1e-6 * np.linalg.norm(so_[:, i])
