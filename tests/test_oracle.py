"""The oracle against its committed golden vectors + the algebraic properties the reference's
algorithm must satisfy (the reference ships no tests: SURVEY.md section 4 / 8(c))."""
import numpy as np
import scipy.sparse as sp
import pytest

import tmc_oracle as O


def test_oracle_matches_golden_tiny(golden_tiny):
    a, z = golden_tiny
    b2 = O.b1_to_b2(a)
    assert np.allclose(b2.data[:4096], z["tfidf_head"], rtol=1e-14, atol=0)
    tr = O.hierarchical_spectral_cluster(b2)
    assert np.array_equal(np.array(tr.left), z["left"]) and np.array_equal(np.array(tr.right), z["right"])
    assert np.allclose(np.array(tr.q), z["q"], rtol=1e-9, atol=1e-13)
    leaf_of_row = np.full(a.shape[0], -1)
    for i in tr.leaves():
        leaf_of_row[tr.items[i]] = i
    assert np.array_equal(leaf_of_row, z["leaf_of_row"])


def test_tfidf_definition():
    a = sp.csr_matrix(np.array([[1.0, 0, 2], [0, 3, 4], [5, 0, 6], [0, 0, 7]]))
    b2 = O.b1_to_b2(a).toarray()
    idf = np.log(4 / np.array([2, 1, 4]))
    assert np.allclose(b2, a.toarray() * idf)


def test_top_singular_pair_is_trivial(golden_tiny):
    a, _ = golden_tiny
    b = O.b2_to_b(O.b1_to_b2(a))[:300]
    d = O.b_to_d(b)
    c = O.bd_to_c(b, d).toarray()
    u, s, _ = np.linalg.svd(c, full_matrices=False)
    assert abs(s[0] - 1.0) < 1e-12
    u1 = np.sqrt(d) / np.sqrt(d.sum())
    assert min(np.linalg.norm(u[:, 0] - u1), np.linalg.norm(u[:, 0] + u1)) < 1e-10
    u2, th, _ = O.second_left(sp.csr_matrix(c), d)
    assert abs(th - s[1] ** 2) < 1e-12
    assert min(np.linalg.norm(u2 - u[:, 1]), np.linalg.norm(u2 + u[:, 1])) < 1e-8


def test_dense_and_arpack_agree(golden_small):
    a, _ = golden_small
    b = O.b2_to_b(O.b1_to_b2(a))[:1200]
    d = O.b_to_d(b)
    c = O.bd_to_c(b, d)
    u_d, th_d, _ = O.second_left(c, d)
    old = O.DENSE_MAX_ROWS
    O.DENSE_MAX_ROWS = 10
    try:
        u_a, th_a, _ = O.second_left(c, d)
    finally:
        O.DENSE_MAX_ROWS = old
    assert abs(th_d - th_a) < 1e-12
    assert min(np.linalg.norm(u_d - u_a), np.linalg.norm(u_d + u_a)) < 1e-9


def test_modularity_properties(golden_tiny):
    a, _ = golden_tiny
    b = O.b2_to_b(O.b1_to_b2(a))[:200]
    rng = np.random.default_rng(0)
    lab = rng.integers(0, 2, 200)
    q = O.get_b_modularity(lab, b)
    assert abs(q - O.get_b_modularity(1 - lab, b)) < 1e-14            # label swap
    p = rng.permutation(200)
    assert abs(q - O.get_b_modularity(lab[p], b[p])) < 1e-13          # row permutation
    assert q <= 0.5
    # direct dense definition on A = B B^T - I
    g = (b @ b.T).toarray() - np.eye(200)
    k = g.sum(1); L = g.sum()
    qd = sum(g[np.ix_(lab == c, lab == c)].sum() / L - (k[lab == c].sum() / L) ** 2 for c in (0, 1))
    assert abs(q - qd) < 1e-12


def test_block_diagonal_split_positive_q():
    rng = np.random.default_rng(1)
    a = sp.block_diag([sp.random(40, 30, 0.4, random_state=1, data_rvs=lambda k: rng.integers(1, 5, k).astype(float)) + sp.eye(40, 30),
                       sp.random(50, 35, 0.4, random_state=2, data_rvs=lambda k: rng.integers(1, 5, k).astype(float)) + sp.eye(50, 35)]).tocsr()
    b = O.b2_to_b(a)
    lab, *_ = O.spectral_cluster(b)
    truth = np.r_[np.zeros(40), np.ones(50)]
    assert np.array_equal(lab, truth) or np.array_equal(lab, 1 - truth)
    assert O.get_b_modularity(lab, b) > 0.4


def test_tree_invariant_under_feature_permutation(golden_tiny):
    a, _ = golden_tiny
    sub = a[:250]
    t1 = O.hierarchical_spectral_cluster(sub, normalize=True)
    p = np.random.default_rng(2).permutation(sub.shape[1])
    t2 = O.hierarchical_spectral_cluster(sub[:, p], normalize=True)
    assert t1.leaf_sets() == t2.leaf_sets()


def test_zero_row_is_an_error():
    a = sp.csr_matrix(np.array([[1.0, 2.0], [0.0, 0.0], [3.0, 1.0]]))
    with pytest.raises(ValueError):
        O.hierarchical_spectral_cluster(a)


def test_c_port_agrees_with_python_oracle(golden_tiny, golden_small):
    """Two independent restatements (numpy/LAPACK exact solve vs the C Lanczos port that bench.py times) give the same trees."""
    import tmc_port as P
    for (a, z), tol in ((golden_tiny, 1e-10), (golden_small, 1e-6)):
        pt = P.make_tree(a, normalize=True, tol=tol)
        gold = {}
        for r, l in enumerate(z["leaf_of_row"]):
            gold.setdefault(int(l), []).append(r)
        assert pt.leaf_sets() == sorted(tuple(v) for v in gold.values())
        assert pt.n_nodes == z["left"].size
        assert abs(pt.q[0] - z["q"][0]) < 1e-7


# ---- SURVEY 8(f) N4 conventions (KMeansGroup, permutation test): properties the restatement must have
def test_kmeans2_properties():
    rng = np.random.default_rng(3)
    two = np.concatenate([rng.normal(-3, 0.3, (40, 2)), rng.normal(3, 0.3, (25, 2))])
    lab = O.kmeans2_labels(two)
    assert np.array_equal(lab, np.r_[np.zeros(40), np.ones(25)].astype(np.int8))      # label 1 = the cluster at the high end
    assert np.array_equal(O.kmeans2_labels(-two), 1 - lab)                            # sign flip of the vectors swaps the labels ...
    flipped = two * np.array([1.0, -1.0])
    assert np.array_equal(O.kmeans2_labels(flipped), lab)                             # ... flipping a later column changes nothing
    perm = rng.permutation(65)
    assert np.array_equal(O.kmeans2_labels(two[perm]), lab[perm])                     # order of the cells does not matter
    assert np.array_equal(O.kmeans2_labels(7.5 * two), lab)                           # nor does the scale
    assert O.kmeans2_labels(np.array([[0.2], [0.9]])).tolist() == [0, 1]
    assert O.kmeans2_labels(np.ones((4, 1))).tolist() == [1, 1, 1, 1]                  # no spread: a single cluster
    assert O.kmeans_num_eigen(5, 100, 1000) == 5 and O.kmeans_num_eigen(5, 12, 1000) == 2 and O.kmeans_num_eigen(5, 3, 1000) == 1


def test_permutation_p_value_properties(golden_tiny):
    a, _ = golden_tiny
    b = O.b2_to_b(O.b1_to_b2(a))[:80]
    rows = np.arange(80)
    labels, *_ = O.spectral_cluster(b)
    p = O.permutation_p_value(labels, b, rows, 30, seed=11)
    assert 0.0 <= p <= 1.0 and p == O.permutation_p_value(labels, b, rows, 30, seed=11)       # deterministic given the seed
    assert p == O.permutation_p_value(1 - labels, b, rows, 30, seed=11)                           # a split, not a labelling, is tested
    assert p <= 0.1                                                                               # the spectral split beats shuffles
    rnd = (np.random.default_rng(0).random(80) < 0.5).astype(np.int8)
    assert O.permutation_p_value(rnd, b, rows, 60, seed=11) > 0.1                                 # a random split does not
    # every shuffle of a 2-cell node is the same split: p = 1 exactly (the 1e-12 slack absorbs rounding)
    two = b[:2]
    assert O.permutation_p_value(np.array([0, 1]), two, np.array([0, 1]), 10, seed=5) == 1.0
    # shuffling keeps the cluster sizes
    lab = np.array([0] * 5 + [1] * 9, dtype=np.int8)
    r = O.SplitMix64(1)
    for _ in range(20):
        O.shuffle_labels(lab, r)
        assert lab.sum() == 9
    # different nodes draw different streams, the same node the same stream wherever it sits in a tree
    t = O.hierarchical_spectral_cluster(a[:120], normalize=True, num_runs=5, seed=3)
    sub = O.hierarchical_spectral_cluster(a[:120], normalize=True, num_runs=5, seed=3, min_modularity=0.0)
    assert t.p_val == sub.p_val


def test_oracle_variants_match_golden(golden_tiny):
    """tests/golden/variants_tiny.npz (tests/golden/make_golden.py) pins the restated conventions of the non-default variants:
    the GPU tests compare libtmc with the live oracle, this keeps the oracle itself from drifting."""
    import os
    from conftest import GOLDEN
    a, _ = golden_tiny
    z = np.load(os.path.join(GOLDEN, "variants_tiny.npz"))
    assert int(z["splitmix_1234567"][0]) == 6457827717110365317
    for tag, kw in (("km1", dict(eigen_group="KMeansGroup", num_eigen=1)), ("km3", dict(eigen_group="KMeansGroup", num_eigen=3)),
                    ("perm", dict(num_runs=20))):
        tr = O.hierarchical_spectral_cluster(a, normalize=True, **kw)
        assert np.array_equal(np.array(tr.left), z[f"{tag}_left"]) and np.array_equal(np.array(tr.right), z[f"{tag}_right"])
        assert np.allclose(np.array(tr.q), z[f"{tag}_q"], rtol=1e-9, atol=1e-13)
        leaf_of_row = np.full(a.shape[0], -1)
        for i in tr.leaves():
            leaf_of_row[tr.items[i]] = i
        assert np.array_equal(leaf_of_row, z[f"{tag}_leaf_of_row"])
        p = np.array([np.nan if x is None else x for x in tr.p_val])
        assert np.array_equal(p, z[f"{tag}_p_val"], equal_nan=True)
    sat = O.hierarchical_spectral_cluster(a[:60], normalize=True, num_runs=10, min_modularity=-1.0, seed=7)
    assert np.array_equal(np.array(sat.left), z["sat_left"])
    assert np.array_equal(np.array([np.nan if x is None else x for x in sat.p_val]), z["sat_p_val"], equal_nan=True)


def test_port_and_oracle_agree_on_config1_through_the_tree_signature():
    """BASELINE.json configs[0] at full size on the CPU: the C port (Lanczos, 1e-10) and the exact-math oracle build the same
    tree; `tree_signature` (the hash bench.py prints and compares across rank counts) sees them as equal and tells a
    different tree apart."""
    import scipy.sparse as sp
    import tmc_port as P
    import too_many_cells_b200 as T
    from too_many_cells_b200 import synth
    ip, ix, dv, _ = synth.synth_config("c1")
    n, g = synth.CONFIGS["c1"][:2]
    a = sp.csr_matrix((dv, ix, ip), shape=(n, g))
    pt = P.make_tree(a, normalize=True, tol=1e-10)
    ot = O.hierarchical_spectral_cluster(a, normalize=True)
    assert pt.leaf_sets() == ot.leaf_sets()
    begin, end, perm = [0] * ot.n_nodes, [0] * ot.n_nodes, []
    def rec(i):
        begin[i] = len(perm)
        if ot.left[i] < 0:
            perm.extend(int(x) for x in ot.items[i])
        else:
            rec(ot.left[i]); rec(ot.right[i])
        end[i] = len(perm)
    rec(0)
    sig_o = T.tree_signature_strict(ot.left, ot.right, begin, end, perm)
    sig_p = T.tree_signature_strict(pt.left, pt.right, pt.begin, pt.end, pt.perm)
    assert sig_o == sig_p
    # swapping the children of the root leaves the signature alone; moving one cell to another leaf changes it
    l2, r2 = list(ot.left), list(ot.right)
    l2[0], r2[0] = r2[0], l2[0]
    b2, e2, p2 = [0] * ot.n_nodes, [0] * ot.n_nodes, []
    def rec2(i):
        b2[i] = len(p2)
        if l2[i] < 0:
            p2.extend(int(x) for x in ot.items[i])
        else:
            rec2(l2[i]); rec2(r2[i])
        e2[i] = len(p2)
    rec2(0)
    assert T.tree_signature_strict(l2, r2, b2, e2, p2) == sig_o
    p3 = list(perm); p3[0], p3[-1] = p3[-1], p3[0]
    assert T.tree_signature_strict(ot.left, ot.right, begin, end, p3) != sig_o


def test_signed_input_semantics_of_the_oracle():
    """Negative entries (`--svd`, `--pca`, `--dense` inputs): the degrees come from |B|, everything else from B
    itself, and u2 is the plain second left singular vector of C = D^{-1/2} B (nothing to deflate)."""
    import scipy.sparse as sp
    import too_many_cells_b200 as T
    from too_many_cells_b200 import synth
    ip, ix, dv, _ = synth.synth_counts(120, 400, 60, 2, 5)
    a = sp.csr_matrix((dv, ix, ip), shape=(120, 400))
    cs = O.center_scale_sparse(a)
    mine = T.center_scale_sparse_cell(a)
    assert abs(cs - mine).max() < 1e-13
    dense = a.toarray()
    mu, sd = dense.mean(axis=1, keepdims=True), dense.std(axis=1, ddof=1, keepdims=True)
    want = np.where(dense != 0, (dense - mu) / sd, 0.0)
    assert np.allclose(cs.toarray(), want, atol=1e-12)
    # a genuinely signed matrix: log counts centred per feature over the stored entries (sparsity kept)
    sg = a.tocsc(copy=True)
    sg.data = np.log1p(sg.data)
    for j in range(sg.shape[1]):
        seg = slice(sg.indptr[j], sg.indptr[j + 1])
        if seg.stop - seg.start > 1:
            sg.data[seg] -= sg.data[seg].mean() - 0.01
    cs = sp.csr_matrix(sg)
    assert cs.data.min() < 0 < cs.data.max()
    b = O.b2_to_b(cs)
    d = O.b_to_d(b)
    assert np.allclose(d, abs(b) @ (abs(b).T @ np.ones(b.shape[0])))
    lab, u2, th2, gap, _ = O.spectral_cluster(b)
    c = O.bd_to_c(b, d).toarray()
    uu, ss, _ = np.linalg.svd(c, full_matrices=False)
    assert abs(th2 - ss[1] ** 2) < 1e-12 and abs(abs(u2 @ uu[:, 1]) - 1.0) < 1e-9
    tr = O.hierarchical_spectral_cluster(cs)
    assert sum(len(x) for x in tr.leaf_sets()) == 120
    u, sv = O.svd_sparse_mat(a, 4)
    assert u.shape == (120, 4) and np.all(np.diff(sv) <= 0)
    sc, sv2 = O.pca_sparse_mat(a, 3)
    assert sc.shape == (120, 3)
