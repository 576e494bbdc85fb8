"""Dense panel (DESIGN.md "dense panel"): frequent columns stored as one byte per (cell, feature), the rest as a residual
CSR with renumbered columns.  The storage must be invisible: every export and the whole tree equal the plain-CSR path
(TMC_NO_PANEL=1) and the oracle, including the corner cases of the split (counts that do not fit a byte, duplicated entries,
all-ones matrices, every column in the panel)."""
import os

import numpy as np
import scipy.sparse as sp
import pytest

import tmc_oracle as O
import too_many_cells_b200 as T

pytestmark = pytest.mark.gpu


def rel(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / max(1e-300, float(np.max(np.abs(b)))))


class env:
    def __init__(self, **kw):
        self.kw = kw

    def __enter__(self):
        self.old = {k: os.environ.get(k) for k in self.kw}
        for k, v in self.kw.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = str(v)

    def __exit__(self, *a):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def same_tree(t1, t2):
    return O.canonical_tree(t1.left, t1.right, t1.items) == O.canonical_tree(t2.left, t2.right, t2.items)


def same_as_oracle(tree, ot):
    return O.canonical_tree(tree.left, tree.right, tree.items) == O.canonical_tree(ot.left, ot.right, lambda i: ot.items[i])


def dense_cols_counts(n=700, g=900, seed=5, binary=False, big=False):
    """Two planted groups; 60 features present in most cells (the panel), the rest sparse."""
    rng = np.random.default_rng(seed)
    grp = rng.integers(0, 2, n)
    lam = np.full((2, g), 0.02)
    lam[:, :60] = rng.uniform(0.3, 3.0, size=(2, 60))
    lam[0, 60:200] *= 6.0
    lam[1, 200:340] *= 6.0
    cnt = rng.poisson(lam[grp]).astype(np.float64)
    cnt[cnt.sum(axis=1) == 0, 0] = 1
    if big:        # counts that do not fit a byte, in panel columns and outside
        cnt[::7, 3] = 300.0
        cnt[::11, 5] = 65535.0
        cnt[::13, 500] = 1000.0
        cnt[::5, 7] = 16.0          # does not fit a nibble, fits a byte
        cnt[::3, 9] = 15.0
        cnt[1::3, 9] = 255.0
    if binary:
        cnt = (cnt > 0).astype(np.float64)
    return sp.csr_matrix(cnt), grp


@pytest.mark.parametrize("bits", [4, 8])
@pytest.mark.parametrize("name", ["tiny", "small"])
def test_panel_is_used_and_invisible(name, bits, golden_tiny, golden_small):
    a, _ = golden_tiny if name == "tiny" else golden_small
    with env(TMC_PANEL_BITS=bits):
        db = T.DeviceB.from_host(a, normalize=True)
    info = db.info()
    assert info["value_type"] == "u16" and info["panel_cols"] >= 8 and info["panel_bits"] == bits
    assert info["panel_row_bytes"] % 256 == 0 and info["panel_row_bytes"] * 8 // bits >= info["panel_cols"]
    assert 0 < info["residual_nnz"] < info["nnz"] == a.nnz
    with env(TMC_NO_PANEL=1):
        dp = T.DeviceB.from_host(a, normalize=True)
    assert dp.info()["panel_cols"] == 0 and dp.info()["residual_nnz"] == a.nnz
    b = O.b2_to_b(O.b1_to_b2(a))
    d0, d1 = T.b_to_d(db), T.b_to_d(dp)
    assert rel(d0, d1) < 1e-13 and rel(d0, O.b_to_d(b)) < 1e-13
    rows = np.arange(3, a.shape[0], 4)
    assert rel(T.b_to_d(db, rows), O.b_to_d(b[rows])) < 1e-13
    x = np.random.default_rng(1).standard_normal(a.shape[0])
    assert rel(T.spmv_pair(db, x), T.spmv_pair(dp, x)) < 1e-12
    u0, th0, _ = T.second_left(db)
    u1, th1, _ = T.second_left(dp)
    s = 1.0 if u0 @ u1 >= 0 else -1.0
    assert abs(th0 - th1) < 1e-12 and np.linalg.norm(s * u0 - u1) < 1e-9
    t0 = T.hierarchical_spectral_cluster(db)
    t1 = T.hierarchical_spectral_cluster(dp)
    assert same_tree(t0, t1) and float(np.max(np.abs(t0.q - t1.q))) < 1e-12
    db.free(); dp.free()


def test_default_cell_format_follows_the_counts(golden_tiny):
    a, _ = golden_tiny
    db = T.DeviceB.from_host(a, normalize=True)              # UMI-like counts: nibbles
    assert db.info()["panel_bits"] == 4
    db.free()
    big = a.copy(); big.data = big.data * 20.0                # nothing fits a nibble any more: bytes
    db = T.DeviceB.from_host(big, normalize=True)
    assert db.info()["panel_bits"] == 8 and db.info()["panel_cols"] >= 8
    db.free()


@pytest.mark.parametrize("bits", [4, 8])
@pytest.mark.parametrize("big", [False, True])
def test_panel_cell_overflow_and_oracle(big, bits):
    a, _ = dense_cols_counts(big=big)
    with env(TMC_PANEL_BITS=bits):
        db = T.DeviceB.from_host(a, normalize=True)
    info = db.info()
    assert info["panel_cols"] >= 60 and info["panel_bits"] == bits
    cap = 15 if bits == 4 else 255
    dense = np.asarray(a.todense())
    assert info["residual_nnz"] >= int((dense > cap).sum())           # at least the counts that do not fit a cell
    b = O.b2_to_b(O.b1_to_b2(a))
    assert rel(T.b_to_d(db), O.b_to_d(b)) < 1e-13
    tree = T.hierarchical_spectral_cluster(db)
    ot = O.hierarchical_spectral_cluster(a, normalize=True)
    assert same_as_oracle(tree, ot)
    db.free()


def test_panel_all_ones_matrix():
    a, _ = dense_cols_counts(binary=True, seed=9)
    db = T.DeviceB.from_host(a, normalize=True)
    info = db.info()
    assert info["value_type"] == "one" and info["panel_cols"] >= 60
    b = O.b2_to_b(O.b1_to_b2(a))
    assert rel(T.b_to_d(db), O.b_to_d(b)) < 1e-13
    assert same_as_oracle(T.hierarchical_spectral_cluster(db), O.hierarchical_spectral_cluster(a, normalize=True))
    db.free()


def test_panel_without_idf_and_lsa_mode():
    a, _ = dense_cols_counts(seed=11)
    db = T.DeviceB.from_host(a, normalize=False)         # raw counts, row L2 only: the column factor is 1
    assert db.info()["panel_cols"] >= 60
    b = O.b2_to_b(a)
    assert rel(T.b_to_d(db), O.b_to_d(b)) < 1e-13
    db.free()
    u, sig, conv = T.lsa_sparse_mat(a, 4)                 # truncated SVD of the idf-scaled matrix (no row factor)
    with env(TMC_NO_PANEL=1):
        u1, sig1, conv1 = T.lsa_sparse_mat(a, 4)
    assert conv and conv1 and rel(sig, sig1) < 1e-10
    for k in range(u.shape[1]):
        s = 1.0 if u[:, k] @ u1[:, k] >= 0 else -1.0
        assert np.linalg.norm(s * u[:, k] - u1[:, k]) < 1e-7


def test_duplicate_entries_drop_the_panel():
    a, _ = dense_cols_counts(seed=13)
    a = a.tocsr()
    # duplicate the first entry of every 5th row (same row, same column, value 1): scipy keeps duplicates in raw CSR arrays
    indptr, indices, data = [0], [], []
    for r in range(a.shape[0]):
        s, e = a.indptr[r], a.indptr[r + 1]
        cols, vals = list(a.indices[s:e]), list(a.data[s:e])
        if r % 5 == 0:
            cols.append(cols[0]); vals.append(1.0)
        indices += cols; data += vals; indptr.append(len(indices))
    dup = sp.csr_matrix((np.array(data), np.array(indices, dtype=np.int32), np.array(indptr, dtype=np.int64)), shape=a.shape)
    assert dup.has_canonical_format is False or dup.nnz > a.nnz
    db = T.DeviceB.from_host(dup, normalize=True)
    assert db.info()["panel_cols"] == 0 and db.info()["residual_nnz"] == dup.nnz
    with env(TMC_NO_PANEL=1):
        dp = T.DeviceB.from_host(dup, normalize=True)
    assert rel(T.b_to_d(db), T.b_to_d(dp)) < 1e-14
    db.free(); dp.free()


@pytest.mark.parametrize("density", ["0.0001", "0.5"])
def test_panel_density_knob_extremes(density, golden_tiny):
    a, _ = golden_tiny
    ref = T.hierarchical_spectral_cluster(a, normalize=True)
    with env(TMC_PANEL_DENSITY=density):
        db = T.DeviceB.from_host(a, normalize=True)
    info = db.info()
    if density == "0.0001":
        assert info["panel_cols"] >= 0.9 * (np.diff(a.tocsc().indptr) > 0).sum()      # (almost) every used column is in the panel
    tree = T.hierarchical_spectral_cluster(db)
    assert same_tree(tree, ref)
    db.free()


# ---------------------------------------------------------------- factor recovery from an already tf-idf-scaled matrix
@pytest.mark.parametrize("where", ["numpy", "gpu"])
def test_recover_counts_from_tfidf_matrix(where, golden_tiny):
    """The reference hands hierarchicalSpectralCluster the tf-idf matrix (Cluster.hs:147, normFlag False).  B2 = A diag(idf) is
    recognised bit for bit, so the compact storage (and the panel) applies to the reference's own call pattern too."""
    a, z = golden_tiny
    b2 = O.b1_to_b2(a) if where == "numpy" else T.tfidf_scale_sparse_mat(a)
    db = T.DeviceB.from_host(b2)
    info = db.info()
    assert info["value_type"] == "u16" and info["panel_cols"] >= 8
    with env(TMC_NO_RECOVER=1):
        d64 = T.DeviceB.from_host(b2)
    assert d64.info()["value_type"] == "f64" and d64.info()["panel_cols"] == 0
    assert rel(T.b_to_d(db), T.b_to_d(d64)) < 1e-13
    x = np.random.default_rng(2).standard_normal(a.shape[0])
    assert rel(T.spmv_pair(db, x), T.spmv_pair(d64, x)) < 1e-12
    t0, t1 = T.hierarchical_spectral_cluster(db), T.hierarchical_spectral_cluster(d64)
    assert same_tree(t0, t1) and float(np.max(np.abs(t0.q - t1.q))) < 1e-12
    gold = {}
    for r, l in enumerate(z["leaf_of_row"]):
        gold.setdefault(int(l), []).append(r)
    assert t0.leaf_sets() == sorted(tuple(v) for v in gold.values())
    cl, tree = T.h_spec_clust(b2, [f"c{i}" for i in range(a.shape[0])])          # the documented drop-in call
    assert same_tree(tree, t0)
    db.free(); d64.free()


def test_recover_handles_columns_without_a_unit_count_and_rejects_inexact_input():
    a, _ = dense_cols_counts(seed=21)
    a = a.tolil()
    a[:, 4] = 0; a[::2, 4] = 2; a[1::4, 4] = 6           # every count of this feature is even: w = 2 idf, counts 1 and 3
    a = sp.csr_matrix(a)
    b2 = O.b1_to_b2(a)
    db = T.DeviceB.from_host(b2)
    assert db.info()["value_type"] == "u16"
    b = O.b2_to_b(b2)
    assert rel(T.b_to_d(db), O.b_to_d(b)) < 1e-13
    db.free()
    a2 = a.tolil(); a2[0, 4] = 3; a2 = sp.csr_matrix(a2)  # 3 is not a multiple of the smallest count 2: not representable
    b2 = O.b1_to_b2(a2)
    db = T.DeviceB.from_host(b2)
    assert db.info()["value_type"] == "f64"
    assert rel(T.b_to_d(db), O.b_to_d(O.b2_to_b(b2))) < 1e-13
    db.free()
    noisy = b2.copy(); noisy.data = noisy.data * (1.0 + 1e-9 * np.random.default_rng(3).random(noisy.nnz))
    db = T.DeviceB.from_host(noisy)
    assert db.info()["value_type"] == "f64"
    db.free()


def test_compact_host_types_build_the_same_tree_bit_for_bit(golden_small):
    """tmc_csr.val_type / idx_type: uint16 / uint8 / float32 counts and uint16 gene ids are widened on the device; every result
    equals the double / int32 call bit for bit (the widened arrays ARE the fp64 arrays)."""
    a, _ = golden_small
    ref = T.hierarchical_spectral_cluster(a, normalize=True)
    for vdt, idt in ((np.uint16, np.int32), (np.uint16, np.uint16), (np.float32, np.uint16), (np.uint8, np.int32)):
        if a.data.max() > np.iinfo(vdt).max if np.issubdtype(vdt, np.integer) else False:
            continue
        m = (a.indptr, a.indices.astype(idt), a.data.astype(vdt), a.shape)
        t = T.hierarchical_spectral_cluster(m, normalize=True)
        assert t.tree_hash_strict() == ref.tree_hash_strict(), (vdt, idt)
        assert np.array_equal(t.perm, ref.perm)
        assert np.max(np.abs(t.q - ref.q)) <= 1e-12
        db = T.DeviceB.from_host(m, normalize=True)
        assert db.info()["value_type"] == "u16"
        db.free()
    # uint16 indices cannot address more than 65536 columns; the host-value entry points want the default types
    wide = (np.array([0, 1], dtype=np.int64), np.array([3], dtype=np.uint16), np.array([1.0]), (1, 70000))
    with pytest.raises(T.TmcError) as ei:
        T.hierarchical_spectral_cluster(wide)
    assert ei.value.code == T._lib.ERR_INVALID


def test_shared_memory_gather_matches_the_plain_gather(golden_small):
    """k_gather_smem (large nodes: the residual part of y resident in shared memory) against k_gather_u (TMC_GS=0): same products,
    same tree.  TMC_GS_ROWS is lowered so that most nodes of `small` take the shared-memory kernel, with and without a panel."""
    a, _ = golden_small
    for extra in ({}, {"TMC_NO_PANEL": "1"}):
        outs = []
        for gs in ({"TMC_GS": "0"}, {"TMC_GS": "1", "TMC_GS_ROWS": "48", "TMC_GS_CHUNK": "256"}):
            env = dict(extra, **gs)
            for k, v in env.items():
                os.environ[k] = v
            try:
                db = T.DeviceB.from_host(a, normalize=True)
                d = T.b_to_d(db)
                x = np.random.default_rng(5).standard_normal(a.shape[0])
                px = T.spmv_pair(db, x)
                tree = T.hierarchical_spectral_cluster(db)
                db.free()
            finally:
                for k in env:
                    del os.environ[k]
            outs.append((d, px, tree))
        (d0, p0, t0), (d1, p1, t1) = outs
        assert np.max(np.abs(d0 - d1)) <= 1e-12 * np.max(np.abs(d0))
        assert np.max(np.abs(p0 - p1)) <= 1e-12 * np.max(np.abs(p0))
        assert t0.tree_hash_strict() == t1.tree_hash_strict()
        assert np.max(np.abs(t0.q - t1.q)) <= 1e-12


@pytest.mark.parametrize("env", [{"TMC_GU_LPR": "16"}, {"TMC_GU_LPR": "8"}, {"TMC_TP_NARROW": "8"}, {"TMC_TP_NARROW": "4"},
                                 {"TMC_TP_NR": "4"}, {"TMC_TP_NR": "16"}, {"TMC_GP_NR": "2"}, {"TMC_GP_NR": "8"},
                                 {"TMC_GS": "1", "TMC_GS_ROWS": "64"}, {"TMC_GS": "1", "TMC_TS": "1", "TMC_GS_ROWS": "64"},
                                 {"TMC_NO_CERT": "1"}, {"TMC_XS_ROWS": "64"}])
def test_experimental_kernel_switches_build_the_same_tree(env, golden_small):
    """Every kernel variant kept behind an environment switch (the measured-and-dropped experiments of DESIGN.md section 4 / 7)
    computes the same products: same tree, same Q, same SpMV pair as the default kernels."""
    a, _ = golden_small
    x = np.random.default_rng(9).standard_normal(a.shape[0])

    def run():
        db = T.DeviceB.from_host(a, normalize=True)
        assert db.info()["panel_cols"] > 0 and db.info()["panel_bits"] == 4
        out = (T.spmv_pair(db, x), T.b_to_d(db), T.hierarchical_spectral_cluster(db))
        db.free()
        return out

    p0, d0, t0 = run()
    for k, v in env.items():
        os.environ[k] = v
    try:
        p1, d1, t1 = run()
    finally:
        for k in env:
            del os.environ[k]
    assert np.max(np.abs(p0 - p1)) <= 1e-12 * np.max(np.abs(p0))
    assert np.max(np.abs(d0 - d1)) <= 1e-12 * np.max(np.abs(d0))
    assert t0.tree_hash_strict() == t1.tree_hash_strict() and np.max(np.abs(t0.q - t1.q)) <= 1e-12
