"""CPU-side checks of the product: the C-ABI library loads and exports every symbol include/tmc.h
declares, the host/device tridiagonal helpers agree with LAPACK, struct layouts match, and the
product fails loudly (no CPU fallback) when there is no CUDA device."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT

import too_many_cells_b200 as T
from too_many_cells_b200 import _lib


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "tmc.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(tmc_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = _declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"libtmc.so does not export {n}"
    assert set(names) == set(_lib.SIGNATURES), "ctypes binding and header disagree"
    # ... and every prototype has as many parameters, of the same width class, as its ctypes signature
    hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "tmc.h")).read(), flags=re.S)
    for m in re.finditer(r"\b(tmc_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", hdr):
        name, params = m.group(1), m.group(2).strip()
        plist = [] if params in ("", "void") else [x.strip() for x in params.split(",")]
        res, args = _lib.SIGNATURES[name]
        assert len(plist) == len(args), (name, plist)
        for decl, ct in zip(plist, args):
            is_ptr = "*" in decl or "[" in decl
            ct_ptr = ct in (C.c_void_p, C.c_char_p) or hasattr(ct, "contents") or isinstance(ct, type(C.POINTER(C.c_int)))
            assert is_ptr == bool(ct_ptr), (name, decl, ct)
            if not is_ptr:
                want = {"int64_t": C.c_int64, "uint64_t": C.c_uint64, "int": C.c_int, "double": C.c_double}[decl.split()[-2] if len(decl.split()) > 1 else decl]
                assert ct is want, (name, decl, ct)


def test_opts_default_and_layout():
    o = T.make_opts()
    assert o.eigen_group == 0 and o.num_eigen == 1 and o.min_modularity == 0.0 and o.min_size == 1
    assert o.normalize == 0 and o.num_runs == 0 and o.tol == 1e-10 and o.max_lanczos == 320 and o.device == -1
    assert C.sizeof(_lib.Opts) == 96 and C.sizeof(_lib.Tree) == 152      # static_assert-ed in tmc_engine.cu


def test_engine_knobs_of_round_two():
    """tmc_opts.warm_start / small_rows (taken from the reserved words: the defaults, 0, mean "on")."""
    o = T.make_opts()
    assert o.sign_stop == 0 and o.warm_start == 0 and o.small_rows == 0
    o = T.make_opts(sign_stop=False, warm_start=False, small_rows=0)          # the plain engine
    assert o.sign_stop == -1 and o.warm_start == -1 and o.small_rows == -1
    assert T.make_opts(small_rows=16).small_rows == 16 and T.make_opts(small_rows=None).small_rows == 0


def test_tree_vertex_names_and_signature():
    """A vertex is named by (smallest row id, number of cells); the signature hashes leaf membership + the sorted names, so it is
    blind to child order and numbering but not to membership or topology."""
    #        0            perm = [4 5 | 0 2 | 1 3]
    #      /   \
    #     1     2         1 = {4,5} (leaf), 2 = {0,2,1,3}, 3 = {0,2}, 4 = {1,3}
    #          / \
    #         3   4
    left = [1, -1, 3, -1, -1]; right = [2, -1, 4, -1, -1]
    begin = [0, 0, 2, 2, 4]; end = [6, 2, 6, 4, 6]; perm = [4, 5, 0, 2, 1, 3]
    names = T.tree_vertex_names(left, right, begin, end, perm)
    assert names.tolist() == [[0, 6], [4, 2], [0, 4], [0, 2], [1, 2]]
    h = T.tree_signature(left, right, begin, end, perm)
    # the same tree with the children of the root and of vertex 2 swapped, renumbered in pre-order
    left2 = [1, 2, -1, -1, -1]; right2 = [4, 3, -1, -1, -1]
    begin2 = [0, 0, 0, 2, 4]; end2 = [6, 4, 2, 4, 6]; perm2 = [1, 3, 0, 2, 4, 5]
    assert T.tree_signature(left2, right2, begin2, end2, perm2) == h
    # another topology on the same leaves: {4,5} merged with {0,2} first
    left3 = [1, 2, -1, -1, -1]; right3 = [4, 3, -1, -1, -1]
    begin3 = [0, 0, 0, 2, 4]; end3 = [6, 4, 2, 4, 6]; perm3 = [4, 5, 0, 2, 1, 3]
    # the (smallest row, size) names cannot tell these two apart ({4,5,0,2} and {0,2,1,3} are both (0, 4)); the strict ones can
    assert T.tree_signature(left3, right3, begin3, end3, perm3) == h
    hs = T.tree_signature_strict(left, right, begin, end, perm)
    assert T.tree_signature_strict(left2, right2, begin2, end2, perm2) == hs
    assert T.tree_signature_strict(left3, right3, begin3, end3, perm3) != hs
    assert T.tree_signature_strict(left, right, begin, end, [4, 5, 0, 3, 1, 2]) != hs
    ns = T.tree_vertex_names(left, right, begin, end, perm, strict=True)
    assert ns.shape == (5, 3) and len({tuple(r) for r in ns.tolist()}) == 5
    assert ns[2, 2] == (ns[3, 2] + ns[4, 2]) and ns[0, 2] == (ns[1, 2] + ns[2, 2])      # set hashes add up (mod 2^64)
    # another membership: cells 2 and 3 exchanged between the leaves
    assert T.tree_signature(left, right, begin, end, [4, 5, 0, 3, 1, 2]) != h


def test_no_cpu_fallback_without_device():
    lib = _lib.load()
    if lib.tmc_device_count() > 0:
        pytest.skip("a CUDA device is present")
    import scipy.sparse as sp
    a = sp.random(20, 10, 0.5, format="csr", random_state=0)
    with pytest.raises(T.TmcError) as ei:
        T.hierarchical_spectral_cluster(a)
    assert ei.value.code == _lib.ERR_NO_DEVICE
    with pytest.raises(T.TmcError):
        T.b1_to_b2(a)


def test_variant_flags_fail_loudly_without_device():
    lib = _lib.load()
    if lib.tmc_device_count() > 0:
        pytest.skip("the variants run on the GPU box in test_gpu_variants")
    # KMeansGroup / dense / num_runs go through the same device check: an error, never a host fallback or a crash
    import scipy.sparse as sp
    a = sp.random(20, 10, 0.5, format="csr", random_state=0)
    import numpy as np
    for kw in (dict(eigen_group="KMeansGroup"), dict(eigen_group="KMeansGroup", num_eigen=3), dict(num_runs=5)):
        with pytest.raises(T.TmcError) as ei:
            T.hierarchical_spectral_cluster(a, **kw)
        assert ei.value.code == _lib.ERR_NO_DEVICE
    with pytest.raises(T.TmcError) as ei:
        T.hierarchical_spectral_cluster(np.asarray(a.todense()))           # tmc_make_tree_dense
    assert ei.value.code == _lib.ERR_NO_DEVICE


@pytest.fixture(scope="module")
def tridiag_lib(tmp_path_factory):
    out = tmp_path_factory.mktemp("th") / "libth.so"
    src = os.path.join(ROOT, "tests", "helpers", "tridiag_host.cpp")
    subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-x", "c++", src, "-o", str(out)], check=True)
    lib = C.CDLL(str(out))
    P = C.POINTER(C.c_double)
    lib.th_top_eig.restype = C.c_double
    lib.th_top_eig.argtypes = [P, P, C.c_int, C.c_double]
    lib.th_vector.argtypes = [P, P, C.c_int, C.c_double, P, P, P]
    return lib


def test_tridiagonal_helpers_match_lapack(tridiag_lib):
    from scipy.linalg import eigh_tridiagonal
    P = C.POINTER(C.c_double)
    rng = np.random.default_rng(0)
    for trial in range(120):
        j = int(rng.integers(1, 330))
        if trial % 3 == 0:
            a = rng.random(j); b = rng.random(j) * 0.3
        elif trial % 3 == 1:
            a = np.sort(rng.random(j))[::-1].copy(); b = rng.random(j) * 10.0 ** (-rng.integers(0, 12, size=j))
        else:
            a = np.full(j, 0.5); b = np.full(j, 0.25)
        b[0] = 0
        w = eigh_tridiagonal(a, b[1:], eigvals_only=True) if j > 1 else a
        th = tridiag_lib.th_top_eig(a.ctypes.data_as(P), b.ctypes.data_as(P), j, -1e300)
        assert abs(th - w[-1]) <= 3e-14 * max(1.0, abs(w[-1]))
        z, dp, dm = np.zeros(j), np.zeros(j), np.zeros(j)
        tridiag_lib.th_vector(a.ctypes.data_as(P), b.ctypes.data_as(P), j, th, dp.ctypes.data_as(P), dm.ctypes.data_as(P),
                              z.ctypes.data_as(P))
        tm = np.diag(a) + np.diag(b[1:], 1) + np.diag(b[1:], -1)
        assert np.linalg.norm(tm @ z - th * z) < 1e-13 and abs(np.linalg.norm(z) - 1) < 1e-12


def test_synth_is_deterministic(golden_tiny):
    from too_many_cells_b200 import synth
    a, z = golden_tiny
    ip, ix, dv, _ = synth.synth_config("tiny")
    assert np.array_equal(ip, z["indptr"]) and np.array_equal(ix, z["indices"]) and np.array_equal(dv, z["counts"])


def _build_cpp_example(tmp_path):
    exe = tmp_path / "host_cpp_example"
    pkg = os.path.join(ROOT, "too-many-cells_b200")
    subprocess.run(["g++", "-std=c++17", "-O1", os.path.join(ROOT, "tests", "helpers", "host_cpp_example.cpp"), "-o", str(exe),
                    "-L", pkg, "-ltmc", f"-Wl,-rpath,{pkg}"], check=True)
    return exe


def _write_csr_text(a, path):
    with open(path, "w") as f:
        f.write(f"{a.shape[0]} {a.shape[1]} {a.nnz}\n")
        f.write(" ".join(map(str, a.indptr.tolist())) + "\n" + " ".join(map(str, a.indices.tolist())) + "\n")
        f.write(" ".join(repr(float(x)) for x in a.data.tolist()) + "\n")


def test_cpp_host_layer_compiles_and_fails_loudly_without_device(tmp_path, golden_tiny):
    """include/tmc.hpp (the compiled-host mirror of the reference's call surface) builds against libtmc.so; on a box without a
    GPU the call surfaces TMC_ERR_NO_DEVICE instead of falling back to anything."""
    a, _ = golden_tiny
    exe = _build_cpp_example(tmp_path)
    _write_csr_text(a[:60], tmp_path / "m.txt")
    out = subprocess.run([str(exe), str(tmp_path / "m.txt")], capture_output=True, text=True, check=True).stdout
    if _lib.load().tmc_device_count() == 0:
        assert out.strip() == "no-device"
    else:
        assert out.startswith("cell,cluster,path\n")


@pytest.mark.gpu
def test_cpp_host_layer_matches_python_mirror(tmp_path, golden_tiny):
    a, _ = golden_tiny
    sub = a[:300]
    exe = _build_cpp_example(tmp_path)
    _write_csr_text(sub, tmp_path / "m.txt")
    out = subprocess.run([str(exe), str(tmp_path / "m.txt")], capture_output=True, text=True, check=True).stdout
    names = [f"CELL{i}" for i in range(300)]
    cl, tree = T.h_spec_clust(T.tfidf_scale_sparse_mat(sub), names)
    want = T.tree_io.clusters_csv(tree, names).replace("\r\n", "\n").rstrip("\n") + "\n"
    # eigenvector sign may differ between two runs only through the (fixed) seed: same seed => same child order
    assert out == want


def test_panel_cells_as_denormal_doubles_are_exact(tmp_path):
    """The dense-panel kernels feed a stored count to the FMA as the denormal double (hi = 0, lo = count [<< 4k]) against an
    operand pre-scaled by 2^960: tests/helpers/denormal_cells.c replays that arithmetic on the host (same IEEE fma) and counts
    the sums that differ from converting the count first -- there must be none."""
    exe = tmp_path / "denormal_cells"
    subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-o", str(exe), os.path.join(ROOT, "tests", "helpers", "denormal_cells.c"), "-lm"],
                   check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.strip() == "0"


# ---------------------------------------------------------------------------------------------------------------------------
# host logic of the non-default variants (csrc/tmc_variants.cuh), compiled against oracle-backed stand-ins of
# the six exports it composes (tests/helpers/variants_host.cpp).  The same trees / p-values come out of the GPU library in
# tests/test_gpu_variants.py.
@pytest.fixture(scope="module")
def variants_host(tmp_path_factory):
    import scipy.sparse as sp
    import tmc_oracle as O
    out = tmp_path_factory.mktemp("vh") / "libvh.so"
    subprocess.run(["g++", "-std=c++17", "-O1", "-shared", "-fPIC", os.path.join(ROOT, "tests", "helpers", "variants_host.cpp"),
                    "-o", str(out)], check=True)
    lib = C.CDLL(str(out))
    state = {}

    def csr_of(ap):
        a = ap.contents
        n, nnz = int(a.n_rows), int(a.nnz)
        ip = np.ctypeslib.as_array(C.cast(a.row_ptr, C.POINTER(C.c_int64)), shape=(n + 1,)).copy()
        ix = np.ctypeslib.as_array(C.cast(a.col_idx, C.POINTER(C.c_int32)), shape=(max(nnz, 1),))[:nnz].copy()
        v = np.ctypeslib.as_array(C.cast(a.val, C.POINTER(C.c_double)), shape=(max(nnz, 1),))[:nnz].copy()
        return sp.csr_matrix((v, ix, ip), shape=(n, int(a.n_cols)))

    def rows_of(p, n):
        return np.ctypeslib.as_array(p, shape=(n,)).copy()

    I64P, U8P, F64P = C.POINTER(C.c_int64), C.POINTER(C.c_uint8), C.POINTER(C.c_double)

    @C.CFUNCTYPE(C.c_int, I64P, C.c_int64, U8P, F64P)
    def cb_mod(rows, n, lab, q):
        q[0] = O.get_b_modularity(np.ctypeslib.as_array(lab, shape=(n,)).astype(np.int8), state["b"][rows_of(rows, n)])
        return 0

    @C.CFUNCTYPE(C.c_int, I64P, C.c_int64, F64P, F64P)
    def cb_sl(rows, n, u2, th):
        _, u, t2, _, _ = O.spectral_cluster(state["b"][rows_of(rows, n)])
        np.ctypeslib.as_array(u2, shape=(n,))[:] = -u if state.get("flip") else u        # the sign of u2 is arbitrary
        th[0] = t2
        return 0

    @C.CFUNCTYPE(C.c_int, I64P, C.c_int64, F64P)
    def cb_deg(rows, n, d):
        np.ctypeslib.as_array(d, shape=(n,))[:] = O.b_to_d(state["b"][rows_of(rows, n)])
        return 0

    @C.CFUNCTYPE(C.c_int, C.POINTER(_lib.Csr), C.c_int, C.c_int, F64P, F64P)
    def cb_ls(ap, normalize, n_vec, u, sig):
        c = csr_of(ap)
        assert normalize == 0
        w, v = np.linalg.eigh((c @ c.T).toarray())
        top = v[:, ::-1][:, :n_vec]
        if state.get("flip"):
            top = -top
        np.ctypeslib.as_array(u, shape=(c.shape[0] * n_vec,))[:] = top.ravel()
        np.ctypeslib.as_array(sig, shape=(n_vec,))[:] = np.sqrt(np.maximum(w[::-1][:n_vec], 0))
        return 0

    @C.CFUNCTYPE(C.c_int, C.POINTER(_lib.Csr), F64P)
    def cb_tfidf(ap, v):
        a = csr_of(ap)
        np.ctypeslib.as_array(v, shape=(a.nnz,))[:] = O.b1_to_b2(a).data
        return 0

    @C.CFUNCTYPE(C.c_int, C.POINTER(_lib.Csr), F64P)
    def cb_l2(ap, v):
        a = csr_of(ap)
        np.ctypeslib.as_array(v, shape=(a.nnz,))[:] = O.b2_to_b(a).data
        return 0

    lib.vh_set_callbacks(cb_mod, cb_sl, cb_deg, cb_ls, cb_tfidf, cb_l2)
    lib.vh_last_error.restype = C.c_char_p
    lib.vh_splitmix.restype = C.c_uint64
    lib.vh_splitmix.argtypes = [C.c_uint64, C.c_int]
    lib.vh_dense_to_csr.restype = C.c_int64
    lib._keep = (cb_mod, cb_sl, cb_deg, cb_ls, cb_tfidf, cb_l2)

    def kmeans_tree(a, normalize=True, **kw):
        """Runs vh_kmeans_tree on host CSR `a`; returns a T.ClusteringTree."""
        state["b"] = O.b2_to_b(O.b1_to_b2(a) if normalize else a)
        state["flip"] = kw.pop("flip", False)
        h = T.spectral.HostCsr(a)
        o = T.make_opts(normalize=normalize, eigen_group="KMeansGroup", **kw)
        tp = C.POINTER(_lib.Tree)()
        st = lib.vh_kmeans_tree(C.byref(h.c), C.byref(o), C.byref(tp))
        assert st == 0, lib.vh_last_error()
        t = tp.contents
        n, m = int(t.n_nodes), int(t.n_rows)
        g = lambda ptr, k, dt: np.ctypeslib.as_array(ptr, shape=(k,)).astype(dt, copy=True)
        tree = T.ClusteringTree(left=g(t.left, n, np.int64), right=g(t.right, n, np.int64), parent=g(t.parent, n, np.int64),
                                q=g(t.q, n, np.float64), theta=g(t.theta, n, np.float64), iters=g(t.iters, n, np.int32),
                                item_begin=g(t.item_begin, n, np.int64), item_end=g(t.item_end, n, np.int64),
                                perm=g(t.perm, m, np.int64), stats={}, p_val=g(t.p_val, n, np.float64))
        lib.vh_tree_free(tp)
        return tree

    def permutation_test(a, tree, num_runs, seed, normalize=True):
        state["b"] = O.b2_to_b(O.b1_to_b2(a) if normalize else a)
        t, keep = T.spectral._tree_to_c(tree)
        p = np.empty(tree.n_nodes)
        st = lib.vh_permutation_test(C.c_int64(a.shape[0]), C.byref(t), int(num_runs), C.c_uint64(seed), p.ctypes.data_as(F64P))
        assert st == 0, lib.vh_last_error()
        return p

    lib.kmeans_tree, lib.permutation_test = kmeans_tree, permutation_test
    return lib


def _same_tree(tree, ot):
    import tmc_oracle as O
    return O.canonical_tree(tree.left, tree.right, tree.items) == O.canonical_tree(ot.left, ot.right, lambda i: ot.items[i])


def _p_by_items(items_of, n, p):
    return {frozenset(int(x) for x in items_of(i)): p[i] for i in range(n)}


def test_variants_splitmix_and_kmeans_match_the_oracle(variants_host):
    import tmc_oracle as O
    r = O.SplitMix64(12345)
    ref = [r.next() for _ in range(5)]
    assert [variants_host.vh_splitmix(12345, k) for k in range(5)] == ref
    r = O.SplitMix64(1234567)                                # known-answer vectors of Vigna's splitmix64.c for seed 1234567
    assert [r.next() for _ in range(3)] == [6457827717110365317, 3203168211198807973, 9817491932198370423]
    assert variants_host.vh_splitmix(1234567, 2) == 9817491932198370423
    rng = np.random.default_rng(0)
    for n, e in [(2, 1), (3, 1), (50, 1), (64, 3), (7, 2)]:
        pts = rng.standard_normal((n, e))
        if n == 50:
            pts[:25] += 4.0
        lab = np.zeros(n, dtype=np.uint8)
        variants_host.vh_kmeans2(pts.ctypes.data_as(C.POINTER(C.c_double)), C.c_int64(n), e, lab.ctypes.data_as(C.POINTER(C.c_uint8)))
        assert np.array_equal(lab, O.kmeans2_labels(pts).astype(np.uint8))
        assert 0 < lab.sum() < n
    same = np.ones((5, 2))
    lab = np.zeros(5, dtype=np.uint8)
    variants_host.vh_kmeans2(same.ctypes.data_as(C.POINTER(C.c_double)), C.c_int64(5), 2, lab.ctypes.data_as(C.POINTER(C.c_uint8)))
    assert lab.sum() == 5 and O.kmeans2_labels(same).sum() == 5       # no spread: one cluster


def test_variants_dense_to_csr(variants_host):
    import scipy.sparse as sp
    rng = np.random.default_rng(1)
    d = rng.random((7, 9)) * (rng.random((7, 9)) < 0.4)
    d[3] = 0.0
    rp = np.zeros(8, dtype=np.int64); ci = np.zeros(63, dtype=np.int32); v = np.zeros(63)
    nnz = variants_host.vh_dense_to_csr(d.ctypes.data_as(C.POINTER(C.c_double)), C.c_int64(7), C.c_int64(9),
                                        rp.ctypes.data_as(C.POINTER(C.c_int64)), ci.ctypes.data_as(C.POINTER(C.c_int32)),
                                        v.ctypes.data_as(C.POINTER(C.c_double)))
    ref = sp.csr_matrix(d)
    assert nnz == ref.nnz and np.array_equal(rp, ref.indptr) and np.array_equal(ci[:nnz], ref.indices) and np.array_equal(v[:nnz], ref.data)


@pytest.mark.parametrize("num_eigen", [1, 3])
def test_variants_kmeans_tree_layout_and_oracle(variants_host, golden_tiny, num_eigen):
    import tmc_oracle as O
    a, _ = golden_tiny
    a = a[:240]
    ot = O.hierarchical_spectral_cluster(a, normalize=True, eigen_group="KMeansGroup", num_eigen=num_eigen)
    for flip in (False, True):                                # the tree must not depend on the sign the solver returns
        t = variants_host.kmeans_tree(a, num_eigen=num_eigen, flip=flip)
        assert t.n_nodes == ot.n_nodes and _same_tree(t, ot)
        # flat pre-order contract of tmc_tree (include/tmc.h)
        assert t.parent[0] == -1 and t.item_begin[0] == 0 and t.item_end[0] == a.shape[0]
        assert sorted(t.perm.tolist()) == list(range(a.shape[0]))
        for i in range(t.n_nodes):
            if t.left[i] >= 0:
                l, r = t.left[i], t.right[i]
                assert l == i + 1 and t.parent[l] == i and t.parent[r] == i
                assert t.item_begin[l] == t.item_begin[i] and t.item_end[l] == t.item_begin[r] and t.item_end[r] == t.item_end[i]
            else:
                assert np.all(np.diff(t.items(i)) > 0)
        qa, qb = _p_by_items(t.items, t.n_nodes, t.q), _p_by_items(lambda i: ot.items[i], ot.n_nodes, ot.q)
        assert all(abs(qa[k] - qb[k]) < 1e-12 for k in qb)
    assert ot.n_nodes > 5
    t5 = variants_host.kmeans_tree(a, num_eigen=num_eigen, min_size=7, min_modularity=0.02)
    o5 = O.hierarchical_spectral_cluster(a, normalize=True, eigen_group="KMeansGroup", num_eigen=num_eigen, min_size=7, min_modularity=0.02)
    assert _same_tree(t5, o5) and min(len(t5.items(i)) for i in t5.leaves()) >= 7


def test_variants_permutation_p_values_match_the_oracle(variants_host, golden_tiny):
    import tmc_oracle as O
    a, _ = golden_tiny
    a = a[:90]
    for kw in (dict(), dict(min_modularity=-1.0)):
        ot = O.hierarchical_spectral_cluster(a, normalize=True, eigen_group="KMeansGroup", num_runs=12, seed=99, **kw)
        t = variants_host.kmeans_tree(a, num_runs=12, seed=99, **kw)
        assert _same_tree(t, ot)
        pa, pb = _p_by_items(t.items, t.n_nodes, t.p_val), _p_by_items(lambda i: ot.items[i], ot.n_nodes, ot.p_val)
        for k, v in pb.items():
            assert (np.isnan(pa[k]) and v is None) or pa[k] == v
    assert np.nanmax(t.p_val) > 0                            # the saturated tree splits noise
    # on a finished SignGroup tree, with children deliberately swapped: p-values belong to the split, not to its numbering
    os_ = O.hierarchical_spectral_cluster(a, normalize=True, num_runs=8, seed=5)
    left = [os_.right[i] if os_.left[i] >= 0 else -1 for i in range(os_.n_nodes)]
    right = [os_.left[i] if os_.left[i] >= 0 else -1 for i in range(os_.n_nodes)]
    # re-number pre-order with the swapped children
    order, st = [], [0]
    while st:
        u = st.pop(); order.append(u)
        if left[u] >= 0:
            st.append(right[u]); st.append(left[u])
    new_id = {u: i for i, u in enumerate(order)}
    n = len(order)
    L = np.full(n, -1, dtype=np.int64); R = np.full(n, -1, dtype=np.int64); P = np.full(n, -1, dtype=np.int64)
    ib = np.zeros(n, dtype=np.int64); ie = np.zeros(n, dtype=np.int64); perm = []
    for u in order:
        i = new_id[u]
        ib[i] = len(perm)
        if left[u] >= 0:
            L[i], R[i] = new_id[left[u]], new_id[right[u]]; P[L[i]] = P[R[i]] = i
        else:
            perm.extend(int(x) for x in os_.items[u])
    for u in reversed(order):
        i = new_id[u]
        ie[i] = ib[i] + len(os_.items[u])
    tree = T.ClusteringTree(left=L, right=R, parent=P, q=np.array([os_.q[u] for u in order]), theta=np.zeros(n), iters=np.zeros(n, np.int32),
                            item_begin=ib, item_end=ie, perm=np.array(perm, dtype=np.int64), stats={})
    p = variants_host.permutation_test(a, tree, 8, 5)
    for u in order:
        if os_.p_val[u] is None:
            assert np.isnan(p[new_id[u]])
        else:
            assert p[new_id[u]] == os_.p_val[u]


def test_haskell_binding_offsets_match_the_header():
    """integration/haskell/TmcFFI.hs pokes / peeks the C structs by byte offset (no hsc2hs): every off_* / size_* constant in it
    must equal the layout of include/tmc.h (taken from the ctypes mirror, whose sizes the library static_asserts)."""
    import re
    src = open(os.path.join(ROOT, "integration", "haskell", "TmcFFI.hs")).read()
    structs = {"csr": _lib.Csr, "opts": _lib.Opts, "tree": _lib.Tree}
    seen = 0
    for m in re.finditer(r"^(off|size)_(csr|opts|tree)(?:_(\w+))? = (\d+)$", src, re.M):
        kind, st, field, val = m.group(1), structs[m.group(2)], m.group(3), int(m.group(4))
        if kind == "size":
            assert C.sizeof(st) == val, m.group(0)
        else:
            assert getattr(st, field).offset == val, m.group(0)
        seen += 1
    assert seen >= 25
    # every foreign import names a real export with that many arguments
    lib = _lib.load()
    for m in re.finditer(r'foreign import ccall safe "(\w+)"\s+\w+\s*::\s*(.*)$', src, re.M):
        name, sig = m.group(1), m.group(2)
        assert hasattr(lib, name)
        n_args = sig.count("->")
        assert n_args == len(_lib.SIGNATURES[name][1]), (name, sig)


def test_ctypes_mirror_matches_the_header_field_by_field(tmp_path):
    """Compile include/tmc.h with gcc and compare offsetof / sizeof of every struct field with the ctypes mirror in _lib.py (which
    the Haskell binding's offsets are in turn checked against)."""
    import re
    hdr = open(os.path.join(ROOT, "include", "tmc.h")).read()
    structs = {"tmc_csr": _lib.Csr, "tmc_opts": _lib.Opts, "tmc_tree": _lib.Tree, "tmc_csr_out": _lib.CsrOut}
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{os.path.join(ROOT, "include", "tmc.h")}"', "int main(void) {"]
    for name, st in structs.items():
        assert re.search(r"}\s*" + name + r"\s*;", hdr), name
        lines.append(f'  printf("{name} size %zu\\n", sizeof({name}));')
        for field, _ in st._fields_:
            lines.append(f'  printf("{name} {field} %zu\\n", offsetof({name}, {field}));')
    lines += ["  return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", str(src), "-o", str(exe)], check=True)      # the header is plain C
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split("\n")
    seen = 0
    for ln in filter(None, out):
        name, field, val = ln.split()
        st = structs[name]
        assert (C.sizeof(st) if field == "size" else getattr(st, field).offset) == int(val), ln
        seen += 1
    assert seen == sum(len(s._fields_) + 1 for s in structs.values())
