"""Host-side mirror of the reference's operator interface for the make-tree hot path.

The reference's host is Haskell (no GHC in this image), so this is the Python
mirror of the same call surface, bound to libtmc.so over the C ABI in
include/tmc.h.  Names, argument meaning and error behaviour follow the
reference call sites:

  tfidf_scale_sparse_mat        src/TooManyCells/Matrix/Preprocess.hs:204-205   (unB2 . b1ToB2 . B1)
  b1_to_b2 / b2_to_b            Math.Clustering.Spectral.Sparse (spectral-clustering)
  b_to_d / second_left /
  spectral_cluster              Math.Clustering.Spectral.Sparse
  get_b_modularity              Math.Modularity.Sparse (modularity)
  hierarchical_spectral_cluster Math.Clustering.Hierarchical.Spectral.Sparse, called at
                                src/TooManyCells/MakeTree/Cluster.hs:147-156,169
  h_spec_clust                  src/TooManyCells/MakeTree/Cluster.hs:135-187

Every function runs on the GPU through libtmc; there is no CPU path here.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib
from ._lib import Csr, Opts, Tree, TmcError, check


# ----------------------------------------------------------------------------- marshalling
class HostCsr:
    """Borrowed host CSR (cells x features) in the dtypes the C ABI wants."""

    def __init__(self, mat, row_offset=0, n_rows_global=0, compact=False):
        if hasattr(mat, "indptr"):              # scipy.sparse CSR
            indptr, indices, data, shape = mat.indptr, mat.indices, mat.data, mat.shape
        else:
            indptr, indices, data, shape = mat[:4]
            if len(mat) == 6:                   # (indptr, indices, data, shape, row_offset, n_rows_global): a rank's row block
                row_offset, n_rows_global = mat[4], mat[5]
        self.indptr = np.ascontiguousarray(indptr, dtype=np.int64)
        # compact element types travel as they are (tmc_csr.val_type / idx_type); anything else is converted to the defaults
        idt = np.asarray(indices).dtype.name if compact else "int32"
        vdt = np.asarray(data).dtype.name if compact else "float64"
        idt = idt if idt in _lib.IDX_TYPES else "int32"
        vdt = vdt if vdt in _lib.VAL_TYPES else "float64"
        self.indices = np.ascontiguousarray(indices, dtype=idt)
        self.data = np.ascontiguousarray(data, dtype=vdt)
        self.shape = (int(shape[0]), int(shape[1]))
        self.c = Csr(self.shape[0], self.shape[1], int(self.data.size),
                     self.indptr.ctypes.data, self.indices.ctypes.data, self.data.ctypes.data,
                     int(row_offset), int(n_rows_global), _lib.VAL_TYPES[vdt], _lib.IDX_TYPES[idt])

    @property
    def nnz(self):
        return int(self.data.size)


def make_opts(min_modularity=0.0, min_size=1, normalize=False, eigen_group="SignGroup", num_eigen=1,
              num_runs=0, seed=None, tol=None, max_lanczos=None, max_restarts=None, max_active=None,
              device=-1, profile=False, scatter_mode=0, sign_stop=True, warm_start=True, small_rows=None) -> Opts:
    lib = _lib.load()
    o = Opts()
    lib.tmc_opts_default(C.byref(o))
    o.eigen_group = {"SignGroup": 0, "KMeansGroup": 1}[eigen_group] if isinstance(eigen_group, str) else int(eigen_group)
    o.num_eigen = int(num_eigen)
    o.min_modularity = float(min_modularity)
    o.min_size = int(min_size)
    o.normalize = int(bool(normalize))
    o.num_runs = int(num_runs or 0)
    if seed is not None:
        o.seed = int(seed)
    if tol is not None:
        o.tol = float(tol)
    if max_lanczos is not None:
        o.max_lanczos = int(max_lanczos)
    if max_restarts is not None:
        o.max_restarts = int(max_restarts)
    if max_active is not None:
        o.max_active = int(max_active)
    o.device = int(device)
    o.profile = int(bool(profile))
    o.scatter_mode = int(scatter_mode)
    o.sign_stop = 0 if sign_stop else -1      # sign-certified Lanczos stop (tmc.h); False = always iterate to `tol`
    o.warm_start = 0 if warm_start else -1    # children start from the parent's Krylov space (tmc.h); False = pseudo-random start
    if small_rows is not None:                # direct solve of small nodes (tmc.h): None = default threshold, 0 / False = off
        o.small_rows = int(small_rows) if small_rows else -1
    return o


# ----------------------------------------------------------------------------- device matrix handle
class DeviceB:
    """B = b2ToB(B2) resident in HBM (optionally idf-scaled first: getB True)."""

    def __init__(self, handle, shape, nnz, keepalive=None):
        self.handle = handle
        self.shape = shape
        self.nnz = nnz
        self._keep = keepalive

    @classmethod
    def from_host(cls, b2, normalize=False, device=-1):
        lib = _lib.load()
        h = HostCsr(b2, compact=True)          # uint16 / uint8 / float32 values and uint16 indices travel as they are
        out = C.c_void_p()
        check(lib.tmc_matrix_upload(C.byref(h.c), int(bool(normalize)), int(device), C.byref(out)))
        return cls(out, h.shape, h.nnz)

    @classmethod
    def adopt_device(cls, n_rows, n_cols, row_ptr, col, val, normalize=False, device=-1, row_offset=0, n_rows_global=0):
        """Adopt torch CUDA tensors (int64 row_ptr, int32 col, float64 val) without copying; val is normalised in place.
        Under a communicator pass this rank's row block with row_offset / n_rows_global."""
        lib = _lib.load()
        out = C.c_void_p()
        check(lib.tmc_matrix_adopt_dev(int(n_rows), int(n_cols), int(val.numel()), row_ptr.data_ptr(), col.data_ptr(),
                                       val.data_ptr(), int(row_offset), int(n_rows_global), int(bool(normalize)), int(device),
                                       C.byref(out)))
        return cls(out, (int(n_rows), int(n_cols)), int(val.numel()), keepalive=(row_ptr, col, val))

    def info(self):
        """Storage of the device copy: value type, dense-panel width, residual CSR entries (tmc_matrix_info)."""
        v = (C.c_int64 * 7)()
        check(_lib.load().tmc_matrix_info(self.handle, v, 7))
        return {"value_type": ("f64", "u16", "one")[int(v[0])], "panel_cols": int(v[1]), "panel_row_bytes": int(v[2]),
                "residual_nnz": int(v[3]), "nnz": int(v[4]), "n_rows": int(v[5]), "panel_bits": int(v[6])}

    def free(self):
        if self.handle:
            _lib.load().tmc_matrix_free(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def _rows_arg(rows):
    if rows is None:
        return None, 0, None
    r = np.ascontiguousarray(rows, dtype=np.int64)
    return r.ctypes.data, int(r.size), r


# ----------------------------------------------------------------------------- A1 / A2
def b1_to_b2(b1):
    """idf column scaling (b1ToB2).  Returns (values, idf) for the same sparsity pattern."""
    lib = _lib.load()
    h = HostCsr(b1)
    val = np.empty(h.nnz, dtype=np.float64)
    idf = np.empty(h.shape[1], dtype=np.float64)
    check(lib.tmc_tfidf(C.byref(h.c), val.ctypes.data, idf.ctypes.data))
    return val, idf


def tfidf_scale_sparse_mat(mat):
    """Preprocess.hs:204-205 — returns a scipy CSR with the tf-idf scaled values."""
    import scipy.sparse as sp
    val, _ = b1_to_b2(mat)
    h = HostCsr(mat)
    return sp.csr_matrix((val, h.indices.copy(), h.indptr.copy()), shape=h.shape)


def b2_to_b(b2):
    """Row L2 normalisation (b2ToB).  Returns (values, row_norms)."""
    lib = _lib.load()
    h = HostCsr(b2)
    val = np.empty(h.nnz, dtype=np.float64)
    nrm = np.empty(h.shape[0], dtype=np.float64)
    check(lib.tmc_row_l2(C.byref(h.c), val.ctypes.data, nrm.ctypes.data))
    return val, nrm


# ----------------------------------------------------------------------------- A3 - A8 (per node)
def b_to_d(b: DeviceB, rows=None):
    lib = _lib.load()
    p, n, keep = _rows_arg(rows)
    out = np.empty(n if rows is not None else b.shape[0], dtype=np.float64)
    check(lib.tmc_degree(b.handle, p, n, out.ctypes.data))
    return out


def second_left(b: DeviceB, rows=None, **opt_kw):
    """u2 of C = D^{-1/2} B_S.  Returns (u2, theta = sigma_2^2, lanczos_steps)."""
    lib = _lib.load()
    p, n, keep = _rows_arg(rows)
    m = n if rows is not None else b.shape[0]
    u2 = np.empty(m, dtype=np.float64)
    th = C.c_double()
    it = C.c_int32()
    o = make_opts(**opt_kw)
    check(lib.tmc_second_left(b.handle, p, n, C.byref(o), u2.ctypes.data, C.byref(th), C.byref(it)))
    return u2, th.value, it.value


def spectral_cluster(b: DeviceB, rows=None, **opt_kw):
    """0/1 labels: 1 iff u2_i >= 0 (spectralCluster)."""
    u2, _, _ = second_left(b, rows, **opt_kw)
    return (u2 >= 0).astype(np.uint8)


def get_b_modularity(labels, b: DeviceB, rows=None) -> float:
    lib = _lib.load()
    p, n, keep = _rows_arg(rows)
    lab = np.ascontiguousarray(labels, dtype=np.uint8)
    q = C.c_double()
    check(lib.tmc_modularity(b.handle, p, n, lab.ctypes.data, C.byref(q)))
    return q.value


def partition_rows(rows, labels):
    """Stable split of sorted row ids into [label 0 | label 1]. Returns (rows_out, n_left)."""
    lib = _lib.load()
    r = np.ascontiguousarray(rows, dtype=np.int64)
    lab = np.ascontiguousarray(labels, dtype=np.uint8)
    out = np.empty_like(r)
    nl = C.c_int64()
    check(lib.tmc_partition(r.ctypes.data, int(r.size), lab.ctypes.data, out.ctypes.data, C.byref(nl)))
    return out, nl.value


def spmv_pair(b: DeviceB, x, rows=None):
    """x' = C_S (C_S^T x)  (the las2 operator)."""
    lib = _lib.load()
    p, n, keep = _rows_arg(rows)
    xx = np.ascontiguousarray(x, dtype=np.float64)
    out = np.empty_like(xx)
    check(lib.tmc_spmv_pair(b.handle, p, n, xx.ctypes.data, out.ctypes.data))
    return out


def bench_spmv_pair(b: DeviceB, reps=20):
    lib = _lib.load()
    sc, ga = C.c_double(), C.c_double()
    check(lib.tmc_bench_spmv_pair(b.handle, int(reps), C.byref(sc), C.byref(ga)))
    return sc.value, ga.value


def lsa_sparse_mat(mat, n_dim, normalize=True, **opt_kw):
    """lsaSparseMat (Preprocess.hs:560-571): cells x n_dim matrix of the first n_dim left singular vectors of the idf-scaled
    matrix (`secondLeft 1 N . b1ToB2`).  Returns (U, singular_values, converged)."""
    lib = _lib.load()
    h = HostCsr(mat)
    u = np.empty((h.shape[0], int(n_dim)), dtype=np.float64)
    sig = np.empty(int(n_dim), dtype=np.float64)
    conv = C.c_int()
    o = make_opts(**opt_kw)
    check(lib.tmc_left_singular(C.byref(h.c), int(bool(normalize)), int(n_dim), C.byref(o), u.ctypes.data, sig.ctypes.data, C.byref(conv)))
    return u, sig, bool(conv.value)


def center_scale_sparse_cell(mat):
    """centerScaleSparseCell on every row (Preprocess.hs:263-271): stored entries -> (x - mu) / sigma with the mean and the
    sample standard deviation of the whole (dense) row; zeros stay zeros; sigma = 0 -> 0.  O(nnz) host pass (scipy CSR in / out)."""
    import scipy.sparse as sp
    a = sp.csr_matrix(mat, dtype=np.float64)
    n = a.shape[1]
    s1 = np.asarray(a.sum(axis=1)).ravel()
    s2 = np.asarray(a.multiply(a).sum(axis=1)).ravel()
    mu = s1 / n
    sigma = np.sqrt(np.maximum((s2 - n * mu * mu) / max(n - 1, 1), 0.0))
    rows = np.repeat(np.arange(a.shape[0]), np.diff(a.indptr))
    out = a.copy()
    with np.errstate(divide="ignore", invalid="ignore"):
        out.data = np.where(sigma[rows] == 0, 0.0, (a.data - mu[rows]) / sigma[rows])
    out.eliminate_zeros()
    return out


def svd_sparse_mat(mat, n_dim, **opt_kw):
    """svdSparseMat (`--svd N`, Preprocess.hs:573-585): `secondLeft 1 N` of the cell-standardised matrix — its first N left
    singular vectors, through the same GPU Lanczos as `--lsa` (the matrix is signed: tmc_left_singular takes it as it is).
    Returns (U [cells x N], singular_values, converged)."""
    return lsa_sparse_mat(center_scale_sparse_cell(mat), n_dim, normalize=False, **opt_kw)


def pca_sparse_mat(mat, n_dim, **opt_kw):
    """pcaSparseMat (`--pca N`, Preprocess.hs:543-556): features standardised, then `scaled #~# V_N` = U_N diag(sigma_N), the
    scores on the first N principal directions.  Returns (scores [cells x N], singular_values, converged)."""
    import scipy.sparse as sp
    scaled = center_scale_sparse_cell(sp.csr_matrix(mat).T.tocsr()).T.tocsr()
    u, sig, conv = lsa_sparse_mat(scaled, n_dim, normalize=False, **opt_kw)
    return u * sig[None, :], sig, conv


# ----------------------------------------------------------------------------- A0 / A9: the tree
@dataclass
class ClusteringTree:
    """Flat pre-order ClusteringTree (children order [label 0, label 1])."""
    left: np.ndarray
    right: np.ndarray
    parent: np.ndarray
    q: np.ndarray            # _ngMod of every vertex (leaf Q is discarded by clusteringTreeToTree)
    theta: np.ndarray
    iters: np.ndarray
    item_begin: np.ndarray
    item_end: np.ndarray
    perm: np.ndarray         # original row ids; node i holds perm[item_begin[i]:item_end[i]] (ascending inside leaves)
    stats: dict
    p_val: np.ndarray = None  # _pVal of every split (NaN = Nothing: leaves, or no permutation test was asked for)

    @property
    def n_nodes(self):
        return int(self.left.size)

    def items(self, i):
        return self.perm[self.item_begin[i]:self.item_end[i]]

    def is_leaf(self, i):
        return self.left[i] < 0

    def leaves(self):
        return np.nonzero(self.left < 0)[0]

    def leaf_sets(self):
        return sorted(tuple(int(x) for x in self.items(i)) for i in self.leaves())

    def tree_hash(self):
        """Canonical SHA-1 of leaf membership + topology (see `tree_signature`)."""
        return tree_signature(self.left, self.right, self.item_begin, self.item_end, self.perm)

    def tree_hash_strict(self):
        """The same with every vertex named by its SET of cells (see `tree_signature_strict`)."""
        return tree_signature_strict(self.left, self.right, self.item_begin, self.item_end, self.perm)

    def tree_hashes(self):
        """(tree_hash(), tree_hash_strict()) from one pass over the tree."""
        r = _tree_names(self.left, self.right, self.item_begin, self.item_end, self.perm)
        return r[0], r[2]


def tree_signature(left, right, begin, end, perm) -> str:
    """SHA-1 of a flat tree that does not depend on the eigenvector sign (child order), the node numbering or the rank
    count: every vertex is named by (smallest row id among its cells, number of cells) -- distinct vertices of a tree
    differ in that pair -- and the hash covers (1) for every cell the name of its leaf (membership) and (2) the sorted
    list of all vertex names (topology: which sets of leaves are merged).  Two trees have the same signature iff they
    have identical leaf membership and identical topology up to swapping children -- with one blind spot: vertices of two
    DIFFERENT trees can share (smallest row, size) although their cells differ ({a, b} | {c} against {a, c} | {b} inside a
    three-cell vertex of a saturated tree).  `tree_signature_strict` closes it; this one is kept because the hashes of rounds
    1 and 2 in profiles/ were computed with it."""
    return _tree_names(left, right, begin, end, perm)[0]


def tree_signature_strict(left, right, begin, end, perm) -> str:
    """Like `tree_signature`, with every vertex named by (smallest row id, number of cells, 64-bit hash of its SET of cells:
    the sum mod 2^64 of a mixed row id over the cells): equal iff the two trees consist of the same sets of cells, i.e. have
    identical leaf membership and identical topology up to swapping children (up to a 2^-64 collision)."""
    return _tree_names(left, right, begin, end, perm)[2]


def tree_vertex_names(left, right, begin, end, perm, strict=False):
    """(smallest row id, number of cells) of every vertex, in the tree's own numbering: an (n_nodes, 2) int64 array;
    strict=True appends the 64-bit hash of the vertex's set of cells as a third column (as int64 bit patterns)."""
    r = _tree_names(left, right, begin, end, perm)
    return np.concatenate([r[1], r[3].view(np.int64)[:, None]], axis=1) if strict else r[1]


def _tree_names(left, right, begin, end, perm):
    import hashlib
    left = np.asarray(left, dtype=np.int64); right = np.asarray(right, dtype=np.int64)
    begin = np.asarray(begin, dtype=np.int64); end = np.asarray(end, dtype=np.int64); perm = np.asarray(perm, dtype=np.int64)
    n = left.size
    mn = np.zeros(n, dtype=np.int64)
    leaves = np.nonzero(left < 0)[0]
    order = leaves[np.argsort(begin[leaves], kind="stable")]
    nonempty = order[end[order] > begin[order]]
    mn[nonempty] = np.minimum.reduceat(perm, begin[nonempty]) if nonempty.size else 0
    # reduceat runs to the next start: correct because the leaves tile perm in ascending `begin` order
    # inner vertices: min over the children, deepest level first (vectorised per level: a saturated 1 M-cell tree has 2 M vertices)
    parent = np.full(n, -1, dtype=np.int64)
    inner = np.nonzero(left >= 0)[0]
    parent[left[inner]] = inner; parent[right[inner]] = inner
    depth = np.zeros(n, dtype=np.int64)
    frontier = np.array([0], dtype=np.int64) if n else np.zeros(0, dtype=np.int64)
    levels = []
    while frontier.size:
        levels.append(frontier)
        kids = frontier[left[frontier] >= 0]
        frontier = np.concatenate([left[kids], right[kids]]) if kids.size else np.zeros(0, dtype=np.int64)
    big = np.iinfo(np.int64).max
    mn[inner] = big
    for lev in reversed(levels[1:]):
        np.minimum.at(mn, parent[lev], mn[lev])
    sizes = (end - begin)[nonempty]
    if int(sizes.sum()) != perm.size or (nonempty.size and (begin[nonempty[0]] != 0 or np.any(begin[nonempty][1:] != end[nonempty][:-1]))):
        raise ValueError("the leaves do not tile perm")
    leaf_name = np.empty(perm.size, dtype=np.int64)
    leaf_name[perm] = np.repeat(mn[nonempty], sizes)
    names0 = np.stack([mn, end - begin], axis=1)
    names = names0[np.lexsort((names0[:, 1], names0[:, 0]))]
    h = hashlib.sha1()
    h.update(np.ascontiguousarray(leaf_name).tobytes())
    h.update(np.ascontiguousarray(names).tobytes())
    # set hash of every vertex: sum mod 2^64 of splitmix64(row id) over its cells (leaves by reduceat, inner vertices bottom-up)
    x = perm.astype(np.uint64) + np.uint64(0x9E3779B97F4A7C15)
    x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    x = x ^ (x >> np.uint64(31))
    sh = np.zeros(n, dtype=np.uint64)
    if nonempty.size:
        sh[nonempty] = np.add.reduceat(x, begin[nonempty])
    for lev in reversed(levels[1:]):
        np.add.at(sh, parent[lev], sh[lev])
    leaf_set = np.empty(perm.size, dtype=np.uint64)
    leaf_set[perm] = np.repeat(sh[nonempty], sizes)
    trip = np.stack([mn.astype(np.uint64), (end - begin).astype(np.uint64), sh], axis=1)
    trip = trip[np.lexsort((trip[:, 2], trip[:, 1], trip[:, 0]))]
    h2 = hashlib.sha1()
    h2.update(np.ascontiguousarray(leaf_set).tobytes())
    h2.update(np.ascontiguousarray(trip).tobytes())
    return h.hexdigest(), names0, h2.hexdigest(), sh


def _tree_from_c(tp) -> ClusteringTree:
    t = tp.contents
    n, m = int(t.n_nodes), int(t.n_rows)

    def arr(ptr, count, dtype):
        return np.ctypeslib.as_array(ptr, shape=(count,)).astype(dtype, copy=True) if count else np.zeros(0, dtype)

    out = ClusteringTree(
        left=arr(t.left, n, np.int64), right=arr(t.right, n, np.int64), parent=arr(t.parent, n, np.int64),
        q=arr(t.q, n, np.float64), theta=arr(t.theta, n, np.float64), iters=arr(t.iters, n, np.int32),
        item_begin=arr(t.item_begin, n, np.int64), item_end=arr(t.item_end, n, np.int64), perm=arr(t.perm, m, np.int64),
        stats=dict(n_sweeps=int(t.n_sweeps), spmv_pairs_nnz=int(t.spmv_pairs_nnz), kernel_launches=int(t.kernel_launches),
                   engine_ms=float(t.engine_ms), spmv_ms=float(t.spmv_ms), spmv_alg_bytes=float(t.spmv_alg_bytes),
                   value_bytes=int(t.value_bytes)),
        p_val=arr(t.p_val, n, np.float64))
    _lib.load().tmc_tree_free(tp)
    return out


def _tree_to_c(tree: ClusteringTree):
    """Borrowed C view of a ClusteringTree (only the topology fields tmc_permutation_test reads)."""
    keep = [np.ascontiguousarray(getattr(tree, k), dtype=np.int64) for k in ("left", "right", "parent", "item_begin", "item_end", "perm")]
    q = np.ascontiguousarray(tree.q, dtype=np.float64)
    t = Tree()
    t.n_nodes, t.n_rows = tree.n_nodes, int(tree.perm.size)
    for name, a in zip(("left", "right", "parent", "item_begin", "item_end", "perm"), keep):
        setattr(t, name, a.ctypes.data_as(C.POINTER(C.c_int64)))
    t.q = q.ctypes.data_as(C.POINTER(C.c_double))
    return t, keep + [q]


def permutation_test(b: DeviceB, tree: ClusteringTree, num_runs, seed=None) -> np.ndarray:
    """The `runs` argument of hierarchicalSpectralCluster (--numRuns, Options.hs:174) on a finished tree: empirical p-value
    of every split (NaN for leaves)."""
    lib = _lib.load()
    t, keep = _tree_to_c(tree)
    p = np.empty(tree.n_nodes, dtype=np.float64)
    seed = make_opts().seed if seed is None else int(seed)
    check(lib.tmc_permutation_test(b.handle, C.byref(t), int(num_runs), seed, p.ctypes.data))
    return p


def hierarchical_spectral_cluster(mat, eigen_group="SignGroup", normalize=False, num_eigen=None, min_size=None,
                                  min_modularity=None, num_runs=None, **engine_kw) -> ClusteringTree:
    """hierarchicalSpectralCluster eigenGroup normFlag numEigen minSize minMod runs items (Left mat).

    `mat` is a host CSR (scipy / (indptr, indices, data, shape)) — copied to the GPU inside the call —
    a DeviceB already resident in HBM, or a dense numpy array (cells x features: the `--dense` call,
    HSD.hierarchicalSpectralCluster, Cluster.hs:157-167).  `None` arguments mean Haskell `Nothing`.
    `eigen_group="KMeansGroup"` (with `num_eigen`) and `num_runs` select the non-default variants (single GPU).
    """
    lib = _lib.load()
    o = make_opts(min_modularity=0.0 if min_modularity is None else min_modularity,
                  min_size=1 if min_size is None else min_size, normalize=normalize, eigen_group=eigen_group,
                  num_eigen=1 if num_eigen is None else num_eigen, num_runs=num_runs or 0, **engine_kw)
    tp = C.POINTER(Tree)()
    if isinstance(mat, DeviceB):
        check(lib.tmc_make_tree_dev(mat.handle, C.byref(o), C.byref(tp)))
    elif isinstance(mat, np.ndarray):
        if mat.ndim != 2:
            raise TmcError(_lib.ERR_INVALID, "dense matrix must be 2-D (cells x features)")
        d = np.ascontiguousarray(mat, dtype=np.float64)
        check(lib.tmc_make_tree_dense(d.ctypes.data, int(d.shape[0]), int(d.shape[1]), C.byref(o), C.byref(tp)))
    else:
        h = HostCsr(mat, compact=True)
        check(lib.tmc_make_tree(C.byref(h.c), C.byref(o), C.byref(tp)))
    return _tree_from_c(tp)


def h_spec_clust(matrix, row_names, eigen_group="SignGroup", num_eigen=None, min_modularity=None, num_runs=None,
                 dense=False, **engine_kw):
    """hSpecClust (Cluster.hs:135-187): returns (cluster_list, tree) where cluster_list is
    [(barcode, row, [leaf, ..., 0])] in DFS-leaf order and tree is the ClusteringTree."""
    from .tree_io import cluster_list
    if dense and not isinstance(matrix, (np.ndarray, DeviceB)):       # sparseToHMat (Cluster.hs:166)
        matrix = np.asarray(matrix.todense()) if hasattr(matrix, "todense") else np.asarray(matrix, dtype=np.float64)
    tree = hierarchical_spectral_cluster(matrix, eigen_group, False, num_eigen, None, min_modularity, num_runs, **engine_kw)
    return cluster_list(tree, row_names), tree
