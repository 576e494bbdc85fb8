"""CPU oracle for the too-many-cells `make-tree` spectral-bipartition path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package may import this
module; only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of `bench.py` do, and only as the checker.

PARITY UNPINNED.  The arithmetic of this path is not in /root/reference: the
reference only *calls* it (`hierarchicalSpectralCluster`,
src/TooManyCells/MakeTree/Cluster.hs:147-156; `b1ToB2`,
src/TooManyCells/Matrix/Preprocess.hs:204-205).  The bodies live in un-vendored
Hackage packages (hierarchical-spectral-clustering 0.5.0.1, spectral-clustering
>= 0.3.0.2 @18729ec5, modularity 0.2.1.1, sparse-linear-algebra @dbad792f,
hmatrix-svdlibc 0.5.0.1 = SVDLIBC las2), none of which is in this container,
and the reference ships no tests or numeric fixtures for the path (its only
result artefacts, workshop/out/*, have no input matrix).  This file therefore
restates the *published* algorithm (Shu et al. 2011, "Efficient spectral
neighborhood blocking for entity resolution"; Newman-Girvan modularity on the
implicit graph A = B B^T - I) in exact-math fp64, anchored on the reference's
call sites.  Every function cites the call site / package function it follows.
Switches for the unsettled upstream details are module constants below.

The eigen-solve is *exact* (dense LAPACK `eigh` for small nodes, ARPACK at
tol ~ 1e-14 for large ones), so both the reference binary (SVDLIBC, kappa=1e-6)
and the GPU engine are approximations of this oracle.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

# ---- switches for [UPSTREAM-RECALL] details -------------
IDF_LOG = math.log          # (1) natural log of m/df, df = #entries > 0
DEGREE_USES_ABS = True      # (2) d = |B| (|B|^T 1)
SIGN_RULE_GE_ZERO_IS_ONE = True   # (3) label 1 iff u2_i >= 0
STRICT_Q_GT_MINMOD = True   # (5) split iff Q > minMod
DENSE_MAX_ROWS = 2600       # nodes up to this many rows use dense eigh


# ---- A1: b1ToB2 (spectral-clustering; call site Preprocess.hs:204-205) ---------
def b1_to_b2(b1: sp.csr_matrix) -> sp.csr_matrix:
    """idf column scaling: B2[i,j] = B1[i,j] * ln(m / df_j), df_j = #{i: B1[i,j] > 0}."""
    b1 = sp.csr_matrix(b1, dtype=np.float64)
    m, n = b1.shape
    df = np.zeros(n, dtype=np.int64)
    np.add.at(df, b1.indices[b1.data > 0], 1)
    with np.errstate(divide="ignore"):
        idf = np.where(df > 0, np.log(m / np.maximum(df, 1)), 0.0)
    out = b1.copy()
    out.data = b1.data * idf[b1.indices]
    return out


# ---- A2: b2ToB (spectral-clustering; evidence Classify/Classify.hs:44-51) ------
def b2_to_b(b2: sp.csr_matrix) -> sp.csr_matrix:
    """L2-normalise every row (done once at the root; Cluster.hs:150 passes normFlag False)."""
    b2 = sp.csr_matrix(b2, dtype=np.float64)
    norms = np.sqrt(np.asarray(b2.multiply(b2).sum(axis=1)).ravel())
    if np.any(norms == 0):
        raise ValueError("zero-norm row: the reference yields NaN / hangs (LoadMatrix.hs:189-193)")
    out = b2.copy()
    out.data = b2.data / np.repeat(norms, np.diff(b2.indptr))
    return out


# ---- A3: bToD / bdToC ------------------------------------------------------------
def b_to_d(b: sp.csr_matrix) -> np.ndarray:
    """d = |B| (|B|^T 1): row sums of the implicit cosine-similarity matrix (self included)."""
    a = abs(b) if DEGREE_USES_ABS else b
    t = np.asarray(a.sum(axis=0)).ravel()
    return a @ t


def bd_to_c(b: sp.csr_matrix, d: np.ndarray) -> sp.csr_matrix:
    return sp.diags(1.0 / np.sqrt(d)) @ b


# ---- A4/A5: secondLeft -> sparseSvd(2) -> u2 (twin idiom Preprocess.hs:512-520) --
def second_left(c: sp.csr_matrix, d: np.ndarray | None = None):
    """Second left singular vector of C (exact).  Returns (u2, sigma2^2, gap).

    sigma_1 = 1 with u_1 ~ D^{1/2} 1 always; u2 is the top eigenvector of
    C C^T restricted to the complement of u_1.
    """
    m = c.shape[0]
    if m <= DENSE_MAX_ROWS:
        g = (c @ c.T).toarray()
        if d is not None:
            # deflate the known trivial pair (sigma_1 = 1, u_1 = D^{1/2} 1 / ||.||) so that u2 stays
            # well defined when sigma_2 = 1 too (disconnected similarity graph)
            u1 = np.sqrt(d) / math.sqrt(d.sum())
            gu = g @ u1
            g = g - np.outer(u1, gu) - np.outer(gu, u1) + (u1 @ gu) * np.outer(u1, u1)
            w, v = np.linalg.eigh(g)
            gap = w[-1] - (w[-2] if m >= 3 else 0.0)
            # the deflated direction shows up as an extra ~0 eigenvalue; it never is the top one unless theta2 ~ 0
            return v[:, -1], float(w[-1]), float(gap)
        w, v = np.linalg.eigh(g)
        u2 = v[:, -2]
        th2 = w[-2]
        gap = w[-2] - (w[-3] if m >= 3 else 0.0)
        return u2, float(th2), float(gap)
    # large node: ARPACK on the deflated operator
    u1 = np.sqrt(d) / math.sqrt(d.sum())
    ct = c.T.tocsr()

    def mv(x):
        x = x - u1 * (u1 @ x)
        y = c @ (ct @ x)
        return y - u1 * (u1 @ y)

    op = spla.LinearOperator((m, m), matvec=mv, dtype=np.float64)
    rng = np.random.default_rng(12345)
    v0 = rng.standard_normal(m)
    w, v = spla.eigsh(op, k=2, which="LA", tol=1e-14, v0=v0, ncv=min(m - 1, 64), maxiter=200000)
    order = np.argsort(w)
    return v[:, order[-1]], float(w[order[-1]]), float(w[order[-1]] - w[order[-2]])


def second_left_signed(c: sp.csr_matrix):
    """Second left singular vector of C with nothing deflated (signed input).  Returns (u2, sigma2^2, gap to sigma3^2)."""
    m = c.shape[0]
    if m <= DENSE_MAX_ROWS:
        w, v = np.linalg.eigh((c @ c.T).toarray())
        return v[:, -2], float(w[-2]), float(w[-2] - (w[-3] if m >= 3 else 0.0))
    u, sv, _ = spla.svds(c, k=3, tol=1e-14, v0=np.random.default_rng(12345).standard_normal(min(c.shape)))
    order = np.argsort(sv)
    return u[:, order[-2]], float(sv[order[-2]] ** 2), float(sv[order[-2]] ** 2 - sv[order[-3]] ** 2)


# ---- A6: spectralCluster ----------------------------------------------------------
def spectral_cluster(b: sp.csr_matrix):
    """labels (0/1) from the sign of u2.  Returns (labels, u2, theta2, gap, d)."""
    m = b.shape[0]
    if m < 1:
        return np.zeros(0, dtype=np.int8), np.zeros(0), 0.0, 0.0, np.zeros(0)
    d = b_to_d(b)
    if m == 1:
        return np.zeros(1, dtype=np.int8), np.zeros(1), 0.0, 0.0, d
    c = bd_to_c(b, d)
    if b.nnz and b.data.min() < 0:
        # signed B (a centred / PCA-reduced matrix): d comes from |B| but C = D^{-1/2} B keeps the signs, so D^{1/2} 1 is not a
        # singular vector any more -- `secondLeft 2 1` is then literally the second left singular vector of C
        u2, th2, gap = second_left_signed(c)
    else:
        u2, th2, gap = second_left(c, d)
    if SIGN_RULE_GE_ZERO_IS_ONE:
        labels = (u2 >= 0).astype(np.int8)
    else:
        labels = (u2 > 0).astype(np.int8)
    return labels, u2, th2, gap, d


# ---- A7: getBModularity (modularity 0.2.1.1, Math.Modularity.Sparse) ---------------
def get_b_modularity(labels: np.ndarray, b: sp.csr_matrix) -> float:
    """Newman-Girvan Q of the 2-way split on A = B B^T - I (self loops removed)."""
    m = b.shape[0]
    ones = np.ones(m)
    t = b.T @ ones
    d = b @ t
    big_l = d.sum() - m
    if big_l == 0:
        return 0.0
    q = 0.0
    for cval in (0, 1):
        ind = (labels == cval).astype(np.float64)
        tc = b.T @ ind
        o_c = tc @ tc - ind.sum()
        l_c = ind @ d - ind.sum()
        q += o_c / big_l - (l_c / big_l) ** 2
    return float(q)


# ---- A6 (KMeansGroup): spectralClusterK (spectral-clustering; flag help text Options.hs:172-173) -------------------------
def kmeans2_labels(pts: np.ndarray) -> np.ndarray:
    """2-means on the rows of `pts` (n x e).  The reference's k-means is un-vendored and its "starting points are
    arbitrary" (Options.hs:172), so the convention is fixed here (and in csrc/tmc_variants.cuh): Lloyd from the rows with the
    smallest / largest first coordinate, ties to cluster 1, at most 100 rounds, label 1 = the cluster started from the largest."""
    pts = np.asarray(pts, dtype=np.float64).reshape(len(pts), -1)
    n = pts.shape[0]
    lo, hi = int(np.argmin(pts[:, 0])), int(np.argmax(pts[:, 0]))
    if pts[lo, 0] == pts[hi, 0]:
        return np.ones(n, dtype=np.int8)
    c0, c1 = pts[lo].copy(), pts[hi].copy()
    lab = np.zeros(n, dtype=np.int8)
    for rnd in range(100):
        d0 = ((pts - c0) ** 2).sum(axis=1)
        d1 = ((pts - c1) ** 2).sum(axis=1)
        new = (d1 <= d0).astype(np.int8)
        changed = rnd == 0 or bool(np.any(new != lab))
        lab = new
        n1 = int(lab.sum())
        if not changed or n1 == 0 or n1 == n:
            break
        c0, c1 = pts[lab == 0].mean(axis=0), pts[lab == 1].mean(axis=0)
    return lab


def kmeans_num_eigen(num_eigen: int, m: int, n_cols: int) -> int:
    """Eigenvectors a node of m cells uses: min(num_eigen, m/4 - 1, n_cols/4 - 1), at least 1 (the engine's rule)."""
    e = max(1, min(int(num_eigen), m // 4 - 1))
    if num_eigen > 1:
        e = max(1, min(e, n_cols // 4 - 1))
    return e


def spectral_cluster_k(b: sp.csr_matrix, num_eigen: int = 1):
    """labels from 2-means on the rows of singular vectors 2..e+1 of C (`spectral 2 e`).  Returns (labels, coords, theta2)."""
    m = b.shape[0]
    d = b_to_d(b)
    c = bd_to_c(b, d)
    e = kmeans_num_eigen(num_eigen, m, b.shape[1])
    g = (c @ c.T).toarray()
    w, v = np.linalg.eigh(g)
    coords = v[:, ::-1][:, 1:e + 1]
    return kmeans2_labels(coords), coords, float(w[-2])


# ---- A10: the `runs` argument (--numRuns, Options.hs:174; _pVal, MakeTree/Types.hs:124-130) ---------------------------------
_M64 = (1 << 64) - 1


class SplitMix64:
    def __init__(self, seed):
        self.s = seed & _M64

    def next(self):
        self.s = (self.s + 0x9E3779B97F4A7C15) & _M64
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M64
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M64
        return z ^ (z >> 31)


def shuffle_labels(lab: np.ndarray, rng: SplitMix64):
    """In-place Fisher-Yates, i = n-1 .. 1, j = next() mod (i+1)."""
    for i in range(len(lab) - 1, 0, -1):
        j = rng.next() % (i + 1)
        lab[i], lab[j] = lab[j], lab[i]


DEFAULT_SEED = 0x2545F4914F6CDD1D       # tmc_opts_default


def permutation_p_value(labels: np.ndarray, b: sp.csr_matrix, rows: np.ndarray, num_runs: int, seed: int = DEFAULT_SEED) -> float:
    """Empirical p of one split: the node's labels (ascending row order) are shuffled num_runs times and Q recomputed;
    p = #{Q_shuffled >= Q_observed - 1e-12} / num_runs.  The stream is seeded from (seed, smallest row id, node size) so
    that it does not depend on how a tree numbers the node ([UPSTREAM-RECALL]: the upstream body is un-vendored)."""
    q_obs = get_b_modularity(labels, b)
    rng = SplitMix64(seed ^ ((int(rows[0]) * 0x9E3779B97F4A7C15 + len(rows)) & _M64))
    lab = np.array(labels, dtype=np.int8)
    ge = 0
    for _ in range(num_runs):
        shuffle_labels(lab, rng)
        if get_b_modularity(lab, b) >= q_obs - 1e-12:
            ge += 1
    return ge / num_runs


# ---- N3: --svd / --pca pre-reductions (src/TooManyCells/Matrix/Preprocess.hs:263-271, 543-556, 573-585) ----------------------
def center_scale_sparse(vec_rows: sp.csr_matrix) -> sp.csr_matrix:
    """centerScaleSparseCell on every row (Preprocess.hs:263-271): the STORED entries become (x - mu) / sigma with mu, sigma the
    mean and the sample standard deviation (n - 1) of the whole dense row, zeros stay zeros (`fmap` over a sparse vector);
    sigma = 0 gives 0.  [UPSTREAM-RECALL: `S.toVector` is the dense vector; `stdDev` of `statistics` is the n - 1 estimator.]"""
    a = sp.csr_matrix(vec_rows, dtype=np.float64)
    n = a.shape[1]
    s1 = np.asarray(a.sum(axis=1)).ravel()
    s2 = np.asarray(a.multiply(a).sum(axis=1)).ravel()
    mu = s1 / n
    var = (s2 - n * mu * mu) / max(n - 1, 1)
    sigma = np.sqrt(np.maximum(var, 0.0))
    out = a.copy()
    rows = np.repeat(np.arange(a.shape[0]), np.diff(a.indptr))
    with np.errstate(divide="ignore", invalid="ignore"):
        out.data = np.where(sigma[rows] == 0, 0.0, (a.data - mu[rows]) / sigma[rows])
    return out


def svd_sparse_mat(mat, n_dim):
    """svdSparseMat (Preprocess.hs:573-585): `secondLeft 1 N` of the cell-standardised matrix = its first N left singular vectors,
    one N-vector per cell.  Exact (LAPACK).  Returns (U [cells x N], sigma [N])."""
    x = center_scale_sparse(mat).toarray()
    u, sv, _ = np.linalg.svd(x, full_matrices=False)
    return u[:, :n_dim], sv[:n_dim]


def pca_sparse_mat(mat, n_dim):
    """pcaSparseMat (Preprocess.hs:543-556): features standardised (`centerScaleSparseCell` on the columns), then the matrix times
    its first N right singular vectors = U_N diag(sigma_N).  Returns the cells x N score matrix and sigma."""
    scaled = center_scale_sparse(sp.csr_matrix(mat).T.tocsr()).T.tocsr()
    x = scaled.toarray()
    u, sv, vt = np.linalg.svd(x, full_matrices=False)
    return x @ vt[:n_dim].T, sv[:n_dim]


# ---- A0: hierarchicalSpectralCluster (call site Cluster.hs:147-156,169) ----------
@dataclass
class OracleTree:
    """Flat pre-order tree, children order [label 0, label 1] (A9)."""
    left: list = field(default_factory=list)
    right: list = field(default_factory=list)
    q: list = field(default_factory=list)          # Q of every vertex (leaf Q is discarded by the host)
    theta: list = field(default_factory=list)      # sigma_2^2 of C at the node
    gap: list = field(default_factory=list)        # theta2 - theta3 (conditioning of u2)
    margin: list = field(default_factory=list)     # min |u2_i| / max |u2_i|  (sign margin)
    items: list = field(default_factory=list)      # ascending original row ids of every vertex
    u2: list = field(default_factory=list)         # optional per-node u2 (only if keep_u2)
    p_val: list = field(default_factory=list)      # permutation p-value of the vertex's split (None without num_runs / for leaves)

    @property
    def n_nodes(self):
        return len(self.left)

    def leaves(self):
        return [i for i in range(self.n_nodes) if self.left[i] < 0]

    def leaf_sets(self):
        return sorted(tuple(int(x) for x in self.items[i]) for i in self.leaves())


def hierarchical_spectral_cluster(b2, min_modularity=0.0, min_size=1, normalize=False,
                                  keep_u2=False, max_nodes=None, eigen_group="SignGroup", num_eigen=1,
                                  num_runs=0, seed=DEFAULT_SEED) -> OracleTree:
    """Divisive recursion `go items B` (Appendix B of SURVEY.md).

    `b2` is what `hSpecClust` hands over as `Left mat`: already tf-idf scaled by
    `loadAllSSM` (LoadMatrix.hs:218) unless normalize=True asks for idf here.
    Rows are L2-normalised once (getB False = b2ToB . B2); children take rows of
    the already-normalised B.
    """
    b2 = sp.csr_matrix(b2, dtype=np.float64)
    if normalize:
        b2 = b1_to_b2(b2)
    b0 = b2_to_b(b2)
    tree = OracleTree()

    def go(rows: np.ndarray) -> int:
        idx = tree.n_nodes
        for lst in (tree.left, tree.right):
            lst.append(-1)
        tree.items.append(rows)
        tree.p_val.append(None)
        b = b0[rows]
        m = rows.size
        if m <= 1:
            tree.q.append(0.0); tree.theta.append(0.0); tree.gap.append(0.0)
            tree.margin.append(0.0); tree.u2.append(None)
            return idx
        if eigen_group == "KMeansGroup":
            labels, coords, th2 = spectral_cluster_k(b, num_eigen)
            u2, gap = coords[:, 0], 0.0
        else:
            labels, u2, th2, gap, _ = spectral_cluster(b)
        q = get_b_modularity(labels, b)
        tree.q.append(q); tree.theta.append(th2); tree.gap.append(gap)
        au = np.abs(u2)
        tree.margin.append(float(au.min() / au.max()) if au.max() > 0 else 0.0)
        tree.u2.append(u2 if keep_u2 else None)
        li = rows[labels == 0]
        ri = rows[labels == 1]
        ok = (q > min_modularity) if STRICT_Q_GT_MINMOD else (q >= min_modularity)
        if ok and li.size >= max(1, min_size) and ri.size >= max(1, min_size) and m > 1:
            if max_nodes is not None and tree.n_nodes >= max_nodes:
                return idx
            if num_runs:
                tree.p_val[idx] = permutation_p_value(labels, b, rows, num_runs, seed)
            tree.left[idx] = go(li)
            tree.right[idx] = go(ri)
        return idx

    import sys
    old = sys.getrecursionlimit()
    sys.setrecursionlimit(max(old, 100000))
    try:
        go(np.arange(b0.shape[0], dtype=np.int64))
    finally:
        sys.setrecursionlimit(old)
    return tree


# ---- helpers for the parity checker -------------------------------------------------
def canonical_tree(left, right, items_of):
    """Canonical nested form of a tree modulo child swap: frozenset of leaf item tuples per vertex."""
    def rec(i):
        if left[i] < 0:
            return ("L", tuple(sorted(int(x) for x in items_of(i))))
        a, b = rec(left[i]), rec(right[i])
        return ("N",) + tuple(sorted((a, b)))
    import sys
    old = sys.getrecursionlimit()
    sys.setrecursionlimit(max(old, 100000))
    try:
        return rec(0)
    finally:
        sys.setrecursionlimit(old)
