"""CPU mirror of the arithmetic of the direct small-node solver (too-many-cells_b200/csrc/tmc_small.cuh, k_small_subtree): the same
formulas transcribed serially in numpy -- Gram matrix G, d = G 1, M = D^{-1/2} G D^{-1/2} - u1 u1^T, Householder tridiagonalisation
with the reflectors kept for the back-transformation, the PRODUCT's Sturm multisection + twisted factorisation (tmc_tridiag.cuh,
compiled for the host), sign rule, Q from G -- checked against the exact-math oracle (LAPACK eigh on the same node) on a few
thousand random sub-nodes, and probed with degenerate inputs (duplicate cells, mutually orthogonal cells) for NaN / inf.
The kernel itself is tested on the GPU (tests/test_gpu_parity.py::test_direct_solve_of_small_nodes_same_tree); this file pins
the METHOD on CPU, where the driver runs it every round."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import tmc_oracle as O                                  # noqa: E402
from too_many_cells_b200 import synth                   # noqa: E402

P = C.POINTER(C.c_double)


@pytest.fixture(scope="module")
def tridiag_lib(tmp_path_factory):
    out = tmp_path_factory.mktemp("th2") / "libth.so"
    src = os.path.join(ROOT, "tests", "helpers", "tridiag_host.cpp")
    subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-x", "c++", src, "-o", str(out)], check=True)
    lib = C.CDLL(str(out))
    lib.th_top_eig.restype = C.c_double
    lib.th_top_eig.argtypes = [P, P, C.c_int, C.c_double]
    lib.th_vector.argtypes = [P, P, C.c_int, C.c_double, P, P, P]
    return lib


def mirror_split(g, lib):
    """One sub-node of k_small_subtree on its Gram matrix g (t x t): returns (labels, theta, q, u2)."""
    t = g.shape[0]
    d = g.sum(axis=1)
    sq = np.sqrt(d); sumd = d.sum()
    m = g / np.outer(sq, sq) - np.outer(sq, sq) / sumd
    a = np.zeros(t); b = np.zeros(t); tau = np.zeros(max(t - 2, 0)); refl = []
    for kk in range(t - 2):
        x = m[kk + 1:, kk].copy()
        nrm2 = float(x @ x); x0 = x[0]
        tail2 = max(nrm2 - x0 * x0, 0.0)
        if not tail2 > 1e-300:
            a[kk] = m[kk, kk]; b[kk + 1] = x0; refl.append(None); continue
        alpha = -np.sqrt(nrm2) if x0 >= 0 else np.sqrt(nrm2)
        v = x.copy(); v[0] -= alpha
        vn2 = nrm2 - x0 * x0 + (x0 - alpha) ** 2
        tk = 2.0 / vn2
        p = tk * (m[kk + 1:, kk + 1:] @ v)
        kc = 0.5 * tk * float(p @ v)
        w = p - kc * v
        m[kk + 1:, kk + 1:] -= np.outer(v, w) + np.outer(w, v)
        a[kk] = m[kk, kk]; b[kk + 1] = alpha; tau[kk] = tk; refl.append(v)
    if t >= 2:
        a[t - 2] = m[t - 2, t - 2]; b[t - 1] = m[t - 1, t - 2]
    a[t - 1] = m[t - 1, t - 1]; b[0] = 0.0
    theta = lib.th_top_eig(a.ctypes.data_as(P), b.ctypes.data_as(P), t, -1e300)
    z, dp, dm = np.zeros(t), np.zeros(t), np.zeros(t)
    lib.th_vector(a.ctypes.data_as(P), b.ctypes.data_as(P), t, theta, dp.ctypes.data_as(P), dm.ctypes.data_as(P), z.ctypes.data_as(P))
    for kk in range(t - 3, -1, -1):
        v = refl[kk]
        if v is None:
            continue
        z[kk + 1:] -= tau[kk] * float(v @ z[kk + 1:]) * v
    lab = z >= 0.0
    n1 = float(lab.sum()); n0 = t - n1
    sd1 = float(d[lab].sum()); t1 = float(g[np.ix_(lab, lab)].sum())
    big_l = sumd - t
    q = 0.0
    if big_l != 0.0:
        o1 = t1 - n1; o0 = sumd - 2.0 * sd1 + t1 - n0
        l1 = sd1 - n1; l0 = (sumd - sd1) - n0
        q = (o0 / big_l - (l0 / big_l) ** 2) + (o1 / big_l - (l1 / big_l) ** 2)
    return lab, theta, q, z


def test_mirror_of_the_direct_solver_matches_the_oracle(tridiag_lib):
    ip, ix, dv, _ = synth.synth_config("tiny")
    n, gcols = synth.CONFIGS["tiny"][:2]
    bm = O.b2_to_b(O.b1_to_b2(sp.csr_matrix((dv, ix, ip), shape=(n, gcols))))
    rng = np.random.default_rng(3)
    checked = 0
    for trial in range(1500):
        t = int(rng.integers(2, 65))
        # half of the nodes from one planted programme's neighbourhood (structure), half random cells (noise)
        rows = np.sort(rng.choice(n, size=t, replace=False)) if trial % 2 else np.sort(rng.choice(np.arange(n)[:: int(rng.integers(1, 5))][: 4 * t], size=t, replace=False))
        b = bm[rows]
        g = (b @ b.T).toarray()
        lab, theta, q, u2 = mirror_split(g.copy(), tridiag_lib)
        olab, ou2, oth, ogap, od = O.spectral_cluster(b)
        assert np.all(np.isfinite(u2)) and np.isfinite(q)
        assert abs(theta - oth) <= 1e-11 * max(1.0, oth), (t, theta, oth)
        margin = np.min(np.abs(ou2))
        if ogap > 1e-6 and margin > 1e-7:            # away from degenerate pairs and from entries decided by rounding
            same = np.array_equal(lab, olab.astype(bool)) or np.array_equal(~lab, olab.astype(bool))
            assert same, (t, ogap, margin)
            assert abs(q - O.get_b_modularity(olab, b)) <= 1e-10
            assert min(np.linalg.norm(u2 - ou2), np.linalg.norm(u2 + ou2)) <= 1e-7 / min(1.0, ogap * 1e3)
            checked += 1
    assert checked > 1300


def test_mirror_survives_degenerate_nodes(tridiag_lib):
    """Duplicate cells (rank-deficient G), mutually orthogonal cells (G = I), two disconnected groups: nothing non-finite, a
    split with both sides non-empty where one exists."""
    rng = np.random.default_rng(0)
    base = np.abs(rng.standard_normal((6, 40))); base /= np.linalg.norm(base, axis=1, keepdims=True)
    cases = {
        "duplicates": base[[0, 0, 0, 1, 1, 2]],
        "all identical": base[[3, 3, 3, 3]],
        "orthogonal": np.eye(5, 40),
        "two groups": np.vstack([np.hstack([base[:3, :20], np.zeros((3, 20))]), np.hstack([np.zeros((3, 20)), base[3:, 20:]])]),
        "pair": base[:2],
    }
    for name, b in cases.items():
        b = b / np.linalg.norm(b, axis=1, keepdims=True)
        g = b @ b.T
        lab, theta, q, u2 = mirror_split(g.copy(), tridiag_lib)
        assert np.all(np.isfinite(u2)) and np.isfinite(q) and np.isfinite(theta), name
        assert -1e-12 <= theta <= 1.0 + 1e-12, (name, theta)
        if name == "two groups":
            assert lab[:3].all() != lab[3:].all() and len(set(lab[:3])) == 1 and len(set(lab[3:])) == 1      # the groups, cleanly
            assert abs(theta - 1.0) < 1e-12 and q > 0.4
        if name == "pair":
            assert lab[0] != lab[1]
