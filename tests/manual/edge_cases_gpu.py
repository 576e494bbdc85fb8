"""Edge cases written after round 1's GPU minutes were spent: run this once on a B200 at the start of round 2
(`python tests/manual/edge_cases_gpu.py`), then promote the checks that hold into tests/test_gpu_parity.py /
tests/test_gpu_variants.py.  Every check prints its outcome; the script exits 1 if any expectation fails (nothing here may hang
or crash the process — that is the point)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, scipy.sparse as sp
import tmc_oracle as O
import too_many_cells_b200 as T
from too_many_cells_b200 import _lib

fails = []


def expect_error(name, code, fn):
    try:
        fn()
    except T.TmcError as e:
        ok = e.code == code
        print(f"{'ok  ' if ok else 'FAIL'} {name}: TmcError {e.code} ({e.msg})")
        if not ok:
            fails.append(name)
        return
    print(f"FAIL {name}: no error raised")
    fails.append(name)


def expect(name, cond, info=""):
    print(f"{'ok  ' if cond else 'FAIL'} {name} {info}")
    if not cond:
        fails.append(name)


rng = np.random.default_rng(0)
# -- empty and degenerate shapes
expect_error("0 x 5 matrix", _lib.ERR_INVALID, lambda: T.hierarchical_spectral_cluster(sp.csr_matrix((0, 5))))
expect_error("3 x 0 matrix (every row is a zero row)", _lib.ERR_ZERO_ROW, lambda: T.hierarchical_spectral_cluster(sp.csr_matrix((3, 0))))
expect_error("dense with an all-zero row", _lib.ERR_ZERO_ROW, lambda: T.hierarchical_spectral_cluster(np.array([[1.0, 0.0], [0.0, 0.0]])))
expect_error("dense 0 x 0", _lib.ERR_INVALID, lambda: T.hierarchical_spectral_cluster(np.zeros((0, 0))))
# -- malformed CSR
a = sp.random(40, 30, 0.3, format="csr", random_state=1); a.data[:] = 1.0 + rng.integers(0, 4, a.nnz)
a = a[np.asarray(a.sum(axis=1)).ravel() > 0]
bad = (a.indptr.copy(), a.indices.copy(), a.data.copy(), a.shape)
bad[1][3] = 30
expect_error("column index == n_cols", _lib.ERR_INVALID, lambda: T.hierarchical_spectral_cluster(bad))
bad = (a.indptr.copy(), a.indices.copy(), a.data.copy(), a.shape)
bad[0][-1] -= 1
expect_error("row_ptr does not end at nnz", _lib.ERR_INVALID, lambda: T.hierarchical_spectral_cluster(bad))
# -- unsorted columns inside rows are allowed (include/tmc.h): same tree as the sorted matrix
ip, ix, dv = a.indptr.copy(), a.indices.copy(), a.data.copy()
for i in range(a.shape[0]):
    p = rng.permutation(ip[i + 1] - ip[i]) + ip[i]
    ix[ip[i]:ip[i + 1]], dv[ip[i]:ip[i + 1]] = ix[p], dv[p]
t_sorted = T.hierarchical_spectral_cluster(a, normalize=True)
t_shuf = T.hierarchical_spectral_cluster((ip, ix, dv, a.shape), normalize=True)
expect("unsorted columns give the sorted matrix's tree", t_sorted.leaf_sets() == t_shuf.leaf_sets())
# -- ragged rows: one-entry cells next to a cell that has every feature
# (seed 101 was picked on the CPU so that no node has a degenerate sigma_2 or an exactly-zero u2 entry: min gap 1.1e-2,
#  min sign margin 3.3e-3 in the oracle — otherwise any two correct solvers may disagree)
rr = np.random.default_rng(101)
rows_, cols_, vals_ = [], [], []
for i in range(59):
    for j in rr.choice(12, size=1 + (i % 3), replace=False) + (0 if i < 30 else 8):
        rows_.append(i); cols_.append(int(j)); vals_.append(float(rr.integers(1, 6)))
for j in range(20):
    rows_.append(59); cols_.append(j); vals_.append(1.0)
r = sp.csr_matrix((vals_, (rows_, cols_)), shape=(60, 20))
tr = T.hierarchical_spectral_cluster(r, normalize=False)
ot = O.hierarchical_spectral_cluster(r, normalize=False)
expect("ragged rows vs oracle", tr.leaf_sets() == ot.leaf_sets(), f"nodes {tr.n_nodes} / {ot.n_nodes}")
# -- variants: invalid flags, mismatched tree
expect_error("eigen_group = 2", _lib.ERR_INVALID, lambda: T.hierarchical_spectral_cluster(a, eigen_group=2))
expect_error("num_runs < 0", _lib.ERR_INVALID, lambda: T.hierarchical_spectral_cluster(a, num_runs=-1))
expect_error("KMeansGroup num_eigen = 0", _lib.ERR_INVALID, lambda: T.hierarchical_spectral_cluster(a, eigen_group="KMeansGroup", num_eigen=0))
db = T.DeviceB.from_host(a[:20], normalize=False)
expect_error("permutation test with a tree of another matrix", _lib.ERR_INVALID, lambda: T.permutation_test(db, t_sorted, 3))
db.free()
# -- KMeansGroup with more eigenvectors than a small matrix can give, binary matrix, tiny matrices
tk = T.hierarchical_spectral_cluster(a[:9], eigen_group="KMeansGroup", num_eigen=50, normalize=True)
ok_ = O.hierarchical_spectral_cluster(a[:9], eigen_group="KMeansGroup", num_eigen=50, normalize=True)
expect("KMeansGroup, num_eigen far above the node size", tk.leaf_sets() == ok_.leaf_sets(), f"nodes {tk.n_nodes} / {ok_.n_nodes}")
b = sp.csr_matrix((rng.random((80, 40)) < 0.3).astype(np.float64)); b = b[np.asarray(b.sum(axis=1)).ravel() > 0]
tb = T.hierarchical_spectral_cluster(b, eigen_group="KMeansGroup", num_eigen=2, num_runs=4)
ob = O.hierarchical_spectral_cluster(b, eigen_group="KMeansGroup", num_eigen=2, num_runs=4)
expect("KMeansGroup + numRuns on a binary matrix", tb.leaf_sets() == ob.leaf_sets(), f"nodes {tb.n_nodes} / {ob.n_nodes}")
# -- --matrix-output -> --matrix-path round trip through the GPU MatrixMarket parser, with values in every notation repr() produces
with __import__("tempfile").TemporaryDirectory() as td:
    d = np.zeros((6, 5)); d[0, 0] = 1e-05; d[1, 2] = 0.1 + 0.2; d[2, 4] = 12345678.9; d[3, 1] = 3.0; d[4, 3] = 2.5e+17; d[5, 0] = 7.0
    m0 = sp.csr_matrix(d)
    T.write_matrix_like(os.path.join(td, "mo"), m0, [f"BC{i}-1" for i in range(6)], [f"G{j}" for j in range(5)])
    m1, bcs, fts = T.load_cellranger_data(os.path.join(td, "mo"))
    # the device parser is exactly rounded for <= 15 significant digits (every count matrix); 17-digit reprs may be 1 ulp off
    expect("write_matrix_like -> load_cellranger_data round trip", m1.shape == m0.shape and np.allclose(m1.toarray(), m0.toarray(), rtol=3e-16, atol=0)
           and bcs == [f"BC{i}-1" for i in range(6)] and fts == [f"G{j}" for j in range(5)],
           f"bit-exact: {np.array_equal(m1.toarray(), m0.toarray())}")
    mm, bb, ff = T.load_all([os.path.join(td, "mo"), os.path.join(td, "mo")])
    expect("two --matrix-path with equal features stack", mm.shape == (12, 5) and np.allclose(mm.toarray()[6:], m0.toarray(), rtol=3e-16, atol=0) and len(bb) == 12)
# -- the C++ host layer's variant overloads (tests/helpers/host_cpp_example.cpp, mode "variants")
import subprocess, tempfile
with tempfile.TemporaryDirectory() as td:
    pkg = os.path.join(ROOT, "too-many-cells_b200")
    exe = os.path.join(td, "host_cpp_example")
    subprocess.run(["g++", "-std=c++17", "-O1", os.path.join(ROOT, "tests", "helpers", "host_cpp_example.cpp"), "-o", exe, "-L", pkg, "-ltmc",
                    f"-Wl,-rpath,{pkg}"], check=True)
    with open(os.path.join(td, "m.txt"), "w") as f:
        f.write(f"{a.shape[0]} {a.shape[1]} {a.nnz}\n" + " ".join(map(str, a.indptr.tolist())) + "\n" + " ".join(map(str, a.indices.tolist())) + "\n"
                + " ".join(repr(float(x)) for x in a.data.tolist()) + "\n")
    out = subprocess.run([exe, os.path.join(td, "m.txt"), "variants"], capture_output=True, text=True)
    tkm = T.hierarchical_spectral_cluster(a, normalize=True, eigen_group="KMeansGroup", num_runs=4)
    tde = T.hierarchical_spectral_cluster(a, normalize=True)
    want = f"nodes {tkm.n_nodes} splits-with-p {(tkm.n_nodes - 1) // 2}\nnodes {tde.n_nodes} splits-with-p 0\n"
    expect("C++ host layer: KMeansGroup + runs, dense overload", out.returncode == 0 and out.stdout == want, repr(out.stdout) + out.stderr[-200:])
print("FAILED:" if fails else "all edge cases behave as expected", fails or "")
sys.exit(1 if fails else 0)
