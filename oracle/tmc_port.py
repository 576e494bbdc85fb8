"""ctypes wrapper of oracle/tmc_port.c (single-threaded C port of the reference's CPU path).
TEST INFRASTRUCTURE / CPU BASELINE ONLY — see the header of tmc_port.c."""
import ctypes as C
import os
import subprocess

import numpy as np
import scipy.sparse as sp

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libtmc_port.so")


class _Tree(C.Structure):
    _fields_ = [("n_nodes", C.c_int64), ("cap", C.c_int64), ("left", C.POINTER(C.c_int64)), ("right", C.POINTER(C.c_int64)),
                ("begin", C.POINTER(C.c_int64)), ("end", C.POINTER(C.c_int64)), ("q", C.POINTER(C.c_double)),
                ("theta", C.POINTER(C.c_double)), ("iters", C.POINTER(C.c_int32))]


_PROBE = """
import ctypes as C, numpy as np, sys
sys.path.insert(0, HERE)
import tmc_port as P, scipy.sparse as sp
P._checked = True
a = sp.random(40, 30, 0.3, format="csr", random_state=0); a.data[:] = 1.0 + np.arange(a.nnz) % 3
a = sp.vstack([a, sp.csr_matrix(np.ones((1, 30)))]).tocsr()
assert P.make_tree(a[np.asarray(a.sum(axis=1)).ravel() > 0], normalize=False).n_nodes >= 1
"""
_checked = False


def _runs_on_this_host():
    """The library is compiled with -march=native and travels between machines with the repository snapshot: make sure it does
    not die with SIGILL on this CPU (a tiny tree in a child process), so that a rebuild can fix it instead of the caller crashing."""
    import sys
    return subprocess.run([sys.executable, "-c", _PROBE.replace("HERE", repr(_HERE))], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL).returncode == 0


def _load():
    global _checked
    if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(os.path.join(_HERE, "tmc_port.c")):
        subprocess.run(["make", "-C", _HERE], check=True, stdout=subprocess.DEVNULL)
    if not _checked:
        if not _runs_on_this_host():
            subprocess.run(["make", "-B", "-C", _HERE], check=True, stdout=subprocess.DEVNULL)     # built for another CPU: rebuild here
            if not _runs_on_this_host():
                raise RuntimeError("oracle/libtmc_port.so does not run on this host even after a rebuild")
        _checked = True
    lib = C.CDLL(_SO)
    lib.tmc_port_make_tree.restype = C.c_int64
    lib.tmc_port_make_tree.argtypes = [C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_int,
                                       C.c_void_p, C.POINTER(_Tree), C.POINTER(C.c_double)]
    lib.tmc_port_tree_free.argtypes = [C.POINTER(_Tree)]
    lib.tmc_port_set_threads.argtypes = [C.c_int]
    lib.tmc_port_get_threads.restype = C.c_int
    return lib


class PortTree:
    def __init__(self, left, right, begin, end, q, theta, iters, perm, steps, work_nnz):
        self.left, self.right, self.begin, self.end, self.q, self.theta, self.iters = left, right, begin, end, q, theta, iters
        self.perm, self.steps, self.work_nnz = perm, steps, work_nnz

    @property
    def n_nodes(self):
        return int(self.left.size)

    def items(self, i):
        return self.perm[self.begin[i]:self.end[i]]

    def leaf_sets(self):
        return sorted(tuple(int(x) for x in self.items(i)) for i in range(self.n_nodes) if self.left[i] < 0)


def make_tree(mat, normalize=False, min_modularity=0.0, tol=1e-6, kmax=400, threads=1) -> PortTree:
    """The whole path on one CPU core (threads=1, like the reference) or with the mat-vecs and the re-orthogonalisation spread over
    `threads` cores.  tf-idf and row-L2 are done with numpy first (as loadAllSSM does before the engine)."""
    import tmc_oracle as O
    b2 = sp.csr_matrix(mat, dtype=np.float64)
    if normalize:
        b2 = O.b1_to_b2(b2)
    b = O.b2_to_b(b2)
    lib = _load()
    lib.tmc_port_set_threads(int(threads))
    m, n = b.shape
    ptr = np.ascontiguousarray(b.indptr, dtype=np.int64); col = np.ascontiguousarray(b.indices, dtype=np.int32)
    val = np.ascontiguousarray(b.data, dtype=np.float64)
    perm = np.empty(m, dtype=np.int64)
    t = _Tree(); work = C.c_double()
    steps = lib.tmc_port_make_tree(m, n, ptr.ctypes.data, col.ctypes.data, val.ctypes.data, float(min_modularity), float(tol), int(kmax),
                                   perm.ctypes.data, C.byref(t), C.byref(work))
    k = int(t.n_nodes)
    arr = lambda p, dt: np.ctypeslib.as_array(p, shape=(k,)).astype(dt, copy=True)
    out = PortTree(arr(t.left, np.int64), arr(t.right, np.int64), arr(t.begin, np.int64), arr(t.end, np.int64), arr(t.q, np.float64),
                   arr(t.theta, np.float64), arr(t.iters, np.int32), perm, int(steps), float(work.value))
    lib.tmc_port_tree_free(C.byref(t))
    return out
