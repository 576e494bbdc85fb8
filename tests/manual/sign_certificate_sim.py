"""CPU simulation behind DESIGN.md section 7 item 2: how much Lanczos work a *sign-certified* stop would save.

The tree depends on u2 only through the signs of its entries (labels; Q is a function of the labels), yet the engine iterates
every node to ||M x - theta x|| <= 1e-10 * theta.  For a Ritz pair (theta, x) with residual r and a gap delta between theta2 and the rest
of the spectrum, Davis-Kahan gives ||x - u2|| <= ~ r / delta, so every label is final once  min_i |x_i| > SAFE * r / delta.  This script
replays Lanczos (full re-orthogonalisation, deflated, like the engine) on every node of the oracle's tree and reports, weighted by
rows x steps (= SpMV work):

  tol1e-10       the engine's current stop (= 1.0)            tol1e-6      SVDLIBC's kappa
  sign_stable    first step after which no label changes any more (hindsight: the lower bound)
  certified      the rule above with delta = theta_1(T) - theta_2(T) - r_2 from the Ritz values, r_2 < delta / 4, nodes of >= MINM cells
  cert_wrong     nodes where the certified stop would have produced a label different from the converged one (must be 0)

Measured (round 1, this container):   default tree   c1 certified 0.64 / sign_stable 0.39 / tol1e-6 0.72, cert_wrong 0
                                                     `small` 0.67 / 0.39 / 0.73, 0 wrong
                                      saturated      `tiny`  0.77 / 0.41 / 0.77, 0 wrong;  `small` 0.78 / 0.44 / 0.77, 0 wrong
                                      other shapes   binarised 3000 x 8000 (scATAC-like) 0.66 / 0.36 / 0.72, 0 wrong;  deeper planted
                                                     hierarchy 4000 x 6000, depth 7 (377 nodes) 0.73 / 0.37 / 0.74, 0 wrong
  Without the size floor (MINM=0) the rule fails on a handful of 6-12-cell nodes whose theta2 and theta3 are nearly degenerate
  (the Krylov space is then almost the whole space and the second Ritz value still far from theta3): hence MINM = 64.

    python tests/manual/sign_certificate_sim.py [tiny|small|c1]      env: MINMOD (min_modularity), SAFE (default 4), MINM (default 64)
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))      # checker scripts live under tests/: only tests may import the oracle
import numpy as np, scipy.sparse as sp
import tmc_oracle as O
from too_many_cells_b200 import synth

name = sys.argv[1] if len(sys.argv) > 1 else "small"
SAFE, MINM = float(os.environ.get("SAFE", "4")), int(os.environ.get("MINM", "64"))
ip, ix, dv, _ = synth.synth_config(name); n, g = synth.CONFIGS[name][:2]
a = sp.csr_matrix((dv, ix, ip), shape=(n, g))
b0 = O.b2_to_b(O.b1_to_b2(a))
tr = O.hierarchical_spectral_cluster(a, normalize=True, min_modularity=float(os.environ.get("MINMOD", "0")))
rng = np.random.default_rng(0)
tot = dict(certified=0, cert_wrong=0, tol10=0, tol6=0, sign_stable=0)
nodes = 0
for u in range(tr.n_nodes):
    rows = tr.items[u]; m = len(rows)
    if m < 3:
        continue
    b = b0[rows]; d = O.b_to_d(b); c = O.bd_to_c(b, d).tocsr(); ct = c.T.tocsr()
    u1 = np.sqrt(d) / np.sqrt(d.sum())
    q = rng.standard_normal(m); q -= u1 * (u1 @ q); q /= np.linalg.norm(q)
    V, al, be, labels, res, cert = [u1, q], [], [0.0], [], [], None
    for j in range(1, min(m - 1, 300) + 1):
        w = c @ (ct @ V[j]); alpha = 0.0
        for _ in range(2):
            h = np.array([V[l] @ w for l in range(j + 1)])
            for l in range(j + 1):
                w -= h[l] * V[l]
            alpha += h[j]
        al.append(alpha); beta = np.linalg.norm(w)
        ev, evec = np.linalg.eigh(np.diag(al) + np.diag(be[1:], 1) + np.diag(be[1:], -1))
        z, th = evec[:, -1], ev[-1]
        x = sum(z[l] * V[l + 1] for l in range(j))
        labels.append(x >= 0); r = beta * abs(z[-1]); res.append(r / max(th, 1e-300))
        if cert is None and j >= 3 and m >= MINM:
            r2 = beta * abs(evec[-1, -2]); gap = ev[-1] - ev[-2] - r2
            if gap > 0 and r2 < 0.25 * (ev[-1] - ev[-2]) and np.min(np.abs(x)) / np.linalg.norm(x) > SAFE * r / gap:
                cert = j
        if res[-1] <= 1e-10 or beta < 1e-13:
            break
        be.append(beta); V.append(w / beta)
    final, J = labels[-1], len(labels)
    same = lambda p: np.array_equal(p, final) or np.array_equal(p, ~final)
    k = J
    while k > 1 and same(labels[k - 2]):
        k -= 1
    cj = cert if cert is not None else J
    tot["certified"] += cj * m; tot["cert_wrong"] += 0 if same(labels[cj - 1]) else 1
    tot["tol10"] += J * m; tot["tol6"] += next((i + 1 for i, x_ in enumerate(res) if x_ <= 1e-6), J) * m; tot["sign_stable"] += k * m
    nodes += 1
print(name, "nodes", nodes, {k: (v if k == "cert_wrong" else round(v / tot["tol10"], 3)) for k, v in tot.items()})
