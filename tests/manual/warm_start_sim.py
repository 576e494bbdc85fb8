"""CPU simulation: does a WARM START of a child's Lanczos from the parent's Krylov space save SpMV work?

The parent's Krylov space, built for u2, also holds rough copies of the next singular directions, and in a hierarchy those are
the children's own second vectors.  At the parent's stop step take the Ritz vectors of the 2nd and 3rd Ritz value of T_j (singular
directions 3 and 4 of C), bring them to the random-walk picture f = D^{-1/2} x (piecewise constant on clusters whatever the degrees)
and start the child from D_child^{1/2} f restricted to its rows instead of a random vector.  The stop rules are the engine's
(tolerance 1e-10, or the sign certificate for nodes of >= MINM cells); work = rows x steps summed over the nodes.

    python tests/manual/warm_start_sim.py [tiny|small|c1|deep]      env: MINMOD, SAFE (4), MINM (64), MIX (0.7), NV (Ritz vectors in the
    mix, 2), MODE (mix | w), ENTRY=1 (entrywise certificate through the purified Ritz vector), ADAPT=e (safety factor scaled by
    (|z_1| sqrt(m))^e: the bias of the start vector towards u2)

Measured (round 2, this container; ratio = work with warm starts / work with random starts, both under the sign certificate):
    small 0.88, deep (4000 x 6000, planted depth 7) 0.88: 0.70 - 0.78 on the planted levels, 0.92 - 0.95 on the structureless ones;
    NV = 1 .. 4 and the weighted mix change nothing (0.88 - 0.91); saturated trees 0.90 - 0.91; no wrong label in any of these runs
    (a few hundred certified nodes -- the failure the GPU runs found has a rate of about 1 in 50 000 small certified nodes:
    DESIGN.md section 3); ENTRY=1 saves another 2.5 %; ADAPT=1 / 0.5 leave 0.94 / 0.91 of the gain.
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, scipy.sparse as sp
import tmc_oracle as O
from too_many_cells_b200 import synth

name = sys.argv[1] if len(sys.argv) > 1 else "small"
SAFE, MINM, MIX = float(os.environ.get("SAFE", "4")), int(os.environ.get("MINM", "64")), float(os.environ.get("MIX", "0.7"))
if name == "deep":
    ip, ix, dv, _ = synth.synth_counts(4000, 6000, 300, 7, 5); n, g = 4000, 6000
else:
    ip, ix, dv, _ = synth.synth_config(name); n, g = synth.CONFIGS[name][:2]
a = sp.csr_matrix((dv, ix, ip), shape=(n, g))
b0 = O.b2_to_b(O.b1_to_b2(a))
tr = O.hierarchical_spectral_cluster(a, normalize=True, min_modularity=float(os.environ.get("MINMOD", "0")))
rng = np.random.default_rng(0)


def lanczos(c, ct, d, start):
    """returns (steps at the engine's stop, labels there, V, T eigen-decomposition at the stop, converged labels)"""
    m = len(d)
    u1 = np.sqrt(d) / np.sqrt(d.sum())
    q = start - u1 * (u1 @ start); q /= np.linalg.norm(q)
    V, al, be = [u1, q], [], [0.0]
    stop = None
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
        r = beta * abs(z[-1])
        if stop is None and j >= 3 and m >= MINM:
            r2 = beta * abs(evec[-1, -2]); gap = ev[-1] - ev[-2] - r2
            ok = gap > 0 and r2 < 0.25 * (ev[-1] - ev[-2])
            if os.environ.get("ADAPT"):
                # a start vector biased towards u2 by K = |z_1| sqrt(m) over the random level hides a near-degenerate partner
                # K times better: scale the safety factor by it
                kb = max(1.0, abs(z[0]) * np.sqrt(m)) ** float(os.environ.get("ADAPT"))
                ok = ok and np.min(np.abs(x)) / np.linalg.norm(x) > kb * SAFE * r / gap
            if ok and os.environ.get("ENTRY") and beta > 0:
                # entrywise bound through one more (free) product: x' = M x / theta = x + (beta z_j / theta) v_{j+1};
                # |u_i sigma^2/theta - x'_i| <= ||c_i|| eps / theta = eps / (theta sqrt(d_i))
                eps = SAFE * r / gap
                xp = x + (beta * z[-1] / th) * (w / beta)
                good = (np.abs(x) > eps) | ((np.abs(xp) * np.sqrt(d) * th > eps) & (np.sign(xp) == np.sign(x)))
                if good.all():
                    stop = (j, x >= 0, list(V), ev.copy(), evec.copy())
            elif ok and np.min(np.abs(x)) / np.linalg.norm(x) > SAFE * r / gap:
                stop = (j, x >= 0, list(V), ev.copy(), evec.copy())
        if r / max(th, 1e-300) <= 1e-10 or beta < 1e-13:
            if stop is None:
                stop = (j, x >= 0, list(V), ev.copy(), evec.copy())
            return stop + (x >= 0,)
        be.append(beta); V.append(w / beta)
    if stop is None:
        stop = (j, x >= 0, list(V), ev.copy(), evec.copy())
    return stop + (x >= 0,)


warm = {0: None}
work = dict(cold=0, warm=0, wrong=0)
per_level = {}
depth = {0: 0}
for u in range(tr.n_nodes):
    rows = np.asarray(tr.items[u]); m = len(rows)
    l_, r_ = tr.left[u], tr.right[u]
    if l_ >= 0:
        depth[l_] = depth[r_] = depth[u] + 1
    if m < 3:
        if l_ >= 0:
            warm[l_] = warm[r_] = None
        continue
    b = b0[rows]; d = O.b_to_d(b); c = O.bd_to_c(b, d).tocsr(); ct = c.T.tocsr()
    jc, labc, _, _, _, final = lanczos(c, ct, d, rng.standard_normal(m))
    w0 = warm.get(u)
    if w0 is None:
        jw, labw, V, ev, evec = jc, labc, None, None, None
        jw, labw, V, ev, evec, _ = lanczos(c, ct, d, rng.standard_normal(m))
    else:
        jw, labw, V, ev, evec, _ = lanczos(c, ct, d, w0 * np.sqrt(d))
    same = lambda p: np.array_equal(p, final) or np.array_equal(p, ~final)
    work["cold"] += jc * m; work["warm"] += jw * m
    work["wrong"] += (0 if same(labw) else 1)
    lv = per_level.setdefault(depth[u], [0, 0, 0]); lv[0] += jc * m; lv[1] += jw * m; lv[2] += 1
    if l_ >= 0:
        # warm vectors for the children from this node's Krylov space at its stop step
        j = len(ev)
        NV = int(os.environ.get("NV", "2")); MODE = os.environ.get("MODE", "mix")
        ks = [k for k in range(2, 2 + NV) if k <= j]
        F = [sum(evec[l, -k] * V[l + 1] for l in range(j)) / np.sqrt(d) for k in ks]
        pos = {int(r): i for i, r in enumerate(rows)}
        for ch in (l_, r_):
            idx = np.array([pos[int(r)] for r in tr.items[ch]])
            if not F:
                warm[ch] = None; continue
            if MODE == "mix":
                fw = sum((MIX ** i) * f[idx] for i, f in enumerate(F))
            else:
                # weight every restricted vector by how much of it varies inside the child (its norm after removing the mean)
                fw = np.zeros(len(idx))
                for i, f in enumerate(F):
                    g_ = f[idx] - f[idx].mean()
                    fw += (1 if i % 2 == 0 else 0.9) * g_ * max(ev[-ks[i]], 0) / max(np.linalg.norm(f[idx]) , 1e-300) * (np.linalg.norm(g_) / max(np.linalg.norm(f[idx]), 1e-300))
            warm[ch] = fw if np.linalg.norm(fw) > 0 else None
print(name, "nodes", tr.n_nodes, "work cold", work["cold"], "warm", work["warm"], "ratio", round(work["warm"] / work["cold"], 3), "wrong", work["wrong"])
for lv in sorted(per_level):
    a_, b_, k_ = per_level[lv]
    print(f"  depth {lv:2d}: nodes {k_:4d}  cold {a_:9d}  warm {b_:9d}  ratio {b_ / max(a_, 1):.3f}")
