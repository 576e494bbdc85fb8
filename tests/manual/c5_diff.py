"""Which vertices of the SATURATED c3 tree (BASELINE configs[4]) depend on how a node is solved, and who is right?

Builds the tree (min_modularity = -1, every cell its own leaf) under several engine settings on the same device matrix, compares
the vertex sets (a vertex = its set of cells: smallest row id, size and a 64-bit set hash), and for parents whose children differ between two settings recomputes
the split on the host in exact-math fp64 (LAPACK eigh on the node's cosine-similarity matrix, the oracle's arithmetic) -- with
the eigen-gap and the smallest |u2| entry, which tell a genuine disagreement from a sign decided by rounding.

    python tests/manual/c5_diff.py [c3|c2]        (one B200; c3 needs ~60 GB)
"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))      # checker scripts live under tests/: only tests may import the oracle
import numpy as np, scipy.sparse as sp, torch
import too_many_cells_b200 as T
from too_many_cells_b200 import synth
import tmc_oracle as O

name = sys.argv[1] if len(sys.argv) > 1 else "c3"
n, g, k, d, s, binary = synth.CONFIGS[name]
dev = torch.device("cuda", 0)
rp, col, val = synth.synth_counts_torch(n, g, k, d, s, dev, binary)
b = T.DeviceB.adopt_device(n, g, rp, col, val, normalize=True, device=0)
settings = {"direct<=64 + warm": dict(), "sweeps + warm": dict(small_rows=0),
            "sweeps, cold": dict(small_rows=0, warm_start=False), "sweeps, cold, tolerance stop": dict(small_rows=0, warm_start=False, sign_stop=False),
            "sweeps, cold, seed 7": dict(small_rows=0, warm_start=False, seed=7)}
trees, names = {}, {}
for key, kw in settings.items():
    t0 = time.perf_counter()
    tr = T.hierarchical_spectral_cluster(b, min_modularity=-1.0, max_active=16384, **kw)
    trees[key] = tr
    nm = T.tree_vertex_names(tr.left, tr.right, tr.item_begin, tr.item_end, tr.perm, strict=True)      # (smallest row, size, set hash)
    names[key] = nm
    print(f"{key:22s} {1e3 * (time.perf_counter() - t0):8.1f} ms  nodes {tr.n_nodes}  Lanczos steps {int(tr.iters.sum())}  hash {tr.tree_hash_strict()}", flush=True)

df = torch.bincount(col.long(), minlength=g).double()
idf = torch.log(torch.tensor(float(n), dtype=torch.float64, device=dev) / df.clamp(min=1.0))


def node_rows_b(rows):
    """rows of B (tf-idf, L2-normalised) for the given cells, as a host CSR"""
    rows_t = torch.as_tensor(rows, device=dev)
    lo, hi = rp[rows_t], rp[rows_t + 1]
    parts_c, parts_v, ip = [], [], [0]
    for a_, b_ in zip(lo.tolist(), hi.tolist()):
        parts_c.append(col[a_:b_].long()); parts_v.append(val[a_:b_] * idf[col[a_:b_].long()]); ip.append(ip[-1] + (b_ - a_))
    m = sp.csr_matrix((torch.cat(parts_v).cpu().numpy(), torch.cat(parts_c).cpu().numpy(), np.array(ip)), shape=(len(rows), g))
    return O.b2_to_b(m)


def as_set(nm):
    return set(map(tuple, nm.tolist()))


keys = list(settings)
ref = keys[0]
for other in keys[1:]:
    sa, sb = as_set(names[ref]), as_set(names[other])
    only_a, only_b = sa - sb, sb - sa
    print(f"\n{ref}  vs  {other}: {len(only_a)} / {len(only_b)} vertices only in one of them")
    if not only_a:
        continue
    # parents present in both whose children differ: walk tree `ref`, smallest parents first
    tr = trees[ref]; nm = names[ref]
    inner = np.nonzero(tr.left >= 0)[0]
    bad = [u for u in inner if tuple(nm[u]) in sb and (tuple(nm[tr.left[u]]) not in sb or tuple(nm[tr.right[u]]) not in sb)]
    bad.sort(key=lambda u: -nm[u][1])      # largest first: a different split high up changes the cell sets of everything below it
    print(f"  parents with different children: {len(bad)}; sizes {[int(nm[u][1]) for u in bad[:20]]}")
    tro = trees[other]; nmo = names[other]; index_o = {tuple(v): i for i, v in enumerate(nmo.tolist())}
    for u in bad[:4]:
        rows = np.sort(tr.items(u))
        bm = node_rows_b(rows)
        lab, u2, th2, gap, dd = O.spectral_cluster(bm)
        side1 = set(rows[lab == 1].tolist()); side0 = set(rows[lab == 0].tolist())
        def agrees(t_, node):
            l_ = set(t_.items(t_.left[node]).tolist())
            return l_ == side0 or l_ == side1
        uo = index_o[tuple(nm[u])]
        if set(tro.items(uo).tolist()) != set(rows.tolist()):
            print(f"  parent (min row {int(nm[u][0])}, {int(nm[u][1])} cells): same name, different cells (below an earlier difference)")
            continue
        second = np.sort(np.abs(u2))[:3]
        print(f"  parent (min row {int(nm[u][0])}, {int(nm[u][1])} cells): theta2 {th2:.6g} gap {gap:.3g} smallest |u2| entries {second}  "
              f"exact split {len(side0)}|{len(side1)}  '{ref}' agrees: {agrees(tr, u)}  '{other}' agrees: {agrees(tro, uo)}", flush=True)
