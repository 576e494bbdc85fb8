"""Stress test of the warm start + sign certificate against the plain tolerance stop from cold starts: saturated trees
(min_modularity = -1: every node down to single cells is split, ~2 inner vertices of 64+ cells per 100 cells go through the
certificate) of many synthetic matrices; any vertex that differs is a wrong certificate (or a sign decided by rounding).

    python tests/manual/warm_stress.py [n_seeds] [cells] [genes]        env: as the engine's (TMC_CERT_WARM_SCALE, TMC_WARM_MIX, ...)
"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch
import too_many_cells_b200 as T
from too_many_cells_b200 import synth

n_seeds = int(sys.argv[1]) if len(sys.argv) > 1 else 20
n = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
g = int(sys.argv[3]) if len(sys.argv) > 3 else 30000
dev = torch.device("cuda", 0)
bad = 0; nodes_cert = 0; steps = [0, 0]
t00 = time.perf_counter()
for seed in range(100, 100 + n_seeds):
    rp, col, val = synth.synth_counts_torch(n, g, 2000 if g >= 20000 else 300, 8, seed, dev, False)
    b = T.DeviceB.adopt_device(n, g, rp, col, val, normalize=True, device=0)
    t_w = T.hierarchical_spectral_cluster(b, min_modularity=-1.0, max_active=16384)
    t_c = T.hierarchical_spectral_cluster(b, min_modularity=-1.0, max_active=16384, warm_start=False, sign_stop=False)
    sizes = t_w.item_end - t_w.item_begin
    nodes_cert += int(((sizes >= 64) & (t_w.left >= 0)).sum())
    steps[0] += int(t_w.iters.sum()); steps[1] += int(t_c.iters.sum())
    same = t_w.tree_hash_strict() == t_c.tree_hash_strict()
    if not same:
        bad += 1
        a_ = set(map(tuple, T.tree_vertex_names(t_w.left, t_w.right, t_w.item_begin, t_w.item_end, t_w.perm, strict=True).tolist()))
        c_ = set(map(tuple, T.tree_vertex_names(t_c.left, t_c.right, t_c.item_begin, t_c.item_end, t_c.perm, strict=True).tolist()))
        only = sorted(a_ - c_, key=lambda v: -v[1])
        print(f"seed {seed}: DIFFERENT, {len(only)} vertices; largest {only[:3]}", flush=True)
    b.free(); del rp, col, val
print(f"{n_seeds} matrices of {n} x {g}: {bad} trees differ; {nodes_cert} inner vertices of >= 64 cells; Lanczos steps warm+certificate {steps[0]}, "
      f"cold+tolerance {steps[1]}; {time.perf_counter() - t00:.1f} s")
