"""Short profiling target: the root-level SpMV pair on a synthetic config (for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import too_many_cells_b200 as T
from too_many_cells_b200 import synth
name = sys.argv[1] if len(sys.argv) > 1 else "c2"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
n, g, k, d, s, binary = synth.CONFIGS[name]
dev = torch.device("cuda", 0)
row_ptr, col, val = synth.synth_counts_torch(n, g, k, d, s, dev, binary)
b = T.DeviceB.adopt_device(n, g, row_ptr, col, val, normalize=True, device=0)
sc, ga = T.bench_spmv_pair(b, reps)
nnz = val.numel()
bv = 8 if os.environ.get("TMC_NO_FACTORED") else (0 if binary else 2)
pair = 2.0 * (bv + 4) * nnz + 16.0 * (n + 1) + 32.0 * n + 16.0 * g
print(f"{name}: nnz {nnz} b_v={bv} scatter {sc:.4f} ms ({0.5*pair/sc/1e6:.0f} GB/s alg) gather {ga:.4f} ms ({0.5*pair/ga/1e6:.0f} GB/s alg)")
if len(sys.argv) > 3:
    t = T.hierarchical_spectral_cluster(b, profile=True)
    print("tree", t.n_nodes, t.stats)
