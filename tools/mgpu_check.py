"""torchrun --nproc-per-node N tools/mgpu_check.py : the N-rank tree must equal the 1-rank tree."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, scipy.sparse as sp, torch
import torch.distributed as dist
import too_many_cells_b200 as T
from too_many_cells_b200 import synth, dist as tdist

rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); lr = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
names = sys.argv[1:] or ["tiny", "small", "c1"]
single = {}
mats = {}
for name in names:
    ip, ix, dv, _ = synth.synth_config(name)
    n, g = synth.CONFIGS[name][:2]
    mats[name] = sp.csr_matrix((dv, ix, ip), shape=(n, g))
    single[name] = T.hierarchical_spectral_cluster(mats[name], normalize=True, device=lr)      # before the communicator exists
dist.init_process_group("nccl", device_id=dev)
tdist.init_comm(rank, world, lr, dev)
ok = True
for name in names:
    a = mats[name]
    t0 = time.time()
    tr = T.hierarchical_spectral_cluster(a, normalize=True, device=lr)       # whole host matrix on every rank, sliced inside
    t1 = time.time()
    lo, hi = tdist.row_block(a.shape[0], rank, world)
    blk = a[lo:hi]
    tr2 = T.hierarchical_spectral_cluster((blk.indptr, blk.indices, blk.data, blk.shape, lo, a.shape[0]), normalize=True, device=lr)  # own block
    same = tr.leaf_sets() == single[name].leaf_sets() and tr2.leaf_sets() == single[name].leaf_sets()
    dq = float(np.max(np.abs(np.sort(tr.q) - np.sort(single[name].q))))
    ok &= same
    print(f"[rank {rank}] {name}: nodes {tr.n_nodes} vs {single[name].n_nodes}  leaf sets equal {same}  max|dQ| {dq:.2e}  "
          f"engine {tr.stats['engine_ms']:.1f} ms (1 GPU {single[name].stats['engine_ms']:.1f} ms) sweeps {tr.stats['n_sweeps']}", flush=True)
tdist.destroy_comm()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
