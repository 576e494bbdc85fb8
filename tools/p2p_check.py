"""torchrun --nproc-per-node N tools/p2p_check.py [sizes...] : the one-shot all-reduce over NVLink peer memory
(csrc/tmc_p2p.cuh) against ncclAllReduce — values and device time per call — then the N-rank tree with TMC_P2P=1 against the
1-rank tree.  Round-2 entry point: the peer path was written without GPU minutes left and is off by default until this passes.

    timeout 300 torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/p2p_check.py
"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
os.environ["TMC_P2P"] = "1"
import numpy as np, scipy.sparse as sp, torch
import torch.distributed as dist
import too_many_cells_b200 as T
from too_many_cells_b200 import _lib, synth, dist as tdist

rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); lr = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
sizes = [int(x) for x in sys.argv[1:]] or [326, 30208, 8 * 30208, 256 * 30208]     # scalars of one node ... the whole Y arena
names = ["small", "c1"]
mats, single = {}, {}
for name in names:                                  # 1-rank trees before the communicator exists
    ip, ix, dv, _ = synth.synth_config(name)
    n, g = synth.CONFIGS[name][:2]
    mats[name] = sp.csr_matrix((dv, ix, ip), shape=(n, g))
    single[name] = T.hierarchical_spectral_cluster(mats[name], normalize=True, device=lr)
dist.init_process_group("nccl", device_id=dev)
tdist.init_comm(rank, world, lr, dev)
lib = _lib.load()
ok = True
for n in sizes:
    d, pu, nu = C.c_double(), C.c_double(), C.c_double()
    _lib.check(lib.tmc_comm_selftest(n, 50, C.byref(d), C.byref(pu), C.byref(nu)))
    good = pu.value > 0 and d.value <= 1e-12
    ok &= good
    if rank == 0:
        print(f"all-reduce of {n:>9d} doubles: max|p2p - nccl| {d.value:.2e}   p2p {pu.value:8.1f} us   nccl {nu.value:8.1f} us   {'ok' if good else 'FAIL'}", flush=True)
for name in names:
    tr = T.hierarchical_spectral_cluster(mats[name], normalize=True, device=lr)
    same = tr.leaf_sets() == single[name].leaf_sets()
    dq = float(np.max(np.abs(np.sort(tr.q) - np.sort(single[name].q))))
    ok &= same
    print(f"[rank {rank}] {name} with TMC_P2P=1: leaf sets equal {same}  max|dQ| {dq:.2e}  engine {tr.stats['engine_ms']:.1f} ms "
          f"(1 GPU {single[name].stats['engine_ms']:.1f} ms)", flush=True)
tdist.destroy_comm()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
