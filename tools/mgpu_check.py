"""torchrun --nproc-per-node N tools/mgpu_check.py [names...] : the N-rank tree must equal the 1-rank tree (canonical hash,
Q to 1e-12) through every multi-GPU entry: (a) the whole host matrix on every rank (replicated mode: own block uploaded and
prepared, prepared blocks exchanged), (b) every rank passes only its own row block (row-sharded throughout), (c) device-resident
arrays adopted on every rank, and (a) once more with TMC_REPL_PREPARE=1 (every rank prepares the whole matrix: the first version).

    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 tools/mgpu_check.py
"""
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
single, mats = {}, {}
for name in names:
    ip, ix, dv, _ = synth.synth_config(name)
    n, g = synth.CONFIGS[name][:2]
    mats[name] = sp.csr_matrix((dv, ix, ip), shape=(n, g))
    single[name] = T.hierarchical_spectral_cluster(mats[name], normalize=True, device=lr)      # before the communicator exists
vname = names[0]
v_single = (T.hierarchical_spectral_cluster(mats[vname][:300], normalize=True, eigen_group="KMeansGroup"),
            T.hierarchical_spectral_cluster(mats[vname][:300], normalize=True, num_runs=6))
# other value kinds through the load-time fallbacks: a tf-idf matrix handed over as fp64 (factor recovery), an fp64 matrix that is
# not a count matrix times column weights (B materialised), and a signed matrix (two degree passes, no deflation)
a0 = mats[names[0]]
b2 = T.tfidf_scale_sparse_mat(a0)                                   # B2 = counts * diag(idf): recognised bit for bit
rng = np.random.default_rng(4)
noisy = b2.copy(); noisy.data = noisy.data * (1.0 + 0.05 * rng.random(noisy.nnz))      # no exact factorisation: fp64 storage
signed = b2.copy(); signed.data = signed.data - 0.9 * np.median(signed.data)          # entries of both signs
kinds = {"tf-idf fp64 (recovered counts)": b2, "fp64, not factorable": noisy, "signed fp64": signed}
kind_single = {k: T.hierarchical_spectral_cluster(m, device=lr) for k, m in kinds.items()}
dist.init_process_group("nccl", device_id=dev)
tdist.init_comm(rank, world, lr, dev)
ok = True


def check(name, how, tr, ms):
    global ok
    ref = single[name]
    same = tr.tree_hash_strict() == ref.tree_hash_strict()
    dq = float(np.max(np.abs(np.sort(tr.q) - np.sort(ref.q)))) if tr.n_nodes == ref.n_nodes else float("inf")
    good = same and dq <= 1e-12
    ok &= good
    print(f"[rank {rank}] {name:6s} {how:34s} nodes {tr.n_nodes:5d} vs {ref.n_nodes:5d}  hash equal {same}  max|dQ| {dq:.1e}  "
          f"wall {ms:7.1f} ms  engine {tr.stats['engine_ms']:7.1f} ms (1 GPU {ref.stats['engine_ms']:.1f})  {'ok' if good else 'FAIL'}", flush=True)


for name in names:
    a = mats[name]
    t0 = time.time(); tr = T.hierarchical_spectral_cluster(a, normalize=True, device=lr)
    check(name, "host matrix, replicated", tr, 1e3 * (time.time() - t0))
    u16 = (a.indptr, a.indices.astype(np.uint16) if a.shape[1] <= 65536 else a.indices, a.data.astype(np.uint16), a.shape)
    t0 = time.time(); tr = T.hierarchical_spectral_cluster(u16, normalize=True, device=lr)
    check(name, "host matrix, uint16 counts", tr, 1e3 * (time.time() - t0))
    lo, hi = tdist.row_block(a.shape[0], rank, world)
    blk = a[lo:hi]
    t0 = time.time(); tr = T.hierarchical_spectral_cluster((blk.indptr, blk.indices, blk.data, blk.shape, lo, a.shape[0]), normalize=True, device=lr)
    check(name, "own row block per rank", tr, 1e3 * (time.time() - t0))
    rp = torch.from_numpy(a.indptr.astype(np.int64)).to(dev); ci = torch.from_numpy(a.indices.astype(np.int32)).to(dev)
    va = torch.from_numpy(a.data.astype(np.float64)).to(dev)
    t0 = time.time()
    b = T.DeviceB.adopt_device(a.shape[0], a.shape[1], rp, ci, va, normalize=True, device=lr)
    tr = T.hierarchical_spectral_cluster(b); b.free()
    check(name, "device-resident arrays adopted", tr, 1e3 * (time.time() - t0))
    os.environ["TMC_REPL_PREPARE"] = "1"
    t0 = time.time(); tr = T.hierarchical_spectral_cluster(a, normalize=True, device=lr)
    check(name, "host matrix, TMC_REPL_PREPARE=1", tr, 1e3 * (time.time() - t0))
    del os.environ["TMC_REPL_PREPARE"]
for k, m in kinds.items():
    t0 = time.time(); tr = T.hierarchical_spectral_cluster(m, device=lr)
    ref = kind_single[k]
    good = tr.tree_hash_strict() == ref.tree_hash_strict() and float(np.max(np.abs(np.sort(tr.q) - np.sort(ref.q)))) <= 1e-10
    ok &= good
    print(f"[rank {rank}] {names[0]:6s} {k:34s} nodes {tr.n_nodes:5d} vs {ref.n_nodes:5d}  wall {1e3 * (time.time() - t0):7.1f} ms  {'ok' if good else 'FAIL'}", flush=True)

# a zero-norm row in ONE rank's block must fail on EVERY rank (no rank may go on into the tree's collectives), in
# both multi-GPU modes; and a slot budget of 1 must not silently truncate handed-off subtrees
a = mats[names[0]].tolil(copy=True)
a[a.shape[0] - 1, :] = 0          # the last rank's block gets an empty cell
bad = sp.csr_matrix(a)
bad.eliminate_zeros()
codes = []
for how in ("whole", "block"):
    try:
        if how == "whole":
            T.hierarchical_spectral_cluster(bad, normalize=True, device=lr)
        else:
            lo, hi = tdist.row_block(bad.shape[0], rank, world)
            blk = bad[lo:hi]
            T.hierarchical_spectral_cluster((blk.indptr, blk.indices, blk.data, blk.shape, lo, bad.shape[0]), normalize=True, device=lr)
        codes.append(0)
    except T.TmcError as e:
        codes.append(e.code)
good = codes == [T._lib.ERR_ZERO_ROW, T._lib.ERR_ZERO_ROW]
ok &= good
print(f"[rank {rank}] zero-norm row in the last rank's block -> status on this rank {codes}: {'ok' if good else 'FAIL'}", flush=True)
t1 = T.hierarchical_spectral_cluster(mats[names[0]], normalize=True, device=lr, max_active=1)
good = t1.tree_hash_strict() == single[names[0]].tree_hash_strict()
ok &= good
print(f"[rank {rank}] max_active = 1 under {world} ranks builds the whole tree: {'ok' if good else 'FAIL'}", flush=True)

# the non-default variants under a communicator: every rank runs them as a replica, results checked across ranks inside libtmc
tk = T.hierarchical_spectral_cluster(mats[vname][:300], normalize=True, eigen_group="KMeansGroup", device=lr)
tp = T.hierarchical_spectral_cluster(mats[vname][:300], normalize=True, num_runs=6, device=lr)
good = tk.tree_hash_strict() == v_single[0].tree_hash_strict() and tp.tree_hash_strict() == v_single[1].tree_hash_strict() and \
    np.array_equal(np.nan_to_num(tp.p_val, nan=-1.0), np.nan_to_num(v_single[1].p_val, nan=-1.0))
ok &= good
print(f"[rank {rank}] variants under {world} ranks (KMeansGroup, numRuns): {'ok' if good else 'FAIL'}", flush=True)
tdist.destroy_comm()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
