"""torchrun --nproc-per-node N tools/mgpu_sweep.py [workload] : one process group, the matrix generated once, several engine
configurations (environment knobs read at every tmc_make_tree* call) timed back to back on the device-resident input:
NCCL vs the peer-memory all-reduce, hand-off size, own-block vs replicated prepare.  Prints ms per tree (max over ranks, CUDA
events), Lanczos steps, sweeps and the canonical tree hash of every configuration (they must all be equal).

    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 tools/mgpu_sweep.py c3
"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import too_many_cells_b200 as T
from too_many_cells_b200 import synth, dist as tdist

rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); lr = int(os.environ.get("LOCAL_RANK", 0))
name = sys.argv[1] if len(sys.argv) > 1 else "c3"
steps = int(os.environ.get("SWEEP_STEPS", "3"))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
n, g, k, d, s, binary = synth.CONFIGS[name]
rp, col, val = synth.synth_counts_torch(n, g, k, d, s, dev, binary)
work = torch.empty_like(val)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
    tdist.init_comm(rank, world, lr, dev)

CONFIGS = [("default", {}), ("p2p", {"TMC_P2P": "1"}), ("handoff 2048", {"TMC_HANDOFF_ROWS": "2048"}), ("handoff 8192", {"TMC_HANDOFF_ROWS": "8192"}),
           ("handoff 16384", {"TMC_HANDOFF_ROWS": "16384"}), ("handoff 32768", {"TMC_HANDOFF_ROWS": "32768"}), ("handoff 65536", {"TMC_HANDOFF_ROWS": "65536"}),
           ("handoff 131072", {"TMC_HANDOFF_ROWS": "131072"}), ("p2p + handoff 32768", {"TMC_P2P": "1", "TMC_HANDOFF_ROWS": "32768"}),
           ("replicated prepare", {"TMC_REPL_PREPARE": "1"}), ("no sign stop", {"TMC_NO_CERT": "1"})]
if os.environ.get("SWEEP_ONLY"):
    CONFIGS = [c for c in CONFIGS if c[0] in os.environ["SWEEP_ONLY"].split(",")]


load_ms, tree_ms = [], []


def one(profile=False):
    work.copy_(val); torch.cuda.synchronize()
    t0 = time.perf_counter()
    b = T.DeviceB.adopt_device(n, g, rp, col, work, normalize=True, device=lr)      # load: prepare (+ exchange of the prepared blocks)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    t = T.hierarchical_spectral_cluster(b, profile=profile)
    t2 = time.perf_counter()
    b.free()
    load_ms.append(1e3 * (t1 - t0)); tree_ms.append(1e3 * (t2 - t1))
    return t


out = []
for label, env in CONFIGS:
    for kk, v in env.items():
        os.environ[kk] = v
    try:
        one()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        del load_ms[:], tree_ms[:]
        for _ in range(steps):
            t = one(True)
        e1.record(); torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / steps, sum(load_ms) / steps, sum(tree_ms) / steps], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        eng_all = [None] * world
        if world > 1:
            dist.all_gather_object(eng_all, round(t.stats["engine_ms"], 1))
        rec = {"config": label, "ms": float(ms[0].item()), "load_ms": float(ms[1].item()), "tree_wall_ms": float(ms[2].item()), "engine_ms_by_rank": eng_all, "engine_ms": t.stats["engine_ms"], "spmv_ms": t.stats["spmv_ms"], "steps": int(t.iters.sum()),
               "sweeps": t.stats["n_sweeps"], "nodes": t.n_nodes, "hash": t.tree_hash()[:12]}
    except Exception as e:      # a configuration that fails must not take the others down (all ranks fail alike)
        rec = {"config": label, "error": str(e)[:200]}
    for kk in env:
        del os.environ[kk]
    out.append(rec)
    if rank == 0:
        print(json.dumps(rec), flush=True)
if rank == 0:
    hs = {r.get("hash") for r in out if "hash" in r}
    print("hashes equal:", len(hs) == 1, flush=True)
if world > 1:
    tdist.destroy_comm()
    dist.destroy_process_group()
