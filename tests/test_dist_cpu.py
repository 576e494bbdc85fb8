"""N>1 host-side logic on CPU (world_size 2, gloo): row-block arithmetic, block-seeded generation, the
unique-id exchange over torch.distributed, and that tmc_comm_init fails loudly without a device."""
import ctypes as C
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from too_many_cells_b200 import _lib, dist as tdist, synth
    res = {}
    # 1. row blocks tile [0, m) exactly, in rank order
    blocks = [tdist.row_block(1001, r, world) for r in range(world)]
    res["blocks_ok"] = blocks[0][0] == 0 and blocks[-1][1] == 1001 and all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
    # 2. every rank generates exactly its own rows of the same matrix
    lo, hi = tdist.row_block(6000, rank, world)
    rp, col, val = synth.synth_counts_torch(6000, 300, 30, 3, 9, torch.device("cpu"), row_range=(lo, hi))
    nnz = torch.tensor([col.numel()]); dist.all_reduce(nnz)
    chk = torch.tensor([float(val.sum()), float(col.double().sum())], dtype=torch.float64); dist.all_reduce(chk)
    res["nnz"] = int(nnz.item()); res["chk"] = chk.tolist()
    # 3. unique id exchange: both ranks end up with the same 128 bytes (needs libnccl; skip quietly if absent)
    lib = _lib.load()
    try:
        uid = tdist.exchange_unique_id(rank, world)
        gathered = [None] * world
        dist.all_gather_object(gathered, uid)
        res["uid_same"] = all(g == gathered[0] for g in gathered) and len(uid) == 128 and any(b != 0 for b in gathered[0])
        # 4. no device here => tmc_comm_init must fail loudly (no CPU fallback for the N>1 path either)
        rc = lib.tmc_comm_init(C.create_string_buffer(uid, 128), rank, world, rank) if lib.tmc_device_count() == 0 else None
        res["init_rc"] = rc
    except _lib.TmcError as e:
        res["uid_same"] = None; res["init_rc"] = e.code
    out[rank] = res
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_world2_gloo():
    world, port = 2, _free_port()
    mgr = mp.Manager(); out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    r0, r1 = out[0], out[1]
    assert r0["blocks_ok"] and r1["blocks_ok"]
    assert r0["nnz"] == r1["nnz"] and r0["chk"] == r1["chk"]
    from too_many_cells_b200 import synth
    full = synth.synth_counts_torch(6000, 300, 30, 3, 9, torch.device("cpu"))
    assert r0["nnz"] == full[1].numel() and abs(r0["chk"][0] - float(full[2].sum())) < 1e-6
    if r0["uid_same"] is not None:
        assert r0["uid_same"] and r1["uid_same"]
    for r in (r0, r1):
        if r["init_rc"] is not None:
            assert r["init_rc"] != 0          # ERR_NO_DEVICE / ERR_COMM, never a silent success


def test_peer_window_exchange_protocol_model(tmp_path):
    """The exchange protocol of csrc/tmc_p2p.cuh (publish into the own double-buffered window, release-signal every peer, wait,
    pull and sum) modelled with threads and std::atomic (tests/helpers/p2p_protocol_sim.cpp): with random stalls between the
    steps no rank may ever sum a buffer that a faster rank has already rewritten.  (The CUDA side — cudaIpc windows, .sys-scope
    ordering — needs N >= 2 GPUs: tools/p2p_check.py.)"""
    import subprocess
    ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = tmp_path / "p2psim"
    subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", os.path.join(ROOT, "tests", "helpers", "p2p_protocol_sim.cpp"), "-o", str(exe)],
                   check=True)
    for ranks, epochs, seed in ((2, 400, 1), (4, 300, 2), (8, 200, 3)):
        out = subprocess.run([str(exe), str(ranks), str(epochs), "48", "2", str(seed)], capture_output=True, text=True, check=True, timeout=120)
        assert out.stdout.strip() == "mismatches 0", (ranks, out.stdout)
