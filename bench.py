#!/usr/bin/env python
"""bench.py — `ms per full make-tree` on synthetic 10x-like matrices (BASELINE.json metric).

One "step" = one complete make-tree hot path over the workload: tf-idf + row-L2
normalisation, then the whole recursive spectral bipartition (degree, Lanczos SpMV
pairs, sign split, modularity, stable row partition) down to the leaves.

  value  : device-resident input (raw counts already in HBM), ms per tree
  e2e    : the same through the host-buffer C-ABI call `tmc_make_tree` (pinned host CSR ->
           H2D -> tree -> D2H of the flat tree), ms per tree
  roofline: SpMV pair kernels (scatter y=C^T x, gather x'=C y): algorithmic bytes
            / CUDA-event time of those launches inside the timed steps
  cpu_baseline / --impl reference: the oracle's CPU implementation of the same path.

N>1 (torchrun, one rank per GPU): the tree build is row-sharded across ranks with NCCL
all-reduces of the feature-dimension vectors; total work is fixed => "strong" scaling.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "ms per full make-tree"
UNIT = "ms"
# ncu dram bytes (read + write) of the root-level SpMV pair on c3, 1 GPU (profiles/, see the roofline comment below)
TRAFFIC_C3_ROOT_PAIR = 13.84e9


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=os.environ.get("TMC_BENCH_WORKLOAD", "c3"))
    ap.add_argument("--cpu-sample-cells", type=int, default=10000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def workload_desc(name):
    from too_many_cells_b200 import synth
    n, g, k, d, s, binary = synth.CONFIGS[name]
    return {"workload": f"{name}: synthetic 10x-like {n} cells x {g} genes, ~{k} nnz/cell, planted depth {d}, seed {s}, "
                        f"tf-idf + default min-modularity 0 stop",
            "cells": n, "features": g, "l2": "inputs larger than L2 (CSR >> 126 MB)" if n * k * 12 > 2.5e8 else
            "matrix smaller than L2; no flush (the tree build re-reads it by design)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm = sorted(float(r[1]) for r in self.rows if len(r) > 2 and r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for k, nm in enumerate(names):
                if len(r) > 5 + k and r[5 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


def cpu_reference_run(name, sample_cells, steps=1, warmup=0):
    """The reference's CPU execution of the path, restated in C (oracle/tmc_port.c: one node at a time, materialised
    sub-matrices, single-vector Lanczos at SVDLIBC's kappa = 1e-6, one core like the reference), timed on a bounded sample
    (the first `sample_cells` cells of the workload) and scaled linearly in cells to the metric's unit.
    Returns (scaled ms per full-size tree, sample cells, sample nnz, description)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import scipy.sparse as sp
    from too_many_cells_b200 import synth
    import tmc_port as P
    n, g, k, d, s, binary = synth.CONFIGS[name]
    ns = min(n, sample_cells)
    ip, ix, dv, _ = synth.synth_counts(ns, g, k, d, s, binary)
    a = sp.csr_matrix((dv, ix, ip), shape=(ns, g))
    P._load()          # build / host-compatibility probe of the port happen here, outside the timed region
    times = []
    for it in range(warmup + steps):
        t = time.perf_counter()
        tr = P.make_tree(a, normalize=True, tol=1e-6)
        dt = time.perf_counter() - t
        if it >= warmup:
            times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    scale = n / ns
    # second, more generous figure: the same port with its mat-vecs and re-orthogonalisation on every host core (the reference
    # itself cannot do that: sequential recursion, single-threaded SVDLIBC)
    all_cores = None
    ncpu = min(os.cpu_count() or 1, 32)
    if ncpu > 1:
        t = time.perf_counter()
        P.make_tree(a, normalize=True, tol=1e-6, threads=ncpu)
        all_cores = {"value": 1e3 * (time.perf_counter() - t) * scale, "unit": "ms", "cores": ncpu,
                     "note": "pthread mode of the port (not something the reference can do); same sample, same scaling"}
        P.make_tree(a[:50], normalize=True, tol=1e-6, threads=1)      # back to the 1-thread pool
    desc = (f"C port of the reference's CPU path (oracle/tmc_port.c: sequential recursion, per-node materialised matrices, "
            f"Lanczos at las2's kappa=1e-6, 1 thread like the reference; the Haskell binary cannot be built here) on the first "
            f"{ns} cells x {g} genes of {name} ({a.nnz} nnz, {tr.n_nodes} nodes, {tr.steps} Lanczos steps): {ms:.0f} ms measured, "
            f"scaled x{scale:.0f} linearly in cells to the full workload (conservative: the full tree is deeper)")
    return ms * scale, ns, a.nnz, desc, all_cores


def tr_bv(trees):
    return float(trees[-1].stats["value_bytes"])


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    name = args.workload
    cfg = workload_desc(name)

    if args.impl == "reference":
        if rank != 0:
            return
        ncores = 1          # the reference path is sequential; the port uses one thread
        ms, ns, nnz, desc, all_cores = cpu_reference_run(name, args.cpu_sample_cells, steps=max(1, args.steps), warmup=min(args.warmup, 1))
        line = {"impl": "reference", "metric": METRIC, "value": ms, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": False, "scaling": "strong", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": cfg,
                "cpu_baseline": {"value": ms, "unit": UNIT, "cores": ncores, "kind": "port", "sample": desc, "all_cores": all_cores},
                "e2e": {"value": ms, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import numpy as np
    import torch
    import too_many_cells_b200 as T
    from too_many_cells_b200 import synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (libtmc has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    lib = T._lib.load()
    if world > 1:
        import torch.distributed as dist
        from too_many_cells_b200 import dist as tdist
        dist.init_process_group("nccl", device_id=dev)
        tdist.init_comm(rank, world, local_rank, dev)          # libtmc's own NCCL communicator (ids travel over torch.distributed)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    n, g, k, d, s, binary = synth.CONFIGS[name]
    t0 = time.time()
    # Every rank holds the WHOLE matrix in HBM (replicated mode: the shared top of the tree is row-sharded over the ranks'
    # blocks [lo, hi), subtrees below the hand-off size go to single ranks); it is generated on the device, identically
    # on every rank (block-seeded generator).
    lo, hi = (n * rank // world, n * (rank + 1) // world)          # this rank's block of the shared phase
    row_ptr, col, val_raw = synth.synth_counts_torch(n, g, k, d, s, dev, binary)
    torch.cuda.synchronize()
    nnz = int(val_raw.numel())
    gen_s = time.time() - t0
    work = torch.empty_like(val_raw)

    def step_dev(profile):
        work.copy_(val_raw)                       # raw counts resident in HBM; tf-idf + L2 happen inside the step
        torch.cuda.synchronize()
        b = T.DeviceB.adopt_device(n, g, row_ptr, col, work, normalize=True, device=local_rank)
        tree = T.hierarchical_spectral_cluster(b, profile=profile)
        b.free()
        return tree

    for _ in range(args.warmup):
        step_dev(False)
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    trees = [step_dev(True) for _ in range(args.steps)]
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms_step = e0.elapsed_time(e1) / args.steps
    if world > 1:
        tt = torch.tensor([ms_step], device=dev, dtype=torch.float64)
        torch.distributed.all_reduce(tt, op=torch.distributed.ReduceOp.MAX)
        ms_step = float(tt.item())

    # ---- end-to-end through the host-buffer C-ABI (rank 0's numbers; every rank participates when N>1)
    # host CSR: the C ABI reads the full row_ptr and only this rank's block of col/val (then replicates over NVLink), so the
    # arrays are full-size but only the own block is backed by (pinned) memory
    h_ptr = row_ptr.cpu().pin_memory()
    k0, k1 = int(h_ptr[lo]), int(h_ptr[hi])
    if world == 1:
        h_col = col.cpu().pin_memory(); h_val = val_raw.cpu().pin_memory()
        np_col, np_val = h_col.numpy(), h_val.numpy()
    else:
        np_col = np.empty(nnz, dtype=np.int32); np_val = np.empty(nnz, dtype=np.float64)      # untouched pages stay uncommitted
        np_col[k0:k1] = col[k0:k1].cpu().numpy(); np_val[k0:k1] = val_raw[k0:k1].cpu().numpy()
        rt = torch.cuda.cudart()
        rt.cudaHostRegister(np_col[k0:k1].ctypes.data, (k1 - k0) * 4, 0)
        rt.cudaHostRegister(np_val[k0:k1].ctypes.data, (k1 - k0) * 8, 0)
    host = (h_ptr.numpy(), np_col, np_val, (n, g))
    e2e_steps = max(1, min(args.steps, 2))
    T.hierarchical_spectral_cluster(host, normalize=True, device=local_rank)     # warm
    barrier()
    t = time.perf_counter()
    for _ in range(e2e_steps):
        tr_e = T.hierarchical_spectral_cluster(host, normalize=True, device=local_rank)
    barrier()
    e2e_ms = 1e3 * (time.perf_counter() - t) / e2e_steps
    h2d = int(h_ptr.numel() * 8 + (k1 - k0) * 12)
    d2h = int(tr_e.n_nodes * (8 * 5 + 8 * 2 + 4) + n * 8)
    if world > 1:
        th = torch.tensor([h2d], device=dev, dtype=torch.int64)
        torch.distributed.all_reduce(th)
        h2d = int(th.item())
        te = torch.tensor([e2e_ms], device=dev, dtype=torch.float64)
        torch.distributed.all_reduce(te, op=torch.distributed.ReduceOp.MAX)
        e2e_ms = float(te.item())

    if rank != 0:
        if world > 1:
            tdist.destroy_comm()
            torch.distributed.destroy_process_group()
        return

    tr = trees[-1]
    spmv_ms = sum(t_.stats["spmv_ms"] for t_ in trees)
    alg = sum(t_.stats["spmv_alg_bytes"] for t_ in trees)
    launches = sum(t_.stats["kernel_launches"] for t_ in trees)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak = float(json.load(open(peaks_path))["hbm_gbs"]); peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak = 6650.0; peak_src = "fallback (B200_PROFILING.md)"
    achieved = alg / (spmv_ms * 1e-3) / 1e9 if spmv_ms > 0 else None
    # the same launches counted with the nominal fp64 value size of SURVEY.md 8(d) (b_v = 8): what a plain fp64 CSR
    # implementation would have to move; reported for context only, `achieved` / `frac` use the bytes actually stored
    nnz_pairs = sum(t_.stats["spmv_pairs_nnz"] for t_ in trees)
    alg_nominal = alg + 2.0 * (8.0 - tr_bv(trees)) * nnz_pairs
    achieved_nominal = alg_nominal / (spmv_ms * 1e-3) / 1e9 if spmv_ms > 0 else None
    # root-level pair in isolation (largest single launch pair) -- single-rank only
    if world == 1:
        work.copy_(val_raw)
        b = T.DeviceB.adopt_device(n, g, row_ptr, col, work, normalize=True, device=local_rank)
        sc_ms, ga_ms = T.bench_spmv_pair(b, 10)
        storage = b.info()
        b.free()
        bv = tr.stats["value_bytes"]          # 8: fp64 B; 2: raw counts as uint16 + diagonal factors; 0: all-ones matrix
        # stored bytes of one pair: (value + column index) per residual CSR entry and half, one byte per panel cell and half
        pair_bytes = (2.0 * (bv + 4.0) * storage["residual_nnz"] + 2.0 * n * storage["panel_cols"] * storage["panel_bits"] / 8.0
                      + 16.0 * (n + 1) + 32.0 * n + 16.0 * g)
        root_pair = {"scatter_ms": sc_ms, "gather_ms": ga_ms, "bytes": pair_bytes, "storage": storage,
                     "scatter_gbs": 0.5 * pair_bytes / (sc_ms * 1e-3) / 1e9, "gather_gbs": 0.5 * pair_bytes / (ga_ms * 1e-3) / 1e9}
    else:
        root_pair = None
    # dram__bytes_read + write of the root-level pair launches from the committed `ncu --set full` capture
    # (profiles/r1_v7_nibble_pair_c3_ncu_raw.csv: k_tgather_t + k_tpanel + k_gather_panel + k_gather_u); the stored bytes of that
    # pair are 2*(6 B * residual entries + panel bytes): the column-sorted residual copy also stores the cell position of every
    # entry (10 B per entry instead of 6), that is the whole difference - nothing is read twice.
    traffic = TRAFFIC_C3_ROOT_PAIR if (name == "c3" and world == 1 and tr.stats["value_bytes"] == 2 and root_pair and root_pair["storage"]["panel_cols"] > 0) else None
    roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
            "traffic": traffic, "traffic_scope": "root-level pair launch (largest launch), ncu dram bytes" if traffic else None, "peak_source": peak_src, "kernel": "SpMV pair: y = C^T x (k_tgather_t on the column-sorted residual blocks + k_tpanel on the dense byte panel), x' = C y (k_gather_panel + k_gather_u on the residual CSR); all launches of the timed steps",
            "algorithmic_bytes_per_step": alg / len(trees), "value_bytes_b_v": tr.stats["value_bytes"],
            "achieved_if_counted_with_fp64_values_b_v8": achieved_nominal,
            "frac_if_counted_with_fp64_values_b_v8": (achieved_nominal / peak) if achieved_nominal else None,
            "algorithmic_bytes_formula": "per SpMV pair 2*nnz_residual*(b_v+4) + 2*m*P*c + 2*(m+1)*8 + (4*m+2*n)*8 (BASELINE.md section 4 with the storage actually used: b_v = bytes per stored value, P = dense panel columns at one byte or half a byte per cell)", "spmv_ms_per_step": spmv_ms / len(trees),
            "spmv_share_of_step": spmv_ms / len(trees) / ms_step,
            "per_gpu": True, "root_pair": root_pair}
    line = {"metric": METRIC, "value": ms_step, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": dict(cfg, nnz=nnz, parallelism=("1 GPU" if world == 1 else
                                             f"{world} GPUs: matrix replicated, shared top of the tree row-sharded with NCCL all-reduce of y = B^T x "
                                             f"per half-step, subtrees below the hand-off size owned by single GPUs (no collectives)")),
            "clocks": clocks,
            "e2e": {"value": e2e_ms, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": launches, "roofline": roof,
            "tree": {"nodes": tr.n_nodes, "leaves": int(len(tr.leaves())), "sweeps": tr.stats["n_sweeps"],
                     "lanczos_steps": int(tr.iters.sum()), "engine_ms_last": tr.stats["engine_ms"], "gen_s": gen_s,
                     "hash": tr.tree_hash()}}
    if not args.no_cpu_baseline and world == 1:          # reported on rank 0 at N=1 only
        ms_c, ns, nnz_c, desc, all_cores = cpu_reference_run(name, args.cpu_sample_cells)
        line["cpu_baseline"] = {"value": ms_c, "unit": UNIT, "cores": 1, "kind": "port", "sample": desc, "all_cores": all_cores}
    print(json.dumps(line))
    if world > 1:
        tdist.destroy_comm()
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
