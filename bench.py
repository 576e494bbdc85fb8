#!/usr/bin/env python
"""bench.py — `ms per full make-tree` on synthetic 10x-like matrices (BASELINE.json metric).

One "step" = one complete make-tree hot path over the workload: tf-idf + row-L2
normalisation, then the whole recursive spectral bipartition (degree, Lanczos SpMV
pairs, sign split, modularity, stable row partition) down to the leaves.

  value  : device-resident input (raw counts already in HBM), ms per tree
  e2e    : the same through the host-buffer C-ABI call `tmc_make_tree` (pinned host CSR ->
           H2D -> tree -> D2H of the flat tree), ms per tree
  roofline: SpMV pair kernels (y = C^T x, x' = C y): algorithmic bytes
           / CUDA-event time of those launches inside the timed steps
  tree.hash: canonical SHA-1 of leaf membership + topology (too_many_cells_b200.tree_signature; tree.hash_strict: the same with
           every vertex named by its set of cells, tree_signature_strict); at N > 1 both are asserted equal on every rank and
           equal to the tree one GPU builds alone (`hash_single_gpu`)
  tree.direct_vertices: vertices of >= 2 cells that never iterated in the sweeps -- nodes of <= 64 cells on one rank are solved
           directly (Gram matrix + dense eigen-solve of the whole subtree in one CTA, csrc/tmc_small.cuh)
  cpu_baseline / --impl reference: the C port of the reference's CPU path (oracle/tmc_port.c), MEASURED on a bounded
           sample of the workload; the b200 arm also measures the GPU on that very sample (`same_input`).

Workloads (--workload): c1 .. c4 = BASELINE.json configs[0..3]; c5 = configs[4] (the c3 matrix, min_modularity = -1:
saturated tree, every cell its own leaf).  Default c3 (the config the metric is quoted on).

N>1 (torchrun, one rank per GPU): the tree build is row-sharded across ranks with all-reduces of the
feature-dimension vectors; total work is fixed => "strong" scaling.
"""
from __future__ import annotations

import argparse
import glob
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


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=os.environ.get("TMC_BENCH_WORKLOAD", "c3"), choices=["c1", "c2", "c3", "c4", "c5"])
    ap.add_argument("--cpu-sample-cells", type=int, default=10000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-host-types", default="compact", choices=["compact", "f64"],
                    help="host arrays of the e2e call: uint16 counts / uint16 gene ids where they fit (default), or double / int32")
    ap.add_argument("--no-verify-hash", action="store_true", help="N > 1: skip building the single-GPU tree the N-rank tree is compared with")
    return ap.parse_args()


def workload(name):
    """(synthetic matrix config, engine options, description)"""
    from too_many_cells_b200 import synth
    base = "c3" if name == "c5" else name
    n, g, k, d, s, binary = synth.CONFIGS[base]
    min_mod = -1.0 if name == "c5" else 0.0
    stop = "min-modularity -1 (saturated tree: every cell its own leaf)" if name == "c5" else "default min-modularity 0 stop"
    kind = "binarised scATAC-like" if binary else "10x-like"
    cfg = {"workload": f"{name}: synthetic {kind} {n} cells x {g} features, ~{k} nnz/cell, planted depth {d}, seed {s}, tf-idf + {stop}",
           "cells": n, "features": g, "l2": "inputs larger than L2 (CSR >> 126 MB)" if n * k * 12 > 2.5e8 else
           "matrix smaller than L2; no flush (the tree build re-reads it by design)"}
    eng = {"min_modularity": min_mod}
    if name == "c5":
        eng["max_active"] = 16384      # a saturated level has 10^5 live nodes: more of them per sweep
    return (n, g, k, d, s, binary), eng, cfg


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


# ----------------------------------------------------------------------------- the CPU arm (measured, never extrapolated into `value`)
def cpu_sample(name, sample_cells):
    """The bounded sample both arms use: the workload's generator at `ns` cells (same genes, seed, density, hierarchy).
    c1 is taken whole (it IS the reference's CPU-runnable config); larger workloads are cut to `sample_cells` cells and to
    at most ~2e7 non-zeros so that one CPU step stays within seconds."""
    import scipy.sparse as sp
    from too_many_cells_b200 import synth
    (n, g, k, d, s, binary), eng, _ = workload(name)
    ns = int(min(n, sample_cells, max(500, 2e7 // k)))
    ip, ix, dv, _ = synth.synth_counts(ns, g, k, d, s, binary)
    return sp.csr_matrix((dv, ix, ip), shape=(ns, g)), n, eng


def cpu_reference_run(name, sample_cells, steps=1, warmup=0):
    """The reference's CPU execution of the path, restated in C (oracle/tmc_port.c: one node at a time, materialised
    sub-matrices, single-vector Lanczos at SVDLIBC's kappa = 1e-6, one core like the reference — sequential recursion,
    single-threaded SVDLIBC), timed on the bounded sample.  Returns a dict with the MEASURED ms per step on the sample and,
    separately and labelled as such, a linear-in-cells extrapolation to the full workload."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import tmc_port as P
    a, n_full, eng = cpu_sample(name, sample_cells)
    ns = a.shape[0]
    P._load()          # build / host-compatibility probe of the port happen here, outside the timed region
    times = []
    for it in range(warmup + steps):
        t = time.perf_counter()
        tr = P.make_tree(a, normalize=True, tol=1e-6, min_modularity=eng["min_modularity"])      # (engine-only knobs are not the port's)
        dt = time.perf_counter() - t
        if it >= warmup:
            times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    scale = n_full / ns
    # second figure: the same port with its mat-vecs and re-orthogonalisation on every host core (the reference itself cannot
    # do that); one run
    all_cores = None
    ncpu = min(os.cpu_count() or 1, 32)
    if ncpu > 1:
        t = time.perf_counter()
        P.make_tree(a, normalize=True, tol=1e-6, min_modularity=eng["min_modularity"], threads=ncpu)
        all_cores = {"value": 1e3 * (time.perf_counter() - t), "unit": "ms", "cores": ncpu,
                     "note": "pthread mode of the port (not something the reference can do); same sample, measured"}
        P.make_tree(a[:50], normalize=True, tol=1e-6, threads=1)      # back to the 1-thread pool
    whole = ns == n_full
    desc = (f"C port of the reference's CPU path (oracle/tmc_port.c: sequential recursion, per-node materialised matrices, Lanczos at "
            f"las2's kappa=1e-6, 1 thread like the reference; the Haskell binary cannot be built here) on "
            f"{'the whole workload' if whole else f'a {ns}-cell sample of the workload (same generator, genes, seed)'}: "
            f"{ns} cells x {a.shape[1]} features, {a.nnz} nnz, {tr.n_nodes} nodes, {tr.steps} Lanczos steps, {ms:.0f} ms measured per tree")
    return {"ms": ms, "cells": ns, "nnz": int(a.nnz), "scale": scale, "desc": desc, "all_cores": all_cores, "matrix": a, "eng": eng,
            "whole": whole, "nodes": tr.n_nodes}


def ncu_traffic(name, storage_bits):
    """dram__bytes_read.sum + dram__bytes_write.sum of the root-level SpMV pair launches, read from the newest committed
    ncu summary (profiles/*_root_pair_traffic.json, written by tools/ncu_summary.py from an `ncu --set full` capture of
    tools/prof_pair.py).  None when no capture matches the workload / storage format."""
    best = None
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "*_root_pair_traffic.json"))):
        try:
            d = json.load(open(path))
        except Exception:
            continue
        if d.get("workload") == name and d.get("panel_bits") == storage_bits:
            best = (d, os.path.relpath(path, ROOT))
    if not best:
        return None, None
    d, path = best
    return float(d["pair_dram_bytes"]), {"source": path, "kernels": d.get("kernels")}


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    name = args.workload
    (n, g, k, d, s, binary), eng, cfg = workload(name)

    if args.impl == "reference":
        # Rank 0 alone times the CPU path (the other ranks have nothing to do).  `value` is what was MEASURED: ms per tree on
        # the bounded sample named in config.reference_sample; the linear extrapolation to the full workload is a separate,
        # labelled key and is never reported as the value.
        if rank != 0:
            return
        r = cpu_reference_run(name, args.cpu_sample_cells, steps=max(1, args.steps), warmup=min(args.warmup, 1))
        line = {"impl": "reference", "metric": METRIC, "value": r["ms"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": r["ms"], "higher_is_better": False, "scaling": "strong", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": dict(cfg, reference_sample=("whole workload" if r["whole"] else f"{r['cells']} of {n} cells ({r['nnz']} nnz); value is the measured time of THIS sample"),
                               reference_sample_cells=r["cells"]),
                "cpu_baseline": {"value": r["ms"], "unit": UNIT, "cores": 1, "kind": "port", "sample": r["desc"], "all_cores": r["all_cores"]},
                "extrapolated_full_workload_ms": None if r["whole"] else {"value": r["ms"] * r["scale"], "how": f"measured sample time x {r['scale']:.0f} (linear in cells; "
                                                  "an under-estimate: the full tree is deeper) -- NOT a measurement"},
                "e2e": {"value": r["ms"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
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
    T._lib.load()

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
        tree = T.hierarchical_spectral_cluster(b, profile=profile, **eng)
        b.free()
        return tree

    # N > 1: the tree ONE GPU builds alone (before the communicator exists), to compare the N-rank tree with
    hash_single = None
    if world > 1 and not args.no_verify_hash:
        hash_single = "/".join(step_dev(False).tree_hashes())
    if world > 1:
        import torch.distributed as dist
        from too_many_cells_b200 import dist as tdist
        dist.init_process_group("nccl", device_id=dev)
        tdist.init_comm(rank, world, local_rank, dev)          # libtmc's own NCCL communicator (ids travel over torch.distributed)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_dev(False)
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0 and not os.environ.get("TMC_BENCH_NO_SAMPLER"):
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

    # the tree of every timed step on every rank must be THE tree: same canonical hash everywhere, equal to the single-GPU one
    # (both signatures: the (smallest row, size) one of the earlier rounds' records and the strict one that names every vertex by its
    # set of cells; "legacy/strict")
    hashes = sorted({"/".join(t_.tree_hashes()) for t_ in trees})
    if world > 1:
        allh = [None] * world
        torch.distributed.all_gather_object(allh, hashes)
        hashes = sorted({h for hs in allh for h in hs})
    if len(hashes) != 1:
        raise SystemExit(f"bench.py: ranks / steps disagree about the tree: {hashes}")
    if hash_single is not None and hashes[0] != hash_single:
        raise SystemExit(f"bench.py: the {world}-rank tree {hashes[0]} differs from the single-GPU tree {hash_single}")

    # ---- end-to-end through the host-buffer C-ABI (every rank participates when N>1)
    # host CSR: the C ABI reads the full row_ptr and only this rank's block of col/val (then replicates over NVLink), so the
    # arrays are full-size but only the own block is backed by (pinned) memory
    # Host element types: the make-tree input is a matrix of UMI counts, which the shim marshals as uint16 (and gene indices as
    # uint16 when they fit) -- tmc_csr.val_type / idx_type; libtmc widens them on the device, results are bit-identical to the
    # fp64 call (tests/test_gpu_panel.py).  --e2e-host-types f64 times the plain double / int32 arrays instead.
    h_ptr = row_ptr.cpu().pin_memory()
    k0, k1 = int(h_ptr[lo]), int(h_ptr[hi])
    compact = args.e2e_host_types == "compact" and float(val_raw.max().item()) <= 65535.0
    vdt = np.uint16 if compact else np.float64
    cdt = np.uint16 if (compact and g <= 65536) else np.int32
    np_col = np.empty(nnz, dtype=cdt); np_val = np.empty(nnz, dtype=vdt)      # untouched pages stay uncommitted (N > 1: own block only)
    CH = 1 << 27
    for o in range(k0, k1, CH):
        e = min(k1, o + CH)
        np_col[o:e] = col[o:e].cpu().numpy().astype(cdt); np_val[o:e] = val_raw[o:e].cpu().numpy().astype(vdt)
    rt = torch.cuda.cudart()
    if k1 > k0:
        rt.cudaHostRegister(np_col[k0:k1].ctypes.data, (k1 - k0) * np_col.itemsize, 0)      # pinned host memory
        rt.cudaHostRegister(np_val[k0:k1].ctypes.data, (k1 - k0) * np_val.itemsize, 0)
    host = (h_ptr.numpy(), np_col, np_val, (n, g))
    e2e_steps = max(1, args.steps)
    tr_e = T.hierarchical_spectral_cluster(host, normalize=True, device=local_rank, **eng)     # warm
    barrier()
    t = time.perf_counter()
    for _ in range(e2e_steps):
        tr_e = T.hierarchical_spectral_cluster(host, normalize=True, device=local_rank, **eng)
    barrier()
    e2e_ms = 1e3 * (time.perf_counter() - t) / e2e_steps
    if "/".join(tr_e.tree_hashes()) != hashes[0]:
        raise SystemExit("bench.py: the host-buffer call built a different tree than the device-resident call")
    h2d = int(h_ptr.numel() * 8 + (k1 - k0) * (np_col.itemsize + np_val.itemsize))
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
    bv = float(tr.stats["value_bytes"])          # 8: fp64 B; 2: raw counts as uint16 + diagonal factors; 0: all-ones matrix
    nnz_pairs = sum(t_.stats["spmv_pairs_nnz"] for t_ in trees)
    alg_nominal = alg + 2.0 * (8.0 - bv) * nnz_pairs
    achieved_nominal = alg_nominal / (spmv_ms * 1e-3) / 1e9 if spmv_ms > 0 else None
    # root-level pair in isolation (largest single launch pair) -- single-rank only
    root_pair, traffic, traffic_src = None, None, None
    if world == 1:
        work.copy_(val_raw)
        b = T.DeviceB.adopt_device(n, g, row_ptr, col, work, normalize=True, device=local_rank)
        sc_ms, ga_ms = T.bench_spmv_pair(b, 10)
        storage = b.info()
        b.free()
        # stored bytes of one pair: (value + column index) per residual CSR entry and half, the panel cells once per half
        pair_bytes = (2.0 * (bv + 4.0) * storage["residual_nnz"] + 2.0 * n * storage["panel_cols"] * storage["panel_bits"] / 8.0
                      + 16.0 * (n + 1) + 32.0 * n + 16.0 * g)
        root_pair = {"scatter_ms": sc_ms, "gather_ms": ga_ms, "bytes": pair_bytes, "storage": storage,
                     "scatter_gbs": 0.5 * pair_bytes / (sc_ms * 1e-3) / 1e9, "gather_gbs": 0.5 * pair_bytes / (ga_ms * 1e-3) / 1e9}
        traffic, traffic_src = ncu_traffic("c3" if name == "c5" else name, storage["panel_bits"])
    roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
            "traffic": traffic, "traffic_scope": "root-level pair (largest launches), ncu dram__bytes_read.sum + dram__bytes_write.sum" if traffic else None,
            "traffic_source": traffic_src, "peak_source": peak_src,
            "kernel": "SpMV pair: y = C^T x (residual + dense panel kernels), x' = C y (panel + residual CSR kernels); all launches of the timed steps",
            "algorithmic_bytes_per_step": alg / len(trees), "value_bytes_b_v": tr.stats["value_bytes"],
            "achieved_if_counted_with_fp64_values_b_v8": achieved_nominal,
            "frac_if_counted_with_fp64_values_b_v8": (achieved_nominal / peak) if achieved_nominal else None,
            "algorithmic_bytes_formula": "per SpMV pair 2*nnz_residual*(b_v+4) + 2*m*P*c + 2*(m+1)*8 + (4*m+2*n)*8 (BASELINE.md section 4 with the storage actually used: b_v = bytes per stored value, P = dense panel columns at one byte or half a byte per cell)",
            "spmv_ms_per_step": spmv_ms / len(trees), "spmv_share_of_step": spmv_ms / len(trees) / ms_step,
            "per_gpu": True, "root_pair": root_pair}
    line = {"metric": METRIC, "value": ms_step, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": dict(cfg, nnz=nnz, parallelism=("1 GPU" if world == 1 else
                                             f"{world} GPUs: matrix replicated, shared top of the tree row-sharded with an all-reduce of y = B^T x "
                                             f"per half-step, subtrees below the hand-off size owned by single GPUs (no collectives)")),
            "clocks": clocks,
            "e2e": {"value": e2e_ms, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                    "host_types": f"row_ptr int64, col_idx {np_col.dtype.name}, val {np_val.dtype.name} (pinned)"},
            "gpu_launches": launches, "roofline": roof,
            "tree": {"nodes": tr.n_nodes, "leaves": int(len(tr.leaves())), "sweeps": tr.stats["n_sweeps"],
                     "lanczos_steps": int(tr.iters.sum()), "engine_ms_last": tr.stats["engine_ms"], "gen_s": gen_s,
                     # vertices of >= 2 cells that never iterated: solved directly, Gram matrix + dense eigen-solve (tmc_small.cuh)
                     "direct_vertices": int(((tr.iters == 0) & ((tr.item_end - tr.item_begin) >= 2)).sum()),
                     "hash": hashes[0].split("/")[0], "hash_strict": hashes[0].split("/")[1],
                     "hash_single_gpu": hash_single.split("/")[0] if hash_single else None,
                     "hash_check": ("all ranks and steps equal" + (", equal to the single-GPU tree" if hash_single else "")) }}
    if not args.no_cpu_baseline and world == 1:          # reported on rank 0 at N=1 only
        r = cpu_reference_run(name, args.cpu_sample_cells)
        # the GPU on the very same input, so that one CPU/GPU pair is measured-vs-measured (host CSR -> tree, and device-resident)
        a = r["matrix"]
        T.hierarchical_spectral_cluster(a, normalize=True, device=local_rank, **eng)
        t = time.perf_counter()
        for _ in range(3):
            tg = T.hierarchical_spectral_cluster(a, normalize=True, device=local_rank, **eng)
        gpu_e2e = 1e3 * (time.perf_counter() - t) / 3
        db = T.DeviceB.from_host(a, normalize=True, device=local_rank)
        torch.cuda.synchronize()
        t = time.perf_counter()
        for _ in range(3):
            T.hierarchical_spectral_cluster(db, **eng)
        gpu_dev = 1e3 * (time.perf_counter() - t) / 3
        db.free()
        line["cpu_baseline"] = {"value": r["ms"] * r["scale"], "unit": UNIT, "cores": 1, "kind": "port", "sample": r["desc"] + (
                                    "" if r["whole"] else f"; value = measured x {r['scale']:.0f} (scaled linearly in cells to the metric's unit)"),
                                "measured_ms": r["ms"], "sample_cells": r["cells"], "scale": r["scale"], "all_cores": r["all_cores"],
                                "same_input": {"cells": r["cells"], "nnz": r["nnz"], "cpu_ms": r["ms"], "gpu_e2e_ms": gpu_e2e, "gpu_device_resident_ms": gpu_dev,
                                               "gpu_nodes": tg.n_nodes, "cpu_nodes": r["nodes"],
                                               "note": "both arms measured on this box on the identical matrix (no scaling)"}}
    print(json.dumps(line))
    if world > 1:
        tdist.destroy_comm()
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
