"""Deterministic synthetic 10x-like count matrices.

Planted binary hierarchy of 2^depth leaf programmes: gene base profile ~ Gamma(0.3);
every branch of the hierarchy up-weights a random 2 % of the genes by U(3,8);
per-cell library size ~ LogNormal(mu, 0.3) with mu solved for the target density;
counts ~ Poisson(library * profile) (the multinomial's Poisson limit); every row
is guaranteed non-empty.  `synth_counts` is the numpy/CPU generator used by tests
and the c1 config; `synth_counts_torch` is the same recipe on a torch device for the
large configs (c2+), generated in row blocks so that 1M x 30k fits.
"""
from __future__ import annotations

import math

import numpy as np

CONFIGS = {
    # name: (cells, genes, nnz_per_cell, depth, seed, binary)
    "tiny": (600, 2000, 120, 3, 7, False),
    "small": (2000, 5000, 250, 4, 11, False),
    "c1": (5000, 20000, 1000, 4, 1, False),
    "c2": (100000, 30000, 2000, 8, 2, False),
    "c3": (1000000, 30000, 2000, 10, 3, False),
    "c4": (200000, 500000, 7500, 8, 4, True),
}


def _programmes(rng, n_genes, depth):
    base = rng.gamma(0.3, 1.0, size=n_genes) + 1e-12
    n_leaf = 1 << depth
    prof = np.tile(base, (n_leaf, 1))
    for level in range(depth):
        n_branch = 1 << (level + 1)
        span = n_leaf // n_branch
        for br in range(n_branch):
            genes = rng.choice(n_genes, size=max(1, n_genes // 50), replace=False)
            mult = rng.uniform(3.0, 8.0, size=genes.size)
            prof[br * span:(br + 1) * span, genes] *= mult
    prof /= prof.sum(axis=1, keepdims=True)
    return prof


def _solve_library(prof_mean, target_nnz):
    lo, hi = 1.0, 1e9
    for _ in range(200):
        mid = math.sqrt(lo * hi)
        nnz = float(np.sum(1.0 - np.exp(-mid * prof_mean)))
        if nnz < target_nnz:
            lo = mid
        else:
            hi = mid
    return math.sqrt(lo * hi)


def synth_counts(n_cells, n_genes, nnz_per_cell, depth, seed, binary=False, block=1024):
    """Returns (indptr int64, indices int32, data float64, leaf_of_cell) of a cells x genes CSR of raw counts."""
    rng = np.random.default_rng(seed)
    prof = _programmes(rng, n_genes, depth)
    lib0 = _solve_library(prof.mean(axis=0), nnz_per_cell)
    leaf = rng.integers(0, prof.shape[0], size=n_cells)
    lib = lib0 * rng.lognormal(0.0, 0.3, size=n_cells)
    indptr = np.zeros(n_cells + 1, dtype=np.int64)
    idx_parts, val_parts = [], []
    for b0 in range(0, n_cells, block):
        b1 = min(n_cells, b0 + block)
        lam = prof[leaf[b0:b1]] * lib[b0:b1, None]
        cnt = rng.poisson(lam)
        empty = cnt.sum(axis=1) == 0
        if empty.any():
            cnt[empty, np.argmax(lam[empty], axis=1)] = 1
        r, c = np.nonzero(cnt)
        indptr[b0 + 1:b1 + 1] = np.bincount(r, minlength=b1 - b0)
        idx_parts.append(c.astype(np.int32))
        v = cnt[r, c].astype(np.float64)
        val_parts.append(np.ones_like(v) if binary else v)
    np.cumsum(indptr, out=indptr)
    return indptr, np.concatenate(idx_parts), np.concatenate(val_parts), leaf


def synth_config(name):
    n, g, k, d, s, binary = CONFIGS[name]
    return synth_counts(n, g, k, d, s, binary)


def synth_counts_torch(n_cells, n_genes, nnz_per_cell, depth, seed, device, binary=False, block=None, row_range=None):
    """Same recipe on a torch device, seeded PER ROW BLOCK so that any rank can generate exactly its own rows:
    rows [lo, hi) = row_range (default: all).  Returns (row_ptr int64, col int32, val float64) tensors on `device`
    for the local rows (row_ptr rebased to 0)."""
    import torch
    if block is None:          # ~1.2e8 dense Poisson draws per block (fixed by the shape, so every rank uses the same blocks)
        block = max(64, min(4096, (1 << 27) // n_genes // 64 * 64))
    rng = np.random.default_rng(seed)
    prof = _programmes(rng, n_genes, depth)
    lib0 = _solve_library(prof.mean(axis=0), nnz_per_cell)
    lo, hi = (0, n_cells) if row_range is None else row_range
    prof_t = torch.from_numpy(prof).to(device=device, dtype=torch.float32)
    g = torch.Generator(device=device)
    counts_per_row = torch.zeros(hi - lo + 1, dtype=torch.int64, device=device)
    cols, vals = [], []
    for bi in range(lo // block, (hi + block - 1) // block):
        b0, b1 = bi * block, min(n_cells, (bi + 1) * block)
        g.manual_seed(seed * 1000003 + bi)
        leaf = torch.randint(0, prof.shape[0], (b1 - b0,), generator=g, device=device)
        lib = lib0 * torch.exp(0.3 * torch.randn(b1 - b0, generator=g, device=device))
        lam = prof_t[leaf] * lib[:, None].float()
        cnt = torch.poisson(lam, generator=g)
        empty = cnt.sum(dim=1) == 0
        if bool(empty.any()):
            er = torch.nonzero(empty).flatten()
            cnt[er, lam[er].argmax(dim=1)] = 1
        s0, s1 = max(lo, b0) - b0, min(hi, b1) - b0          # the part of this block inside [lo, hi)
        cnt = cnt[s0:s1]
        nz = torch.nonzero(cnt)
        counts_per_row[b0 + s0 - lo + 1:b0 + s1 - lo + 1] = torch.bincount(nz[:, 0], minlength=s1 - s0)
        cols.append(nz[:, 1].to(torch.int32))
        v = cnt[nz[:, 0], nz[:, 1]].to(torch.float64)
        vals.append(torch.ones_like(v) if binary else v)
        del lam, cnt, nz
    row_ptr = torch.cumsum(counts_per_row, 0)
    return row_ptr, torch.cat(cols), torch.cat(vals)
