"""SURVEY.md 8(f) N1 — the steps right before the hot path, through libtmc on the GPU.

  load_cellranger_data   src/TooManyCells/Matrix/Load.hs:85-152 (10x directory: matrix.mtx[.gz], features.tsv[.gz] /
                         genes.tsv, barcodes.tsv[.gz]; the MTX file is features x cells and is transposed to cells x features)
  filter_num_sparse_mat  src/TooManyCells/Matrix/Preprocess.hs:342-386 (--filter-thresholds "(MINCELL, MINFEATURE)")
  merge_single_cells     src/TooManyCells/Matrix/Types.hs:214-254 (several --matrix-path: `mconcat` of SingleCells)
  write_matrix_like      src/TooManyCells/Matrix/Utility.hs:220-250 (--matrix-output FOLDER: the 10x directory written back)
"""
from __future__ import annotations

import ctypes as C
import gzip
import os

import numpy as np

from . import _lib
from ._lib import CsrOut, check
from .spectral import HostCsr


def _csr_from_out(op):
    import scipy.sparse as sp
    o = op.contents
    m, n, nnz = int(o.n_rows), int(o.n_cols), int(o.nnz)
    arr = lambda p, k, dt: (np.ctypeslib.as_array(p, shape=(k,)).astype(dt, copy=True) if k else np.zeros(0, dt))
    mat = sp.csr_matrix((arr(o.val, nnz, np.float64), arr(o.col_idx, nnz, np.int32), arr(o.row_ptr, m + 1, np.int64)), shape=(m, n))
    kr = arr(o.kept_rows, int(o.n_kept_rows), np.int64) if o.kept_rows else None
    kc = arr(o.kept_cols, int(o.n_kept_cols), np.int64) if o.kept_cols else None
    _lib.load().tmc_csr_out_free(op)
    return mat, kr, kc


def parse_matrix_market(data: bytes, transpose=True):
    """MatrixMarket coordinate text -> scipy CSR (cells x features when transpose=True, as loadCellrangerData does)."""
    lib = _lib.load()
    op = C.POINTER(CsrOut)()
    check(lib.tmc_mtx_parse(data, len(data), int(bool(transpose)), C.byref(op)))
    return _csr_from_out(op)[0]


def _read(path):
    with open(path, "rb") as f:
        raw = f.read()
    return gzip.decompress(raw) if path.endswith(".gz") else raw


def _first_existing(d, names):
    for nm in names:
        p = os.path.join(d, nm)
        if os.path.exists(p):
            return p
    raise FileNotFoundError(f"none of {names} in {d}")


def load_cellranger_data(folder, feature_column=1):
    """Returns (matrix cells x features, barcodes, features) like SingleCells{_matrix,_rowNames,_colNames}.
    `feature_column` is 1-based as the reference's --feature-column."""
    mtx = _first_existing(folder, ["matrix.mtx.gz", "matrix.mtx"])
    feat = _first_existing(folder, ["features.tsv.gz", "features.tsv", "genes.tsv.gz", "genes.tsv"])
    bar = _first_existing(folder, ["barcodes.tsv.gz", "barcodes.tsv"])
    mat = parse_matrix_market(_read(mtx), transpose=True)
    features = [ln.split("\t")[min(feature_column - 1, len(ln.split("\t")) - 1)] for ln in _read(feat).decode().splitlines() if ln]
    barcodes = [ln.split("\t")[0] for ln in _read(bar).decode().splitlines() if ln]
    if mat.shape != (len(barcodes), len(features)):
        raise ValueError(f"matrix is {mat.shape} but there are {len(barcodes)} barcodes and {len(features)} features")
    return mat, barcodes, features


def filter_num_sparse_mat(mat, row_thresh, col_thresh):
    """filterNumSparseMat: returns (filtered matrix, kept row ids, kept column ids)."""
    lib = _lib.load()
    h = HostCsr(mat)
    op = C.POINTER(CsrOut)()
    check(lib.tmc_filter_thresholds(C.byref(h.c), float(row_thresh), float(col_thresh), C.byref(op)))
    return _csr_from_out(op)


def merge_single_cells(parts):
    """`mconcat` of SingleCells (Semigroup instance, Matrix/Types.hs:214-254) for several `--matrix-path`: `parts` is a list of
    (matrix, barcodes, features).  Matrices that share their feature list are stacked vertically as they are; otherwise the
    features become the ascending union of the names (`Set.toAscList . Set.union`) and every matrix is re-indexed into it
    (`mergeFeaturesSingleCells`).  Cells keep the order of the inputs.  Index arithmetic only: no value is touched."""
    import scipy.sparse as sp
    parts = [p for p in parts if p is not None]
    if not parts:
        raise ValueError("no matrices to merge")
    mat, barcodes, features = parts[0]
    mat, barcodes, features = sp.csr_matrix(mat), list(barcodes), list(features)
    for m2, b2, f2 in parts[1:]:
        m2, f2 = sp.csr_matrix(m2), list(f2)
        if not features or not f2 or features == f2:
            features = features or f2
            if mat.shape[1] != m2.shape[1]:
                raise ValueError("matrices with equal feature lists must have equal widths")
        else:
            if len(set(features)) != len(features) or len(set(f2)) != len(f2):
                raise ValueError("mergeFeaturesSingleCells: duplicated feature names cannot be merged")
            new = sorted(set(features) | set(f2))
            pos = {name: j for j, name in enumerate(new)}

            def reindex(m, names):
                m = m.tocoo()
                cols = np.fromiter((pos[nm] for nm in names), dtype=np.int64, count=len(names))
                return sp.csr_matrix((m.data, (m.row, cols[m.col])), shape=(m.shape[0], len(new)))
            mat, m2, features = reindex(mat, features), reindex(m2, f2), new
        mat = sp.vstack([mat, m2], format="csr")
        barcodes += list(b2)
    mat.sort_indices()
    return mat, barcodes, features


def load_all(folders, feature_column=1):
    """Every `--matrix-path` loaded on the GPU and merged (loadAllSSM's `mconcat`, Program/LoadMatrix.hs:213-217)."""
    return merge_single_cells([load_cellranger_data(f, feature_column) for f in folders])


def write_matrix_like(folder, mat, barcodes, features):
    """`writeSparseMatrixLike (MatrixTranspose False)` (Matrix/Utility.hs:220-250; `--matrix-output FOLDER`): the cells x features
    matrix written back as a 10x directory — `matrix.mtx.gz` with features as rows ("like input"), `features.tsv.gz` with the
    reference's three columns `name<TAB>name<TAB>NA`, `barcodes.tsv.gz`.  Coordinate / real / general, 1-based, one entry per line in
    (cell, feature) order, values in shortest round-trip notation, so that `load_cellranger_data` reads back the same matrix."""
    import scipy.sparse as sp
    m = sp.csr_matrix(mat)
    m.sort_indices()
    if m.shape != (len(barcodes), len(features)):
        raise ValueError(f"matrix is {m.shape} but there are {len(barcodes)} barcodes and {len(features)} features")
    os.makedirs(folder, exist_ok=True)
    rows = np.repeat(np.arange(m.shape[0], dtype=np.int64), np.diff(m.indptr))
    lines = ["%%MatrixMarket matrix coordinate real general", f"{m.shape[1]} {m.shape[0]} {m.nnz}"]
    lines += [f"{j + 1} {i + 1} {repr(float(v))}" for i, j, v in zip(rows.tolist(), m.indices.tolist(), m.data.tolist())]
    with gzip.open(os.path.join(folder, "matrix.mtx.gz"), "wt") as f:
        f.write("\n".join(lines) + "\n")
    with gzip.open(os.path.join(folder, "features.tsv.gz"), "wt") as f:
        f.write("".join(f"{x}\t{x}\tNA\n" for x in features))
    with gzip.open(os.path.join(folder, "barcodes.tsv.gz"), "wt") as f:
        f.write("".join(f"{x}\n" for x in barcodes))
