// sm_100a kernels of the spectral-bipartition engine.
//
// Data model (DESIGN.md "HBM layout"): one immutable CSR of B (row-L2-normalised
// tf-idf values) and a permutation `perm` of row ids; every live tree node owns a
// contiguous position range [begin,end) of `perm` (quicksort-style), so all
// per-cell vectors (degree, Lanczos basis, labels) are position-indexed and a
// single launch advances EVERY live node by one step ("sweep").  No matrix is
// ever re-materialised per node, unlike the reference's IntMap copies
// (SURVEY.md 8(a) A3/A8).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "tmc_tridiag.cuh"

namespace tmc {

enum ValType : int { VT_F64 = 0, VT_U16 = 1, VT_ONE = 2 };
// ST_KM (KMeansGroup only): Lloyd rounds of the 2-means on u2, one per sweep, between the eigen-solve and the modularity step
enum SlotState : int { ST_IDLE = 0, ST_DEG = 1, ST_ORTH0 = 2, ST_LAN = 3, ST_KM = 4, ST_MOD = 5, ST_DONE = 6 };
enum SlotAction : int { ACT_NONE = 0, ACT_INIT = 1, ACT_NEXT = 2, ACT_RITZ_LABEL = 3, ACT_RITZ_RESTART = 4, ACT_FINISH = 5, ACT_RITZ_MULTI = 6,
                        ACT_KM_ASSIGN = 7, ACT_KM_DONE = 8 };
// per-slot scalars of the 2-means (Ctx::km): centroids, sums of the last assignment, label changes, round, min / max keys of u2
enum { KM_C0 = 0, KM_C1 = 1, KM_S0 = 2, KM_S1 = 3, KM_CHANGED = 4, KM_ROUND = 5, KM_MINKEY = 6, KM_MAXKEY = 7, KM_N = 8 };

// scalar slots at the tail of a reduction row
enum { R_NRM2 = 0, R_SUMD = 1, R_YNORM2 = 2, R_N1 = 3, R_SD1 = 4, R_SGN = 5, R_NSCAL = 8 };
// The scalars live at the tail of the pass-1 coefficient row so that ONE all-reduce per sweep carries both.

struct Slot {              // one live tree node (device resident)
  int begin, end;          // this rank's position range in its local perm (rows are owned statically per rank)
  int gsize;               // global number of cells in the node (sum over ranks)
  int n0_local;            // label-0 cells of this rank (set by k_partition)
  int state, action;
  int j;                   // Lanczos vectors q_1..q_j stored in V[1..j]
  int iters, restarts;
  int node;                // host node id
  int split, n0;
  int stop_after;          // 0 = full lifecycle, else stop when reaching this state (fine-grained API)
  int tpar;                // which of the two transposed buffers holds this node's column-sorted block
  int priv;                // 1 = the whole node lives on this rank (handed-off subtree): no collectives, gsize = local size
  int nev;                 // > 0: truncated-SVD mode: top `nev` left singular vectors of B, no deflation, no split
  long long tb, te;        // the block: entries [tb, te) of tcol/tpos/tval[tpar], sorted by column
  long long tnnz0;         // entries of label-0 cells (set by k_tpart_scan): children blocks are [tb, tb+tnnz0) and [tb+tnnz0, te)
  int tchunk0, ntchunk;    // this sweep's chunks of the block (set by k_prepare)
  double sumd, inv_sqrt_sumd, beta, theta, resid, q, theta_hint;
  // sign-certified stop (DESIGN.md "sign-certified stop"): cert_req = this sweep probes min|x_i| of the tentative Ritz vector;
  // cert_need = SAFE * r / gap it has to exceed; cert_thr = r / gap below which the next probe is worth its pass over V
  int cert_req;
  int deg_phase;           // signed matrices: 0 = degree pass on |B| (d, for D^{-1/2}), 1 = on B itself (d2, for the modularity)
  double cert_need, cert_thr;
  double sumd2;            // signed matrices: sum of d2 = ||B^T 1||^2
  // warm start (DESIGN.md "warm start"): warm = this node's start vector comes from its parent's Krylov space (Ctx::ws) instead
  // of the hash; have_z2 = zvec2 holds the Ritz vector of T's SECOND value at this node's stop step (the children's start)
  int warm, have_z2;
  double cert_k;           // warm-started node: how far its start vector favoured u2 over its random share (k_finish); the
                           // certificate's safety factor of small nodes is scaled by it (cert_passed)
  int flip;                // canonical eigenvector sign: the labels written at the stop step are to be inverted before the split
};

struct Report {            // what the host reads back after every sweep
  int state, split, n0, iters;   // n0 = global label-0 count
  int n0_local, have_z2;         // have_z2: Ctx::ws holds a start vector for the children
  long long tnnz0;
  double q, theta;
};

#define TMC_MAX_CHUNK 256
struct Chunk { int slot, begin, len; };
struct TChunk { long long begin; int slot, pad; };        // entries [begin, begin + tchunk) of the slot's transposed block
struct Activation { int slot, begin, end, gsize, node, stop_after, tpar, priv; long long tb, te; int nev, warm; };

struct Ctx {
  // matrix B (CSR, rows = cells)
  const int64_t* row_ptr; const int32_t* col; const void* val;   // val: double (B itself) | uint16 | absent, see vt
  int vt;                          // VT_F64: val = B.  VT_U16 / VT_ONE: B = diag(rrow) * A * diag(wcol), A = small integer counts / all ones
  const double* rrow;              // [n_rows] 1/||row of A diag(w)||  (factored modes)
  const double* w2col;             // [n_cols] w_j^2
  const double* invw;              // [n_cols] 1/w_j (0 where w_j = 0)
  int n_rows, n_cols;              // local rows, features
  int64_t row_offset;              // global id of local row 0
  // dense panel (factored modes): the P most frequent columns are renumbered to [0, P) and stored as one byte per (row, column),
  // panel + row * pw; row_ptr/col/val then hold only the residual (other columns, counts that do not fit).  P = 0: no panel.
  const uint8_t* panel; int P, Ppad;      // Ppad: P rounded up to whole warps of 512 columns
  int pnib, pw;                            // pnib: 1 = one nibble per cell (counts <= 15), 0 = one byte (<= 255); pw: bytes per panel row
  const int64_t* row_ptr0;         // row pointers of the caller's CSR (nnz accounting only)
  double* pacc;                    // [m] panel half of u = A y per position, consumed by k_gather_u
  // position-indexed state
  int* perm; int* perm_tmp;
  double* d; double* sdeg; double* xin; double* w; double* u2; uint8_t* label;
  // signed matrices (entries < 0: a centred / PCA-reduced input; fp64 storage only).  The reference takes d = |B| (|B|^T 1) for
  // D^{-1/2} but B itself everywhere else, so D^{1/2} 1 is no longer a singular vector of C: no deflation, u2 = the SECOND Ritz
  // vector, and the modularity needs the degrees of B itself (d2) from a second degree pass.
  int signed_mode; double* d2;
  double* V; int64_t vstride;      // basis arena V[l*vstride + p], l = 0 is the locked u1
  double* Y; int64_t ystride;      // per-slot feature vectors y = B_S^T x
  // per-slot state
  Slot* st; Report* rep; int max_active;
  double* redA; double* redB;      // CGS pass-0 / pass-1 coefficients h_l = <V_l, w>: redA[slot][K+2], redB[slot][K+2+R_NSCAL]
  int n_slots;                     // slots [0, n_slots) may be live this sweep
  double* alpha; double* beta; double* zvec; double* scratch;   // [slot][K+2] (scratch: 2*(K+2))
  int K;
  int* list_offs;                                    // [3][max_active+1] per-slot offsets into the three chunk lists
  Chunk* chunks; int* n_chunks; int chunk_rows;      // SpMV kernels: few rows per CTA (memory-level parallelism)
  Chunk* vchunks; int* n_vchunks; int vchunk_rows;   // vector kernels (dots/update/advance): up to 256 positions per CTA
  // transposed (column-sorted) per-node copy of B: y = B_S^T x as a gather with run merging instead of nnz atomics
  int use_t;
  // residual gather with the dense vector resident in shared memory (k_gather_smem): nodes of at least gs_min_rows local rows,
  // when the residual part of y (n_cols - P doubles) fits one CTA's shared memory; their rows leave the plain chunk list
  int gs_on, gs_min_rows, gs_chunk_rows;
  int ts_on;                                         // the same nodes form y[P, n) from their CSR rows into shared memory (k_tscatter_smem)
  Chunk* gchunks; int* n_gchunks;
  int xs_min_rows;                                   // nodes with at least this many local rows read x in the strided mapping
  int32_t* tcol[2]; int32_t* tpos[2]; void* tval[2];   // tval element type follows vt
  TChunk* tchunks; int* n_tchunks; int tchunk;       // entries per chunk (multiple of 2048)
  int* tile_cnt; long long* tile_off;                // per chunk: label-1 entries / exclusive prefix inside the node
  int* newpos;                                       // [m] where k_partition moved every position
  int* tb_cnt; long long* tb_colstart;               // [n_cols(+1)] scratch for building a block from CSR rows
  int* err;                        // device error flag
  unsigned long long* stats;       // [0] nnz gathered in Lanczos pairs, [1] nnz gathered in degree pairs
  // options
  double tol, min_modularity; int min_size, max_restarts; uint64_t seed;
  int rank, world;
  int collectives;                 // 0 once only private (handed-off) nodes remain: ranks run free
  // sign-certified stop: certmin[slot] = min_i |x_i| of the probed Ritz vector (bit pattern of a non-negative double, atomicMin)
  unsigned long long* certmin; int cert_on, cert_min_rows; double cert_safe;
  // KMeansGroup with one eigenvector inside the sweeps (Options.hs:172-173): labels from Lloyd's 2-means on u2 instead of its sign
  int kmeans; double* km;          // km[slot][KM_N]
  // truncated-SVD mode (one slot): Ritz vectors of T for the top nev values [nev][K+2], output U [m][nev] (cells x nev), sigma^2 [nev]
  double* zmulti; double* uout; double* ev_out; int* svd_conv;
  // warm start: a node that splits leaves f = D^{-1/2} (V z2) on its positions (z2 = Ritz vector of the second Ritz value of T,
  // i.e. the next singular direction after u2), k_partition moves it with the cells, and a child starts from D_child^{1/2} f
  int warm_on, warm_min_rows; double* ws; double* ws_tmp; double* zvec2;
  int canon_sign;                  // canonical sign of u2 (k_finish, ST_MOD)
  int cert_warm_scale;             // scale the certificate's safety factor of warm-started nodes by their start bias (k_finish)
  double cert_s0, warm_mix, cert_kfac;
};

__device__ __forceinline__ double* slot_scal(const Ctx& c, int slot) { return c.redB + (int64_t)slot * (c.K + 2 + R_NSCAL) + (c.K + 2); }
__device__ __forceinline__ double* slot_redB(const Ctx& c, int slot) { return c.redB + (int64_t)slot * (c.K + 2 + R_NSCAL); }

// Value access by storage type.  In the factored modes the kernels work with A (raw counts) and the diagonal factors
// ride on the vectors: xin carries rrow, y is kept as y' = w^2 (A^T xin) so the gather needs no column factor.
template <int VT> __device__ __forceinline__ double ld_val(const void* base, int64_t k) {
  if (VT == VT_F64) return __ldg(reinterpret_cast<const double*>(base) + k);
  if (VT == VT_U16) return (double)__ldg(reinterpret_cast<const unsigned short*>(base) + k);
  return 1.0;
}
template <int VT> __device__ __forceinline__ void ld_val4(const void* base, long long k0, double* t) {   // k0 % 4 == 0
  if (VT == VT_F64) {
    const double2 a = __ldg(reinterpret_cast<const double2*>(reinterpret_cast<const double*>(base) + k0));
    const double2 b = __ldg(reinterpret_cast<const double2*>(reinterpret_cast<const double*>(base) + k0 + 2));
    t[0] = a.x; t[1] = a.y; t[2] = b.x; t[3] = b.y;
  } else if (VT == VT_U16) {
    const ushort4 a = __ldg(reinterpret_cast<const ushort4*>(reinterpret_cast<const unsigned short*>(base) + k0));
    t[0] = a.x; t[1] = a.y; t[2] = a.z; t[3] = a.w;
  } else { t[0] = t[1] = t[2] = t[3] = 1.0; }
}
template <int VT> __device__ __forceinline__ void ld_val2(const void* base, long long k0, double* t) {   // k0 % 2 == 0
  if (VT == VT_F64) {
    const double2 a = __ldg(reinterpret_cast<const double2*>(reinterpret_cast<const double*>(base) + k0));
    t[0] = a.x; t[1] = a.y;
  } else if (VT == VT_U16) {
    const ushort2 a = __ldg(reinterpret_cast<const ushort2*>(reinterpret_cast<const unsigned short*>(base) + k0));
    t[0] = a.x; t[1] = a.y;
  } else { t[0] = t[1] = 1.0; }
}
template <int E> __device__ __forceinline__ void ld_idx(const int32_t* p, int* out) {     // E consecutive int32, p aligned to E*4 bytes
  if (E == 2) { const int2 a = __ldg(reinterpret_cast<const int2*>(p)); out[0] = a.x; out[1] = a.y; }
  else {
#pragma unroll
    for (int q = 0; q < E / 4; ++q) { const int4 a = __ldg(reinterpret_cast<const int4*>(p) + q); out[4 * q] = a.x; out[4 * q + 1] = a.y; out[4 * q + 2] = a.z; out[4 * q + 3] = a.w; }
  }
}
template <int E, int VT> __device__ __forceinline__ void ld_vals(const void* base, long long k0, double* t) {
  if (E == 2) ld_val2<VT>(base, k0, t);
  else {
#pragma unroll
    for (int q = 0; q < E / 4; ++q) ld_val4<VT>(base, k0 + 4 * q, t + 4 * q);
  }
}
template <int VT> __device__ __forceinline__ void st_val(void* base, long long k, double v) {
  if (VT == VT_F64) reinterpret_cast<double*>(base)[k] = v;
  else if (VT == VT_U16) reinterpret_cast<unsigned short*>(base)[k] = (unsigned short)v;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-wide sum for 256-thread CTAs; result valid on thread 0
__device__ __forceinline__ double block_sum_256(double v, double* sm8) {
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) sm8[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x < 32) {
    t = (threadIdx.x < (blockDim.x >> 5)) ? sm8[threadIdx.x] : 0.0;
    t = warp_sum(t);
  }
  __syncthreads();
  return t;
}

__device__ __forceinline__ unsigned long long f64_key(double v) {       // order-preserving map double -> uint64
  const unsigned long long b = (unsigned long long)__double_as_longlong(v);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double key_f64(unsigned long long k) {
  return __longlong_as_double((long long)((k >> 63) ? (k & 0x7fffffffffffffffull) : ~k));
}
__device__ __forceinline__ double hash_unit(uint64_t seed, uint64_t row) {
  uint64_t x = seed + (row + 1) * 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  x = x ^ (x >> 31);
  return (double)(x >> 11) * (2.0 / 9007199254740992.0) - 1.0;
}

// ------------------------------------------------------------------ load-time kernels (K1, K2)
// K1: df_j = #{entries > 0 in column j}  (b1ToB2)
__global__ void k_df_count(const int32_t* __restrict__ col, const double* __restrict__ val, int64_t nnz,
                           int n_cols, int* __restrict__ df, int* __restrict__ err) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < nnz; i += stride) {
    int c = col[i];
    if (c < 0 || c >= n_cols) { atomicMax(err, 1); continue; }
    if (val[i] > 0) atomicAdd(&df[c], 1);
  }
}
__global__ void k_idf(const int* __restrict__ df, int n_cols, double m, double* __restrict__ idf) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n_cols) idf[j] = df[j] > 0 ? log(m / (double)df[j]) : 0.0;
}
__global__ void k_idf_scale(const int32_t* __restrict__ col, double* __restrict__ val, int64_t nnz,
                            const double* __restrict__ idf) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < nnz; i += stride) val[i] *= idf[col[i]];
}
// K2: row L2 normalisation in place (b2ToB, A2) + validation (NaN, negative, zero rows, column range)
__global__ void k_row_l2(const int64_t* __restrict__ row_ptr, const int32_t* __restrict__ col,
                         double* __restrict__ val, int n_rows, int n_cols, double* __restrict__ norm_out,
                         int* __restrict__ err) {
  int warp = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  int lane = threadIdx.x & 31;
  int nwarps = (int)(((int64_t)gridDim.x * blockDim.x) >> 5);
  for (int r = warp; r < n_rows; r += nwarps) {
    int64_t rs = row_ptr[r], re = row_ptr[r + 1];
    double ss = 0.0; int bad = 0;
    for (int64_t k = rs + lane; k < re; k += 32) {
      double v = val[k]; int c = col[k];
      if (!(v == v) || isinf(v)) bad |= 1;
      if (v < 0) bad |= 2;
      if (c < 0 || c >= n_cols) bad |= 4;
      ss += v * v;
    }
    ss = warp_sum(ss);
    bad = __reduce_or_sync(0xffffffffu, bad);
    double nrm = sqrt(ss);
    if (bad & 1) atomicMax(err, 5); else if (bad & 4) atomicMax(err, 1);
    else if (!(nrm > 0)) atomicMax(err, 4);
    if (lane == 0 && norm_out) norm_out[r] = nrm;
    double inv = nrm > 0 ? 1.0 / nrm : 0.0;
    for (int64_t k = rs + lane; k < re; k += 32) val[k] *= inv;
  }
}

// column renumbering by descending frequency: the hot part of every feature-dimension vector (y, w^2, 1/w) becomes a
// compact prefix, which is what the L1 sees in the gathers (measured +10 % on k_gather, tools/exp/run_exp3.py)
__global__ void k_col_hist(const int32_t* __restrict__ col, int64_t nnz, int n_cols, int* __restrict__ cnt, int* __restrict__ flags) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < nnz; i += stride) { const int cj = col[i]; if (cj < 0 || cj >= n_cols) atomicOr(flags, 4); else atomicAdd(cnt + cj, 1); }
}
__global__ void k_iota_i32(int* __restrict__ p, int n) { int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) p[i] = i; }
__global__ void k_rank_from_order(const int* __restrict__ order, int n, int* __restrict__ rank) {
  int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) rank[order[i]] = i;
}
__global__ void k_remap_cols(int32_t* __restrict__ col, int64_t nnz, const int* __restrict__ rank) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < nnz; i += stride) col[i] = __ldg(rank + col[i]);
}

// compact host element types (tmc_csr.val_type / idx_type) widened on arrival
template <class SRC, class DST>
__global__ void k_widen(const SRC* __restrict__ src, DST* __restrict__ dst, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) dst[i] = (DST)src[i];
}

// adopted device arrays: row_ptr must be a non-decreasing sequence from 0 to nnz before any kernel walks the rows
__global__ void k_check_row_ptr(const int64_t* __restrict__ rp, int64_t n_rows, int64_t nnz, int* __restrict__ bad) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0 && (rp[0] != 0 || rp[n_rows] != nnz)) atomicOr(bad, 1);
  if (i < n_rows && rp[i + 1] < rp[i]) atomicOr(bad, 1);
}

// value scan: validity + can the values be stored as small integers / are they all ones?
// flags: 1 NaN/Inf, 2 negative, 4 column out of range, 8 not an integer <= 65535, 16 not all ones
__global__ void k_scan_values(const int32_t* __restrict__ col, const double* __restrict__ val, int64_t nnz, int n_cols, int* __restrict__ flags,
                              unsigned long long* __restrict__ n_small) {      // n_small[0] = #{0 < v <= 15}, [1] = #{0 < v <= 255}
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int f = 0; int n15 = 0, n255 = 0;
  for (; i < nnz; i += stride) {
    const double v = val[i]; const int cj = col[i];
    n15 += (v > 0.0 && v <= 15.0); n255 += (v > 0.0 && v <= 255.0);
    if (!(v == v) || isinf(v)) f |= 1;
    if (v < 0) f |= 2;
    if (cj < 0 || cj >= n_cols) f |= 4;
    if (!(v >= 0 && v <= 65535.0 && v == floor(v))) f |= 8;
    if (v != 1.0) f |= 16;
  }
  f = __reduce_or_sync(0xffffffffu, f);
  if ((threadIdx.x & 31) == 0 && f) atomicOr(flags, f);
  if (n_small) {
    n15 = __reduce_add_sync(0xffffffffu, n15); n255 = __reduce_add_sync(0xffffffffu, n255);
    if ((threadIdx.x & 31) == 0) { if (n15) atomicAdd(n_small, (unsigned long long)n15); if (n255) atomicAdd(n_small + 1, (unsigned long long)n255); }
  }
}
// ---- factor recovery: a caller-supplied B2 = A diag(w) (tf-idf already applied, the reference's own call pattern at
// Cluster.hs:147) is recognised so that the compact storage applies: w_j = smallest positive value of column j, A = B2 / w.
// Accepted only if EVERY entry is reproduced bit for bit by fl(round(v / w_j) * w_j) with a count in [1, 65535].
__global__ void k_colmin(const int32_t* __restrict__ col, const double* __restrict__ val, int64_t nnz, unsigned long long* __restrict__ cmin) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < nnz; i += stride) {
    const double v = val[i];
    if (v > 0.0) {                                 // positive doubles order like their bit patterns
      const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
      unsigned long long* p = cmin + col[i];
      if (bits < *p) atomicMin(p, bits);           // racy pre-check: only skips when the stored value is already smaller
    }
  }
}
__global__ void k_colmin_fix(double* __restrict__ w, int n) {        // columns without a positive entry: weight 0
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n && isinf(w[j])) w[j] = 0.0;
}
// cnt_out == nullptr: check only.  flags |= 8 when some entry is not an exact small multiple, |= 16 when some count != 1.
__global__ void k_factor_check(const int32_t* __restrict__ col, const double* __restrict__ val, int64_t nnz, const double* __restrict__ w,
                               double* __restrict__ cnt_out, int* __restrict__ flags, unsigned long long* __restrict__ n_small) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int f = 0, n15 = 0, n255 = 0;
  for (; i < nnz; i += stride) {
    const double v = val[i];
    double r = 0.0;
    if (v > 0.0) {
      const double wj = __ldg(w + col[i]);
      r = rint(v / wj);
      if (!(r >= 1.0 && r <= 65535.0 && __dmul_rn(r, wj) == v)) f |= 8;
      if (r != 1.0) f |= 16;
      n15 += (r <= 15.0); n255 += (r <= 255.0);
    }
    if (cnt_out) cnt_out[i] = r;
  }
  f = __reduce_or_sync(0xffffffffu, f);
  if ((threadIdx.x & 31) == 0 && f) atomicOr(flags, f);
  if (n_small) {
    n15 = __reduce_add_sync(0xffffffffu, n15); n255 = __reduce_add_sync(0xffffffffu, n255);
    if ((threadIdx.x & 31) == 0) { if (n15) atomicAdd(n_small, (unsigned long long)n15); if (n255) atomicAdd(n_small + 1, (unsigned long long)n255); }
  }
}
__global__ void k_to_u16(const double* __restrict__ val, unsigned short* __restrict__ out, int64_t nnz) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < nnz; i += stride) out[i] = (unsigned short)val[i];
}
// rrow_i = 1 / || (A diag(w))_i ||_2 ; flags zero rows
__global__ void k_row_factor(const int64_t* __restrict__ row_ptr, const int32_t* __restrict__ col, const double* __restrict__ val,
                             const double* __restrict__ w, int n_rows, double* __restrict__ rrow, int* __restrict__ err) {
  int warp = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  const int nwarps = (int)(((int64_t)gridDim.x * blockDim.x) >> 5);
  for (int r = warp; r < n_rows; r += nwarps) {
    const int64_t rs = row_ptr[r], re = row_ptr[r + 1];
    double ss = 0.0;
    for (int64_t k = rs + lane; k < re; k += 32) { const double v = val[k] * (w ? w[col[k]] : 1.0); ss += v * v; }
    ss = warp_sum(ss);
    if (lane == 0) { if (!(ss > 0)) { atomicMax(err, 4); rrow[r] = 0.0; } else rrow[r] = 1.0 / sqrt(ss); }
  }
}
__global__ void k_fill_f64(double* __restrict__ p, int64_t n, double v) { int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; if (i < n) p[i] = v; }
__global__ void k_col_factors(const double* __restrict__ w, int n_cols, double* __restrict__ w2, double* __restrict__ invw) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n_cols) { const double x = w ? w[j] : 1.0; w2[j] = x * x; invw[j] = x != 0.0 ? 1.0 / x : 0.0; }
}

__global__ void k_col_factors_perm(const double* __restrict__ w, const int* __restrict__ new2old, int n_cols, double* __restrict__ w2, double* __restrict__ invw) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n_cols) { const double x = w ? w[new2old[j]] : 1.0; w2[j] = x * x; invw[j] = x != 0.0 ? 1.0 / x : 0.0; }
}

// ------------------------------------------------------------------ dense panel construction (load time)
// An entry goes to the panel when its (renumbered) column is < P and its count fits the cell (cap = 255 or 15); everything else
// stays in the residual CSR.  k_panel_count: residual entries per row (cnt[n_rows] = 0 closes the scan).
__global__ void __launch_bounds__(256) k_panel_count(const int64_t* __restrict__ row_ptr, const int32_t* __restrict__ col, const double* __restrict__ val,
                                                     int n_rows, const int* __restrict__ old2new, int P, double cap, long long* __restrict__ cnt) {
  const int warp = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  const int nw = (int)(((int64_t)gridDim.x * blockDim.x) >> 5);
  for (int r = warp; r <= n_rows; r += nw) {
    if (r == n_rows) { if (lane == 0) cnt[r] = 0; continue; }
    const int64_t rs = row_ptr[r], re = row_ptr[r + 1];
    int n = 0;
    for (int64_t k = rs + lane; k < re; k += 32) {
      const double v = val[k];
      n += !(__ldg(old2new + col[k]) < P && v <= cap);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
    if (lane == 0) cnt[r] = n;
  }
}
// k_panel_fill: one warp per row writes the panel bytes and compacts the residual (order kept).  A duplicated (row, column)
// would overwrite a panel byte instead of adding: the warp re-counts its row and raises flag 32 (the caller then drops the panel).
__global__ void __launch_bounds__(256) k_panel_fill(const int64_t* __restrict__ row_ptr, const int32_t* __restrict__ col, const double* __restrict__ val,
                                                    int n_rows, const int* __restrict__ old2new, int P, int pw, int nib, double cap,
                                                    const long long* __restrict__ row_ptr_r,
                                                    int32_t* __restrict__ col_r, unsigned short* __restrict__ val_r, uint8_t* __restrict__ panel,
                                                    int* __restrict__ flags) {
  const int warp = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  const int nw = (int)(((int64_t)gridDim.x * blockDim.x) >> 5);
  for (int r = warp; r < n_rows; r += nw) {
    const int64_t rs = row_ptr[r], re = row_ptr[r + 1];
    long long base = row_ptr_r[r];
    uint8_t* prow = panel + (size_t)r * pw;
    int np = 0;
    for (int64_t k0 = rs; k0 < re; k0 += 32) {
      const int64_t k = k0 + lane;
      const bool valid = k < re;
      const double v = valid ? val[k] : 0.0;
      const int nc = valid ? __ldg(old2new + col[k]) : 0;
      const bool inp = valid && nc < P && v <= cap;
      if (inp && v > 0.0) {
        if (nib) atomicOr(reinterpret_cast<unsigned*>(prow) + (nc >> 3), (unsigned)v << (4 * (nc & 7)));    // two columns share a byte
        else prow[nc] = (uint8_t)v;
        np++;
      }
      const bool res = valid && !inp;
      const unsigned m = __ballot_sync(0xffffffffu, res);
      if (res) {
        const long long o = base + __popc(m & ((1u << lane) - 1u));
        col_r[o] = nc;
        if (val_r) val_r[o] = (unsigned short)v;
      }
      base += __popc(m);
    }
    __syncwarp();
    int nzb = 0;
    const unsigned* pwd = reinterpret_cast<const unsigned*>(prow);
    for (int i = lane; i < (pw >> 2); i += 32) {
      const unsigned wv = __ldcg(pwd + i);
      if (nib) {
        unsigned m = wv | (wv >> 1); m |= m >> 2;            // bit 0 of every nibble = "nibble is non-zero"
        nzb += __popc(m & 0x11111111u);
      } else nzb += ((wv & 0xffu) != 0) + ((wv & 0xff00u) != 0) + ((wv & 0xff0000u) != 0) + ((wv & 0xff000000u) != 0);
    }
    int d = np - nzb;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
    if (lane == 0 && d != 0) atomicOr(flags, 32);
  }
}

// ------------------------------------------------------------------ per-sweep bookkeeping
__global__ void k_activate(Ctx c, const Activation* __restrict__ acts, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Activation a = acts[i];
  Slot s;
  s.begin = a.begin; s.end = a.end; s.gsize = a.gsize; s.n0_local = 0;
  s.state = ST_DEG; s.action = ACT_NONE; s.j = 0; s.iters = 0; s.restarts = 0; s.node = a.node;
  s.split = 0; s.n0 = 0; s.stop_after = a.stop_after;
  s.tpar = a.tpar; s.priv = a.priv; s.nev = a.nev; s.tb = a.tb; s.te = a.te; s.tnnz0 = 0; s.tchunk0 = 0; s.ntchunk = 0;
  s.sumd = 0; s.inv_sqrt_sumd = 0; s.beta = 0; s.theta = 0; s.resid = 0; s.q = 0; s.theta_hint = -1e300;
  s.cert_req = 0; s.deg_phase = 0; s.cert_need = 0.0; s.cert_thr = 0.02; s.sumd2 = 0.0;
  s.warm = a.warm; s.have_z2 = 0; s.flip = 0; s.cert_k = 1.0;
  c.st[a.slot] = s;
  slot_scal(c, a.slot)[R_N1] = 0.0; slot_scal(c, a.slot)[R_SD1] = 0.0; slot_scal(c, a.slot)[R_SGN] = 0.0;
}

// Work lists, rebuilt on the device every sweep.  k_prepare (one CTA) scans the live slots into three offset tables
// (SpMV chunks, vector chunks, transposed-block chunks) and zeroes the reduction rows; k_prepare_fill (many CTAs) writes the
// chunk records (the single-CTA version spent 20-40 us per sweep writing up to 40k records).
__device__ int block_scan_counts(const Ctx& c, int mode, int* offs, int* wsum) {
  // mode 0: chunk_rows positions, 1: vchunk_rows positions, 2: tchunk entries of the transposed block (scatter states only),
  // 3: gs_chunk_rows positions of the large gathering nodes
  const int tid = threadIdx.x;
  const int per = (c.n_slots + 1023) / 1024;        // slots per thread (<= 16: the host caps max_active at 16384)
  int cnt[16]; int mine = 0;
  for (int q = 0; q < per; ++q) {
    int slot = tid * per + q; int n = 0;
    if (slot < c.n_slots) {
      const Slot& s = c.st[slot];
      if (mode == 2) {
        if (s.state == ST_DEG || s.state == ST_LAN || s.state == ST_MOD) n = (int)((s.te - (s.tb & ~7ll) + c.tchunk - 1) / c.tchunk);
      } else if (mode == 3) {
        if ((s.state == ST_DEG || s.state == ST_LAN || (s.state == ST_MOD && c.ts_on)) && s.end - s.begin >= c.gs_min_rows)
          n = (s.end - s.begin + c.gs_chunk_rows - 1) / c.gs_chunk_rows;
      } else if (s.state >= ST_DEG && s.state <= ST_MOD) {
        const int cs = mode == 0 ? c.chunk_rows : c.vchunk_rows;
        n = (s.end - s.begin + cs - 1) / cs;
        if (mode == 0 && c.gs_on && s.end - s.begin >= c.gs_min_rows) n = 0;      // gathered by k_gather_smem
      }
    }
    cnt[q] = n; mine += n;
  }
  int lane = tid & 31, wid = tid >> 5;
  int incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
  if (lane == 31) wsum[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    int v = wsum[lane]; int iv = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, iv, o); if (lane >= o) iv += t; }
    wsum[lane] = iv - v;
  }
  __syncthreads();
  int base = wsum[wid] + incl - mine;
  for (int q = 0; q < per; ++q) {
    int slot = tid * per + q;
    if (slot < c.n_slots) {
      offs[slot] = base;
      if (mode == 2) { c.st[slot].tchunk0 = base; c.st[slot].ntchunk = cnt[q]; }
    }
    base += cnt[q];
  }
  if (tid == 1023) offs[c.n_slots] = base;
  __syncthreads();
  const int total = offs[c.n_slots];
  __syncthreads();
  return total;
}

__global__ void __launch_bounds__(1024) k_prepare(Ctx c) {
  __shared__ int wsum[32];
  int* offsA = c.list_offs; int* offsV = c.list_offs + (c.max_active + 1); int* offsT = c.list_offs + 2 * (c.max_active + 1);
  const int ta = block_scan_counts(c, 0, offsA, wsum);
  const int tv = block_scan_counts(c, 1, offsV, wsum);
  const int tt = c.use_t ? block_scan_counts(c, 2, offsT, wsum) : 0;
  int* offsG = c.list_offs + 3 * (c.max_active + 1);
  const int tg = c.gs_on ? block_scan_counts(c, 3, offsG, wsum) : 0;
  if (threadIdx.x == 0) { *c.n_chunks = ta; *c.n_vchunks = tv; if (c.use_t) *c.n_tchunks = tt; *c.n_gchunks = tg; }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  // zero reduction rows (dots of both passes up to j+1, and the scalars): one warp per slot
  for (int slot = wid; slot < c.n_slots; slot += 32) {
    const Slot& s = c.st[slot];
    double* ra = c.redA + (int64_t)slot * (c.K + 2);
    double* rb = slot_redB(c, slot);
    int nv = min(s.j + 2, c.K + 2);
    if (s.state < ST_DEG || s.state > ST_MOD) {
      // idle slot below the highest live one: its rows still travel through the per-sweep all-reduces (in place), so they
      // are kept at zero instead of growing by a factor `world` per sweep
      if (c.world > 1) { for (int i = lane; i < nv; i += 32) { ra[i] = 0.0; rb[i] = 0.0; } if (lane < R_NSCAL) slot_scal(c, slot)[lane] = 0.0; }
      continue;
    }
    for (int i = lane; i < nv; i += 32) { ra[i] = 0.0; rb[i] = 0.0; }
    if (lane < 3) slot_scal(c, slot)[lane] = 0.0;       // NRM2, SUMD, YNORM2 (N1/SD1 persist from the previous sweep)
  }
}

__device__ __forceinline__ int find_slot(const int* __restrict__ offs, int n_slots, int ci) {
  int lo = 0, hi = n_slots;                 // last slot with offs[slot] <= ci
  while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (__ldg(offs + mid) <= ci) lo = mid; else hi = mid; }
  return lo;
}
__global__ void __launch_bounds__(256) k_prepare_fill(Ctx c) {
  const int* offsA = c.list_offs; const int* offsV = c.list_offs + (c.max_active + 1); const int* offsT = c.list_offs + 2 * (c.max_active + 1);
  const int ta = *c.n_chunks, tv = *c.n_vchunks, tt = c.use_t ? *c.n_tchunks : 0;
  const int stride = gridDim.x * blockDim.x, t0 = blockIdx.x * blockDim.x + threadIdx.x;
  for (int ci = t0; ci < ta; ci += stride) {
    const int sl = find_slot(offsA, c.n_slots, ci);
    const Slot& s = c.st[sl];
    Chunk ch; ch.slot = sl; ch.begin = s.begin + (ci - offsA[sl]) * c.chunk_rows; ch.len = min(c.chunk_rows, s.end - ch.begin);
    c.chunks[ci] = ch;
  }
  for (int ci = t0; ci < tv; ci += stride) {
    const int sl = find_slot(offsV, c.n_slots, ci);
    const Slot& s = c.st[sl];
    Chunk ch; ch.slot = sl; ch.begin = s.begin + (ci - offsV[sl]) * c.vchunk_rows; ch.len = min(c.vchunk_rows, s.end - ch.begin);
    c.vchunks[ci] = ch;
  }
  for (int ci = t0; ci < tt; ci += stride) {
    const int sl = find_slot(offsT, c.n_slots, ci);
    const Slot& s = c.st[sl];
    TChunk ch; ch.slot = sl; ch.pad = 0; ch.begin = (s.tb & ~7ll) + (long long)(ci - offsT[sl]) * c.tchunk;
    c.tchunks[ci] = ch;
  }
  if (c.gs_on) {
    const int* offsG = c.list_offs + 3 * (c.max_active + 1);
    const int tg = *c.n_gchunks;
    for (int ci = t0; ci < tg; ci += stride) {
      const int sl = find_slot(offsG, c.n_slots, ci);
      const Slot& s = c.st[sl];
      Chunk ch; ch.slot = sl; ch.begin = s.begin + (ci - offsG[sl]) * c.gs_chunk_rows; ch.len = min(c.gs_chunk_rows, s.end - ch.begin);
      c.gchunks[ci] = ch;
    }
  }
}

// zero y of every slot that scatters this sweep; grid (max_active_used, ny)
__global__ void k_zero_y(Ctx c) {
  int slot = blockIdx.x;
  int st = c.st[slot].state;
  if (st != ST_DEG && st != ST_LAN && st != ST_MOD) return;
  double2* y = reinterpret_cast<double2*>(c.Y + (int64_t)slot * c.ystride);
  int n2 = (int)(c.ystride >> 1);
  for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < n2; i += gridDim.y * blockDim.x) y[i] = make_double2(0.0, 0.0);
}

// ------------------------------------------------------------------ K4 half 1: y_S = B_S^T (xin)   (scatter)
// xin = D^{-1/2} q_j for Lanczos (so y = C^T q_j), 1 for the degree pass, the 0/1 label for modularity.
template <int VT>
__global__ void __launch_bounds__(256) k_scatter(Ctx c) {
  if ((int)blockIdx.x >= *c.n_chunks) return;
  const Chunk ch = c.chunks[blockIdx.x];
  const int state = c.st[ch.slot].state;
  if (state != ST_DEG && state != ST_LAN && state != ST_MOD) return;
  double* __restrict__ y = c.Y + (int64_t)ch.slot * c.ystride;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = warp; r < ch.len; r += 8) {
    const int p = ch.begin + r;
    const int row = c.perm[p];
    const double xi = (state == ST_DEG) ? (VT == VT_F64 ? 1.0 : c.rrow[row]) : c.xin[p];
    if (xi == 0.0) continue;
    const bool absv = (VT == VT_F64) && c.signed_mode && state == ST_DEG && c.st[ch.slot].deg_phase == 0;
    const int64_t rs = c.row_ptr[row], re = c.row_ptr[row + 1];
    for (int64_t k = rs + lane; k < re; k += 32) {
      const int cj = __ldg(c.col + k);
      double v = ld_val<VT>(c.val, k);
      if (absv) v = fabs(v);
      v *= xi;
      if (VT != VT_F64) v *= __ldg(c.w2col + cj);
      atomicAdd(y + cj, v);
    }
  }
}

// ------------------------------------------------------------------ K4 half 1, transposed form: y_S = B_S^T (xin) as a GATHER
// The node's entries are kept a second time sorted by column ("T-block": tcol, tpos = cell position, tval).
// Each lane takes 8 consecutive entries (two 128-bit loads per array), multiplies by xin[pos] and merges equal
// columns locally; a warp whose 256 entries are one column reduces with shuffles.  Only run totals touch y, so
// the 175 M RED.F64 of the CSR scatter (measured: atomic-throughput bound at 1.6 TB/s) become a few per column.
__device__ __forceinline__ int4 ldg_i4(const int32_t* p) { return __ldg(reinterpret_cast<const int4*>(p)); }
__device__ __forceinline__ double2 ldg_d2(const double* p) { return __ldg(reinterpret_cast<const double2*>(p)); }

// E = consecutive entries per lane (4 or 8); XS = read x in the strided mapping and change mapping through padded smem
template <int E, int XS, int VT>
__global__ void __launch_bounds__(256) k_tgather_t(Ctx c, int cls) {       // cls: 0 all nodes, 1 only large (>= xs_min_rows), 2 only small
  __shared__ double xs_sm[XS ? 8 : 1][XS ? (32 * E + 2 * E) : 1];
  if ((int)blockIdx.x >= *c.n_tchunks) return;
  const TChunk ch = c.tchunks[blockIdx.x];
  const Slot& s = c.st[ch.slot];
  const int state = s.state;
  if (state != ST_DEG && state != ST_LAN && state != ST_MOD) return;
  if (c.ts_on && s.end - s.begin >= c.gs_min_rows) return;          // residual product of this node: k_tscatter_smem
  if (cls) { const bool big = (s.end - s.begin) >= c.xs_min_rows; if (big != (cls == 1)) return; }
  const long long tb = s.tb, te = s.te;
  const int par = s.tpar;
  const int32_t* __restrict__ tcol = par ? c.tcol[1] : c.tcol[0];
  const int32_t* __restrict__ tpos = par ? c.tpos[1] : c.tpos[0];
  const void* __restrict__ tval = par ? c.tval[1] : c.tval[0];
  const double* __restrict__ xin = c.xin;
  double* __restrict__ y = c.Y + (int64_t)ch.slot * c.ystride;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long cend = min(ch.begin + (long long)c.tchunk, te);
  constexpr bool FACT = (VT != VT_F64);
  const bool use_x = FACT || state != ST_DEG;      // factored: even the degree pass multiplies by xin = rrow
  const bool absv = !FACT && c.signed_mode && state == ST_DEG && s.deg_phase == 0;      // d = |B| (|B|^T 1)
  constexpr int W = 32 * E;                 // entries per warp iteration
  constexpr bool xs_on = (XS == 1);
  for (long long base = ch.begin + warp * W; base < cend; base += 8 * W) {
    double* xs = xs_sm[XS ? warp : 0];
    if (XS && xs_on && use_x) {
      // x = xin[pos] in the STRIDED mapping (entry base + 32e + lane): positions ascend inside a column, so a warp-wide
      // gather touches a few consecutive lines; padded layout i -> i + i/16 is conflict-free for both mappings.
#pragma unroll
      for (int e = 0; e < E; ++e) {
        const long long k = base + 32 * e + lane;
        double xv = 0.0;
        if (k >= tb && k < te) xv = __ldg(xin + __ldg(tpos + k));
        const int i = 32 * e + lane;
        xs[i + (i >> 4)] = xv;
      }
      __syncwarp();
    }
    const long long k0 = base + lane * E;
    int cc[E]; double t[E];
    if (k0 < te && k0 + E > tb) {
      ld_idx<E>(tcol + k0, cc);
      ld_vals<E, VT>(tval, k0, t);
      if (absv) {
#pragma unroll
        for (int e = 0; e < E; ++e) t[e] = fabs(t[e]);
      }
      if (use_x) {
        if (XS && xs_on) {
#pragma unroll
          for (int e = 0; e < E; ++e) { const int i = E * lane + e; t[e] *= xs[i + (i >> 4)]; }
        } else {
          int pp[E];
          ld_idx<E>(tpos + k0, pp);
#pragma unroll
          for (int e = 0; e < E; ++e) {
            const bool ok = (k0 + e >= tb) && (k0 + e < te);
            t[e] = ok ? t[e] * __ldg(xin + pp[e]) : 0.0;
          }
        }
      }
#pragma unroll
      for (int e = 0; e < E; ++e) if (!((k0 + e >= tb) && (k0 + e < te))) { cc[e] = -1; t[e] = 0.0; }
    } else {
#pragma unroll
      for (int e = 0; e < E; ++e) { cc[e] = -1; t[e] = 0.0; }
    }
    if (XS && xs_on && use_x) __syncwarp();
    // whole warp inside one column?
    const int cfirst = __shfl_sync(0xffffffffu, cc[0], 0);
    const int clast = __shfl_sync(0xffffffffu, cc[E - 1], 31);
    if (cfirst == clast && cfirst >= 0) {
      double a = 0.0;
#pragma unroll
      for (int e = 0; e < E; ++e) a += t[e];
      a = warp_sum(a);
      if (lane == 0 && a != 0.0) atomicAdd(y + cfirst, FACT ? a * __ldg(c.w2col + cfirst) : a);
      continue;
    }
    double acc = t[0]; int cur = cc[0];
#pragma unroll
    for (int e = 1; e < E; ++e) {
      if (cc[e] == cur) acc += t[e];
      else { if (cur >= 0 && acc != 0.0) atomicAdd(y + cur, FACT ? acc * __ldg(c.w2col + cur) : acc); cur = cc[e]; acc = t[e]; }
    }
    if (cur >= 0 && acc != 0.0) atomicAdd(y + cur, FACT ? acc * __ldg(c.w2col + cur) : acc);
  }
}

// ------------------------------------------------------------------ dense panel halves of the SpMV pair
// Small counts stored densely, columns [0, P) contiguous in y: both halves stream the panel rows with coalesced loads and no
// index traffic.  Two formats (Ctx::pnib): one byte per cell (counts <= 255) or one nibble per cell (counts <= 15, the usual
// case for UMI counts; larger counts live in the residual CSR).
// A stored count b is used as the DENORMAL double with bit pattern (hi = 0, lo = b) = b * 2^-1074 -- exact, one PRMT (byte) or
// one LOP3 (nibble, masked IN PLACE: lo = w & (15 << 4k) = b * 16^k) and no conversion instruction (cvt.f64.u32 runs at a
// fraction of the fp64 rate; the 2^52 trick costs an extra DADD per entry, which made both kernels fp64-issue bound).  The dense
// operand (x or y) is pre-scaled by 2^960 (and y by 16^-k for nibble k) where it is staged in shared memory, the sums are
// scaled back by 2^114 (and 16^-k) once per row / column: power-of-two scalings are exact and every product is formed exactly
// inside the FMA, so the result is bit-identical to converting the count first.  |x| < 2^60 is required (else inf -> the
// non-finite degree / Ritz checks fire); the engine's vectors are bounded by the row factors and the column sums.
#define TMC_PANEL_UP 0x1p960
#define TMC_PANEL_DOWN 0x1p114
__device__ __forceinline__ double u8_f64(unsigned w, int k) { return __hiloint2double(0, (int)__byte_perm(w, 0u, 0x4440u + (unsigned)k)); }
__device__ __forceinline__ double u4_f64(unsigned w, int k) { return __hiloint2double(0, (int)(w & (15u << (4 * k)))); }      // = count * 16^k * 2^-1074
// 16 columns of one panel row as seen by a lane: 16 bytes, or 8 bytes of nibbles
template <int NIB> struct PanelLane { typedef uint4 T; };
template <> struct PanelLane<1> { typedef uint2 T; };
// acc[0..16) += cells * xv; nibble format: acc[8*i + k] carries the extra factor 16^k (removed by the caller)
template <int NIB> __device__ __forceinline__ void panel_fma(const typename PanelLane<NIB>::T& d, double xv, double* acc) {
  if constexpr (NIB != 0) {
    const unsigned w[2] = {d.x, d.y};
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[8 * i + k] += u4_f64(w[i], k) * xv;
  } else {
    const unsigned w[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[4 * i + k] += u8_f64(w[i], k) * xv;
  }
}
// a += one 32-bit word of a lane (4 byte cells / 8 nibble cells) times its y values (2 / 4 double2 holding the format's column scale)
template <int NIB> __device__ __forceinline__ void panel_word_dot(unsigned w, const double2* yv, double& a) {
  if constexpr (NIB != 0) {
#pragma unroll
    for (int k = 0; k < 8; k += 2) { a = fma(u4_f64(w, k), yv[k >> 1].x, a); a = fma(u4_f64(w, k + 1), yv[k >> 1].y, a); }
  } else {
#pragma unroll
    for (int k = 0; k < 4; k += 2) { a = fma(u8_f64(w, k), yv[k >> 1].x, a); a = fma(u8_f64(w, k + 1), yv[k >> 1].y, a); }
  }
}
template <int NIB> __device__ __forceinline__ unsigned panel_word(const typename PanelLane<NIB>::T& d, int wi) {
  if constexpr (NIB != 0) return wi == 0 ? d.x : d.y;
  else return wi == 0 ? d.x : (wi == 1 ? d.y : (wi == 2 ? d.z : d.w));
}
// y[0..P) += w^2 * sum_r panel[r][.] * xin[r] over the rows of G consecutive vector chunks.  Lane = 16 columns (one 16-byte or
// 8-byte load per row), warp = 512 columns, CTA = blockDim/32 warps side by side (the host sizes the CTA so that the slabs divide
// Ppad evenly); grid (ceil(n_vchunks / G), slabs).  Two register sets of NR rows in ping-pong: the next rows are in flight during
// the math; one atomicAdd per column per CTA and slot.  (Measured alternatives, tools/exp/panel_kern.cu + profiles/r1_v5, v6:
// 8-byte lanes with converted bytes 3.1 TB/s, a cp.async ring through shared memory 3.1 TB/s, this one 6.1 TB/s on bytes.)
template <int NR, int NIB>
__global__ void __launch_bounds__(256) k_tpanel(Ctx c, int G) {
  typedef typename PanelLane<NIB>::T LT;
  __shared__ int s_row[TMC_MAX_CHUNK + 2 * NR];
  __shared__ double s_x[TMC_MAX_CHUNK + 2 * NR];
  const int nv = *c.n_vchunks;
  const int i0 = blockIdx.x * G;
  if (i0 >= nv) return;
  const int i1 = min(nv, i0 + G);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int colw = (blockIdx.y * (blockDim.x >> 5) + warp) * 512;
  const bool on = colw < c.Ppad;
  const LT* __restrict__ pb = reinterpret_cast<const LT*>(c.panel + (NIB ? (colw >> 1) : colw)) + lane;
  const size_t ws = (size_t)c.pw / sizeof(LT);                                    // lane units per panel row
  double acc[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) acc[k] = 0.0;
  int cur = -1;
  auto flush = [&]() {
    if (cur >= 0 && on) {
      double* y = c.Y + (int64_t)cur * c.ystride;
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const int cj = colw + lane * 16 + k;
        const double down = NIB ? TMC_PANEL_DOWN / (double)(1u << (4 * (k & 7))) : TMC_PANEL_DOWN;
        if (cj < c.P && acc[k] != 0.0) atomicAdd(y + cj, acc[k] * down * __ldg(c.w2col + cj));
        acc[k] = 0.0;
      }
    }
  };
  for (int i = i0; i < i1; ++i) {
    const Chunk ch = c.vchunks[i];
    const int state = c.st[ch.slot].state;
    if (ch.slot != cur) { flush(); cur = ch.slot; }
    if (state != ST_DEG && state != ST_LAN && state != ST_MOD) continue;
    __syncthreads();
    // rows padded to a multiple of 2*NR with x = 0 (and a valid row id), so that the loop below needs no bounds tests
    const int lenp = (ch.len + 2 * NR - 1) / (2 * NR) * (2 * NR);
    for (int t = threadIdx.x; t < lenp; t += blockDim.x) {
      const bool ok = t < ch.len;
      s_row[t] = c.perm[ch.begin + (ok ? t : ch.len - 1)];
      s_x[t] = ok ? c.xin[ch.begin + t] * TMC_PANEL_UP : 0.0;
    }
    __syncthreads();
    if (!on) continue;
    LT d0[NR], d1[NR];
    auto load = [&](LT* d, int r) {
#pragma unroll
      for (int q = 0; q < NR; ++q) d[q] = __ldg(pb + (size_t)s_row[r + q] * ws);
    };
    auto fma_rows = [&](const LT* d, int r) {
#pragma unroll
      for (int q = 0; q < NR; ++q) panel_fma<NIB>(d[q], s_x[r + q], acc);
    };
    load(d0, 0);
    for (int r = 0; r < lenp; r += 2 * NR) {
      load(d1, r + NR);
      fma_rows(d0, r);
      if (r + 2 * NR < lenp) load(d0, r + 2 * NR);
      fma_rows(d1, r + NR);
    }
  }
  flush();
}
// EXPERIMENT (TMC_TP_NARROW=1): the nibble kernel with 4-byte lanes -- a lane owns ONE 32-bit word (8 columns), a warp 256
// columns, so 8 accumulators and NR words per register set instead of 16 and 2*NR: about 50 registers, 5 CTAs per SM instead of
// 3.  Twice the load instructions for the same bytes (they are 1/8 of the instruction stream), more warps to hide them behind.
template <int NR>
__global__ void __launch_bounds__(256) k_tpanel_narrow(Ctx c, int G) {
  __shared__ int s_row[TMC_MAX_CHUNK + 2 * NR];
  __shared__ double s_x[TMC_MAX_CHUNK + 2 * NR];
  const int nv = *c.n_vchunks;
  const int i0 = blockIdx.x * G;
  if (i0 >= nv) return;
  const int i1 = min(nv, i0 + G);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int colw = (blockIdx.y * (blockDim.x >> 5) + warp) * 256;
  const bool on = colw < c.Ppad;
  const unsigned* __restrict__ pb = reinterpret_cast<const unsigned*>(c.panel + (colw >> 1)) + lane;
  const size_t ws = (size_t)c.pw / sizeof(unsigned);
  double acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.0;
  int cur = -1;
  auto flush = [&]() {
    if (cur >= 0 && on) {
      double* y = c.Y + (int64_t)cur * c.ystride;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int cj = colw + lane * 8 + k;
        const double down = TMC_PANEL_DOWN / (double)(1u << (4 * k));
        if (cj < c.P && acc[k] != 0.0) atomicAdd(y + cj, acc[k] * down * __ldg(c.w2col + cj));
        acc[k] = 0.0;
      }
    }
  };
  for (int i = i0; i < i1; ++i) {
    const Chunk ch = c.vchunks[i];
    const int state = c.st[ch.slot].state;
    if (ch.slot != cur) { flush(); cur = ch.slot; }
    if (state != ST_DEG && state != ST_LAN && state != ST_MOD) continue;
    __syncthreads();
    const int lenp = (ch.len + 2 * NR - 1) / (2 * NR) * (2 * NR);
    for (int t = threadIdx.x; t < lenp; t += blockDim.x) {
      const bool ok = t < ch.len;
      s_row[t] = c.perm[ch.begin + (ok ? t : ch.len - 1)];
      s_x[t] = ok ? c.xin[ch.begin + t] * TMC_PANEL_UP : 0.0;
    }
    __syncthreads();
    if (!on) continue;
    unsigned d0[NR], d1[NR];
    auto load = [&](unsigned* d, int r) {
#pragma unroll
      for (int q = 0; q < NR; ++q) d[q] = __ldg(pb + (size_t)s_row[r + q] * ws);
    };
    auto fma_rows = [&](const unsigned* d, int r) {
#pragma unroll
      for (int q = 0; q < NR; ++q) {
        const double xv = s_x[r + q];
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] += u4_f64(d[q], k) * xv;
      }
    };
    load(d0, 0);
    for (int r = 0; r < lenp; r += 2 * NR) {
      load(d1, r + NR);
      fma_rows(d0, r);
      if (r + 2 * NR < lenp) load(d0, r + 2 * NR);
      fma_rows(d1, r + NR);
    }
  }
  flush();
}

// pacc[p] = sum_{j < P} panel[row(p)][j] * y[j]; one CTA per vector chunk.  y is staged in shared memory in blocks of TMC_GP_CB
// columns, stored so that the eight 16-byte reads of a lane (its 16 columns of a 512-column step) are conflict-free: the pair
// (16*lane + 2j, 16*lane + 2j + 1) sits at double2 index j*32 + lane.  A warp works on NR rows at a time (every y value read
// from shared memory is used NR times); two register sets in ping-pong keep the next step's panel bytes in flight.
#define TMC_GP_CB 4096
template <int NR, int NIB>
__global__ void __launch_bounds__(256) k_gather_panel(Ctx c) {
  typedef typename PanelLane<NIB>::T LT;
  __shared__ double2 ys2[TMC_GP_CB / 2];
  __shared__ double ps[TMC_MAX_CHUNK];
  if ((int)blockIdx.x >= *c.n_vchunks) return;
  const Chunk ch = c.vchunks[blockIdx.x];
  const int state = c.st[ch.slot].state;
  if (state != ST_DEG && state != ST_LAN) return;
  const double* __restrict__ y = c.Y + (int64_t)ch.slot * c.ystride;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int p = threadIdx.x; p < ch.len; p += 256) ps[p] = 0.0;
  for (int cb = 0; cb < c.Ppad; cb += TMC_GP_CB) {
    const int nb = min(TMC_GP_CB, c.Ppad - cb);
    __syncthreads();
    for (int t = threadIdx.x; t < (nb >> 1); t += 256) {
      const int w = t & 255;                       // pair index inside the 512-column step: columns 2w, 2w + 1
      double2 yv = ldg_d2(y + cb + 2 * t);
      // nibble format: column j of a lane sits in nibble (j & 7) of its word, read in place as count * 16^(j & 7)
      const double up0 = NIB ? TMC_PANEL_UP / (double)(1u << (4 * ((2 * w) & 7))) : TMC_PANEL_UP;
      yv.x *= up0; yv.y *= NIB ? up0 * 0.0625 : up0;
      ys2[(t & ~255) + ((w & 7) << 5) + (w >> 3)] = yv;
    }
    __syncthreads();
    const int nit = nb >> 9;
    for (int r0 = warp * NR; r0 < ch.len; r0 += 8 * NR) {
      const int nr = min(NR, ch.len - r0);
      const LT* pr[NR];
#pragma unroll
      for (int q = 0; q < NR; ++q)
        pr[q] = reinterpret_cast<const LT*>(c.panel + (size_t)c.perm[ch.begin + r0 + min(q, nr - 1)] * c.pw + (NIB ? (cb >> 1) : cb)) + lane;
      double acc[NR];
#pragma unroll
      for (int q = 0; q < NR; ++q) acc[q] = 0.0;
      LT d0[NR], d1[NR];
      auto load = [&](LT* d, int i) {
#pragma unroll
        for (int q = 0; q < NR; ++q) d[q] = __ldg(pr[q] + 32 * i);
      };
      auto fma_step = [&](const LT* d, int i) {
        constexpr int NW = NIB ? 2 : 4, NP = 8 / NW;       // words per lane, y pairs per word
#pragma unroll
        for (int wi = 0; wi < NW; ++wi) {
          double2 yv[NP];
#pragma unroll
          for (int j = 0; j < NP; ++j) yv[j] = ys2[i * 256 + (wi * NP + j) * 32 + lane];
#pragma unroll
          for (int q = 0; q < NR; ++q) panel_word_dot<NIB>(panel_word<NIB>(d[q], wi), yv, acc[q]);
        }
      };
      load(d0, 0);
      for (int i = 0; i < nit; i += 2) {
        if (i + 1 < nit) load(d1, i + 1);
        fma_step(d0, i);
        if (i + 2 < nit) load(d0, i + 2);
        if (i + 1 < nit) fma_step(d1, i + 1);
      }
#pragma unroll
      for (int q = 0; q < NR; ++q) {
        const double sq = warp_sum(acc[q]);
        if (lane == 0 && q < nr) ps[r0 + q] += sq * TMC_PANEL_DOWN;
      }
    }
  }
  __syncthreads();
  for (int p = threadIdx.x; p < ch.len; p += 256) c.pacc[ch.begin + p] = ps[p];
}

// sorted construction of the root block (perm = identity): after a stable radix sort of the CSR entries by column,
// tpos holds CSR entry indices; replace them by the row (= position) and fetch the values
// rows [row_lo, row_hi): rowid[k - e0] = position of the row (= r - row_lo), e0 = row_ptr[row_lo]
__global__ void __launch_bounds__(256) k_tb_rowid(const int64_t* __restrict__ row_ptr, int row_lo, int row_hi, int64_t e0, int32_t* __restrict__ rowid) {
  const int warp = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  const int nw = (int)(((int64_t)gridDim.x * blockDim.x) >> 5);
  for (int r = row_lo + warp; r < row_hi; r += nw) {
    const int64_t rs = row_ptr[r], re = row_ptr[r + 1];
    for (int64_t k = rs + lane; k < re; k += 32) rowid[k - e0] = r - row_lo;
  }
}
__global__ void k_tb_iota(int32_t* __restrict__ p, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long st = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += st) p[i] = (int32_t)i;
}
template <int VT>
__global__ void k_tb_finish_sorted(const int32_t* __restrict__ rowid, const void* __restrict__ val, int64_t e0, int32_t* __restrict__ tpos,
                                   void* __restrict__ tval, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long st = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += st) { const int32_t e = tpos[i]; tpos[i] = __ldg(rowid + e); st_val<VT>(tval, i, ld_val<VT>(val, e0 + e)); }
}

// ---- building a node's T-block from its CSR rows (root, and the per-node API entry points)
__global__ void __launch_bounds__(256) k_tb_count(Ctx c, int begin, int end) {
  const int warp = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  const int nw = (int)(((int64_t)gridDim.x * blockDim.x) >> 5);
  for (int p = begin + warp; p < end; p += nw) {
    const int row = c.perm[p];
    const int64_t rs = c.row_ptr[row], re = c.row_ptr[row + 1];
    for (int64_t k = rs + lane; k < re; k += 32) atomicAdd(c.tb_cnt + __ldg(c.col + k), 1);
  }
}
__global__ void __launch_bounds__(1024) k_tb_scan(Ctx c) {      // exclusive scan of tb_cnt -> tb_colstart; tb_cnt := 0 (cursor)
  __shared__ long long wsum[32];
  __shared__ long long carry_s;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < c.n_cols; base += 1024) {
    const int j = base + tid;
    long long v = (j < c.n_cols) ? (long long)c.tb_cnt[j] : 0, incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { long long t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) wsum[wid] = incl;
    __syncthreads();
    if (wid == 0) {
      long long w = wsum[lane], iw = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { long long t = __shfl_up_sync(0xffffffffu, iw, o); if (lane >= o) iw += t; }
      wsum[lane] = iw - w;
    }
    __syncthreads();
    const long long carry = carry_s;
    if (j < c.n_cols) { c.tb_colstart[j] = carry + wsum[wid] + incl - v; c.tb_cnt[j] = 0; }
    __syncthreads();
    if (tid == 1023) carry_s = carry + wsum[31] + incl;
    __syncthreads();
  }
  if (tid == 0) c.tb_colstart[c.n_cols] = carry_s;
}
template <int VT>
__global__ void __launch_bounds__(256) k_tb_fill(Ctx c, int begin, int end, long long tb, int par) {
  const int warp = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  const int nw = (int)(((int64_t)gridDim.x * blockDim.x) >> 5);
  for (int p = begin + warp; p < end; p += nw) {
    const int row = c.perm[p];
    const int64_t rs = c.row_ptr[row], re = c.row_ptr[row + 1];
    for (int64_t k = rs + lane; k < re; k += 32) {
      const int col = __ldg(c.col + k);
      const long long idx = tb + c.tb_colstart[col] + atomicAdd(c.tb_cnt + col, 1);
      (par ? c.tcol[1] : c.tcol[0])[idx] = col; (par ? c.tpos[1] : c.tpos[0])[idx] = p; st_val<VT>(par ? c.tval[1] : c.tval[0], idx, ld_val<VT>(c.val, k));
    }
  }
}

// ---- splitting a T-block: stable partition of the parent's entries by the label of their cell into the other buffer
__device__ __forceinline__ bool slot_splits(const Slot& s) { return s.state == ST_MOD && s.action == ACT_FINISH && s.split; }

__global__ void __launch_bounds__(256) k_tpart_count(Ctx c) {
  __shared__ int sm[8];
  if ((int)blockIdx.x >= *c.n_tchunks) return;
  const TChunk ch = c.tchunks[blockIdx.x];
  const Slot& s = c.st[ch.slot];
  if (!slot_splits(s)) return;
  const long long lo = max(ch.begin, s.tb), hi = min(ch.begin + (long long)c.tchunk, s.te);
  const int32_t* __restrict__ tpos = s.tpar ? c.tpos[1] : c.tpos[0];
  int ones = 0;
  for (long long k = lo + threadIdx.x; k < hi; k += 256) ones += c.label[__ldg(tpos + k)];
  ones = __reduce_add_sync(0xffffffffu, ones);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = ones;
  __syncthreads();
  if (threadIdx.x == 0) { int t = 0; for (int q = 0; q < 8; ++q) t += sm[q]; c.tile_cnt[blockIdx.x] = t; }
}
__global__ void __launch_bounds__(1024) k_tpart_scan(Ctx c) {      // one CTA per slot: prefix of its chunks' label-1 counts
  __shared__ long long wsum[32];
  __shared__ long long carry_s;
  Slot& s = c.st[blockIdx.x];
  if (!slot_splits(s)) return;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int c0 = s.tchunk0, nc = s.ntchunk;
  if (tid == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < nc; base += 1024) {
    const int i = base + tid;
    long long v = (i < nc) ? (long long)c.tile_cnt[c0 + i] : 0, incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { long long t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) wsum[wid] = incl;
    __syncthreads();
    if (wid == 0) {
      long long w = wsum[lane], iw = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { long long t = __shfl_up_sync(0xffffffffu, iw, o); if (lane >= o) iw += t; }
      wsum[lane] = iw - w;
    }
    __syncthreads();
    const long long carry = carry_s;
    if (i < nc) c.tile_off[c0 + i] = carry + wsum[wid] + incl - v;
    __syncthreads();
    if (tid == 1023) carry_s = carry + wsum[31] + incl;
    __syncthreads();
  }
  if (tid == 0) s.tnnz0 = (s.te - s.tb) - carry_s;
}
template <int VT>
__global__ void __launch_bounds__(256) k_tpart_scatter(Ctx c) {
  // Every 2048-entry tile is staged in shared memory in DESTINATION order ([label-0 run | label-1 run]) and then
  // written out as two contiguous, fully coalesced segments (the direct per-thread stores wrote 4 bytes into 32
  // different sectors per instruction: 3.4x the ideal time on c3).
  __shared__ int wcnt[8];
  __shared__ long long run_s;
  __shared__ int s_col[2048], s_pos[2048];
  __shared__ double s_val[VT == VT_F64 ? 2048 : 1];
  __shared__ unsigned short s_v16[VT == VT_U16 ? 2048 : 1];
  if ((int)blockIdx.x >= *c.n_tchunks) return;
  const TChunk ch = c.tchunks[blockIdx.x];
  const Slot& s = c.st[ch.slot];
  if (!slot_splits(s)) return;
  const long long tb = s.tb, te = s.te, nnz0 = s.tnnz0;
  const int par = s.tpar;
  const int32_t* __restrict__ scol = par ? c.tcol[1] : c.tcol[0]; const int32_t* __restrict__ spos = par ? c.tpos[1] : c.tpos[0];
  const void* __restrict__ sval = par ? c.tval[1] : c.tval[0];
  int32_t* __restrict__ dcol = par ? c.tcol[0] : c.tcol[1]; int32_t* __restrict__ dpos = par ? c.tpos[0] : c.tpos[1];
  void* __restrict__ dval = par ? c.tval[0] : c.tval[1];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) run_s = c.tile_off[blockIdx.x];
  __syncthreads();
  const long long cend = min(ch.begin + (long long)c.tchunk, te);
  for (long long base = ch.begin; base < cend; base += 2048) {
    const long long k0 = base + tid * 8;
    int cc[8], pp[8], ff[8]; double vv[8];
    int mine = 0;
    if (k0 < te && k0 + 8 > tb) {
      const int4 ca = ldg_i4(scol + k0), cb = ldg_i4(scol + k0 + 4), pa = ldg_i4(spos + k0), pb = ldg_i4(spos + k0 + 4);
      ld_val4<VT>(sval, k0, vv); ld_val4<VT>(sval, k0 + 4, vv + 4);
      cc[0] = ca.x; cc[1] = ca.y; cc[2] = ca.z; cc[3] = ca.w; cc[4] = cb.x; cc[5] = cb.y; cc[6] = cb.z; cc[7] = cb.w;
      pp[0] = pa.x; pp[1] = pa.y; pp[2] = pa.z; pp[3] = pa.w; pp[4] = pb.x; pp[5] = pb.y; pp[6] = pb.z; pp[7] = pb.w;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const bool ok = (k0 + e >= tb) && (k0 + e < te);
        ff[e] = ok ? (int)c.label[pp[e]] : -1;
        mine += (ff[e] == 1);
      }
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) ff[e] = -1;
    }
    // block exclusive scan of `mine`
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) wcnt[wid] = incl;
    __syncthreads();
    int wbase = 0, tot = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q) { const int w = wcnt[q]; if (q < wid) wbase += w; tot += w; }
    const long long run = run_s;                                 // label-1 entries of the node before this tile
    const long long lo = max(base, tb), hi = min(base + 2048, te);
    const int nvalid = (int)(hi - lo);                           // valid entries of the tile form one contiguous index range
    const int nz = nvalid - tot;                                 // label-0 entries of the tile
    int ones_local = wbase + incl - mine;                        // label-1 entries of the tile before this thread's
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      if (ff[e] < 0) continue;
      const int idx_t = (int)((k0 + e) - lo);                    // index inside the tile's valid range
      const int si = ff[e] ? (nz + ones_local) : (idx_t - ones_local);
      s_col[si] = cc[e]; s_pos[si] = c.newpos[pp[e]];
      if (VT == VT_F64) s_val[si] = vv[e];
      if (VT == VT_U16) s_v16[si] = (unsigned short)vv[e];
      ones_local += ff[e];
    }
    __syncthreads();
    const long long z0 = tb + ((lo - tb) - run);                 // destination of the tile's first label-0 entry
    const long long o0 = tb + nnz0 + run;                        // ... and of its first label-1 entry
    for (int i = tid; i < nvalid; i += 256) {
      const long long dest = (i < nz) ? (z0 + i) : (o0 + (i - nz));
      dcol[dest] = s_col[i]; dpos[dest] = s_pos[i];
      if (VT == VT_F64) reinterpret_cast<double*>(dval)[dest] = s_val[i];
      if (VT == VT_U16) reinterpret_cast<unsigned short*>(dval)[dest] = s_v16[i];
    }
    if (tid == 0) run_s = run + tot;
    __syncthreads();
  }
}

// ------------------------------------------------------------------ K4 half 2: w = D^{-1/2} B_S y   (gather)
// One warp per row at a time, 4 x 32 entries in flight plus the NEXT 4 x 32 prefetched into registers before the
// dependent y[col] gathers are issued, and the next row's extent prefetched while the current row streams
// (measured: 4.2 -> 5.4 TB/s on the c2 root against the non-pipelined loop; tools/exp).  y is read through the
// read-only path (it is not written in this kernel) so the compiler may keep all four chains in flight.

// The gather in use: U x 32 entries in flight, values kept RAW (uint16) through the prefetch stage (U = 6 / 8 measured
// slower: 72 / 80 registers).  With a dense panel the panel half of the row sum arrives in pacc (k_gather_panel).
template <int VT> struct RawVal { typedef double T; };
template <> struct RawVal<VT_U16> { typedef unsigned short T; };
template <> struct RawVal<VT_ONE> { typedef unsigned char T; };
template <int VT> __device__ __forceinline__ typename RawVal<VT>::T ld_raw(const void* base, int64_t k) {
  if (VT == VT_F64) return (typename RawVal<VT>::T)__ldg(reinterpret_cast<const double*>(base) + k);
  if (VT == VT_U16) return (typename RawVal<VT>::T)__ldg(reinterpret_cast<const unsigned short*>(base) + k);
  return (typename RawVal<VT>::T)1;
}
// LPR = lanes per row (32: one row per warp at a time; 16 / 8: two / four rows per warp side by side -- the residual rows are
// short, ~340 entries on c3, and a warp sits behind a dependent chain perm -> row_ptr -> col/val -> y per row, so more rows in
// flight per warp is what hides it; the sub-warp segments are reduced with width-LPR shuffles)
template <int VT, int U, int LPR>
__global__ void __launch_bounds__(256) k_gather_u(Ctx c) {
  __shared__ double sm8[8];
  if ((int)blockIdx.x >= *c.n_chunks) return;
  const Chunk ch = c.chunks[blockIdx.x];
  const int state = c.st[ch.slot].state;
  if (state != ST_DEG && state != ST_LAN) return;
  const double* __restrict__ y = c.Y + (int64_t)ch.slot * c.ystride;
  const int32_t* __restrict__ col = c.col;
  const void* __restrict__ val = c.val;
  constexpr int RPW = 32 / LPR;                       // rows per warp
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sub = lane / LPR, sl = lane % LPR;        // which of the warp's rows, lane inside the row's segment
  typedef typename RawVal<VT>::T RT;
  const int dphase = (VT == VT_F64 && c.signed_mode && state == ST_DEG) ? c.st[ch.slot].deg_phase : -1;
  const bool absv = dphase == 0;
  double* __restrict__ dout = dphase == 1 ? c.d2 : c.d;
  double dsum = 0.0;
  long long nz = 0, nz0 = 0;
  constexpr int RSTEP = 8 * RPW;                      // rows a CTA advances per iteration
  int r = warp * RPW + sub;
  int64_t rs = 0, re = 0;
  int row_cur = 0;
  if (r < ch.len) { row_cur = c.perm[ch.begin + r]; rs = c.row_ptr[row_cur]; re = c.row_ptr[row_cur + 1]; }
  // all lanes of a warp run the same number of outer iterations (the shuffles below are warp-wide): rows past the end are empty
  for (int r0 = warp * RPW; r0 < ch.len; r0 += RSTEP) {
    const bool live = r < ch.len;
    const int p = ch.begin + r;
    const int rn = r + RSTEP;
    int64_t nrs = 0, nre = 0;
    int nrow = 0;
    if (rn < ch.len) { nrow = c.perm[ch.begin + rn]; nrs = c.row_ptr[nrow]; nre = c.row_ptr[nrow + 1]; }
    if (live && sl == 0) { nz += re - rs; if (c.P) nz0 += c.row_ptr0[row_cur + 1] - c.row_ptr0[row_cur]; }
    double acc[U];
#pragma unroll
    for (int u = 0; u < U; ++u) acc[u] = 0.0;
    int64_t k = rs + sl;
    bool have = live && (k + LPR * (U - 1) < re);
    int cc[U]; RT vv[U];
    if (have) {
#pragma unroll
      for (int u = 0; u < U; ++u) { cc[u] = __ldg(col + k + LPR * u); vv[u] = ld_raw<VT>(val, k + LPR * u); }
    }
    while (have) {
      const int64_t kn = k + LPR * U;
      const bool have_n = (kn + LPR * (U - 1) < re);
      int dc[U]; RT dv[U];
      if (have_n) {
#pragma unroll
        for (int u = 0; u < U; ++u) { dc[u] = __ldg(col + kn + LPR * u); dv[u] = ld_raw<VT>(val, kn + LPR * u); }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) acc[u] += (absv ? fabs((double)vv[u]) : (double)vv[u]) * __ldg(y + cc[u]);
      if (have_n) {
#pragma unroll
        for (int u = 0; u < U; ++u) { cc[u] = dc[u]; vv[u] = dv[u]; }
      }
      k = kn; have = have_n;
    }
    if (live) for (; k < re; k += LPR) { const double v = ld_val<VT>(val, k); acc[0] += (absv ? fabs(v) : v) * __ldg(y + __ldg(col + k)); }
    double s = 0.0;
#pragma unroll
    for (int u = 0; u < U; ++u) s += acc[u];
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);      // stays inside the row's LPR-lane segment
    if (live && sl == 0) {
      if (c.P) s += c.pacc[p];
      if (state == ST_DEG) { const double dv2 = (VT == VT_F64) ? s : s * c.rrow[row_cur]; dout[p] = dv2; dsum += dv2; }
      else c.w[p] = c.sdeg[p] * s;
    }
    r = rn; rs = nrs; re = nre; row_cur = nrow;
  }
  if (sl == 0 && (nz || nz0)) {
    atomicAdd(c.stats + (state == ST_DEG ? 1 : 0), (unsigned long long)(c.P ? nz0 : nz));     // nonzeros of the caller's matrix covered
    if (c.P) atomicAdd(c.stats + 2, (unsigned long long)nz);                                   // of which read from the residual CSR
  }
  if (state == ST_DEG) {
    double t = block_sum_256(dsum, sm8);
    if (threadIdx.x == 0) atomicAdd(slot_scal(c, ch.slot) + R_SUMD, t);
  }
}

// The same gather for LARGE nodes with the residual part of y resident in shared memory: one CTA of 32 warps per gs_chunk_rows
// positions stages y[P, n_cols) (up to ~200 KB: a 20-25 k-gene residual fits) once and then streams its rows; the dependent
// y[col] reads cost a shared-memory access instead of an L2 round trip (k_gather_u on the c3 root: 14 long-scoreboard stalls
// per issue, 2.4 TB/s).  Launched only when Ctx::gs_on (the host checks that the vector fits); those nodes' rows are not in
// the plain chunk list.
template <int VT, int U>
__global__ void __launch_bounds__(1024, 1) k_gather_smem(Ctx c) {
  extern __shared__ double gs_y[];
  __shared__ double sm32[32];
  if ((int)blockIdx.x >= *c.n_gchunks) return;
  const Chunk ch = c.gchunks[blockIdx.x];
  const int state = c.st[ch.slot].state;
  if (state != ST_DEG && state != ST_LAN) return;
  const double* __restrict__ y0 = c.Y + (int64_t)ch.slot * c.ystride;      // columns < P (counts too large for a panel cell) stay in HBM / L2
  const double* __restrict__ y = y0 + c.P;
  const int ny = c.n_cols - c.P;
  for (int i = threadIdx.x; i < ny; i += 1024) gs_y[i] = y[i];
  __syncthreads();
  const int32_t* __restrict__ col = c.col;
  const void* __restrict__ val = c.val;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  typedef typename RawVal<VT>::T RT;
  const int P = c.P;
  const int dphase = (VT == VT_F64 && c.signed_mode && state == ST_DEG) ? c.st[ch.slot].deg_phase : -1;
  const bool absv = dphase == 0;
  double* __restrict__ dout = dphase == 1 ? c.d2 : c.d;
  double dsum = 0.0;
  long long nz = 0, nz0 = 0;
  int r = warp;
  int64_t rs = 0, re = 0;
  int row_cur = 0;
  if (r < ch.len) { row_cur = c.perm[ch.begin + r]; rs = c.row_ptr[row_cur]; re = c.row_ptr[row_cur + 1]; }
  while (r < ch.len) {
    const int p = ch.begin + r;
    const int rn = r + 32;
    int64_t nrs = 0, nre = 0;
    int nrow = 0;
    if (rn < ch.len) { nrow = c.perm[ch.begin + rn]; nrs = c.row_ptr[nrow]; nre = c.row_ptr[nrow + 1]; }
    nz += re - rs;
    if (P) nz0 += c.row_ptr0[row_cur + 1] - c.row_ptr0[row_cur];
    double acc[U];
#pragma unroll
    for (int u = 0; u < U; ++u) acc[u] = 0.0;
    int64_t k = rs + lane;
    bool have = (k + 32 * (U - 1) < re);
    int cc[U]; RT vv[U];
    if (have) {
#pragma unroll
      for (int u = 0; u < U; ++u) { cc[u] = __ldg(col + k + 32 * u); vv[u] = ld_raw<VT>(val, k + 32 * u); }
    }
    while (have) {
      const int64_t kn = k + 32 * U;
      const bool have_n = (kn + 32 * (U - 1) < re);
      int dc[U]; RT dv[U];
      if (have_n) {
#pragma unroll
        for (int u = 0; u < U; ++u) { dc[u] = __ldg(col + kn + 32 * u); dv[u] = ld_raw<VT>(val, kn + 32 * u); }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) acc[u] += (absv ? fabs((double)vv[u]) : (double)vv[u]) * (cc[u] >= P ? gs_y[cc[u] - P] : __ldg(y0 + cc[u]));
      if (have_n) {
#pragma unroll
        for (int u = 0; u < U; ++u) { cc[u] = dc[u]; vv[u] = dv[u]; }
      }
      k = kn; have = have_n;
    }
    for (; k < re; k += 32) {
      const double v = ld_val<VT>(val, k); const int cj = __ldg(col + k);
      acc[0] += (absv ? fabs(v) : v) * (cj >= P ? gs_y[cj - P] : __ldg(y0 + cj));
    }
    double sacc = 0.0;
#pragma unroll
    for (int u = 0; u < U; ++u) sacc += acc[u];
    double s = warp_sum(sacc);
    if (lane == 0) {
      if (P) s += c.pacc[p];
      if (state == ST_DEG) { const double dv2 = (VT == VT_F64) ? s : s * c.rrow[row_cur]; dout[p] = dv2; dsum += dv2; }
      else c.w[p] = c.sdeg[p] * s;
    }
    r = rn; rs = nrs; re = nre; row_cur = nrow;
  }
  if (lane == 0 && (nz || nz0)) {
    atomicAdd(c.stats + (state == ST_DEG ? 1 : 0), (unsigned long long)(P ? nz0 : nz));
    if (P) atomicAdd(c.stats + 2, (unsigned long long)nz);
  }
  if (state == ST_DEG) {
    dsum = warp_sum(dsum);
    if (lane == 0) sm32[warp] = dsum;
    __syncthreads();
    if (warp == 0) {
      double t = warp_sum(sm32[lane]);
      if (lane == 0) atomicAdd(slot_scal(c, ch.slot) + R_SUMD, t);
    }
  }
}

// Transposed product of the LARGE nodes' residual, y[P, n) += A_res^T xin, from the CSR rows (no column-sorted copy is read): a
// CTA accumulates its gs_chunk_rows rows into a private copy of y[P, n) in shared memory (fp64 shared-memory atomics: a CAS loop,
// but conflicts are rare -- 20 k columns, 1024 threads) and flushes the non-zero totals with one global atomic per column.
// 6 bytes per entry instead of the T-block's 10, and no gather of x at all (x_r is a per-row scalar).  Experimental: TMC_TS=1.
template <int VT, int U>
__global__ void __launch_bounds__(1024, 1) k_tscatter_smem(Ctx c) {
  extern __shared__ double gs_y[];
  if ((int)blockIdx.x >= *c.n_gchunks) return;
  const Chunk ch = c.gchunks[blockIdx.x];
  const int state = c.st[ch.slot].state;
  if (state != ST_DEG && state != ST_LAN && state != ST_MOD) return;
  const int ny = c.n_cols - c.P, P = c.P;
  for (int i = threadIdx.x; i < ny; i += 1024) gs_y[i] = 0.0;
  __syncthreads();
  const int32_t* __restrict__ col = c.col;
  const void* __restrict__ val = c.val;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  typedef typename RawVal<VT>::T RT;
  constexpr bool FACT = (VT != VT_F64);
  const bool use_x = FACT || state != ST_DEG;
  const bool absv = !FACT && c.signed_mode && state == ST_DEG && c.st[ch.slot].deg_phase == 0;
  double* __restrict__ y0 = c.Y + (int64_t)ch.slot * c.ystride;
  auto add = [&](int cj, double v) {      // columns < P (counts too large for a panel cell): straight to y in global memory
    if (cj >= P) atomicAdd(gs_y + (cj - P), v);
    else atomicAdd(y0 + cj, FACT ? v * __ldg(c.w2col + cj) : v);
  };
  for (int r = warp; r < ch.len; r += 32) {
    const int p = ch.begin + r;
    const double xr = use_x ? c.xin[p] : 1.0;
    if (xr == 0.0) continue;                       // label-0 cells of the modularity product
    const int row = c.perm[p];
    const int64_t rs = c.row_ptr[row], re = c.row_ptr[row + 1];
    int64_t k = rs + lane;
    for (; k + 32 * (U - 1) < re; k += 32 * U) {
      int cc[U]; RT vv[U];
#pragma unroll
      for (int u = 0; u < U; ++u) { cc[u] = __ldg(col + k + 32 * u); vv[u] = ld_raw<VT>(val, k + 32 * u); }
#pragma unroll
      for (int u = 0; u < U; ++u) add(cc[u], (absv ? fabs((double)vv[u]) : (double)vv[u]) * xr);
    }
    for (; k < re; k += 32) { const double v = ld_val<VT>(val, k); add(__ldg(col + k), (absv ? fabs(v) : v) * xr); }
  }
  __syncthreads();
  double* __restrict__ y = y0 + P;
  for (int i = threadIdx.x; i < ny; i += 1024) {
    const double v = gs_y[i];
    if (v != 0.0) atomicAdd(y + i, FACT ? v * __ldg(c.w2col + P + i) : v);
  }
}

// ||y_S||^2 for slots in the modularity state; grid (slots_used, ny). Columns are split over ranks.
__global__ void __launch_bounds__(256) k_ynorm(Ctx c) {
  __shared__ double sm8[8];
  int slot = blockIdx.x;
  if (c.st[slot].state != ST_MOD) return;
  const double* __restrict__ y = c.Y + (int64_t)slot * c.ystride;
  int lo = (int)((int64_t)c.n_cols * c.rank / c.world), hi = (int)((int64_t)c.n_cols * (c.rank + 1) / c.world);
  if (c.st[slot].priv || !c.collectives) { lo = 0; hi = c.n_cols; }
  double s = 0.0;
  for (int i = lo + blockIdx.y * blockDim.x + threadIdx.x; i < hi; i += gridDim.y * blockDim.x) {
    double v = y[i];
    if (c.vt != VT_F64) v *= c.invw[i];            // stored y' = w * y
    s += v * v;
  }
  double t = block_sum_256(s, sm8);
  if (threadIdx.x == 0) atomicAdd(slot_scal(c, slot) + R_YNORM2, t);
}

// ------------------------------------------------------------------ K5: full re-orthogonalisation (CGS2)
// h_l = <V_l, w>, l = 0..j  (l = 0 is the locked trivial vector u1 = D^{1/2}1/||.||, the deflation)
__global__ void __launch_bounds__(256) k_dots(Ctx c, int pass) {
  __shared__ double ws[TMC_MAX_CHUNK];
  if ((int)blockIdx.x >= *c.n_vchunks) return;
  const Chunk ch = c.vchunks[blockIdx.x];
  const Slot& s = c.st[ch.slot];
  const int state = s.state;
  if (state != ST_ORTH0 && state != ST_LAN) return;
  const int j = s.j;
  for (int i = threadIdx.x; i < ch.len; i += 256) ws[i] = c.w[ch.begin + i];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* r = pass ? slot_redB(c, ch.slot) : c.redA + (int64_t)ch.slot * (c.K + 2);
  for (int l = warp; l <= j; l += 8) {
    const double* __restrict__ v = c.V + (int64_t)l * c.vstride + ch.begin;
    double a = 0.0;
    for (int i = lane; i < ch.len; i += 32) a += v[i] * ws[i];
    a = warp_sum(a);
    if (lane == 0) atomicAdd(r + l, a);
  }
}

// w -= sum_l h_l V_l ; the first pass also accumulates ||w||^2 (beta_{j+1}^2 = that minus the pass-1 coefficients squared)
__global__ void __launch_bounds__(256) k_update(Ctx c, int pass) {
  extern __shared__ double smem[];
  double* hs = smem;                       // K+2
  double* part = smem + (c.K + 2);         // 8 * chunk_rows
  __shared__ double sm8[8];
  if ((int)blockIdx.x >= *c.n_vchunks) return;
  const Chunk ch = c.vchunks[blockIdx.x];
  const Slot& s = c.st[ch.slot];
  const int state = s.state;
  if (state != ST_ORTH0 && state != ST_LAN) return;
  const int j = s.j;
  const double* r = pass ? slot_redB(c, ch.slot) : c.redA + (int64_t)ch.slot * (c.K + 2);
  for (int l = threadIdx.x; l <= j; l += 256) hs[l] = r[l];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = lane; i < ch.len; i += 32) {
    double a = 0.0;
    for (int l = warp; l <= j; l += 8) a += hs[l] * c.V[(int64_t)l * c.vstride + ch.begin + i];
    part[warp * c.vchunk_rows + i] = a;
  }
  __syncthreads();
  double nn = 0.0;
  for (int i = threadIdx.x; i < ch.len; i += 256) {
    double t = 0.0;
#pragma unroll
    for (int q = 0; q < 8; ++q) t += part[q * c.vchunk_rows + i];
    double wv = c.w[ch.begin + i] - t;
    c.w[ch.begin + i] = wv;
    nn += wv * wv;
  }
  if (pass == 0) {      // ||w||^2 after the first pass; the second pass' tiny coefficients are subtracted in k_finish
    double t = block_sum_256(nn, sm8);
    if (threadIdx.x == 0) atomicAdd(slot_scal(c, ch.slot) + R_NRM2, t);
  }
}

// ------------------------------------------------------------------ per-node scalar logic: one warp per slot
// 2 warps per CTA; each warp keeps its node's T (alpha, beta) and the scratch of the eigen-solve in shared memory, so the
// serial O(j) recurrences of lane 0 run at shared-memory latency instead of an L2 round trip per element.
#define TMC_FINISH_WARPS 2
__global__ void __launch_bounds__(32 * TMC_FINISH_WARPS) k_finish(Ctx c) {
  extern __shared__ double fsm[];
  const int slot = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  double* sA = fsm + (size_t)(threadIdx.x >> 5) * 5 * (c.K + 2);
  double* sB = sA + (c.K + 2); double* sdp = sB + (c.K + 2); double* sdm = sdp + (c.K + 2); double* sz = sdm + (c.K + 2);
  if (slot >= c.n_slots) return;
  Slot* s = c.st + slot;
  const int state = s->state;
  if (state < ST_DEG || state > ST_MOD) return;
  const int K2 = c.K + 2;
  const double* redA = c.redA + (int64_t)slot * K2;
  const double* redB = slot_redB(c, slot);
  double* scal = slot_scal(c, slot);
  if (state == ST_DEG) {
    if (lane == 0 && c.signed_mode && !s->nev && s->deg_phase == 1) {      // second pass done: degrees of B itself
      s->sumd2 = scal[R_SUMD];
      s->action = ACT_INIT;
    } else if (lane == 0) {
      double sumd = scal[R_SUMD];
      s->sumd = sumd; s->sumd2 = sumd;
      if (s->nev) s->inv_sqrt_sumd = 0;
      else if (!(sumd > 0) || isinf(sumd)) { atomicMax(c.err, 7); s->inv_sqrt_sumd = 0; }
      else s->inv_sqrt_sumd = 1.0 / sqrt(sumd);
      if (c.signed_mode && !s->nev) { s->deg_phase = 1; s->action = ACT_NONE; }      // stay in ST_DEG for the pass on B itself
      else s->action = ACT_INIT;
    }
    return;
  }
  if (state == ST_ORTH0) {
    if (lane == 0) {
      double b = sqrt(fmax(scal[R_NRM2] - redB[0] * redB[0], 0.0));
      s->beta = b;
      // a warm start vector that is (numerically) a multiple of u1 carries no direction: start over from the hash
      if (s->warm && !(b * b > 1e-20 * (scal[R_NRM2] + redA[0] * redA[0]))) { s->warm = 0; s->action = ACT_INIT; return; }
      if (!(b > 0)) { s->action = ACT_FINISH; s->split = 0; s->q = 0.0; s->theta = 0.0; }   // degenerate start (cannot happen for len >= 2)
      else s->action = ACT_NEXT;
    }
    return;
  }
  if (state == ST_LAN) {
    const int j = s->j;
    double* gA = c.alpha + (int64_t)slot * K2;
    double* gB = c.beta + (int64_t)slot * K2;
    double* A = sA; double* B = sB;
    for (int i = lane; i < j - 1; i += 32) { A[i] = gA[i]; B[i + 1] = gB[i + 1]; }     // T_{j-1} from the previous steps
    if (lane == 0) B[0] = 0.0;
    __syncwarp();
    double beta = 0.0;
    if (lane == 0) {
      A[j - 1] = redA[j] + redB[j];
      double g2 = 0.0;
      for (int l = 0; l <= j; ++l) g2 += redB[l] * redB[l];
      beta = sqrt(fmax(scal[R_NRM2] - g2, 0.0));
      B[j] = beta;
      gA[j - 1] = A[j - 1]; gB[j] = beta;
      if (j == 1) gB[0] = 0.0;
      s->beta = beta;
      s->iters += 1;
    }
    __syncwarp();
    beta = __shfl_sync(0xffffffffu, beta, 0);
    const int len = s->gsize;
    const bool exhausted = (j >= len - 1);
    const bool breakdown = !(beta > 1e-13);
    const bool full = (j >= c.K);
    const bool check = exhausted || breakdown || full || j <= 64 || (j & 3) == 0;     // T_j is tiny: test every step early on
    if (s->nev) {
      // truncated SVD: the top nev Ritz pairs of T_j must all have converged (tested every 4th step)
      const int nev = min(s->nev, j);
      const bool test = exhausted || breakdown || full || ((j & 3) == 0 && j >= s->nev);
      if (!test) { if (lane == 0) s->action = ACT_NEXT; return; }
      __threadfence_block();
      double lo, hi, piv;
      tmc_gershgorin(A, B, j, &lo, &hi, &piv);
      bool all_ok = true;
      for (int i = 0; i < nev; ++i) {
        const double th = tmc_top_eig_warp(A, B, j, -1e300, 1e300, i + 1);
        double r = 0.0;
        if (lane == 0) {
          double* z = c.zmulti + (int64_t)i * K2;
          tmc_twisted_vector(A, B, j, th, sdp, sdm, sz, piv);
          for (int q = 0; q < j; ++q) z[q] = sz[q];
          c.ev_out[i] = th;
          r = beta * fabs(sz[j - 1]);
        }
        r = __shfl_sync(0xffffffffu, r, 0);
        if (!(r <= c.tol * fmax(th, 1e-300) || r <= 1e-14)) all_ok = false;
      }
      if (lane == 0) {
        const bool stop = all_ok || exhausted || breakdown || full;
        *c.svd_conv = all_ok || exhausted || breakdown ? 1 : 0;
        s->theta = c.ev_out[0];
        s->action = stop ? ACT_RITZ_MULTI : ACT_NEXT;
      }
      return;
    }
    if (c.signed_mode) {
      // signed matrix: nothing is deflated, the wanted vector belongs to the SECOND largest Ritz value of T_j
      const bool exh = (j >= len);                    // the Krylov space is the whole space
      const bool chk = exh || breakdown || full || j <= 64 || (j & 3) == 0;
      if (j < 2 && !exh && !breakdown) { if (lane == 0) s->action = ACT_NEXT; return; }
      if (!chk) { if (lane == 0) s->action = ACT_NEXT; return; }
      __threadfence_block();
      const double th1 = tmc_top_eig_warp(A, B, j, -1e300, 1e300, 1);
      const double th2 = j >= 2 ? tmc_top_eig_warp(A, B, j, -1e300, th1, 2) : th1;
      if (lane == 0) {
        double lo, hi, piv;
        tmc_gershgorin(A, B, j, &lo, &hi, &piv);
        tmc_twisted_vector(A, B, j, th2, sdp, sdm, sz, piv);
        const double resid = beta * fabs(sz[j - 1]);
        s->theta = th2; s->resid = resid;
        const bool conv = exh || breakdown || resid <= c.tol * fmax(th2, 1e-3 * th1) || resid <= 1e-14;
        if (conv || (full && s->restarts >= c.max_restarts)) {
          s->action = ACT_RITZ_LABEL; scal[R_N1] = 0.0; scal[R_SD1] = 0.0; scal[R_SGN] = 0.0;
          if (c.kmeans) {
            unsigned long long* kb = reinterpret_cast<unsigned long long*>(c.km + (int64_t)slot * KM_N);
            kb[KM_MINKEY] = ~0ull; kb[KM_MAXKEY] = 0ull; c.km[(int64_t)slot * KM_N + KM_ROUND] = 0.0;
          }
        }
        else if (full) { s->action = ACT_RITZ_RESTART; s->restarts += 1; }
        else s->action = ACT_NEXT;
        if (s->action != ACT_NEXT) { double* z = c.zvec + (int64_t)slot * K2; for (int q = 0; q < j; ++q) z[q] = sz[q]; }
      }
      return;
    }
    if (!check) { if (lane == 0) s->action = ACT_NEXT; return; }
    __threadfence_block();
    // bracket from the previous test: theta grows monotonically with j and (when the nearest eigenvalue of T_j is the top
    // one) by no more than the previous residual; both hints are verified with a Sturm count before use
    const double theta = tmc_top_eig_warp(A, B, j, s->theta_hint, (s->theta_hint > -1e299) ? s->theta + 1.0001 * s->resid + 1e-15 : 1e300);
    int want_cert = 0; double resid = 0.0, piv = 0.0;
    if (lane == 0) {
      double lo, hi;
      tmc_gershgorin(A, B, j, &lo, &hi, &piv);
      tmc_twisted_vector(A, B, j, theta, sdp, sdm, sz, piv);
      resid = beta * fabs(sz[j - 1]);
      s->theta = theta; s->resid = resid;
      s->theta_hint = theta - 1e-13 * fmax(1.0, fabs(theta));
      bool conv = exhausted || breakdown || resid <= c.tol * theta || resid <= 1e-14;
      if (conv || (full && s->restarts >= c.max_restarts)) {
        s->action = ACT_RITZ_LABEL;
        scal[R_N1] = 0.0; scal[R_SD1] = 0.0; scal[R_SGN] = 0.0;
        if (c.kmeans) {          // k_advance reduces min / max of u2 for the start centroids
          unsigned long long* kb = reinterpret_cast<unsigned long long*>(c.km + (int64_t)slot * KM_N);
          kb[KM_MINKEY] = ~0ull; kb[KM_MAXKEY] = 0ull; c.km[(int64_t)slot * KM_N + KM_ROUND] = 0.0;
        }
      } else if (full) {
        s->action = ACT_RITZ_RESTART; s->restarts += 1; s->theta_hint = -1e300;
      } else s->action = ACT_NEXT;
      // sign-certified stop: worth a look once r / theta (a lower bound of r / gap) is under the slot's probe threshold
      want_cert = s->action == ACT_NEXT && c.cert_on && s->stop_after == 0 && len >= c.cert_min_rows && j >= 3 && resid < s->cert_thr * theta;
      if (s->action != ACT_NEXT || want_cert) { double* z = c.zvec + (int64_t)slot * K2; for (int q = 0; q < j; ++q) z[q] = sz[q]; }   // k_advance forms the Ritz vector
      s->have_z2 = 0;
    }
    const bool warm_ok = c.warm_on && s->stop_after == 0 && len >= c.warm_min_rows && j >= 2;
    int want_z2 = (lane == 0 && warm_ok && s->action == ACT_RITZ_LABEL) ? 1 : 0;
    want_z2 = __shfl_sync(0xffffffffu, want_z2, 0);
    if (want_z2) {      // stopped on the tolerance: the children's start vector needs the second Ritz pair (three digits are plenty)
      const double th2 = tmc_top_eig_warp(A, B, j, -1e300, theta, 2, 1e-7);
      if (lane == 0) {
        tmc_twisted_vector(A, B, j, th2, sdp, sdm, sz, piv);
        double* z2 = c.zvec2 + (int64_t)slot * K2; for (int q = 0; q < j; ++q) z2[q] = sz[q];
        s->have_z2 = 1;
      }
    }
    want_cert = __shfl_sync(0xffffffffu, want_cert, 0);
    if (want_cert) {
      // Davis-Kahan: ||x - u2|| <= ~ r / gap, gap = distance of theta to the rest of the spectrum.  The second Ritz value
      // bounds that from the Krylov space: gap >= theta - theta_2 - r_2 once r_2 is itself small against the distance.
      const double th2 = tmc_top_eig_warp(A, B, j, -1e300, theta, 2, 1e-7);      // an upper bound of theta_2, good to 1e-7: all the gap needs
      if (lane == 0) {
        tmc_twisted_vector(A, B, j, th2, sdp, sdm, sz, piv);
        if (warm_ok) { double* z2 = c.zvec2 + (int64_t)slot * K2; for (int q = 0; q < j; ++q) z2[q] = sz[q]; s->have_z2 = 1; }   // in case the probe ends the node
        const double r2 = beta * fabs(sz[j - 1]);
        const double sep = theta - th2, gap = sep - r2;
        // A start vector that favours u2 hides a nearly degenerate partner u3 (not yet resolved by the Krylov space: the gap
        // seen from T_j is then too large) behind a residual r ~ g rho, rho = weight of u3 / weight of u2 in the start, while the
        // Ritz vector carries the admixture rho itself.  A random start has rho ~ 1.  A warm start is mixed with a random share
        // of about half its norm (k_advance, ACT_INIT), so rho >~ 1 / K with K = 2 |z_1| sqrt(m) (z_1 = component of the Ritz
        // vector along the start vector, 1 / sqrt(m) = the random level); cert_passed scales the safety factor of SMALL nodes
        // by it (the demand on r / gap of a large node is already far beyond what such a pair could pass).
        s->cert_k = (s->warm && c.cert_warm_scale) ? fmax(1.0, c.cert_kfac * fabs(c.zvec[(int64_t)slot * K2]) * sqrt((double)len)) : 1.0;
        if (gap > 0.0 && r2 < 0.25 * sep && resid < s->cert_thr * gap) {
          s->cert_req = 1; s->cert_need = c.cert_safe * resid / gap;
          c.certmin[slot] = 0x7FF0000000000000ull;                 // +inf
          scal[R_N1] = 0.0; scal[R_SD1] = 0.0; scal[R_SGN] = 0.0;                      // in case the probe succeeds and k_advance labels this sweep
        }
      }
    }
    return;
  }
  if (state == ST_KM) {
    // Lloyd's 2-means on u2, k = 2, start = smallest / largest entry, ties to cluster 1, at most 100 rounds -- the convention of
    // variants::kmeans2_labels / oracle kmeans2_labels.  Round r's assignment is made by k_km_assign in this sweep with the
    // centroids set here; the next sweep evaluates it (changed? one cluster empty?) and either moves the centroids or is done.
    if (lane == 0) {
      double* km = c.km + (int64_t)slot * KM_N;
      const int round = (int)km[KM_ROUND];
      const double n = (double)s->gsize;
      bool assign = true;
      if (round == 0) {
        const unsigned long long* kb = reinterpret_cast<const unsigned long long*>(km);
        const double mn = key_f64(kb[KM_MINKEY]), mx = key_f64(kb[KM_MAXKEY]);
        if (mn == mx) { km[KM_C0] = INFINITY; km[KM_C1] = mn; }      // no spread: every cell to cluster 1 (the node becomes a leaf)
        else { km[KM_C0] = mn; km[KM_C1] = mx; }
      } else {
        const double n1 = scal[R_N1];
        const bool changed = round == 1 || km[KM_CHANGED] > 0.0;
        if (!changed || n1 == 0.0 || n1 == n || round >= 100) assign = false;
        else { km[KM_C0] = km[KM_S0] / (n - n1); km[KM_C1] = km[KM_S1] / n1; }
      }
      if (assign) {
        km[KM_S0] = 0.0; km[KM_S1] = 0.0; km[KM_CHANGED] = 0.0; km[KM_ROUND] = (double)(round + 1);
        scal[R_N1] = 0.0; scal[R_SD1] = 0.0; scal[R_SGN] = 0.0;
        s->action = ACT_KM_ASSIGN;
      } else s->action = ACT_KM_DONE;         // labels, xin, N1, SD1 of the last assignment stand
    }
    return;
  }
  // ST_MOD: Newman-Girvan Q of the split on A = B B^T - I  (getBModularity).
  // With t = B^T 1 (||t||^2 = sum d), t1 = B^T 1_{c1}:  O_1 = ||t1||^2 - n1,
  // O_0 = ||t - t1||^2 - n0 = sum d - 2 sum_{c1} d + ||t1||^2 - n0,  L_c = sum_{c} d - n_c,  L = sum d - m.
  if (lane == 0) {
    const double m = (double)s->gsize;
    const double n1 = scal[R_N1], sd1 = scal[R_SD1];
    const double n0 = m - n1, sd = c.signed_mode ? s->sumd2 : s->sumd, t1 = scal[R_YNORM2];
    const double L = sd - m;
    double q = 0.0;
    if (L != 0.0) {
      const double O1 = t1 - n1, O0 = sd - 2.0 * sd1 + t1 - n0;
      const double L1 = sd1 - n1, L0 = (sd - sd1) - n0;
      q = (O0 / L - (L0 / L) * (L0 / L)) + (O1 / L - (L1 / L) * (L1 / L));
    }
    s->q = q;
    // Canonical eigenvector sign: u2 is defined up to its sign, and the sign a Krylov solver returns depends on its start
    // vector.  Label 1 goes to the side S with sum_{i in S} h_i - sum_{i not in S} h_i >= 0 for a fixed pseudo-random h indexed
    // by the global row id -- a function of the PARTITION alone (not of the entries of a Ritz vector that is only converged as
    // far as its signs need), so the child order does not depend on the seed, the warm start, the stop rule, the storage mode
    // or the number of ranks.  Q is symmetric in the two labels; the labels themselves are inverted by k_advance before the split.
    const bool flip = c.canon_sign && s->stop_after == 0 && scal[R_SGN] < 0.0;
    s->flip = flip ? 1 : 0;
    const int in0 = flip ? (int)(n1 + 0.5) : (int)(n0 + 0.5), in1 = flip ? (int)(n0 + 0.5) : (int)(n1 + 0.5);
    s->n0 = in0;
    bool split = (q > c.min_modularity) && in0 >= 1 && in1 >= 1 && in0 >= c.min_size && in1 >= c.min_size;
    if (s->stop_after == ST_DONE) split = false;     // fine-grained API: report Q only
    s->split = split ? 1 : 0;
    s->action = ACT_FINISH;
  }
}

// ------------------------------------------------------------------ sign-certified stop
// The tree needs u2 only through the signs of its entries.  For the Ritz pair (theta, x = V z) with residual r, every label
// is final once min_i |x_i| > SAFE * r / gap (||x|| = 1: z has unit norm and V is orthonormal).  k_finish requests a probe
// (cert_req, z in zvec), this kernel reduces min_i |x_i| over the node's positions, k_advance / k_commit act on the outcome.
// Safety scale of a warm-started node (Slot::cert_k = K).  A hidden pair passes the probe when g / gap < K mn / SAFE (mn = the
// smallest |x_i|); for a cold start that is mn / SAFE, at most ~2.5e-4 for the smallest certified nodes (64 cells) and falling
// like m^-1.5.  The scale keeps a warm node at or under cert_s0 = 1e-4, never above K: large nodes (tiny mn) are not touched.
__device__ __forceinline__ double cert_scale(const Ctx& c, const Slot& s, double mn) {
  return fmax(1.0, fmin(s.cert_k, s.cert_k * mn / (c.cert_safe * c.cert_s0)));
}
__device__ __forceinline__ bool cert_passed(const Ctx& c, int slot, const Slot& s) {
  if (!s.cert_req) return false;
  const double mn = __longlong_as_double((long long)c.certmin[slot]);
  return mn > s.cert_need * cert_scale(c, s, mn);
}
__global__ void __launch_bounds__(256) k_cert_probe(Ctx c) {
  __shared__ double sm8[8];
  if ((int)blockIdx.x >= *c.n_vchunks) return;
  const Chunk ch = c.vchunks[blockIdx.x];
  const Slot& s = c.st[ch.slot];
  if (s.state != ST_LAN || !s.cert_req) return;
  const int j = s.j;
  const double* __restrict__ z = c.zvec + (int64_t)ch.slot * (c.K + 2);
  double mn = INFINITY;
  for (int i = threadIdx.x; i < ch.len; i += 256) {
    const int p = ch.begin + i;
    double u = 0.0;
    for (int l = 1; l <= j; ++l) u += z[l - 1] * c.V[(int64_t)l * c.vstride + p];
    mn = fmin(mn, fabs(u));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
  if ((threadIdx.x & 31) == 0) sm8[threadIdx.x >> 5] = mn;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int q = 1; q < 8; ++q) mn = fmin(mn, sm8[q]);
    atomicMin(c.certmin + ch.slot, (unsigned long long)__double_as_longlong(mn));      // non-negative doubles order like their bits
  }
}

// ------------------------------------------------------------------ per-position transitions
__global__ void __launch_bounds__(256) k_advance(Ctx c) {
  __shared__ double sm8[8];
  if ((int)blockIdx.x >= *c.n_vchunks) return;
  const Chunk ch = c.vchunks[blockIdx.x];
  const Slot& s = c.st[ch.slot];
  int action = s.action;
  if (action == ACT_NEXT && cert_passed(c, ch.slot, s)) action = ACT_RITZ_LABEL;      // every label certified: stop here
  if (action == ACT_FINISH && s.state == ST_MOD && s.flip && s.split) {      // canonical sign (k_finish): invert the labels before the split
    for (int i = threadIdx.x; i < ch.len; i += 256) c.label[ch.begin + i] ^= 1;
    return;
  }
  if (action == ACT_NONE || action == ACT_FINISH || action == ACT_KM_ASSIGN || action == ACT_KM_DONE) return;
  if (action == ACT_INIT && s.nev) {      // truncated SVD of B itself: no D^{-1/2}, no deflation (V_0 = 0)
    for (int i = threadIdx.x; i < ch.len; i += 256) {
      const int p = ch.begin + i;
      c.sdeg[p] = (c.vt == VT_F64) ? 1.0 : c.rrow[c.perm[p]];
      c.V[p] = 0.0;
      c.w[p] = hash_unit(c.seed, (uint64_t)(c.row_offset + c.perm[p]));
    }
    return;
  }
  if (action == ACT_RITZ_MULTI) {
    const int j = s.j, nev = min(s.nev, j);
    for (int i = threadIdx.x; i < ch.len; i += 256) {
      const int p = ch.begin + i;
      for (int e = 0; e < nev; ++e) {
        const double* z = c.zmulti + (int64_t)e * (c.K + 2);
        double u = 0.0;
        for (int l = 1; l <= j; ++l) u += z[l - 1] * c.V[(int64_t)l * c.vstride + p];
        c.uout[(int64_t)c.perm[p] * s.nev + e] = u;
      }
      for (int e = nev; e < s.nev; ++e) c.uout[(int64_t)c.perm[p] * s.nev + e] = 0.0;
    }
    return;
  }
  if (action == ACT_INIT) {
    const double isd = s.inv_sqrt_sumd;
    for (int i = threadIdx.x; i < ch.len; i += 256) {
      const int p = ch.begin + i;
      const double dd = c.d[p];
      if (!(dd > 0)) atomicMax(c.err, 7);
      const double sq = sqrt(dd);
      c.sdeg[p] = (c.vt == VT_F64 ? 1.0 : c.rrow[c.perm[p]]) / sq;      // D^{-1/2} (times the row factor in the factored modes)
      c.V[p] = c.signed_mode ? 0.0 : sq * isd;             // V_0 = u1 (nothing to deflate when B has negative entries)
      // warm start: D_child^{1/2} f plus a random share of norm ~ warm_mix / 2 (the hash has variance 1/3 per entry; the warm
      // part has norm ~ 0.3 - 0.8), so that every direction keeps at least the weight a cold start would give it
      const double hsh = hash_unit(c.seed, (uint64_t)(c.row_offset + c.perm[p]));
      c.w[p] = s.warm ? c.ws[p] * sq + hsh * (0.5 * c.warm_mix * sqrt(3.0 / (double)s.gsize)) : hsh;
    }
    return;
  }
  if (action == ACT_NEXT) {
    const int jn = (s.state == ST_ORTH0) ? 1 : s.j + 1;
    const double ib = 1.0 / s.beta;
    double* __restrict__ vn = c.V + (int64_t)jn * c.vstride;
    for (int i = threadIdx.x; i < ch.len; i += 256) {
      const int p = ch.begin + i;
      const double v = c.w[p] * ib;
      vn[p] = v;
      c.xin[p] = c.sdeg[p] * v;
    }
    return;
  }
  // Ritz vector u = sum_{l=1..j} z_l V_l
  const int j = s.j;
  const double* __restrict__ z = c.zvec + (int64_t)ch.slot * (c.K + 2);
  const double* __restrict__ z2 = c.zvec2 + (int64_t)ch.slot * (c.K + 2);
  const bool leave_ws = c.warm_on && action == ACT_RITZ_LABEL && s.have_z2;      // the start vector of the children (see Ctx::ws)
  double n1 = 0.0, sd1 = 0.0, sgn = 0.0;
  double umin = INFINITY, umax = -INFINITY;
  for (int i = threadIdx.x; i < ch.len; i += 256) {
    const int p = ch.begin + i;
    double u = 0.0, g = 0.0;
    if (leave_ws) {
      for (int l = 1; l <= j; ++l) { const double v = c.V[(int64_t)l * c.vstride + p]; u += z[l - 1] * v; g += z2[l - 1] * v; }
      c.ws[p] = g / sqrt(c.d[p]);
    } else {
      for (int l = 1; l <= j; ++l) u += z[l - 1] * c.V[(int64_t)l * c.vstride + p];
    }
    if (action == ACT_RITZ_RESTART) { c.w[p] = u; continue; }
    c.u2[p] = u;
    umin = fmin(umin, u); umax = fmax(umax, u);
    const int lab = (u >= 0.0) ? 1 : 0;                   // spectralCluster sign rule (A6)
    if (c.canon_sign) { const double h = hash_unit(0x7A6D635F7369676Eull, (uint64_t)(c.row_offset + c.perm[p])); sgn += lab ? h : -h; }
    c.label[p] = (uint8_t)lab;
    c.xin[p] = lab ? (c.vt == VT_F64 ? 1.0 : c.rrow[c.perm[p]]) : 0.0;
    n1 += lab; sd1 += lab ? (c.signed_mode ? c.d2[p] : c.d[p]) : 0.0;
  }
  if (action == ACT_RITZ_LABEL) {
    double t = block_sum_256(n1, sm8);
    double t2 = block_sum_256(sd1, sm8);
    if (threadIdx.x == 0) { atomicAdd(slot_scal(c, ch.slot) + R_N1, t); atomicAdd(slot_scal(c, ch.slot) + R_SD1, t2); }
    if (c.canon_sign) { const double t3 = block_sum_256(sgn, sm8); if (threadIdx.x == 0) atomicAdd(slot_scal(c, ch.slot) + R_SGN, t3); }
    if (c.kmeans) {          // smallest / largest entry of u2: the start centroids of the 2-means
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) { umin = fmin(umin, __shfl_xor_sync(0xffffffffu, umin, o)); umax = fmax(umax, __shfl_xor_sync(0xffffffffu, umax, o)); }
      if ((threadIdx.x & 31) == 0 && umin <= umax) {
        unsigned long long* kb = reinterpret_cast<unsigned long long*>(c.km + (int64_t)ch.slot * KM_N);
        atomicMin(kb + KM_MINKEY, f64_key(umin)); atomicMax(kb + KM_MAXKEY, f64_key(umax));
      }
    }
  }
}

// One Lloyd round of the 2-means on u2 (slots in ST_KM whose k_finish asked for an assignment): label, xin, cluster sums
__global__ void __launch_bounds__(256) k_km_assign(Ctx c) {
  __shared__ double sm8[8];
  if ((int)blockIdx.x >= *c.n_vchunks) return;
  const Chunk ch = c.vchunks[blockIdx.x];
  const Slot& s = c.st[ch.slot];
  if (s.state != ST_KM || s.action != ACT_KM_ASSIGN) return;
  double* km = c.km + (int64_t)ch.slot * KM_N;
  const double c0 = km[KM_C0], c1 = km[KM_C1];
  double n1 = 0.0, sd1 = 0.0, s0 = 0.0, s1 = 0.0, chg = 0.0;
  for (int i = threadIdx.x; i < ch.len; i += 256) {
    const int p = ch.begin + i;
    const double u = c.u2[p];
    const double d0 = (u - c0) * (u - c0), d1 = (u - c1) * (u - c1);
    const int l = d1 <= d0 ? 1 : 0;                 // ties go to cluster 1
    chg += (l != (int)c.label[p]);
    c.label[p] = (uint8_t)l;
    c.xin[p] = l ? (c.vt == VT_F64 ? 1.0 : c.rrow[c.perm[p]]) : 0.0;
    n1 += l; s1 += l ? u : 0.0; s0 += l ? 0.0 : u;
    sd1 += l ? (c.signed_mode ? c.d2[p] : c.d[p]) : 0.0;
  }
  const double a = block_sum_256(n1, sm8), b2 = block_sum_256(sd1, sm8), e0 = block_sum_256(s0, sm8), e1 = block_sum_256(s1, sm8),
               g = block_sum_256(chg, sm8);
  if (threadIdx.x == 0) {
    atomicAdd(slot_scal(c, ch.slot) + R_N1, a); atomicAdd(slot_scal(c, ch.slot) + R_SD1, b2);
    atomicAdd(km + KM_S0, e0); atomicAdd(km + KM_S1, e1); atomicAdd(km + KM_CHANGED, g);
  }
}

// ------------------------------------------------------------------ K7: stable partition of the node's position range
// one CTA per slot (grid = slots used); children = [label 0 | label 1], original order kept (A8).
__global__ void __launch_bounds__(1024) k_partition(Ctx c) {
  __shared__ int wtot[32];
  __shared__ int run1_s;
  const int slot = blockIdx.x;
  const Slot s = c.st[slot];
  if (s.state != ST_MOD || s.action != ACT_FINISH || !s.split) return;
  const int begin = s.begin, end = s.end;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  __shared__ int n0_s;
  // local label-0 count (the node's cells on other ranks are partitioned by their owners)
  int ones = 0;
  for (int p = begin + tid; p < end; p += 1024) ones += c.label[p];
  ones = __reduce_add_sync(0xffffffffu, ones);
  if (lane == 0) wtot[wid] = ones;
  __syncthreads();
  if (tid == 0) { int t = 0; for (int q = 0; q < 32; ++q) t += wtot[q]; n0_s = (end - begin) - t; run1_s = 0; c.st[slot].n0_local = n0_s; }
  __syncthreads();
  const int n0 = n0_s;
  for (int base = begin; base < end; base += 1024) {
    const int p = base + tid;
    const int f = (p < end) ? (int)c.label[p] : 0;
    unsigned bal = __ballot_sync(0xffffffffu, f);
    int before = __popc(bal & ((1u << lane) - 1u));
    if (lane == 0) wtot[wid] = __popc(bal);
    __syncthreads();
    int run1 = run1_s;
    int wbase = 0;
    for (int q = 0; q < wid; ++q) wbase += wtot[q];
    int ones_before = run1 + wbase + before;               // label-1 positions before p in the node
    if (p < end) {
      int idx = p - begin;
      int dest = f ? (begin + n0 + ones_before) : (begin + (idx - ones_before));
      c.perm_tmp[dest] = c.perm[p];
      if (c.warm_on && s.have_z2) c.ws_tmp[dest] = c.ws[p];
      if (c.newpos) c.newpos[p] = dest;
    }
    __syncthreads();
    if (tid == 0) { int t = 0; for (int q = 0; q < 32; ++q) t += wtot[q]; run1_s = run1 + t; }
    __syncthreads();
  }
  for (int p = begin + tid; p < end; p += 1024) c.perm[p] = c.perm_tmp[p];
  if (c.warm_on && s.have_z2) for (int p = begin + tid; p < end; p += 1024) c.ws[p] = c.ws_tmp[p];
}

// apply the pending action to the slot state and publish the report the host reads
__global__ void k_commit(Ctx c, int n_slots) {
  int slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= n_slots) return;
  Slot* s = c.st + slot;
  int state = s->state;
  if (s->cert_req) {
    if (s->action == ACT_NEXT && cert_passed(c, slot, *s)) s->action = ACT_RITZ_LABEL;
    else {          // the smallest entry decides when the next probe can succeed
      const double mn = __longlong_as_double((long long)c.certmin[slot]);
      s->cert_thr = (mn > 0.0 && mn < 1e300) ? fmin(mn / (c.cert_safe * cert_scale(c, *s, mn)), 0.5 * s->cert_thr) : 0.0;
    }
    s->cert_req = 0;
  }
  switch (s->action) {
    case ACT_INIT: state = ST_ORTH0; s->j = 0; break;
    case ACT_NEXT: if (state == ST_ORTH0) { state = ST_LAN; s->j = 1; } else s->j += 1; break;
    case ACT_RITZ_LABEL: state = c.kmeans ? ST_KM : ST_MOD; break;
    case ACT_KM_ASSIGN: break;
    case ACT_KM_DONE: state = ST_MOD; break;
    case ACT_RITZ_RESTART: state = ST_ORTH0; s->j = 0; break;
    case ACT_FINISH: state = ST_DONE; break;
    case ACT_RITZ_MULTI: state = ST_DONE; break;
    default: break;
  }
  if (s->stop_after != 0 && state == s->stop_after && state != s->state) state = ST_DONE;   // fine-grained API stop
  s->action = ACT_NONE;
  s->state = state;
  Report r; r.state = state; r.split = s->split; r.n0 = s->n0; r.iters = s->iters; r.n0_local = s->n0_local; r.have_z2 = s->have_z2; r.tnnz0 = s->tnnz0;
  r.q = s->q; r.theta = s->theta;
  c.rep[slot] = r;
}

// factored modes: the degree pass needs xin = rrow on the node's positions
__global__ void __launch_bounds__(256) k_deg_xin(Ctx c) {
  if ((int)blockIdx.x >= *c.n_vchunks) return;
  const Chunk ch = c.vchunks[blockIdx.x];
  if (c.st[ch.slot].state != ST_DEG) return;
  for (int i = threadIdx.x; i < ch.len; i += 256) { const int p = ch.begin + i; c.xin[p] = c.rrow[c.perm[p]]; }
}

// helpers for the fine-grained API
__global__ void k_iota(int* p, int n, int offset) { int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) p[i] = i + offset; }
__global__ void k_set_xin_scaled(Ctx c, const double* __restrict__ x, int n) {   // xin = D^{-1/2} x
  int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) c.xin[i] = c.sdeg[i] * x[i];
}
__global__ void k_set_labels(Ctx c, const uint8_t* __restrict__ lab, int n, int slot) {
  __shared__ double sm8[8];
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  double n1 = 0.0, sd1 = 0.0;
  if (i < n) { int l = lab[i] ? 1 : 0; c.label[i] = (uint8_t)l; c.xin[i] = l ? (c.vt == VT_F64 ? 1.0 : c.rrow[c.perm[i]]) : 0.0; n1 = l; sd1 = l ? (c.signed_mode ? c.d2[i] : c.d[i]) : 0.0; }
  double t = block_sum_256(n1, sm8);
  double t2 = block_sum_256(sd1, sm8);
  if (threadIdx.x == 0) { atomicAdd(slot_scal(c, slot) + R_N1, t); atomicAdd(slot_scal(c, slot) + R_SD1, t2); }
}
__global__ void k_force_state(Ctx c, int slot, int state, int j, int stop_after) {
  c.st[slot].state = state; c.st[slot].j = j; c.st[slot].action = ACT_NONE; c.st[slot].stop_after = stop_after;
  slot_scal(c, slot)[R_N1] = 0.0; slot_scal(c, slot)[R_SD1] = 0.0;
}

// Batched permutation test (csrc/tmc_engine.cu, permutation_pvals_batched): many disjoint nodes, one slot each, re-opened in the
// modularity state run after run with shuffled labels.
__global__ void k_force_mod_batch(Ctx c, int n_slots) {
  const int slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= n_slots) return;
  Slot& s = c.st[slot];
  s.state = ST_MOD; s.j = 0; s.action = ACT_NONE; s.stop_after = ST_DONE; s.split = 0;      // stop_after = ST_DONE: report Q only, never split
  slot_scal(c, slot)[R_N1] = 0.0; slot_scal(c, slot)[R_SD1] = 0.0;
}
__global__ void __launch_bounds__(256) k_set_labels_batch(Ctx c, const uint8_t* __restrict__ lab, const int* __restrict__ pos2slot, int n) {
  __shared__ double sm8[8];
  __shared__ int same_s;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  int slot = -1; double n1 = 0.0, sd1 = 0.0;
  if (i < n) {
    slot = pos2slot[i];
    const int l = lab[i] ? 1 : 0;
    c.label[i] = (uint8_t)l;
    c.xin[i] = l ? (c.vt == VT_F64 ? 1.0 : c.rrow[c.perm[i]]) : 0.0;
    n1 = l; sd1 = l ? (c.signed_mode ? c.d2[i] : c.d[i]) : 0.0;
  }
  // a block inside one (large) node adds its totals once; blocks that straddle nodes add per cell (small nodes: spread addresses)
  const int first = pos2slot[min(blockIdx.x * blockDim.x, n - 1)];
  const int same = __syncthreads_and(slot == first || slot < 0);
  if (threadIdx.x == 0) same_s = same;
  __syncthreads();
  if (same_s) {
    const double t = block_sum_256(n1, sm8);
    const double t2 = block_sum_256(sd1, sm8);
    if (threadIdx.x == 0 && t != 0.0) { atomicAdd(slot_scal(c, first) + R_N1, t); atomicAdd(slot_scal(c, first) + R_SD1, t2); }
  } else if (slot >= 0 && n1 != 0.0) {
    atomicAdd(slot_scal(c, slot) + R_N1, n1); atomicAdd(slot_scal(c, slot) + R_SD1, sd1);
  }
}

// tmc_partition API: set up slot `slot` so that k_partition splits [begin,end)
__global__ void k_force_partition(Ctx c, int slot, int begin, int end) {
  Slot& s = c.st[slot];
  s.begin = begin; s.end = end; s.gsize = end - begin; s.state = ST_MOD; s.action = ACT_FINISH; s.split = 1;
}

}  // namespace tmc
