// Direct solve of SMALL nodes (DESIGN.md "small nodes"): a node of at most TMC_SMALL_MAX cells, all on this rank, does not go
// through the sweeps.  One CTA forms the node's Gram matrix G = B_S B_S^T (cosine similarities) once -- every sub-node's
// similarity matrix is a principal sub-matrix of it, because B is normalised once at the root and children take rows of B
//  -- and then bipartitions the WHOLE subtree in shared memory with the reference's own arithmetic:
//   d = G_T 1, M = D^{-1/2} G_T D^{-1/2} with u1 = D^{1/2} 1 deflated (bToD / bdToC, A3), u2 = its top eigenvector (A4/A5: here
//   Householder tridiagonalisation + the engine's Sturm multisection and twisted factorisation instead of Lanczos -- for <= 64
//   cells the Krylov space would be the whole space anyway), labels = sign rule (A6) with the engine's canonical sign,
//   Q from G (A7: L = sum d - m, O_c = 1_c^T G 1_c - m_c, L_c = 1_c^T d - m_c), split iff Q > min_modularity and both sides
//   non-empty and >= min_size (A0), stable partition (A8).
// A saturated tree spends most of its sweeps on such nodes (c5: 5.5 M Lanczos steps on 1 M inner vertices, almost all tiny).
#pragma once
#include "tmc_kernels.cuh"

namespace tmc {

#define TMC_SMALL_MAX 64
#define TMC_SMALL_LD 65            // leading dimension of the shared-memory matrices (doubles)
#define TMC_SMALL_TW 64            // panel words per row and tile
#define TMC_SMALL_TLD 65

struct SmallNode { int begin, len; long long rec0; };      // positions [begin, begin + len) of perm; first record of its subtree
struct SmallRec { int size, leaf; double q, theta; };      // pre-order records of the subtree (the node itself first)

// sum over threads 0..63 of a CTA of 256 (other threads pass anything); every thread gets the result; fixed order
__device__ __forceinline__ double small_sum64(double v, double* sm2) {
  if (threadIdx.x < 64) {
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) sm2[threadIdx.x >> 5] = v;
  }
  __syncthreads();
  const double r = sm2[0] + sm2[1];
  __syncthreads();
  return r;
}

__device__ __forceinline__ unsigned small_panel_cell(const Ctx& c, int row, int col) {
  const uint8_t b = __ldg(c.panel + (size_t)row * c.pw + (c.pnib ? (col >> 1) : col));
  return c.pnib ? ((b >> (4 * (col & 1))) & 15u) : (unsigned)b;
}

__device__ __forceinline__ unsigned small_nzmask(unsigned w) { unsigned m = w | (w >> 1); m |= m >> 2; return m & 0x11111111u; }   // bit 4k = nibble k != 0

// per-CTA scratch in global memory: [n_cols * 8 bytes] column heads (int) of the fast path, aliased with the dense vector (double)
// of the fallback path -- both are all-zero between uses -- then nxt int[cap], erow int[cap], ev double[cap]
__host__ __device__ static inline size_t small_scratch_stride(int n_cols, int cap) { return ((size_t)n_cols * 8 + (size_t)cap * 16 + 255) & ~(size_t)255; }

template <int VT>
__global__ void __launch_bounds__(256) k_small_subtree(Ctx c, const SmallNode* __restrict__ nodes, int n_nodes, SmallRec* __restrict__ recs,
                                                       char* __restrict__ scratch, int cap, unsigned long long* __restrict__ prof) {
  extern __shared__ double ssm[];
  double* Gs = ssm;                                        // [64][65] Gram matrix of the node
  double* Ms = Gs + TMC_SMALL_MAX * TMC_SMALL_LD;          // [64][65] deflated normalised matrix of the current sub-node / Householder vectors
  double* w2s = Ms;                                        // (Gram phase) column weights of the panel tile: <= 512 doubles inside Ms
  unsigned* tile = reinterpret_cast<unsigned*>(Ms + 512);  // (Gram phase) [64][65] panel words, also inside Ms (512 + 2080 doubles <= 4160)
  int* ovf = reinterpret_cast<int*>(Ms);                   // (Gram phase, residual part) entry ids / columns of the residual entries in panel columns
  int* ovc = ovf + 4096;
  double* dd = Ms + TMC_SMALL_MAX * TMC_SMALL_LD;          // degrees of the sub-node
  double* sq = dd + 64; double* ta = sq + 64; double* tb = ta + 64; double* tz = tb + 64; double* tdp = tz + 64; double* tdm = tdp + 64;
  double* hv = tdm + 64; double* hp = hv + 64; double* tau = hp + 64; double* rr = tau + 64; double* sm2 = rr + 64;     // sm2: 4 doubles
  int* rows = reinterpret_cast<int*>(sm2 + 4); int* idx = rows + 64; int* idx2 = idx + 64; int* lab = idx2 + 64;
  int* st_lo = lab + 64; int* st_hi = st_lo + 72; int* ctl = st_hi + 72;      // ctl: [0] stack pointer, [1] records written, [2..] scalars
  int* off = st_lo;                                        // (Gram phase) first entry id of every row, off[k] = entries of the node
  constexpr bool FACT = (VT != VT_F64);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  char* base = scratch + (size_t)blockIdx.x * small_scratch_stride(c.n_cols, cap);
  double* scr = reinterpret_cast<double*>(base);
  int* head = reinterpret_cast<int*>(base);
  int* nxt = reinterpret_cast<int*>(base + (size_t)c.n_cols * 8);
  int* erow = nxt + cap;
  double* ev = reinterpret_cast<double*>(erow + cap);
  const int P = c.P;
  long long tp0 = 0;
  auto tick = [&](int phase) { if (prof && tid == 0) { const long long t = clock64(); atomicAdd(prof + phase, (unsigned long long)(t - tp0)); tp0 = t; } };

  for (int nid = blockIdx.x; nid < n_nodes; nid += gridDim.x) {
    const SmallNode nd = nodes[nid];
    const int k = nd.len;
    __syncthreads();
    if (prof && tid == 0) tp0 = clock64();
    if (tid < k) { const int r = c.perm[nd.begin + tid]; rows[tid] = r; idx[tid] = tid; rr[tid] = FACT ? c.rrow[r] : 1.0; }
    for (int e = tid; e < TMC_SMALL_MAX * TMC_SMALL_LD; e += 256) Gs[e] = 0.0;
    if (tid == 0) { ctl[2] = 0; }
    __syncthreads();
    // entries per row, residual entries in panel columns ("overflow": counts too large for a cell; the cell holds 0 there)
    for (int i = warp; i < k; i += 8) {
      const int64_t rs = c.row_ptr[rows[i]], re = c.row_ptr[rows[i] + 1];
      int no = 0;
      if (P > 0) for (int64_t e = rs + lane; e < re; e += 32) no += (__ldg(c.col + e) < P);
      no = __reduce_add_sync(0xffffffffu, no);
      if (lane == 0) { off[i + 1] = (int)min((int64_t)0x3fffffff, re - rs); if (no) atomicAdd(ctl + 2, no); }
    }
    __syncthreads();
    if (tid == 0) { off[0] = 0; long long t = 0; for (int i = 0; i < k; ++i) { t += off[i + 1]; off[i + 1] = (int)min(t, (long long)0x7fffffff); } ctl[3] = (t <= cap && ctl[2] <= 4096) ? 1 : 0; ctl[2] = 0; }
    __syncthreads();
    const bool fast = ctl[3] != 0;

    if (fast) {
      // ---- residual CSR part, fast path: per-column lists of the node's entries (head / nxt in global scratch); every entry
      // walks the entries inserted before it in its column: each co-occurring pair of rows is visited exactly once
      // (sum_c n_c (n_c - 1) / 2 products instead of a gather per pair of rows and entry).
      for (int i = warp; i < k; i += 8) {
        const int64_t rs = c.row_ptr[rows[i]], re = c.row_ptr[rows[i] + 1];
        double dsum = 0.0;
        for (int64_t e = rs + lane; e < re; e += 32) {
          const int col = __ldg(c.col + e);
          const double v = ld_val<VT>(c.val, e);
          const int id = off[i] + (int)(e - rs);
          __stcg(erow + id, i); __stcg(ev + id, v);
          __stcg(nxt + id, atomicExch(head + col, id + 1));
          dsum += v * v * (FACT ? __ldg(c.w2col + col) : 1.0);
          if (P > 0 && col < P) { const int o = atomicAdd(ctl + 2, 1); ovf[o] = id; ovc[o] = col; }
        }
        dsum = warp_sum(dsum);
        if (lane == 0) Gs[i * TMC_SMALL_LD + i] = dsum;
      }
      __syncthreads();
      for (int i = warp; i < k; i += 8) {
        const int64_t rs = c.row_ptr[rows[i]], re = c.row_ptr[rows[i] + 1];
        for (int64_t e = rs + lane; e < re; e += 32) {
          int p = __ldcg(nxt + off[i] + (int)(e - rs));
          if (!p) continue;
          const int col = __ldg(c.col + e);
          const double vw = ld_val<VT>(c.val, e) * (FACT ? __ldg(c.w2col + col) : 1.0);
          while (p) {
            atomicAdd(Gs + i * TMC_SMALL_LD + __ldcg(erow + p - 1), vw * __ldcg(ev + p - 1));
            p = __ldcg(nxt + p - 1);
          }
        }
      }
      const int n_ovf = ctl[2];
      for (int t = tid; t < n_ovf * k; t += 256) {
        const int o = t / k, i2 = t % k;
        const int id = ovf[o], col = ovc[o];
        const int i = __ldcg(erow + id);
        if (i2 == i) continue;
        const unsigned pc = small_panel_cell(c, rows[i2], col);
        if (pc) atomicAdd(Gs + i * TMC_SMALL_LD + i2, (double)pc * __ldcg(ev + id) * __ldg(c.w2col + col));
      }
      __syncthreads();
      for (int i = warp; i < k; i += 8) {
        const int64_t rs = c.row_ptr[rows[i]], re = c.row_ptr[rows[i] + 1];
        for (int64_t e = rs + lane; e < re; e += 32) __stcg(head + __ldg(c.col + e), 0);
      }
      // every pair landed on one side of the diagonal or the other
      for (int e = tid; e < k * k; e += 256) {
        const int i = e / k, j = e % k;
        if (i < j) { const double sxy = Gs[i * TMC_SMALL_LD + j] + Gs[j * TMC_SMALL_LD + i]; Gs[i * TMC_SMALL_LD + j] = sxy; Gs[j * TMC_SMALL_LD + i] = sxy; }
      }
      __syncthreads();
    } else {
      // ---- residual CSR part, fallback (more entries than the scratch lists hold): row j is scattered into a dense scratch
      // vector, the rows i >= j gather from it; plus the cross terms of j's entries in panel columns with the other rows' cells.
      for (int j = 0; j < k; ++j) {
        const int64_t rs = c.row_ptr[rows[j]], re = c.row_ptr[rows[j] + 1];
        for (int64_t e = rs + tid; e < re; e += 256) {
          const int col = __ldg(c.col + e);
          const double v = ld_val<VT>(c.val, e);
          __stcg(scr + col, FACT ? v * __ldg(c.w2col + col) : v);
        }
        __syncthreads();
        for (int i = j + warp; i < k; i += 8) {
          const int64_t is = c.row_ptr[rows[i]], ie = c.row_ptr[rows[i] + 1];
          double a = 0.0;
          for (int64_t e = is + lane; e < ie; e += 32) a += ld_val<VT>(c.val, e) * __ldcg(scr + __ldg(c.col + e));
          a = warp_sum(a);
          if (lane == 0) { Gs[i * TMC_SMALL_LD + j] += a; if (i != j) Gs[j * TMC_SMALL_LD + i] += a; }
        }
        __syncthreads();
        if (P > 0) {
          for (int i = warp; i < k; i += 8) {
            if (i == j) continue;
            double a = 0.0;
            for (int64_t e = rs + lane; e < re; e += 32) {
              const int col = __ldg(c.col + e);
              if (col < P) { const unsigned pc = small_panel_cell(c, rows[i], col); if (pc) a += (double)pc * __ldcg(scr + col); }
            }
            a = warp_sum(a);
            if (lane == 0 && a != 0.0) { Gs[i * TMC_SMALL_LD + j] += a; Gs[j * TMC_SMALL_LD + i] += a; }
          }
          __syncthreads();
        }
        for (int64_t e = rs + tid; e < re; e += 256) __stcg(scr + __ldg(c.col + e), 0.0);
        __syncthreads();
      }
    }
    tick(0);

    // ---- Gram matrix, dense panel part: sum_c a_ic a_jc w_c^2 over the panel columns, tile by tile through shared memory.
    // Thread t owns column j = t % 64 and the rows i = t / 64 + 4 q (upper triangle only); a warp shares i (broadcast reads).
    // Only cells that are non-zero in both rows are multiplied (nibble masks).
    if (P > 0) {
      const int pwords = c.pw >> 2, cpw = c.pnib ? 8 : 4;
      double acc[16];
#pragma unroll
      for (int q = 0; q < 16; ++q) acc[q] = 0.0;
      const int j = tid & 63, ig = tid >> 6;
      for (int w0 = 0; w0 < pwords; w0 += TMC_SMALL_TW) {
        __syncthreads();
        for (int e = tid; e < k * TMC_SMALL_TW; e += 256) {
          const int i = e / TMC_SMALL_TW, w = e % TMC_SMALL_TW;
          tile[i * TMC_SMALL_TLD + w] = (w0 + w < pwords) ? __ldg(reinterpret_cast<const unsigned*>(c.panel + (size_t)rows[i] * c.pw) + w0 + w) : 0u;
        }
        for (int e = tid; e < TMC_SMALL_TW * cpw; e += 256) { const int col = w0 * cpw + e; w2s[e] = col < P ? __ldg(c.w2col + col) : 0.0; }
        __syncthreads();
        if (j < k) {
          const unsigned* tj = tile + j * TMC_SMALL_TLD;
          for (int w = 0; w < TMC_SMALL_TW; ++w) {
            const unsigned wb = tj[w];
            if (c.pnib) {
              const unsigned mb = small_nzmask(wb);
              if (!mb) continue;
#pragma unroll
              for (int q = 0; q < 16; ++q) {
                const int i = ig + 4 * q;
                if (i < k && i <= j) {
                  const unsigned wa = tile[i * TMC_SMALL_TLD + w];
                  unsigned m = small_nzmask(wa) & mb;            // nibbles that are non-zero in both rows
                  while (m) {
                    const int bit = __ffs(m) - 1; m &= m - 1;
                    acc[q] += (double)(((wa >> bit) & 15u) * ((wb >> bit) & 15u)) * w2s[w * 8 + (bit >> 2)];
                  }
                }
              }
            } else {
              if (!wb) continue;
#pragma unroll
              for (int q = 0; q < 16; ++q) {
                const int i = ig + 4 * q;
                if (i < k && i <= j) {
                  const unsigned wa = tile[i * TMC_SMALL_TLD + w];
#pragma unroll
                  for (int b = 0; b < 4; ++b) {
                    const unsigned pa = (wa >> (8 * b)) & 255u, pb = (wb >> (8 * b)) & 255u;
                    if (pa && pb) acc[q] += (double)(pa * pb) * w2s[w * 4 + b];
                  }
                }
              }
            }
          }
        }
      }
      __syncthreads();
      if (j < k) {
#pragma unroll
        for (int q = 0; q < 16; ++q) {
          const int i = ig + 4 * q;
          if (i < k && i <= j) { Gs[i * TMC_SMALL_LD + j] += acc[q]; if (i != j) Gs[j * TMC_SMALL_LD + i] += acc[q]; }
        }
      }
      __syncthreads();
    }
    if (FACT) {
      for (int e = tid; e < k * k; e += 256) { const int i = e / k, j = e % k; Gs[i * TMC_SMALL_LD + j] *= rr[i] * rr[j]; }
      __syncthreads();
    }
    tick(1);

    // ---- the subtree, depth first (pre-order records)
    if (tid == 0) { st_lo[0] = 0; st_hi[0] = k; ctl[0] = 1; ctl[1] = 0; }
    __syncthreads();
    while (ctl[0] > 0) {
      __syncthreads();
      const int sp = ctl[0] - 1;
      const int lo = st_lo[sp], hi = st_hi[sp], t = hi - lo;
      const int rec_i = ctl[1];
      __syncthreads();
      if (tid == 0) { ctl[0] = sp; ctl[1] = rec_i + 1; }
      SmallRec rec; rec.size = t; rec.leaf = 1; rec.q = 0.0; rec.theta = 0.0;
      if (t < 2) {
        if (tid == 0) recs[nd.rec0 + rec_i] = rec;
        __syncthreads();
        continue;
      }
      // degrees d = G_T 1 (the diagonal, the cell's similarity with itself, included: bToD)
      double dv = 0.0;
      if (tid < t) {
        const double* g = Gs + idx[lo + tid] * TMC_SMALL_LD;
        for (int b = 0; b < t; ++b) dv += g[idx[lo + b]];
        dd[tid] = dv; sq[tid] = sqrt(dv);
        if (!(dv > 0.0) || isinf(dv)) atomicMax(c.err, 7);
      }
      const double sumd = small_sum64(tid < t ? dv : 0.0, sm2);
      // M = D^{-1/2} G_T D^{-1/2} - u1 u1^T,  u1 = D^{1/2} 1 / sqrt(sum d)
      for (int e = tid; e < t * t; e += 256) {
        const int a = e / t, b = e % t;
        Ms[a * TMC_SMALL_LD + b] = Gs[idx[lo + a] * TMC_SMALL_LD + idx[lo + b]] / (sq[a] * sq[b]) - sq[a] * sq[b] / sumd;
      }
      __syncthreads();
      // Householder tridiagonalisation (lower triangle referenced; reflector k kept in column k below the sub-diagonal)
      for (int kk = 0; kk + 2 < t; ++kk) {
        const int n = t - kk - 1;                   // the column below the diagonal: rows kk+1 .. t-1
        double x = (tid < n) ? Ms[(kk + 1 + tid) * TMC_SMALL_LD + kk] : 0.0;
        const double nrm2 = small_sum64(tid < n ? x * x : 0.0, sm2);
        const double x0 = Ms[(kk + 1) * TMC_SMALL_LD + kk];
        const double tail2 = fmax(nrm2 - x0 * x0, 0.0);
        if (!(tail2 > 1e-300)) {                    // already tridiagonal in this column
          if (tid == 0) { ta[kk] = Ms[kk * TMC_SMALL_LD + kk]; tb[kk + 1] = x0; tau[kk] = 0.0; }
          __syncthreads();
          continue;
        }
        const double alpha = (x0 >= 0.0) ? -sqrt(nrm2) : sqrt(nrm2);
        if (tid == 0) x -= alpha;
        const double vn2 = nrm2 - x0 * x0 + (x0 - alpha) * (x0 - alpha);
        const double tk = 2.0 / vn2;
        if (tid < n) hv[tid] = x;
        __syncthreads();
        // p = tau A v over the trailing block (symmetric: read the full stored square)
        double p = 0.0;
        if (tid < n) {
          const double* arow = Ms + (kk + 1 + tid) * TMC_SMALL_LD + (kk + 1);
          for (int b = 0; b < n; ++b) p += arow[b] * hv[b];
          p *= tk;
        }
        const double pv = small_sum64(tid < n ? p * x : 0.0, sm2);
        const double Kc = 0.5 * tk * pv;
        if (tid < n) hp[tid] = p - Kc * x;          // w
        __syncthreads();
        for (int e = tid; e < n * n; e += 256) {
          const int a = e / n, b = e % n;
          Ms[(kk + 1 + a) * TMC_SMALL_LD + (kk + 1 + b)] -= hv[a] * hp[b] + hp[a] * hv[b];
        }
        if (tid < n) Ms[(kk + 1 + tid) * TMC_SMALL_LD + kk] = hv[tid];       // the reflector, for the back-transformation
        if (tid == 0) { ta[kk] = Ms[kk * TMC_SMALL_LD + kk]; tb[kk + 1] = alpha; tau[kk] = tk; }
        __syncthreads();
      }
      if (tid == 0) {
        if (t >= 2) { ta[t - 2] = Ms[(t - 2) * TMC_SMALL_LD + (t - 2)]; tb[t - 1] = Ms[(t - 1) * TMC_SMALL_LD + (t - 2)]; }
        ta[t - 1] = Ms[(t - 1) * TMC_SMALL_LD + (t - 1)];
        tb[0] = 0.0;
      }
      __syncthreads();
      // top eigenpair of T (warp 0), back-transformed through the reflectors: u2 = H_0 ... H_{t-3} z
      if (warp == 0) {
        const double theta = tmc_top_eig_warp(ta, tb, t, -1e300);
        if (lane == 0) {
          double glo, ghi, piv;
          tmc_gershgorin(ta, tb, t, &glo, &ghi, &piv);
          tmc_twisted_vector(ta, tb, t, theta, tdp, tdm, tz, piv);
          reinterpret_cast<double*>(ctl + 4)[0] = theta;
        }
        __syncwarp();
        for (int kk = t - 3; kk >= 0; --kk) {
          const double tk = tau[kk];
          if (tk == 0.0) continue;
          const int n = t - kk - 1;
          double s = 0.0;
          for (int a = lane; a < n; a += 32) s += Ms[(kk + 1 + a) * TMC_SMALL_LD + kk] * tz[kk + 1 + a];
          s = warp_sum(s) * tk;
          for (int a = lane; a < n; a += 32) tz[kk + 1 + a] -= s * Ms[(kk + 1 + a) * TMC_SMALL_LD + kk];
          __syncwarp();
        }
      }
      __syncthreads();
      const double theta = reinterpret_cast<double*>(ctl + 4)[0];
      // labels (sign rule), canonical sign, Q
      int l = 0; double h = 0.0;
      if (tid < t) {
        l = (tz[tid] >= 0.0) ? 1 : 0;
        h = hash_unit(0x7A6D635F7369676Eull, (uint64_t)(c.row_offset + rows[idx[lo + tid]]));
      }
      const double sgn = small_sum64(tid < t ? (l ? h : -h) : 0.0, sm2);
      if (c.canon_sign && sgn < 0.0) l ^= 1;
      if (tid < t) lab[tid] = l;
      __syncthreads();
      double row1 = 0.0;                            // sum over the label-1 cells b of G_ab, for a label-1 cell a
      if (tid < t && l) {
        const double* g = Gs + idx[lo + tid] * TMC_SMALL_LD;
        for (int b = 0; b < t; ++b) if (lab[b]) row1 += g[idx[lo + b]];
      }
      const double t1 = small_sum64(tid < t ? row1 : 0.0, sm2);
      const double n1 = small_sum64(tid < t ? (double)l : 0.0, sm2);
      const double sd1 = small_sum64((tid < t && l) ? dv : 0.0, sm2);
      {
        const double m = (double)t, n0 = m - n1, L = sumd - m;
        double q = 0.0;
        if (L != 0.0) {
          const double O1 = t1 - n1, O0 = sumd - 2.0 * sd1 + t1 - n0;
          const double L1 = sd1 - n1, L0 = (sumd - sd1) - n0;
          q = (O0 / L - (L0 / L) * (L0 / L)) + (O1 / L - (L1 / L) * (L1 / L));
        }
        const int in0 = (int)(n0 + 0.5), in1 = (int)(n1 + 0.5);
        const bool split = (q > c.min_modularity) && in0 >= 1 && in1 >= 1 && in0 >= c.min_size && in1 >= c.min_size;
        rec.q = q; rec.theta = theta; rec.leaf = split ? 0 : 1;
        if (tid == 0) recs[nd.rec0 + rec_i] = rec;
        if (split) {
          // stable partition of idx[lo, hi): label 0 first, original order kept on both sides
          if (tid < t) {
            int before = 0;
            for (int b = 0; b < tid; ++b) before += lab[b];
            const int dest = l ? (in0 + before) : (tid - before);
            idx2[dest] = idx[lo + tid];
          }
          __syncthreads();
          if (tid < t) idx[lo + tid] = idx2[tid];
          if (tid == 0) {
            const int s2 = ctl[0];
            st_lo[s2] = lo + in0; st_hi[s2] = hi;            // right child (label 1), processed after the whole left subtree
            st_lo[s2 + 1] = lo; st_hi[s2 + 1] = lo + in0;
            ctl[0] = s2 + 2;
          }
        }
      }
      __syncthreads();
    }
    __syncthreads();
    tick(2);
    if (tid < k) c.perm[nd.begin + tid] = rows[idx[tid]];
  }
}

// dynamic shared memory of k_small_subtree
static inline size_t small_smem_bytes() {
  return sizeof(double) * (2 * (size_t)TMC_SMALL_MAX * TMC_SMALL_LD + 11 * 64 + 4) + sizeof(int) * (4 * 64 + 2 * 72 + 8);
}

}  // namespace tmc
