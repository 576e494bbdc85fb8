// Experimental SpMV-pair kernel variants (NOT product code): used to pick the design by measurement.
#include <cuda_runtime.h>
#include <stdint.h>
#define DI __device__ __forceinline__
DI double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
DI int ld_stream_i32(const int* p) { int v; asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p)); return v; }
DI double ld_stream_f64(const double* p) { double v; asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p)); return v; }

// ---------------- gather variants: w[p] = s[p] * sum_k val[k] * y[col[k]] over row perm[p]
template <int MODE>
__global__ void __launch_bounds__(256) g_base(const int64_t* __restrict__ rp, const int* __restrict__ col, const double* __restrict__ val,
                                              const int* __restrict__ perm, const double* __restrict__ y, double* __restrict__ w, int n_rows, int rows_per_cta) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r0 = blockIdx.x * rows_per_cta;
  for (int r = warp; r < rows_per_cta; r += 8) {
    int p = r0 + r; if (p >= n_rows) break;
    int row = perm[p];
    int64_t rs = rp[row], re = rp[row + 1];
    double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    int64_t k = rs + lane;
    for (; k + 96 < re; k += 128) {
      int c0, c1, c2, c3; double v0, v1, v2, v3;
      if (MODE == 1) { c0 = ld_stream_i32(col + k); c1 = ld_stream_i32(col + k + 32); c2 = ld_stream_i32(col + k + 64); c3 = ld_stream_i32(col + k + 96);
        v0 = ld_stream_f64(val + k); v1 = ld_stream_f64(val + k + 32); v2 = ld_stream_f64(val + k + 64); v3 = ld_stream_f64(val + k + 96); }
      else { c0 = __ldg(col + k); c1 = __ldg(col + k + 32); c2 = __ldg(col + k + 64); c3 = __ldg(col + k + 96);
        v0 = __ldg(val + k); v1 = __ldg(val + k + 32); v2 = __ldg(val + k + 64); v3 = __ldg(val + k + 96); }
      if (MODE == 2) { a0 += v0 * c0; a1 += v1 * c1; a2 += v2 * c2; a3 += v3 * c3; }     // stream only, no y gather
      else { a0 += v0 * y[c0]; a1 += v1 * y[c1]; a2 += v2 * y[c2]; a3 += v3 * y[c3]; }
    }
    for (; k < re; k += 32) a0 += __ldg(val + k) * (MODE == 2 ? (double)__ldg(col + k) : y[__ldg(col + k)]);
    double s = warp_sum((a0 + a1) + (a2 + a3));
    if (lane == 0) w[p] = s;
  }
}

// pipelined: U-way unroll, register double-buffering, next-row pointer prefetch, streaming loads bypass L1
template <int U, int NOALLOC>
__global__ void __launch_bounds__(256) g_pipe(const int64_t* __restrict__ rp, const int* __restrict__ col, const double* __restrict__ val,
                                              const int* __restrict__ perm, const double* __restrict__ y, double* __restrict__ w, int n_rows, int rows_per_cta) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r0 = blockIdx.x * rows_per_cta;
  int p = r0 + warp;
  if (warp >= rows_per_cta || p >= n_rows) return;
  int row = perm[p];
  int64_t rs = rp[row], re = rp[row + 1];
  for (;;) {
    // prefetch next row's extent
    int pn = p + 8; bool has_next = (pn < r0 + rows_per_cta) && (pn < n_rows);
    int64_t nrs = 0, nre = 0;
    if (has_next) { int nrow = perm[pn]; nrs = rp[nrow]; nre = rp[nrow + 1]; }
    double acc[U];
#pragma unroll
    for (int u = 0; u < U; ++u) acc[u] = 0.0;
    int64_t k = rs + lane;
    int c[U]; double v[U];
    bool have = (k + 32 * (U - 1) < re);
    if (have) {
#pragma unroll
      for (int u = 0; u < U; ++u) { c[u] = NOALLOC ? ld_stream_i32(col + k + 32 * u) : __ldg(col + k + 32 * u); v[u] = NOALLOC ? ld_stream_f64(val + k + 32 * u) : __ldg(val + k + 32 * u); }
    }
    while (have) {
      int64_t kn = k + 32 * U;
      bool have_n = (kn + 32 * (U - 1) < re);
      int cn[U]; double vn[U];
      if (have_n) {
#pragma unroll
        for (int u = 0; u < U; ++u) { cn[u] = NOALLOC ? ld_stream_i32(col + kn + 32 * u) : __ldg(col + kn + 32 * u); vn[u] = NOALLOC ? ld_stream_f64(val + kn + 32 * u) : __ldg(val + kn + 32 * u); }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) acc[u] += v[u] * y[c[u]];
      if (have_n) {
#pragma unroll
        for (int u = 0; u < U; ++u) { c[u] = cn[u]; v[u] = vn[u]; }
      }
      k = kn; have = have_n;
    }
    for (; k < re; k += 32) acc[0] += __ldg(val + k) * y[__ldg(col + k)];
    double s = 0;
#pragma unroll
    for (int u = 0; u < U; ++u) s += acc[u];
    s = warp_sum(s);
    if (lane == 0) w[p] = s;
    if (!has_next) break;
    p = pn; rs = nrs; re = nre;
  }
}

// ---------------- scatter variants: y[col[k]] += val[k] * x[p]
template <int MODE>
__global__ void __launch_bounds__(256) s_base(const int64_t* __restrict__ rp, const int* __restrict__ col, const double* __restrict__ val,
                                              const int* __restrict__ perm, const double* __restrict__ x, double* __restrict__ y, int n_rows, int rows_per_cta,
                                              int n_cols, int replicas, int64_t ystride) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r0 = blockIdx.x * rows_per_cta;
  double* yy = y + (int64_t)(blockIdx.x % replicas) * ystride;
  for (int r = warp; r < rows_per_cta; r += 8) {
    int p = r0 + r; if (p >= n_rows) break;
    int row = perm[p];
    double xi = x[p];
    int64_t rs = rp[row], re = rp[row + 1];
    int64_t k = rs + lane;
    for (; k + 96 < re; k += 128) {
      int c0 = __ldg(col + k), c1 = __ldg(col + k + 32), c2 = __ldg(col + k + 64), c3 = __ldg(col + k + 96);
      double v0 = __ldg(val + k), v1 = __ldg(val + k + 32), v2 = __ldg(val + k + 64), v3 = __ldg(val + k + 96);
      if (MODE == 1) {   // uniform pseudo-random columns: no hot addresses
        c0 = (int)(((unsigned)(k * 2654435761u)) % (unsigned)n_cols); c1 = (int)(((unsigned)((k + 32) * 2654435761u)) % (unsigned)n_cols);
        c2 = (int)(((unsigned)((k + 64) * 2654435761u)) % (unsigned)n_cols); c3 = (int)(((unsigned)((k + 96) * 2654435761u)) % (unsigned)n_cols);
      }
      atomicAdd(yy + c0, v0 * xi); atomicAdd(yy + c1, v1 * xi); atomicAdd(yy + c2, v2 * xi); atomicAdd(yy + c3, v3 * xi);
    }
    for (; k < re; k += 32) atomicAdd(yy + __ldg(col + k), __ldg(val + k) * xi);
  }
}

// transposed product from a column-sorted COO copy (ccol sorted, cpos = row position, cval): each lane owns E
// consecutive entries, merges runs locally, warp-merges when the whole warp is one run; atomics only per run.
template <int E>
__global__ void __launch_bounds__(256) t_coo(const int* __restrict__ ccol, const int* __restrict__ cpos, const double* __restrict__ cval,
                                             const double* __restrict__ x, double* __restrict__ y, int64_t nnz) {
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t base = warp_global * (32 * E) + (int64_t)lane * E;
  if (warp_global * (32 * E) >= nnz) return;
  int c[E]; double t[E];
#pragma unroll
  for (int e = 0; e < E; ++e) {
    int64_t k = base + e;
    if (k < nnz) { c[e] = ld_stream_i32(ccol + k); int p = ld_stream_i32(cpos + k); t[e] = ld_stream_f64(cval + k) * x[p]; }
    else { c[e] = -1; t[e] = 0.0; }
  }
  // whole-warp single run?
  int cfirst = __shfl_sync(0xffffffffu, c[0], 0);
  int clast = __shfl_sync(0xffffffffu, c[E - 1], 31);
  if (cfirst == clast && cfirst >= 0) {
    double s = 0;
#pragma unroll
    for (int e = 0; e < E; ++e) s += t[e];
    s = warp_sum(s);
    if (lane == 0) atomicAdd(y + cfirst, s);
    return;
  }
  double acc = t[0]; int cur = c[0];
#pragma unroll
  for (int e = 1; e < E; ++e) {
    if (c[e] == cur) acc += t[e];
    else { if (cur >= 0) atomicAdd(y + cur, acc); cur = c[e]; acc = t[e]; }
  }
  if (cur >= 0) atomicAdd(y + cur, acc);
}

extern "C" {
#define LAUNCH_G(KERN) KERN<<<(n_rows + rpc - 1) / rpc, 256>>>(rp, col, val, perm, y, w, n_rows, rpc)
void exp_gather(int variant, const int64_t* rp, const int* col, const double* val, const int* perm, const double* y, double* w, int n_rows, int rpc) {
  switch (variant) {
    case 0: LAUNCH_G((g_base<0>)); break;
    case 1: LAUNCH_G((g_base<1>)); break;
    case 2: LAUNCH_G((g_base<2>)); break;
    case 3: LAUNCH_G((g_pipe<4, 0>)); break;
    case 4: LAUNCH_G((g_pipe<4, 1>)); break;
    case 5: LAUNCH_G((g_pipe<8, 1>)); break;
    case 6: LAUNCH_G((g_pipe<2, 1>)); break;
    case 7: LAUNCH_G((g_pipe<8, 0>)); break;
  }
}
void exp_scatter(int variant, const int64_t* rp, const int* col, const double* val, const int* perm, const double* x, double* y, int n_rows, int rpc,
                 int n_cols, int replicas, int64_t ystride) {
  int grid = (n_rows + rpc - 1) / rpc;
  if (variant == 0) s_base<0><<<grid, 256>>>(rp, col, val, perm, x, y, n_rows, rpc, n_cols, replicas, ystride);
  else s_base<1><<<grid, 256>>>(rp, col, val, perm, x, y, n_rows, rpc, n_cols, replicas, ystride);
}
void exp_tcoo(int E, const int* ccol, const int* cpos, const double* cval, const double* x, double* y, int64_t nnz) {
  if (E == 8) { int64_t warps = (nnz + 255) / 256; t_coo<8><<<(unsigned)((warps + 7) / 8), 256>>>(ccol, cpos, cval, x, y, nnz); }
  else if (E == 4) { int64_t warps = (nnz + 127) / 128; t_coo<4><<<(unsigned)((warps + 7) / 8), 256>>>(ccol, cpos, cval, x, y, nnz); }
  else { int64_t warps = (nnz + 511) / 512; t_coo<16><<<(unsigned)((warps + 7) / 8), 256>>>(ccol, cpos, cval, x, y, nnz); }
}
}

// vectorised version of the transposed gather (what the product's k_tgather does), E = 8 consecutive entries per lane
__device__ __forceinline__ int4 ldg_i4(const int* p) { return __ldg(reinterpret_cast<const int4*>(p)); }
__device__ __forceinline__ double2 ldg_d2(const double* p) { return __ldg(reinterpret_cast<const double2*>(p)); }
template <int MODE>   // 0: lane-local merge; 1: + no atomics at all (upper bound); 2: no x gather
__global__ void __launch_bounds__(256) t_coo_v8(const int* __restrict__ tcol, const int* __restrict__ tpos, const double* __restrict__ tval,
                                                const double* __restrict__ xin, double* __restrict__ y, long long nnz) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long cbeg = (long long)blockIdx.x * 8192, cend = min(cbeg + 8192, nnz);
  double sink = 0;
  for (long long base = cbeg + warp * 256; base < cend; base += 2048) {
    const long long k0 = base + lane * 8;
    int cc[8]; double t[8];
    if (k0 + 8 <= nnz) {
      const int4 ca = ldg_i4(tcol + k0), cb = ldg_i4(tcol + k0 + 4);
      const double2 v0 = ldg_d2(tval + k0), v1 = ldg_d2(tval + k0 + 2), v2 = ldg_d2(tval + k0 + 4), v3 = ldg_d2(tval + k0 + 6);
      const int4 pa = ldg_i4(tpos + k0), pb = ldg_i4(tpos + k0 + 4);
      cc[0] = ca.x; cc[1] = ca.y; cc[2] = ca.z; cc[3] = ca.w; cc[4] = cb.x; cc[5] = cb.y; cc[6] = cb.z; cc[7] = cb.w;
      const int pp[8] = {pa.x, pa.y, pa.z, pa.w, pb.x, pb.y, pb.z, pb.w};
      t[0] = v0.x; t[1] = v0.y; t[2] = v1.x; t[3] = v1.y; t[4] = v2.x; t[5] = v2.y; t[6] = v3.x; t[7] = v3.y;
#pragma unroll
      for (int e = 0; e < 8; ++e) t[e] *= (MODE == 2 ? (double)pp[e] : __ldg(xin + pp[e]));
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) { cc[e] = -1; t[e] = 0.0; }
    }
    if (MODE == 1) { for (int e = 0; e < 8; ++e) sink += t[e] * cc[e]; continue; }
    const int cfirst = __shfl_sync(0xffffffffu, cc[0], 0), clast = __shfl_sync(0xffffffffu, cc[7], 31);
    if (cfirst == clast && cfirst >= 0) {
      double a = ((t[0] + t[1]) + (t[2] + t[3])) + ((t[4] + t[5]) + (t[6] + t[7]));
      a = warp_sum(a);
      if (lane == 0) atomicAdd(y + cfirst, a);
      continue;
    }
    double acc = t[0]; int cur = cc[0];
#pragma unroll
    for (int e = 1; e < 8; ++e) {
      if (cc[e] == cur) acc += t[e];
      else { if (cur >= 0) atomicAdd(y + cur, acc); cur = cc[e]; acc = t[e]; }
    }
    if (cur >= 0) atomicAdd(y + cur, acc);
  }
  if (MODE == 1 && sink == 1.2345e-300) y[0] = sink;
}
extern "C" void exp_tcoo_v8(int mode, const int* ccol, const int* cpos, const double* cval, const double* x, double* y, long long nnz) {
  unsigned grid = (unsigned)((nnz + 8191) / 8192);
  if (mode == 0) t_coo_v8<0><<<grid, 256>>>(ccol, cpos, cval, x, y, nnz);
  else if (mode == 1) t_coo_v8<1><<<grid, 256>>>(ccol, cpos, cval, x, y, nnz);
  else t_coo_v8<2><<<grid, 256>>>(ccol, cpos, cval, x, y, nnz);
}

// gather with the hot part of y (columns renumbered by descending frequency, ids < H) staged in shared memory
template <int NT>
__global__ void __launch_bounds__(NT) g_smem(const int64_t* __restrict__ rp, const int* __restrict__ col, const unsigned short* __restrict__ val,
                                             const int* __restrict__ perm, const double* __restrict__ y, double* __restrict__ w, int n_rows,
                                             int rows_per_cta, int H) {
  extern __shared__ double ys[];
  for (int i = threadIdx.x; i < H; i += NT) ys[i] = y[i];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = NT / 32;
  const int r0 = blockIdx.x * rows_per_cta;
  for (int r = warp; r < rows_per_cta; r += nw) {
    int p = r0 + r; if (p >= n_rows) break;
    int row = perm[p];
    int64_t rs = rp[row], re = rp[row + 1];
    double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    int64_t k = rs + lane;
    for (; k + 96 < re; k += 128) {
      int c0 = __ldg(col + k), c1 = __ldg(col + k + 32), c2 = __ldg(col + k + 64), c3 = __ldg(col + k + 96);
      double v0 = __ldg(val + k), v1 = __ldg(val + k + 32), v2 = __ldg(val + k + 64), v3 = __ldg(val + k + 96);
      a0 += v0 * (c0 < H ? ys[c0] : __ldg(y + c0)); a1 += v1 * (c1 < H ? ys[c1] : __ldg(y + c1));
      a2 += v2 * (c2 < H ? ys[c2] : __ldg(y + c2)); a3 += v3 * (c3 < H ? ys[c3] : __ldg(y + c3));
    }
    for (; k < re; k += 32) { int c0 = __ldg(col + k); a0 += (double)__ldg(val + k) * (c0 < H ? ys[c0] : __ldg(y + c0)); }
    double s = warp_sum((a0 + a1) + (a2 + a3));
    if (lane == 0) w[p] = s;
  }
}
// baseline with u16 values (same structure, no smem)
__global__ void __launch_bounds__(256) g_u16(const int64_t* __restrict__ rp, const int* __restrict__ col, const unsigned short* __restrict__ val,
                                            const int* __restrict__ perm, const double* __restrict__ y, double* __restrict__ w, int n_rows, int rows_per_cta) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r0 = blockIdx.x * rows_per_cta;
  for (int r = warp; r < rows_per_cta; r += 8) {
    int p = r0 + r; if (p >= n_rows) break;
    int row = perm[p];
    int64_t rs = rp[row], re = rp[row + 1];
    double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    int64_t k = rs + lane;
    for (; k + 96 < re; k += 128) {
      int c0 = __ldg(col + k), c1 = __ldg(col + k + 32), c2 = __ldg(col + k + 64), c3 = __ldg(col + k + 96);
      double v0 = __ldg(val + k), v1 = __ldg(val + k + 32), v2 = __ldg(val + k + 64), v3 = __ldg(val + k + 96);
      a0 += v0 * __ldg(y + c0); a1 += v1 * __ldg(y + c1); a2 += v2 * __ldg(y + c2); a3 += v3 * __ldg(y + c3);
    }
    for (; k < re; k += 32) a0 += (double)__ldg(val + k) * __ldg(y + __ldg(col + k));
    double s = warp_sum((a0 + a1) + (a2 + a3));
    if (lane == 0) w[p] = s;
  }
}
extern "C" void exp_gather_smem(int nt, const int64_t* rp, const int* col, const unsigned short* val, const int* perm, const double* y, double* w,
                                int n_rows, int rpc, int H) {
  size_t sm = (size_t)H * 8;
  int grid = (n_rows + rpc - 1) / rpc;
  if (nt == 1024) { cudaFuncSetAttribute(g_smem<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm); g_smem<1024><<<grid, 1024, sm>>>(rp, col, val, perm, y, w, n_rows, rpc, H); }
  else if (nt == 512) { cudaFuncSetAttribute(g_smem<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm); g_smem<512><<<grid, 512, sm>>>(rp, col, val, perm, y, w, n_rows, rpc, H); }
  else { cudaFuncSetAttribute(g_smem<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm); g_smem<256><<<grid, 256, sm>>>(rp, col, val, perm, y, w, n_rows, rpc, H); }
}
extern "C" void exp_gather_u16(const int64_t* rp, const int* col, const unsigned short* val, const int* perm, const double* y, double* w, int n_rows, int rpc) {
  g_u16<<<(n_rows + rpc - 1) / rpc, 256>>>(rp, col, val, perm, y, w, n_rows, rpc);
}

// EXPERIMENT: transposed gather with the entries ordered by (position block, column, position) and the x window of the
// block staged in shared memory (PB positions).  One CTA = one contiguous share of one position block.
template <int PB>
__global__ void __launch_bounds__(256) t_posblk(const int* __restrict__ tcol, const int* __restrict__ tpos, const unsigned short* __restrict__ tval,
                                                const long long* __restrict__ blk_ptr, const double* __restrict__ x, double* __restrict__ y,
                                                int n_rows, int shares) {
  extern __shared__ double xs[];
  const int pb = blockIdx.x / shares, sh = blockIdx.x % shares;
  const long long b0 = blk_ptr[pb], b1 = blk_ptr[pb + 1];
  long long per = ((b1 - b0 + shares - 1) / shares + 127) & ~127ll;
  const long long e0 = b0 + sh * per, e1 = min(b1, e0 + per);
  if (e0 >= e1) return;
  const int p0 = pb * PB;
  for (int i = threadIdx.x; i < PB; i += 256) xs[i] = (p0 + i < n_rows) ? x[p0 + i] : 0.0;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (long long base = e0 + warp * 128; base < e1; base += 8 * 128) {
    const long long k0 = base + lane * 4;
    int cc[4]; double t[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const long long k = k0 + e;
      if (k < e1) { cc[e] = __ldg(tcol + k); t[e] = (double)__ldg(tval + k) * xs[__ldg(tpos + k) - p0]; }
      else { cc[e] = -1; t[e] = 0.0; }
    }
    const int cfirst = __shfl_sync(0xffffffffu, cc[0], 0), clast = __shfl_sync(0xffffffffu, cc[3], 31);
    if (cfirst == clast && cfirst >= 0) {
      double a = warp_sum((t[0] + t[1]) + (t[2] + t[3]));
      if (lane == 0) atomicAdd(y + cfirst, a);
      continue;
    }
    double acc = t[0]; int cur = cc[0];
#pragma unroll
    for (int e = 1; e < 4; ++e) {
      if (cc[e] == cur) acc += t[e];
      else { if (cur >= 0) atomicAdd(y + cur, acc); cur = cc[e]; acc = t[e]; }
    }
    if (cur >= 0) atomicAdd(y + cur, acc);
  }
}
extern "C" void exp_tposblk(int pbsize, const int* tcol, const int* tpos, const unsigned short* tval, const long long* blk_ptr, int n_blocks,
                            const double* x, double* y, int n_rows, int shares) {
  if (pbsize == 8192) { cudaFuncSetAttribute(t_posblk<8192>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 8); t_posblk<8192><<<n_blocks * shares, 256, 8192 * 8>>>(tcol, tpos, tval, blk_ptr, x, y, n_rows, shares); }
  else if (pbsize == 4096) { t_posblk<4096><<<n_blocks * shares, 256, 4096 * 8>>>(tcol, tpos, tval, blk_ptr, x, y, n_rows, shares); }
  else { cudaFuncSetAttribute(t_posblk<16384>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * 8); t_posblk<16384><<<n_blocks * shares, 256, 16384 * 8>>>(tcol, tpos, tval, blk_ptr, x, y, n_rows, shares); }
}
