// Candidate structures for the two dense-panel kernels, timed on a synthetic 1M x 5888 byte panel (tools/exp, not product code).
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
__device__ __forceinline__ double u8_f64(unsigned w, int k) {
  return __hiloint2double(0x43300000, (int)__byte_perm(w, 0u, 0x4440u + (unsigned)k)) - 4503599627370496.0;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double2 ldg_d2(const double* p) { return __ldg(reinterpret_cast<const double2*>(p)); }
// ---------------- gather: out[r] = sum_j panel[r][j] * y[j]; CTA = 256 rows; W=16 bytes per lane, NR rows per warp, prefetch
#define CB 4096
template <int NR>
__global__ void __launch_bounds__(256) k_gather16(const uint8_t* __restrict__ panel, const int* __restrict__ perm, const double* __restrict__ y,
                                                  int n_rows, int ppad, double* __restrict__ out) {
  __shared__ double2 ys2[CB / 2];
  __shared__ double ps[256];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int begin = blockIdx.x * 256, len = min(256, n_rows - begin);
  for (int p = threadIdx.x; p < len; p += 256) ps[p] = 0.0;
  for (int cb = 0; cb < ppad; cb += CB) {
    const int nb = min(CB, ppad - cb);
    __syncthreads();
    for (int t = threadIdx.x; t < (nb >> 1); t += 256) {
      const int w = t & 255;
      ys2[(t & ~255) + ((w & 7) << 5) + (w >> 3)] = ldg_d2(y + cb + 2 * t);
    }
    __syncthreads();
    const int nit = nb >> 9;
    for (int r0 = warp * NR; r0 < len; r0 += 8 * NR) {
      const int nr = min(NR, len - r0);
      const uint4* pr[NR];
#pragma unroll
      for (int q = 0; q < NR; ++q) pr[q] = reinterpret_cast<const uint4*>(panel + (size_t)perm[begin + r0 + min(q, nr - 1)] * ppad + cb) + lane;
      double acc[NR];
#pragma unroll
      for (int q = 0; q < NR; ++q) acc[q] = 0.0;
      uint4 d[NR], dn[NR];
#pragma unroll
      for (int q = 0; q < NR; ++q) d[q] = __ldg(pr[q]);
      for (int i = 0; i < nit; ++i) {
        if (i + 1 < nit) {
#pragma unroll
          for (int q = 0; q < NR; ++q) dn[q] = __ldg(pr[q] + 32 * (i + 1));
        }
#pragma unroll
        for (int cpt = 0; cpt < 4; ++cpt) {
          const double2 ya = ys2[i * 256 + (2 * cpt) * 32 + lane], yb = ys2[i * 256 + (2 * cpt + 1) * 32 + lane];
#pragma unroll
          for (int q = 0; q < NR; ++q) {
            const unsigned w = cpt == 0 ? d[q].x : (cpt == 1 ? d[q].y : (cpt == 2 ? d[q].z : d[q].w));
            acc[q] += u8_f64(w, 0) * ya.x + u8_f64(w, 1) * ya.y + u8_f64(w, 2) * yb.x + u8_f64(w, 3) * yb.y;
          }
        }
#pragma unroll
        for (int q = 0; q < NR; ++q) d[q] = dn[q];
      }
#pragma unroll
      for (int q = 0; q < NR; ++q) {
        const double sq = warp_sum(acc[q]);
        if (lane == 0 && q < nr) ps[r0 + q] += sq;
      }
    }
  }
  __syncthreads();
  for (int p = threadIdx.x; p < len; p += 256) out[begin + p] = ps[p];
}
// ---------------- transposed: yout[j] += sum_r panel[r][j] * x[r]; CTA = RB rows x (blockDim/32 * 512) columns
template <int NR>
__global__ void k_tp16(const uint8_t* __restrict__ panel, const int* __restrict__ perm, const double* __restrict__ x, int n_rows, int ppad,
                       int rb, double* __restrict__ yout) {
  __shared__ int s_row[256];
  __shared__ double s_x[256];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int colw = blockIdx.y * (blockDim.x >> 5) * 512 + warp * 512;
  const bool on = colw < ppad;
  const uint4* __restrict__ pb = reinterpret_cast<const uint4*>(panel + colw) + lane;
  const size_t ws = (size_t)ppad >> 4;
  double acc[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) acc[k] = 0.0;
  const int r_begin = blockIdx.x * rb, r_end = min(n_rows, r_begin + rb);
  for (int c0 = r_begin; c0 < r_end; c0 += 256) {
    const int len = min(256, r_end - c0);
    __syncthreads();
    if ((int)threadIdx.x < len) { s_row[threadIdx.x] = perm[c0 + threadIdx.x]; s_x[threadIdx.x] = x[c0 + threadIdx.x]; }
    if (blockDim.x < 256) for (int t = threadIdx.x + blockDim.x; t < len; t += blockDim.x) { s_row[t] = perm[c0 + t]; s_x[t] = x[c0 + t]; }
    __syncthreads();
    if (!on) continue;
    uint4 d[NR], dn[NR];
#pragma unroll
    for (int q = 0; q < NR; ++q) d[q] = __ldg(pb + (size_t)s_row[min(q, len - 1)] * ws);
    for (int r = 0; r < len; r += NR) {
      if (r + NR < len) {
#pragma unroll
        for (int q = 0; q < NR; ++q) dn[q] = __ldg(pb + (size_t)s_row[min(r + NR + q, len - 1)] * ws);
      }
#pragma unroll
      for (int q = 0; q < NR; ++q) {
        const double xv = r + q < len ? s_x[r + q] : 0.0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          acc[k] += u8_f64(d[q].x, k) * xv; acc[4 + k] += u8_f64(d[q].y, k) * xv;
          acc[8 + k] += u8_f64(d[q].z, k) * xv; acc[12 + k] += u8_f64(d[q].w, k) * xv;
        }
      }
#pragma unroll
      for (int q = 0; q < NR; ++q) d[q] = dn[q];
    }
  }
  if (on) {
#pragma unroll
    for (int k = 0; k < 16; ++k) if (acc[k] != 0.0) atomicAdd(yout + colw + lane * 16 + k, acc[k]);
  }
}
int main() {
  const int n_rows = 1000000, ppad = 5888 / 512 * 512 + 512;   // 6144 (multiple of 512)
  const size_t bytes = (size_t)n_rows * ppad;
  uint8_t* p; int* perm; double *y, *x, *out, *yout;
  CK(cudaMalloc(&p, bytes)); CK(cudaMalloc(&perm, 4 * n_rows)); CK(cudaMalloc(&y, 8 * ppad)); CK(cudaMalloc(&x, 8 * n_rows));
  CK(cudaMalloc(&out, 8 * n_rows)); CK(cudaMalloc(&yout, 8 * ppad));
  {
    std::vector<uint8_t> h(1 << 20); for (size_t i = 0; i < h.size(); ++i) h[i] = ((i * 2654435761u) >> 13) % 4 == 0 ? (uint8_t)(1 + (i % 3)) : 0;
    for (size_t o = 0; o < bytes; o += h.size()) CK(cudaMemcpy(p + o, h.data(), std::min(h.size(), bytes - o), cudaMemcpyHostToDevice));
    std::vector<int> hp(n_rows); for (int i = 0; i < n_rows; ++i) hp[i] = i;
    CK(cudaMemcpy(perm, hp.data(), 4 * n_rows, cudaMemcpyHostToDevice));
    std::vector<double> hy(ppad), hx(n_rows); for (int i = 0; i < ppad; ++i) hy[i] = 1.0 + 1e-3 * (i % 7); for (int i = 0; i < n_rows; ++i) hx[i] = 1.0 + 1e-3 * (i % 5);
    CK(cudaMemcpy(y, hy.data(), 8 * ppad, cudaMemcpyHostToDevice)); CK(cudaMemcpy(x, hx.data(), 8 * n_rows, cudaMemcpyHostToDevice));
  }
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  auto time = [&](const char* name, auto launch) {
    launch(); cudaDeviceSynchronize();
    cudaEventRecord(a); for (int i = 0; i < 5; ++i) launch(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); ms /= 5;
    double h[2]; cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost); double hy2[2]; cudaMemcpy(hy2, yout, 16, cudaMemcpyDeviceToHost);
    printf("%-28s %7.3f ms  %6.0f GB/s  out0 %.6f y0 %.3f (%s)\n", name, ms, bytes / ms / 1e6, h[0], hy2[0] / 6, cudaGetErrorString(cudaGetLastError()));
  };
  const int nch = (n_rows + 255) / 256;
  time("gather16 NR=2", [&] { k_gather16<2><<<nch, 256>>>(p, perm, y, n_rows, ppad, out); });
  time("gather16 NR=4", [&] { k_gather16<4><<<nch, 256>>>(p, perm, y, n_rows, ppad, out); });
  for (int rb : {512, 1024, 2048}) {
    printf("rb %d\n", rb);
    const int nb = (n_rows + rb - 1) / rb;
    cudaMemset(yout, 0, 8 * ppad);
    time("tp16 NR=2 8 warps", [&] { k_tp16<2><<<dim3(nb, (ppad + 4095) / 4096), 256>>>(p, perm, x, n_rows, ppad, rb, yout); });
    cudaMemset(yout, 0, 8 * ppad);
    time("tp16 NR=4 8 warps", [&] { k_tp16<4><<<dim3(nb, (ppad + 4095) / 4096), 256>>>(p, perm, x, n_rows, ppad, rb, yout); });
    cudaMemset(yout, 0, 8 * ppad);
    time("tp16 NR=2 6 warps (3072)", [&] { k_tp16<2><<<dim3(nb, (ppad + 3071) / 3072), 192>>>(p, perm, x, n_rows, ppad, rb, yout); });
    cudaMemset(yout, 0, 8 * ppad);
    time("tp16 NR=4 6 warps (3072)", [&] { k_tp16<4><<<dim3(nb, (ppad + 3071) / 3072), 192>>>(p, perm, x, n_rows, ppad, rb, yout); });
    cudaMemset(yout, 0, 8 * ppad);
    time("tp16 NR=2 12 warps (6144)", [&] { k_tp16<2><<<dim3(nb, 1), 384>>>(p, perm, x, n_rows, ppad, rb, yout); });
    cudaMemset(yout, 0, 8 * ppad);
    time("tp16 NR=4 12 warps (6144)", [&] { k_tp16<4><<<dim3(nb, 1), 384>>>(p, perm, x, n_rows, ppad, rb, yout); });
  }
  return 0;
}
