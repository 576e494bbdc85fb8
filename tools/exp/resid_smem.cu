// Microbenchmark: residual halves with the 22k-column dense vector held in shared memory (one CTA per SM).  tools/exp only.
#include <cstdio>
#include <cstdint>
#include <vector>
#include <algorithm>
#include <random>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
// T half: y[col] += val * x[row], y in shared memory (fp64 shared atomics), flushed with global atomics at the end
__global__ void __launch_bounds__(1024) k_t_smem(const int64_t* __restrict__ rp, const int32_t* __restrict__ col, const unsigned short* __restrict__ val,
                                                 const double* __restrict__ x, int n_rows, int ncol, double* __restrict__ y) {
  extern __shared__ double ys[];
  for (int i = threadIdx.x; i < ncol; i += blockDim.x) ys[i] = 0.0;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const int per = (n_rows + gridDim.x - 1) / gridDim.x;
  const int r0 = blockIdx.x * per, r1 = min(n_rows, r0 + per);
  for (int r = r0 + warp; r < r1; r += nw) {
    const int64_t s = rp[r], e = rp[r + 1];
    const double xv = x[r];
    for (int64_t k = s + lane; k < e; k += 32) atomicAdd(&ys[__ldg(col + k)], (double)__ldg(val + k) * xv);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < ncol; i += blockDim.x) if (ys[i] != 0.0) atomicAdd(y + i, ys[i]);
}
// gather half: u[row] = sum val * y[col], y staged in shared memory
__global__ void __launch_bounds__(1024) k_g_smem(const int64_t* __restrict__ rp, const int32_t* __restrict__ col, const unsigned short* __restrict__ val,
                                                 const double* __restrict__ y, int n_rows, int ncol, double* __restrict__ u) {
  extern __shared__ double ys[];
  for (int i = threadIdx.x; i < ncol; i += blockDim.x) ys[i] = y[i];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const int per = (n_rows + gridDim.x - 1) / gridDim.x;
  const int r0 = blockIdx.x * per, r1 = min(n_rows, r0 + per);
  for (int r = r0 + warp; r < r1; r += nw) {
    const int64_t s = rp[r], e = rp[r + 1];
    double a0 = 0, a1 = 0;
    int64_t k = s + lane;
    for (; k + 32 < e; k += 64) {
      const int c0 = __ldg(col + k), c1 = __ldg(col + k + 32);
      const double v0 = (double)__ldg(val + k), v1 = (double)__ldg(val + k + 32);
      a0 += v0 * ys[c0]; a1 += v1 * ys[c1];
    }
    for (; k < e; k += 32) a0 += (double)__ldg(val + k) * ys[__ldg(col + k)];
    a0 += a1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a0 += __shfl_xor_sync(0xffffffffu, a0, o);
    if (lane == 0) u[r] = a0;
  }
}
int main() {
  const int n_rows = 1000000, ncol = 22231, per_row = 340;
  std::vector<int64_t> rp(n_rows + 1); std::vector<int32_t> col((size_t)n_rows * per_row); std::vector<unsigned short> val(col.size(), 1);
  std::mt19937 g(1);
  // skewed columns like the residual: density falls off; sorted within the row
  std::vector<int> tmp(per_row);
  for (int r = 0; r < n_rows; ++r) {
    rp[r] = (int64_t)r * per_row;
    for (int j = 0; j < per_row; ++j) { double u = (g() & 0xffffff) / 16777216.0; tmp[j] = (int)(ncol * u * u); }
    std::sort(tmp.begin(), tmp.end());
    for (int j = 0; j < per_row; ++j) col[(size_t)r * per_row + j] = tmp[j];
  }
  rp[n_rows] = (int64_t)n_rows * per_row;
  int64_t* d_rp; int32_t* d_col; unsigned short* d_val; double *d_x, *d_y, *d_u;
  CK(cudaMalloc(&d_rp, 8 * rp.size())); CK(cudaMalloc(&d_col, 4 * col.size())); CK(cudaMalloc(&d_val, 2 * val.size()));
  CK(cudaMalloc(&d_x, 8 * n_rows)); CK(cudaMalloc(&d_y, 8 * ncol)); CK(cudaMalloc(&d_u, 8 * n_rows));
  CK(cudaMemcpy(d_rp, rp.data(), 8 * rp.size(), cudaMemcpyHostToDevice)); CK(cudaMemcpy(d_col, col.data(), 4 * col.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_val, val.data(), 2 * val.size(), cudaMemcpyHostToDevice));
  std::vector<double> hx(n_rows, 1.0), hy(ncol, 1.0);
  CK(cudaMemcpy(d_x, hx.data(), 8 * n_rows, cudaMemcpyHostToDevice)); CK(cudaMemcpy(d_y, hy.data(), 8 * ncol, cudaMemcpyHostToDevice));
  const int smem = ncol * 8;
  CK(cudaFuncSetAttribute(k_t_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  CK(cudaFuncSetAttribute(k_g_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  const double bytes = 6.0 * col.size();
  auto time = [&](const char* name, auto launch) {
    launch(); cudaDeviceSynchronize();
    cudaEventRecord(a); for (int i = 0; i < 5; ++i) launch(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); ms /= 5;
    double h[2]; cudaMemcpy(h, d_u, 16, cudaMemcpyDeviceToHost); double hy2[2]; cudaMemcpy(hy2, d_y, 16, cudaMemcpyDeviceToHost);
    printf("%-22s %7.3f ms  %6.0f GB/s (6 B/entry)  u0 %.1f y0 %.1f (%s)\n", name, ms, bytes / ms / 1e6, h[0], hy2[0], cudaGetErrorString(cudaGetLastError()));
  };
  for (int thr : {1024, 512}) {
    printf("threads %d\n", thr);
    time("T smem atomics", [&] { k_t_smem<<<148, thr, smem>>>(d_rp, d_col, d_val, d_x, n_rows, ncol, d_y); });
    time("gather smem y", [&] { k_g_smem<<<148, thr, smem>>>(d_rp, d_col, d_val, d_y, n_rows, ncol, d_u); });
  }
  return 0;
}
