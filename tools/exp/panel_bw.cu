// Microbenchmark: what bounds a dense u8 panel x fp64 vector product?  (tools/exp, not part of the product)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o panel_bw panel_bw.cu && ./panel_bw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
__device__ __forceinline__ double u8_f64(unsigned w, int k) {
  return __hiloint2double(0x43300000, (int)__byte_perm(w, 0u, 0x4440u + (unsigned)k)) - 4503599627370496.0;
}
__device__ __forceinline__ double u8_raw(unsigned w, int k) { return __hiloint2double(0x43300000, (int)__byte_perm(w, 0u, 0x4440u + (unsigned)k)); }
// MODE 0: stream only (xor); 1: magic convert + fma; 2: fma without the subtraction (1 fp64 op / byte); 3: cvt.f64.u32 + fma;
// 4: fp32 convert + fp32 fma (pipe comparison only)
template <int MODE>
__global__ void __launch_bounds__(256) k_stream(const uint4* __restrict__ p, size_t n16, double* out) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  double acc[4] = {0, 0, 0, 0}; unsigned x = 0; float facc = 0.f;
  const double y0 = 1.0 + threadIdx.x * 1e-3, y1 = 0.5, y2 = 0.25, y3 = 2.0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride * 2) {
    uint4 v[2];
    v[0] = __ldg(p + i);
    v[1] = (i + stride < n16) ? __ldg(p + i + stride) : make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const unsigned w[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (MODE == 0) x ^= w[q];
        if (MODE == 1) { acc[0] += u8_f64(w[q], 0) * y0; acc[1] += u8_f64(w[q], 1) * y1; acc[2] += u8_f64(w[q], 2) * y2; acc[3] += u8_f64(w[q], 3) * y3; }
        if (MODE == 2) { acc[0] += u8_raw(w[q], 0) * y0; acc[1] += u8_raw(w[q], 1) * y1; acc[2] += u8_raw(w[q], 2) * y2; acc[3] += u8_raw(w[q], 3) * y3; }
        if (MODE == 3) { acc[0] += (double)(w[q] & 255u) * y0; acc[1] += (double)((w[q] >> 8) & 255u) * y1; acc[2] += (double)((w[q] >> 16) & 255u) * y2; acc[3] += (double)(w[q] >> 24) * y3; }
        if (MODE == 4) { facc += (float)(w[q] & 255u) * 1.5f + (float)((w[q] >> 8) & 255u) * 0.5f + (float)((w[q] >> 16) & 255u) * 0.25f + (float)(w[q] >> 24) * 2.f; }
      }
    }
  }
  double s = acc[0] + acc[1] + acc[2] + acc[3] + (double)x + (double)facc;
  if (s == 123.456) out[0] = s;
}
// row-wise: warp per row, 5888-byte rows, 4 rows per warp at a time (like k_gather_panel without y traffic)
template <int MODE>
__global__ void __launch_bounds__(256) k_rows(const uint8_t* __restrict__ p, int n_rows, int ppad, double* out) {
  const int warp = (int)(((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  const int nw = (int)(((size_t)gridDim.x * blockDim.x) >> 5);
  double acc = 0; unsigned x = 0;
  const double y0 = 1.0 + lane * 1e-3;
  for (int r = warp * 4; r < n_rows; r += nw * 4) {
    for (int i = 0; i < (ppad >> 8); ++i) {
      uint2 d[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) d[q] = __ldg(reinterpret_cast<const uint2*>(p + (size_t)min(r + q, n_rows - 1) * ppad) + 32 * i + lane);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (MODE == 0) x ^= d[q].x ^ d[q].y;
        else {
#pragma unroll
          for (int k = 0; k < 4; ++k) acc += u8_f64(d[q].x, k) * y0 + u8_f64(d[q].y, k) * y0;
        }
      }
    }
  }
  if (acc + (double)x == 123.456) out[0] = acc;
}
// R variants: NR rows per warp at a time, lane loads W bytes (8 or 16) per row per iteration, NR independent accumulators
template <int NR, int W>
__global__ void __launch_bounds__(256) k_rows2(const uint8_t* __restrict__ p, int n_rows, int ppad, double* out) {
  const int warp = (int)(((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  const int nw = (int)(((size_t)gridDim.x * blockDim.x) >> 5);
  double acc[NR]; for (int q = 0; q < NR; ++q) acc[q] = 0;
  const double y0 = 1.0 + lane * 1e-3, y1 = 0.5 + lane * 1e-3;
  const int nit = ppad / (32 * W);
  for (int r = warp * NR; r < n_rows; r += nw * NR) {
    for (int i = 0; i < nit; ++i) {
      unsigned d[NR][W / 4];
#pragma unroll
      for (int q = 0; q < NR; ++q) {
        const uint8_t* src = p + (size_t)min(r + q, n_rows - 1) * ppad + (size_t)(32 * i + lane) * W;
        if (W == 16) { uint4 v = __ldg(reinterpret_cast<const uint4*>(src)); d[q][0] = v.x; d[q][1] = v.y; d[q][W / 4 - 2] = v.z; d[q][W / 4 - 1] = v.w; }
        else { uint2 v = __ldg(reinterpret_cast<const uint2*>(src)); d[q][0] = v.x; d[q][1] = v.y; }
      }
#pragma unroll
      for (int q = 0; q < NR; ++q)
#pragma unroll
        for (int j = 0; j < W / 4; ++j)
          acc[q] += u8_f64(d[q][j], 0) * y0 + u8_f64(d[q][j], 1) * y1 + u8_f64(d[q][j], 2) * y0 + u8_f64(d[q][j], 3) * y1;
    }
  }
  double s = 0; for (int q = 0; q < NR; ++q) s += acc[q];
  if (s == 123.456) out[0] = s;
}
// one row per warp, whole row streamed with uint4 loads, U loads in flight along the row, 4 accumulators
template <int U>
__global__ void __launch_bounds__(256) k_row1(const uint8_t* __restrict__ p, int n_rows, int ppad, double* out) {
  const int warp = (int)(((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  const int nw = (int)(((size_t)gridDim.x * blockDim.x) >> 5);
  double acc[4] = {0, 0, 0, 0};
  const double y0 = 1.0 + lane * 1e-3, y1 = 0.5 + lane * 1e-3;
  const int n16 = ppad / 16;
  for (int r = warp; r < n_rows; r += nw) {
    const uint4* src = reinterpret_cast<const uint4*>(p + (size_t)r * ppad);
    for (int i = lane; i < n16; i += 32 * U) {
      uint4 v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) v[u] = (i + 32 * u < n16) ? __ldg(src + i + 32 * u) : make_uint4(0, 0, 0, 0);
#pragma unroll
      for (int u = 0; u < U; ++u) {
        acc[0] += u8_f64(v[u].x, 0) * y0 + u8_f64(v[u].x, 1) * y1 + u8_f64(v[u].x, 2) * y0 + u8_f64(v[u].x, 3) * y1;
        acc[1] += u8_f64(v[u].y, 0) * y0 + u8_f64(v[u].y, 1) * y1 + u8_f64(v[u].y, 2) * y0 + u8_f64(v[u].y, 3) * y1;
        acc[2] += u8_f64(v[u].z, 0) * y0 + u8_f64(v[u].z, 1) * y1 + u8_f64(v[u].z, 2) * y0 + u8_f64(v[u].z, 3) * y1;
        acc[3] += u8_f64(v[u].w, 0) * y0 + u8_f64(v[u].w, 1) * y1 + u8_f64(v[u].w, 2) * y0 + u8_f64(v[u].w, 3) * y1;
      }
    }
  }
  double s = acc[0] + acc[1] + acc[2] + acc[3];
  if (s == 123.456) out[0] = s;
}
int main() {
  const int n_rows = 1000000, ppad = 5888;
  const size_t bytes = (size_t)n_rows * ppad;
  uint8_t* p; double* out;
  CK(cudaMalloc(&p, bytes)); CK(cudaMalloc(&out, 8));
  CK(cudaMemset(p, 1, bytes));
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  auto time = [&](const char* name, auto launch) {
    launch(); cudaDeviceSynchronize();
    cudaEventRecord(a); for (int i = 0; i < 5; ++i) launch(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); ms /= 5;
    printf("%-28s %7.3f ms  %6.0f GB/s  (%s)\n", name, ms, bytes / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
  };
  for (int g : {148 * 16}) {
    printf("grid %d\n", g);
    time("stream xor", [&] { k_stream<0><<<g, 256>>>((const uint4*)p, bytes / 16, out); });
    time("stream magic+fma (2 op/B)", [&] { k_stream<1><<<g, 256>>>((const uint4*)p, bytes / 16, out); });
    time("stream fma only (1 op/B)", [&] { k_stream<2><<<g, 256>>>((const uint4*)p, bytes / 16, out); });
    time("stream cvt.f64 + fma", [&] { k_stream<3><<<g, 256>>>((const uint4*)p, bytes / 16, out); });
    time("stream fp32 cvt + ffma", [&] { k_stream<4><<<g, 256>>>((const uint4*)p, bytes / 16, out); });
  }
  for (int g : {148 * 4, 148 * 8, 148 * 16}) {
    printf("grid %d\n", g);
    time("rows2 NR=4 W=8", [&] { k_rows2<4, 8><<<g, 256>>>(p, n_rows, 5888, out); });
    time("rows2 NR=8 W=8", [&] { k_rows2<8, 8><<<g, 256>>>(p, n_rows, 5888, out); });
    time("rows2 NR=2 W=16", [&] { k_rows2<2, 16><<<g, 256>>>(p, n_rows, 5632, out); });
    time("rows2 NR=4 W=16", [&] { k_rows2<4, 16><<<g, 256>>>(p, n_rows, 5632, out); });
    time("row1 U=2", [&] { k_row1<2><<<g, 256>>>(p, n_rows, ppad, out); });
    time("row1 U=4", [&] { k_row1<4><<<g, 256>>>(p, n_rows, ppad, out); });
  }
  time("rows xor", [&] { k_rows<0><<<148 * 8, 256>>>(p, n_rows, ppad, out); });
  time("rows magic+fma", [&] { k_rows<1><<<148 * 8, 256>>>(p, n_rows, ppad, out); });
  return 0;
}
