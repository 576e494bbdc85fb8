// the steps immediately before the hot path, on the GPU.
//   tmc_mtx_parse          10x MatrixMarket text (features x cells, coordinate) -> cells x features CSR
//                          == readMatrix + matToSpMat + transposeSM + sparsifySM   (src/TooManyCells/Matrix/Load.hs:85-152)
//   tmc_filter_thresholds  --filter-thresholds "(MINCELL, MINFEATURE)"             (src/TooManyCells/Matrix/Preprocess.hs:342-386)
// Byte / integer work bound by HBM: newline scan, one thread per line, radix sort by (cell, feature), compaction.
// Included at the end of tmc_engine.cu (same translation unit: shares TmcError / CK / guarded / dalloc).
#pragma once
#include <cub/device/device_scan.cuh>

namespace tmc {

__constant__ double tmc_p10[23] = {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};

// ---- text -> lines
__global__ void k_nl_count(const char* __restrict__ t, int64_t n, int64_t start, int* __restrict__ blk_cnt) {
  // a "line start" is a byte at position >= start whose predecessor is '\n' (or position == start)
  const int64_t base = start + (int64_t)blockIdx.x * 4096;
  int c = 0;
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) {
    const int64_t p = base + i;
    if (p < n) c += (p == start || t[p - 1] == '\n') ? 1 : 0;
  }
  c = __reduce_add_sync(0xffffffffu, c);
  __shared__ int sm[8];
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) { int s = 0; for (int q = 0; q < (int)(blockDim.x >> 5); ++q) s += sm[q]; blk_cnt[blockIdx.x] = s; }
}
__global__ void k_nl_mark(const char* __restrict__ t, int64_t n, int64_t start, const int64_t* __restrict__ blk_off, int64_t* __restrict__ line_start) {
  // 256 threads x 16 consecutive bytes each; block-wide exclusive scan of per-thread counts
  __shared__ int wsum[8];
  const int64_t base = start + (int64_t)blockIdx.x * 4096 + threadIdx.x * 16;
  int flags = 0, cnt = 0;
  for (int i = 0; i < 16; ++i) { const int64_t p = base + i; if (p < n && (p == start || t[p - 1] == '\n')) { flags |= 1 << i; cnt++; } }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int incl = cnt;
  for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
  if (lane == 31) wsum[wid] = incl;
  __syncthreads();
  int wbase = 0; for (int q = 0; q < wid; ++q) wbase += wsum[q];
  int64_t o = blk_off[blockIdx.x] + wbase + incl - cnt;
  for (int i = 0; i < 16; ++i) if (flags & (1 << i)) line_start[o++] = base + i;
}

// ---- one thread per line: "i j [v]"  (1-based indices; v absent for `pattern`)
__device__ __forceinline__ bool is_ws(char ch) { return ch == ' ' || ch == '\t' || ch == '\r'; }
__global__ void k_parse_lines(const char* __restrict__ t, int64_t n, const int64_t* __restrict__ line_start, int64_t n_lines, int pattern,
                              int64_t* __restrict__ key, double* __restrict__ val, uint8_t* __restrict__ valid, int n_feat, int n_cell,
                              int transpose, int* __restrict__ err) {
  const int64_t li = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (li >= n_lines) return;
  int64_t p = line_start[li];
  const int64_t end = (li + 1 < n_lines) ? line_start[li + 1] : n;
  auto skip = [&]() { while (p < end && is_ws(t[p])) ++p; };
  auto at_eol = [&]() { return p >= end || t[p] == '\n'; };
  skip();
  if (at_eol() || t[p] == '%') { valid[li] = 0; key[li] = 0; val[li] = 0; return; }       // blank / comment line
  long long a[2] = {0, 0};
  for (int q = 0; q < 2; ++q) {
    skip();
    if (at_eol() || t[p] < '0' || t[p] > '9') { atomicMax(err, 2); valid[li] = 0; return; }
    long long x = 0;
    while (p < end && t[p] >= '0' && t[p] <= '9') { x = x * 10 + (t[p] - '0'); ++p; }
    a[q] = x;
  }
  double v = 1.0;
  if (!pattern) {
    skip();
    if (at_eol()) { atomicMax(err, 2); valid[li] = 0; return; }
    bool neg = false;
    if (t[p] == '-') { neg = true; ++p; } else if (t[p] == '+') ++p;
    unsigned long long mant = 0; int dexp = 0, nd = 0; bool any = false;
    while (p < end && t[p] >= '0' && t[p] <= '9') { if (nd < 19) { mant = mant * 10 + (t[p] - '0'); nd += (mant != 0); } else dexp++; ++p; any = true; }
    if (p < end && t[p] == '.') { ++p; while (p < end && t[p] >= '0' && t[p] <= '9') { if (nd < 19) { mant = mant * 10 + (t[p] - '0'); nd += (mant != 0); dexp--; } ++p; any = true; } }
    if (!any) { atomicMax(err, 2); valid[li] = 0; return; }
    if (p < end && (t[p] == 'e' || t[p] == 'E')) {
      ++p; bool eneg = false; if (p < end && (t[p] == '-' || t[p] == '+')) { eneg = t[p] == '-'; ++p; }
      int ex = 0; while (p < end && t[p] >= '0' && t[p] <= '9') { ex = ex * 10 + (t[p] - '0'); if (ex > 10000) ex = 10000; ++p; }
      dexp += eneg ? -ex : ex;
    }
    // one correctly rounded operation for mantissas < 2^53 and |dexp| <= 22 (all count matrices); pow() otherwise
    double m = (double)mant;
    if (dexp == 0) v = m; else if (dexp > 0 && dexp <= 22) v = m * tmc_p10[dexp]; else if (dexp < 0 && dexp >= -22) v = m / tmc_p10[-dexp]; else v = m * pow(10.0, (double)dexp);
    if (neg) v = -v;
  }
  skip();
  if (!at_eol()) { atomicMax(err, 2); valid[li] = 0; return; }
  if (a[0] < 1 || a[0] > n_feat || a[1] < 1 || a[1] > n_cell) { atomicMax(err, 3); valid[li] = 0; return; }
  const long long r = transpose ? a[1] - 1 : a[0] - 1, cidx = transpose ? a[0] - 1 : a[1] - 1;
  key[li] = (r << 32) | cidx;
  val[li] = v;
  valid[li] = (v != 0.0) ? 1 : 0;                    // sparsifySM: explicit zeros are dropped
}

__global__ void k_flag_to_i32(const uint8_t* __restrict__ f, int64_t n, int* __restrict__ o) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; if (i < n) o[i] = f[i];
}
__global__ void k_compact_kv(const int64_t* __restrict__ key, const double* __restrict__ val, const uint8_t* __restrict__ valid, const int64_t* __restrict__ off,
                             int64_t n, int64_t* __restrict__ okey, int64_t* __restrict__ oidx) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && valid[i]) { okey[off[i]] = key[i]; oidx[off[i]] = i; }
}
__global__ void k_unpack_sorted(const int64_t* __restrict__ skey, const int64_t* __restrict__ sidx, const double* __restrict__ val, int64_t n,
                                int32_t* __restrict__ col, double* __restrict__ oval, int* __restrict__ row_cnt, int* __restrict__ dup) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t k = skey[i];
  if (i && skey[i - 1] == k) atomicOr(dup, 1);       // the same (cell, feature) twice: the engine's CSR contract forbids it (tmc.h)
  col[i] = (int32_t)(k & 0xffffffffll); oval[i] = val[sidx[i]];
  atomicAdd(row_cnt + (int)(k >> 32), 1);
}
__global__ void k_i32_to_i64(const int* __restrict__ a, int64_t n, int64_t* __restrict__ o) { int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; if (i < n) o[i] = a[i]; }

// ---- filter thresholds
__global__ void k_row_col_sums(const int64_t* __restrict__ rp, const int32_t* __restrict__ col, const double* __restrict__ val, int n_rows,
                               double* __restrict__ rsum, double* __restrict__ csum) {
  const int warp = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  const int nw = (int)(((int64_t)gridDim.x * blockDim.x) >> 5);
  for (int r = warp; r < n_rows; r += nw) {
    double s = 0;
    for (int64_t k = rp[r] + lane; k < rp[r + 1]; k += 32) { const double v = val[k]; s += v; atomicAdd(csum + col[k], v); }
    s = warp_sum(s);
    if (lane == 0) rsum[r] = s;
  }
}
__global__ void k_thresh_flags(const double* __restrict__ s, int n, double th, int* __restrict__ f) { int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) f[i] = s[i] >= th ? 1 : 0; }
__global__ void k_filter_count(const int64_t* __restrict__ rp, const int32_t* __restrict__ col, const int* __restrict__ rkeep, const int* __restrict__ ckeep,
                               const int64_t* __restrict__ rnew, int n_rows, int* __restrict__ cnt) {
  const int warp = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  const int nw = (int)(((int64_t)gridDim.x * blockDim.x) >> 5);
  for (int r = warp; r < n_rows; r += nw) {
    if (!rkeep[r]) continue;
    int c = 0;
    for (int64_t k = rp[r] + lane; k < rp[r + 1]; k += 32) c += ckeep[col[k]];
    c = __reduce_add_sync(0xffffffffu, c);
    if (lane == 0) cnt[rnew[r]] = c;
  }
}
__global__ void k_filter_fill(const int64_t* __restrict__ rp, const int32_t* __restrict__ col, const double* __restrict__ val, const int* __restrict__ rkeep,
                              const int* __restrict__ ckeep, const int64_t* __restrict__ rnew, const int64_t* __restrict__ cnew,
                              const int64_t* __restrict__ orp, int n_rows, int32_t* __restrict__ ocol, double* __restrict__ oval) {
  const int warp = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  const int nw = (int)(((int64_t)gridDim.x * blockDim.x) >> 5);
  for (int r = warp; r < n_rows; r += nw) {
    if (!rkeep[r]) continue;
    int64_t o = orp[rnew[r]];
    for (int64_t k0 = rp[r]; k0 < rp[r + 1]; k0 += 32) {          // order inside the row is preserved
      const int64_t k = k0 + lane;
      const int keep = (k < rp[r + 1]) ? ckeep[col[k]] : 0;
      const unsigned bal = __ballot_sync(0xffffffffu, keep);
      if (keep) { const int64_t d = o + __popc(bal & ((1u << lane) - 1u)); ocol[d] = (int32_t)cnew[col[k]]; oval[d] = val[k]; }
      o += __popc(bal);
    }
  }
}

}  // namespace tmc

struct CsrOutStorage {
  tmc_csr_out o;
  std::vector<int64_t> row_ptr, kept_rows, kept_cols; std::vector<int32_t> col; std::vector<double> val;
};

template <class T> static void exclusive_scan_dev(const T* in, int64_t* out, int64_t n, int64_t* total_host) {
  // out[i] = sum_{k<i} in[k] with int64 accumulation (cub::DeviceScan, plumbing); T = int
  if (n <= 0) { if (total_host) *total_host = 0; return; }
  int64_t* wide = nullptr; void* tmp = nullptr; size_t bytes = 0;
  CK(cudaMalloc(&wide, sizeof(int64_t) * n));
  try {
    tmc::k_i32_to_i64<<<(unsigned)((n + 255) / 256), 256>>>(in, n, wide);
    CK(cub::DeviceScan::ExclusiveSum(nullptr, bytes, wide, out, n));
    CK(cudaMalloc(&tmp, std::max<size_t>(bytes, 16)));
    CK(cub::DeviceScan::ExclusiveSum(tmp, bytes, wide, out, n));
    if (total_host) {
      int64_t last = 0, lastv = 0;
      CK(cudaMemcpy(&last, out + n - 1, sizeof(int64_t), cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(&lastv, wide + n - 1, sizeof(int64_t), cudaMemcpyDeviceToHost));
      *total_host = last + lastv;
    }
  } catch (...) { cudaFree(wide); if (tmp) cudaFree(tmp); throw; }
  cudaFree(wide); cudaFree(tmp);
}

extern "C" {

int tmc_mtx_parse(const char* text, int64_t n_bytes, int transpose, tmc_csr_out** out) {
  return guarded([&] {
    if (!text || !out || n_bytes <= 0) throw TmcError(TMC_ERR_INVALID, "null / empty MatrixMarket text");
    pick_device(-1);
    // header on the host: banner, comments, size line
    int64_t p = 0; bool pattern = false, have_banner = false;
    auto line_end = [&](int64_t q) { while (q < n_bytes && text[q] != '\n') ++q; return q; };
    while (p < n_bytes && text[p] == '%') {
      int64_t e = line_end(p);
      std::string ln(text + p, text + e);
      if (ln.rfind("%%MatrixMarket", 0) == 0) {
        have_banner = true;
        if (ln.find("coordinate") == std::string::npos) throw TmcError(TMC_ERR_UNSUPPORTED, "only `coordinate` MatrixMarket files are supported");
        if (ln.find("pattern") != std::string::npos) pattern = true;
        if (ln.find("complex") != std::string::npos || ln.find("symmetric") != std::string::npos || ln.find("hermitian") != std::string::npos)
          throw TmcError(TMC_ERR_UNSUPPORTED, "complex / symmetric MatrixMarket files are not supported");
      }
      p = e + 1;
    }
    if (!have_banner) throw TmcError(TMC_ERR_INVALID, "missing %%MatrixMarket banner");
    long long nf = 0, nc = 0, nz = 0;
    { int64_t e = line_end(p); std::string ln(text + p, text + e);
      if (sscanf(ln.c_str(), "%lld %lld %lld", &nf, &nc, &nz) != 3 || nf < 0 || nc < 0 || nz < 0) throw TmcError(TMC_ERR_INVALID, "bad MatrixMarket size line");
      p = e + 1; }
    if (nf >= (1ll << 31) || nc >= (1ll << 31)) throw TmcError(TMC_ERR_INVALID, "dimension exceeds int32");
    const int64_t start = std::min<int64_t>(p, n_bytes);
    const int64_t n_rows = transpose ? nc : nf, n_cols = transpose ? nf : nc;
    std::vector<void*> frees;
    auto A = [&](size_t bytes) { void* q = nullptr; CK(cudaMalloc(&q, std::max<size_t>(bytes, 16))); frees.push_back(q); return q; };
    CsrOutStorage* st = new CsrOutStorage();
    try {
      int64_t kept = 0;
      int32_t* d_col = nullptr; double* d_oval = nullptr; int64_t* d_rp = nullptr;
      if (n_bytes > start) {
        char* d_text = (char*)A(n_bytes);
        CK(cudaMemcpy(d_text, text, n_bytes, cudaMemcpyHostToDevice));
        const int64_t nblk = (n_bytes - start + 4095) / 4096;
        int* d_bc = (int*)A(sizeof(int) * nblk); int64_t* d_bo = (int64_t*)A(sizeof(int64_t) * nblk);
        k_nl_count<<<(unsigned)nblk, 256>>>(d_text, n_bytes, start, d_bc);
        int64_t n_lines = 0;
        exclusive_scan_dev<int>(d_bc, d_bo, nblk, &n_lines);
        int64_t* d_ls = (int64_t*)A(sizeof(int64_t) * std::max<int64_t>(n_lines, 1));
        k_nl_mark<<<(unsigned)nblk, 256>>>(d_text, n_bytes, start, d_bo, d_ls);
        int64_t* d_key = (int64_t*)A(sizeof(int64_t) * std::max<int64_t>(n_lines, 1)); double* d_val = (double*)A(sizeof(double) * std::max<int64_t>(n_lines, 1));
        uint8_t* d_valid = (uint8_t*)A(std::max<int64_t>(n_lines, 1)); int* d_err = (int*)A(sizeof(int));
        CK(cudaMemset(d_err, 0, sizeof(int)));
        if (n_lines) k_parse_lines<<<(unsigned)((n_lines + 255) / 256), 256>>>(d_text, n_bytes, d_ls, n_lines, pattern ? 1 : 0, d_key, d_val, d_valid, (int)nf, (int)nc, transpose ? 1 : 0, d_err);
        int herr = 0; CK(cudaMemcpy(&herr, d_err, sizeof(int), cudaMemcpyDeviceToHost));
        if (herr == 2) throw TmcError(TMC_ERR_INVALID, "malformed MatrixMarket entry line");
        if (herr == 3) throw TmcError(TMC_ERR_INVALID, "MatrixMarket index out of range");
        int* d_vi = (int*)A(sizeof(int) * std::max<int64_t>(n_lines, 1)); int64_t* d_off = (int64_t*)A(sizeof(int64_t) * std::max<int64_t>(n_lines, 1));
        if (n_lines) k_flag_to_i32<<<(unsigned)((n_lines + 255) / 256), 256>>>(d_valid, n_lines, d_vi);
        exclusive_scan_dev<int>(d_vi, d_off, n_lines, &kept);
        if (kept) {
          int64_t* kA = (int64_t*)A(sizeof(int64_t) * kept); int64_t* kB = (int64_t*)A(sizeof(int64_t) * kept);
          int64_t* iA = (int64_t*)A(sizeof(int64_t) * kept); int64_t* iB = (int64_t*)A(sizeof(int64_t) * kept);
          k_compact_kv<<<(unsigned)((n_lines + 255) / 256), 256>>>(d_key, d_val, d_valid, d_off, n_lines, kA, iA);
          cub::DoubleBuffer<int64_t> dk(kA, kB), dv(iA, iB);
          size_t bytes = 0; CK(cub::DeviceRadixSort::SortPairs(nullptr, bytes, dk, dv, kept, 0, 63));
          void* tmp = A(bytes);
          CK(cub::DeviceRadixSort::SortPairs(tmp, bytes, dk, dv, kept, 0, 63));          // stable: by (row, col)
          d_col = (int32_t*)A(sizeof(int32_t) * kept); d_oval = (double*)A(sizeof(double) * kept);
          int* d_rc = (int*)A(sizeof(int) * (n_rows + 1)); CK(cudaMemset(d_rc, 0, sizeof(int) * (n_rows + 1)));
          int* d_dup = (int*)A(sizeof(int)); CK(cudaMemset(d_dup, 0, sizeof(int)));
          k_unpack_sorted<<<(unsigned)((kept + 255) / 256), 256>>>(dk.Current(), dv.Current(), d_val, kept, d_col, d_oval, d_rc, d_dup);
          int dup = 0; CK(cudaMemcpy(&dup, d_dup, sizeof(int), cudaMemcpyDeviceToHost));
          if (dup) throw TmcError(TMC_ERR_INVALID, "duplicate (row, column) entry in the MatrixMarket text");
          d_rp = (int64_t*)A(sizeof(int64_t) * (n_rows + 1));
          exclusive_scan_dev<int>(d_rc, d_rp, n_rows + 1, nullptr);
        }
        CK(cudaGetLastError());
      }
      st->row_ptr.assign(n_rows + 1, 0); st->col.resize(kept); st->val.resize(kept);
      if (kept) {
        CK(cudaMemcpy(st->row_ptr.data(), d_rp, sizeof(int64_t) * (n_rows + 1), cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(st->col.data(), d_col, sizeof(int32_t) * kept, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(st->val.data(), d_oval, sizeof(double) * kept, cudaMemcpyDeviceToHost));
      }
      (void)nz;
      st->o = tmc_csr_out{n_rows, n_cols, kept, st->row_ptr.data(), st->col.data(), st->val.data(), nullptr, nullptr, 0, 0};
    } catch (...) { for (void* q : frees) cudaFree(q); delete st; throw; }
    for (void* q : frees) cudaFree(q);
    *out = &st->o;
  });
}

int tmc_filter_thresholds(const tmc_csr* a, double row_thresh, double col_thresh, tmc_csr_out** out) {
  return guarded([&] {
    if (!out) throw TmcError(TMC_ERR_INVALID, "null out");
    validate_csr(a);
    require_default_types(a, "tmc_filter_thresholds");
    pick_device(-1);
    const int64_t m = a->n_rows, n = a->n_cols, nnz = a->nnz;
    std::vector<void*> frees;
    auto A = [&](size_t bytes) { void* q = nullptr; CK(cudaMalloc(&q, std::max<size_t>(bytes, 16))); frees.push_back(q); return q; };
    CsrOutStorage* st = new CsrOutStorage();
    try {
      int64_t* rp = (int64_t*)A(sizeof(int64_t) * (m + 1)); int32_t* col = (int32_t*)A(sizeof(int32_t) * nnz); double* val = (double*)A(sizeof(double) * nnz);
      CK(cudaMemcpy(rp, a->row_ptr, sizeof(int64_t) * (m + 1), cudaMemcpyHostToDevice));
      if (nnz) { CK(cudaMemcpy(col, a->col_idx, sizeof(int32_t) * nnz, cudaMemcpyHostToDevice)); CK(cudaMemcpy(val, a->val, sizeof(double) * nnz, cudaMemcpyHostToDevice)); }
      double* rs = (double*)A(sizeof(double) * m); double* cs = (double*)A(sizeof(double) * n);
      CK(cudaMemset(cs, 0, sizeof(double) * n));
      k_row_col_sums<<<148 * 8, 256>>>(rp, col, val, (int)m, rs, cs);
      int* rk = (int*)A(sizeof(int) * (m + 1)); int* ck = (int*)A(sizeof(int) * (n + 1));
      k_thresh_flags<<<(unsigned)((m + 255) / 256), 256>>>(rs, (int)m, row_thresh, rk);
      k_thresh_flags<<<(unsigned)((n + 255) / 256), 256>>>(cs, (int)n, col_thresh, ck);
      int64_t* rnew = (int64_t*)A(sizeof(int64_t) * (m + 1)); int64_t* cnew = (int64_t*)A(sizeof(int64_t) * (n + 1));
      int64_t mk = 0, nk = 0;
      exclusive_scan_dev<int>(rk, rnew, m, &mk);
      exclusive_scan_dev<int>(ck, cnew, n, &nk);
      if (mk == 0) throw TmcError(TMC_ERR_INVALID, "no cells pass the filter (reference: emptyMatCheckErr \"cells\", Preprocess.hs:91-97)");
      if (nk == 0) throw TmcError(TMC_ERR_INVALID, "no features pass the filter (reference: emptyMatCheckErr \"features\")");
      int* cnt = (int*)A(sizeof(int) * (mk + 1)); CK(cudaMemset(cnt, 0, sizeof(int) * (mk + 1)));
      k_filter_count<<<148 * 8, 256>>>(rp, col, rk, ck, rnew, (int)m, cnt);
      int64_t* orp = (int64_t*)A(sizeof(int64_t) * (mk + 1));
      int64_t onnz = 0;
      exclusive_scan_dev<int>(cnt, orp, mk + 1, nullptr);
      CK(cudaMemcpy(&onnz, orp + mk, sizeof(int64_t), cudaMemcpyDeviceToHost));
      int32_t* ocol = (int32_t*)A(sizeof(int32_t) * onnz); double* oval = (double*)A(sizeof(double) * onnz);
      k_filter_fill<<<148 * 8, 256>>>(rp, col, val, rk, ck, rnew, cnew, orp, (int)m, ocol, oval);
      CK(cudaGetLastError());
      st->row_ptr.resize(mk + 1); st->col.resize(onnz); st->val.resize(onnz);
      CK(cudaMemcpy(st->row_ptr.data(), orp, sizeof(int64_t) * (mk + 1), cudaMemcpyDeviceToHost));
      if (onnz) { CK(cudaMemcpy(st->col.data(), ocol, sizeof(int32_t) * onnz, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(st->val.data(), oval, sizeof(double) * onnz, cudaMemcpyDeviceToHost)); }
      std::vector<int> hrk(m), hck(n);
      CK(cudaMemcpy(hrk.data(), rk, sizeof(int) * m, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(hck.data(), ck, sizeof(int) * n, cudaMemcpyDeviceToHost));
      for (int64_t i = 0; i < m; ++i) if (hrk[i]) st->kept_rows.push_back(i);
      for (int64_t j = 0; j < n; ++j) if (hck[j]) st->kept_cols.push_back(j);
      st->o = tmc_csr_out{mk, nk, onnz, st->row_ptr.data(), st->col.data(), st->val.data(), st->kept_rows.data(), st->kept_cols.data(), mk, nk};
    } catch (...) { for (void* q : frees) cudaFree(q); delete st; throw; }
    for (void* q : frees) cudaFree(q);
    *out = &st->o;
  });
}

void tmc_csr_out_free(tmc_csr_out* o) { if (o) delete reinterpret_cast<CsrOutStorage*>(o); }

}  // extern "C"
