// One-shot all-reduce over NVLink peer memory: the per-sweep exchanges of the row-sharded
// top of the tree — y = B^T x (n doubles per live shared node), the CGS coefficients and the scalars — are small (0.24-60 MB),
// latency-bound messages.  Instead of three ncclAllReduce calls per sweep, every rank PUBLISHES its partial vector in a window
// of its own HBM that all peers have mapped (cudaIpc over NVSwitch), signals the peers with one release-store each, and then
// PULLS and sums the N windows itself, in rank order: one kernel pair per exchange, no ring, no proxy thread, and a result that
// is bit-identical on every rank by construction (fixed summation order — the determinism requirement of SURVEY 8(e)).
//
//   window of rank r:  [ flags: 64 x u64 | block counter | buf 0: cap doubles | buf 1: cap doubles ]
//   exchange e (epoch e, buffer e & 1):
//     k_p2p_publish   src -> own buf;  last CTA: st.release.sys  flags_of_peer[q][r] = e   for every peer q
//     k_p2p_reduce    wait (ld.acquire.sys) own flags[q] >= e for every q;  dst[i] = sum_q buf_q[i]  (q = 0..N-1)
//   Buffer e & 1 is reused by exchange e+2; by then every peer has finished reading exchange e, because rank r only starts
//   publishing e+2 after its own reduce e+1, which waited for every peer's signal e+1, which each peer issued (stream order)
//   after its reduce e.
//
// STATUS: written at the end of round 1 with no GPU minutes left — NOT yet run on hardware, therefore OFF by default.
// TMC_P2P=1 enables it (every rank must set it); tools/p2p_check.py (torchrun, N >= 2) checks it against ncclAllReduce and times
// both through tmc_comm_selftest.  Spins are bounded: a peer that never signals raises device error flag 9 -> TMC_ERR_COMM
// instead of hanging the GPU.  Any failure while setting the windows up (no P2P, IPC refused) is agreed on by all ranks and
// leaves the NCCL path in place.
#pragma once

namespace p2p {

constexpr int MAX_PEERS = 16;
constexpr size_t HDR_BYTES = 1024;          // 64 flags (512 B) + counter, padded
constexpr unsigned long long SPIN_LIMIT = 1ull << 24;      // ~ 10-20 s of polling (about 1 us per probe); then error flag 9

struct Args {
  const double* win[MAX_PEERS];             // buffer (epoch & 1) of every rank; win[rank] = own
  unsigned long long* peer_flags[MAX_PEERS];// flags array inside the window of rank q
  unsigned long long* my_flags;             // == peer_flags[rank]
  unsigned int* counter;                    // CTA arrival counter of the publish kernel (own window)
  double* own_buf;
  unsigned long long epoch;
  size_t n;
  int world, rank;
  int* err;                                 // engine error flag (may be null)
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// src -> own window buffer; the last CTA to finish tells every peer that exchange `epoch` of this rank is readable
__global__ void k_p2p_publish(Args a, const double* __restrict__ src) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += stride) a.own_buf[i] = src[i];
  __threadfence_system();
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) {
    const unsigned int done = atomicAdd(a.counter, 1u);
    last = (done == gridDim.x - 1);
    if (last) *a.counter = 0u;              // every CTA has arrived: re-arm for the next publish (stream-ordered)
  }
  __syncthreads();
  if (last && (int)threadIdx.x < a.world && (int)threadIdx.x != a.rank) {
    __threadfence_system();
    st_release_sys(a.peer_flags[threadIdx.x] + a.rank, a.epoch);
  }
}

// wait until every rank has published exchange `epoch`, then dst[i] = sum over ranks (rank order) of their windows
__global__ void k_p2p_reduce(Args a, double* __restrict__ dst) {
  if ((int)threadIdx.x < a.world && (int)threadIdx.x != a.rank) {
    unsigned long long spins = 0;
    while (ld_acquire_sys(a.my_flags + threadIdx.x) < a.epoch) {
      if (++spins > SPIN_LIMIT) { if (a.err) atomicCAS(a.err, 0, 9); break; }
      __nanosleep(64);
    }
  }
  __syncthreads();
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += stride) {
    double s = 0.0;
    for (int q = 0; q < a.world; ++q) s += __ldcv(a.win[q] + i);      // .cv: never served from a stale cached line
    dst[i] = s;
  }
}

struct PeerReduce {
  bool tried = false, ok = false;
  int world = 1, rank = 0;
  size_t cap = 0;                           // doubles per buffer
  char* base = nullptr;
  char* peer[MAX_PEERS] = {nullptr};
  unsigned long long epoch = 0;

  static size_t bytes_for(size_t cap_doubles) { return HDR_BYTES + 2 * cap_doubles * sizeof(double); }
  double* buf_of(char* w, int which) const { return reinterpret_cast<double*>(w + HDR_BYTES) + (size_t)which * cap; }

  // Collective: unmap the peers' windows, wait until every rank has done so, then free the own window (an exporter must not
  // free memory that an importer still has mapped).  `in_use`: exchanges may be in flight -> drain them first.
  void release(cudaStream_t s, bool in_use) {
    if (in_use) { cudaDeviceSynchronize(); barrier(s); }
    for (int q = 0; q < world; ++q) if (q != rank && peer[q]) { cudaIpcCloseMemHandle(peer[q]); }
    for (int q = 0; q < MAX_PEERS; ++q) peer[q] = nullptr;
    if (g_nccl.comm && g_nccl.world > 1) barrier(s);
    if (base) cudaFree(base);
    base = nullptr; cap = 0; ok = false;
  }

  // Collective: every rank calls it with the same `want` (doubles).  Returns whether the peer path is usable.
  bool ensure(size_t want, cudaStream_t s) {
    if (!g_nccl.comm || g_nccl.world < 2 || g_nccl.world > MAX_PEERS) return false;
    if (tried && (!ok || cap >= want)) return ok;
    tried = true;
    if (ok) release(s, true);               // growing: drain, unmap everywhere, free, then build the larger windows
    world = g_nccl.world; rank = g_nccl.rank;
    const size_t max_cap = ((size_t)2 << 30) / (2 * sizeof(double));          // at most 2 GB of window; larger messages stay on NCCL
    cap = std::min(std::max(want, (size_t)1 << 16), max_cap);
    int good = 1;
    cudaIpcMemHandle_t mine;
    memset(&mine, 0, sizeof(mine));
    if (cudaMalloc(&base, bytes_for(cap)) != cudaSuccess) { base = nullptr; good = 0; cudaGetLastError(); }
    if (good && cudaMemset(base, 0, HDR_BYTES) != cudaSuccess) good = 0;
    if (good && cudaDeviceSynchronize() != cudaSuccess) good = 0;
    if (good && cudaIpcGetMemHandle(&mine, base) != cudaSuccess) { good = 0; cudaGetLastError(); }
    // exchange the handles (64 bytes each) with one broadcast per rank; this also orders every memset before any peer access
    std::vector<cudaIpcMemHandle_t> all(world);
    char* d_h = nullptr;
    if (cudaMalloc(&d_h, sizeof(cudaIpcMemHandle_t) * world) != cudaSuccess) { cudaGetLastError(); d_h = nullptr; }
    if (d_h) {
      cudaMemcpyAsync(d_h + sizeof(mine) * rank, &mine, sizeof(mine), cudaMemcpyHostToDevice, s);
      for (int r = 0; r < world; ++r)
        if (g_nccl.Broadcast(d_h + sizeof(mine) * r, d_h + sizeof(mine) * r, sizeof(mine), NCCL_UINT8, r, g_nccl.comm, s) != 0) good = 0;
      cudaMemcpyAsync(all.data(), d_h, sizeof(mine) * world, cudaMemcpyDeviceToHost, s);
      if (cudaStreamSynchronize(s) != cudaSuccess) good = 0;
    } else good = 0;
    good = agree(good, s);                  // all ranks hold valid handles, or nobody goes on
    if (good) {
      peer[rank] = base;
      for (int q = 0; q < world && good; ++q) {
        if (q == rank) continue;
        void* p = nullptr;
        if (cudaIpcOpenMemHandle(&p, all[q], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { good = 0; cudaGetLastError(); }
        peer[q] = static_cast<char*>(p);
      }
      good = agree(good, s);
    }
    if (d_h) cudaFree(d_h);
    if (!good) { release(s, false); return false; }
    ok = true;
    return true;
  }

  // min over ranks of a 0/1 flag (NCCL): the ranks must take the same path or the first exchange deadlocks
  int agree(int v, cudaStream_t s) {
    int* d = nullptr;
    if (cudaMalloc(&d, sizeof(int)) != cudaSuccess) { cudaGetLastError(); return 0; }      // (a rank that cannot even do this is lost anyway)
    cudaMemcpyAsync(d, &v, sizeof(int), cudaMemcpyHostToDevice, s);
    int r = g_nccl.AllReduce(d, d, 1, NCCL_INT32, NCCL_MIN, g_nccl.comm, s);
    int out = 0;
    cudaMemcpyAsync(&out, d, sizeof(int), cudaMemcpyDeviceToHost, s);
    if (cudaStreamSynchronize(s) != cudaSuccess || r != 0) out = 0;
    cudaFree(d);
    return out;
  }
  int barrier(cudaStream_t s) { return agree(1, s); }

  // in-place all-reduce (sum) of p[0, n) on stream s; n <= cap, same n on every rank
  void allreduce(double* p, size_t n, cudaStream_t s, int* err) {
    ++epoch;
    const int which = (int)(epoch & 1ull);
    Args a;
    memset(&a, 0, sizeof(a));
    for (int q = 0; q < world; ++q) { a.win[q] = buf_of(peer[q], which); a.peer_flags[q] = reinterpret_cast<unsigned long long*>(peer[q]); }
    a.my_flags = reinterpret_cast<unsigned long long*>(base);
    a.counter = reinterpret_cast<unsigned int*>(base + 64 * sizeof(unsigned long long));
    a.own_buf = buf_of(base, which);
    a.epoch = epoch; a.n = n; a.world = world; a.rank = rank; a.err = err;
    const int threads = 256;
    const int blocks = (int)std::max<size_t>(1, std::min<size_t>(148 * 4, (n + threads - 1) / threads));
    k_p2p_publish<<<blocks, threads, 0, s>>>(a, p);
    k_p2p_reduce<<<blocks, threads, 0, s>>>(a, p);
  }

  // Collective teardown (tmc_comm_destroy)
  void destroy(cudaStream_t s) {
    if (ok || base) release(s, ok);
    tried = false; epoch = 0;
  }
};

static PeerReduce g_p2p;
static bool enabled() { const char* e = getenv("TMC_P2P"); return e && atoi(e) != 0; }

}  // namespace p2p
