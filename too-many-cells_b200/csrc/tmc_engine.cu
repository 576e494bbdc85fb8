// libtmc: host scheduler + C ABI (include/tmc.h) of the B200 spectral-bipartition engine.
//
// Replaces `hierarchicalSpectralCluster ... (Left mat)` as called from
// src/TooManyCells/MakeTree/Cluster.hs:147-156,169 and `b1ToB2` as called from
// src/TooManyCells/Matrix/Preprocess.hs:204-205.  The recursion `go items B` of the
// reference is sequential, one node at a time, re-materialising a sparse matrix per
// node; here the tree is a permutation with segment boundaries and every live node
// advances in the same kernel sweep.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdio.h>
#include <string.h>
#include <algorithm>
#include <chrono>
#include <stdlib.h>
#include <deque>
#include <cmath>
#include <functional>
#include <limits>
#include <memory>
#include <mutex>
#include <queue>
#include <string>
#include <vector>

#include "../../include/tmc.h"
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include "tmc_kernels.cuh"
#include "tmc_small.cuh"

using namespace tmc;

static_assert(sizeof(tmc_csr) == 72, "tmc_csr layout is part of the ABI (mirrored in _lib.py)");
static_assert(sizeof(tmc_opts) == 96, "tmc_opts layout is part of the ABI (mirrored in _lib.py)");
static_assert(sizeof(tmc_tree) == 152, "tmc_tree layout is part of the ABI (mirrored in _lib.py)");

static thread_local std::string g_err;
static int fail(int code, const std::string& msg) { g_err = msg; return code; }
#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) {                                                                       \
      throw TmcError(e_ == cudaErrorMemoryAllocation ? TMC_ERR_NOMEM : TMC_ERR_CUDA,               \
                     std::string(#call) + ": " + cudaGetErrorString(e_));                          \
    }                                                                                              \
  } while (0)

struct TmcError { int code; std::string msg; TmcError(int c, std::string m) : code(c), msg(std::move(m)) {} };

// ----------------------------------------------------------------------------- NCCL via dlopen (no link-time dependency)
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
struct Nccl {
  void* lib = nullptr;
  int (*GetUniqueId)(ncclUniqueId*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Broadcast)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1;
  bool load() {
    if (lib) return true;
    const char* names[] = {"libnccl.so.2", "libnccl.so", nullptr};
    for (int i = 0; names[i] && !lib; ++i) lib = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
    if (!lib) return false;
    GetUniqueId = (decltype(GetUniqueId))dlsym(lib, "ncclGetUniqueId");
    CommInitRank = (decltype(CommInitRank))dlsym(lib, "ncclCommInitRank");
    CommDestroy = (decltype(CommDestroy))dlsym(lib, "ncclCommDestroy");
    AllReduce = (decltype(AllReduce))dlsym(lib, "ncclAllReduce");
    Broadcast = (decltype(Broadcast))dlsym(lib, "ncclBroadcast");
    GetErrorString = (decltype(GetErrorString))dlsym(lib, "ncclGetErrorString");
    return GetUniqueId && CommInitRank && CommDestroy && AllReduce && Broadcast;
  }
};
static Nccl g_nccl;
enum { NCCL_UINT8 = 1, NCCL_INT32 = 2, NCCL_FLOAT64 = 8, NCCL_SUM = 0, NCCL_MAX = 2, NCCL_MIN = 3 };
#include "tmc_p2p.cuh"      // one-shot all-reduce over NVLink peer memory (TMC_P2P=1; off by default until validated on hardware)

// ----------------------------------------------------------------------------- device matrix
struct Workspace;
struct tmc_matrix {
  int64_t n_rows = 0, n_cols = 0, nnz = 0;   // n_rows = rows owned by this rank
  int64_t row_offset = 0, n_rows_global = 0; // static row-block ownership across ranks
  // replicated mode: every rank holds ALL rows (n_rows == n_rows_global) and owns the block [own_lo, own_hi) for the shared
  // top of the tree; small subtrees are handed to single ranks (no collectives below the hand-off size)
  bool replicated = false; int64_t own_lo = 0, own_hi = 0;
  // replicated mode, load time: every rank prepares only its own row block (tf-idf factors, panel rows, residual) and the
  // ranks exchange the PREPARED blocks; the raw col/val of the other ranks' blocks are fetched only if a fallback needs them
  bool raw_complete = true;          // raw col/val of every block are on this device
  bool own_prepare = false;          // prepare_matrix is working on [own_lo, own_hi) and uses collectives
  std::vector<int64_t> blk_e;        // raw entry boundaries of the ranks' row blocks: row_ptr[n_rows * r / world], r = 0..world
  int64_t* row_ptr = nullptr; int32_t* col = nullptr; double* val = nullptr;
  bool owns = false;
  size_t raw_cap[3] = {0, 0, 0};             // owned arrays: byte capacities of row_ptr / col / val (they go back to the raw-buffer cache)
  int device = 0;
  // storage mode (DESIGN.md "factored values"): VT_F64 = val holds B itself; VT_U16 / VT_ONE = B = diag(rrow) A diag(w)
  // with A the caller's small-integer values kept as uint16 (valc) / implicit ones
  int vt = VT_F64;
  bool has_neg = false;                    // entries < 0 (centred / PCA-reduced input): fp64 storage, no deflation, two degree passes
  void* valc = nullptr; double* rrow = nullptr; double* w2col = nullptr; double* invw = nullptr;
  size_t fb_cap_nnz = 0; void* fb_valc_spare = nullptr;
  // dense panel (factored modes, DESIGN.md "dense panel"): the P most frequent columns as bytes, the rest as a residual CSR with
  // renumbered columns (panel columns first); valc then holds the residual values
  int P = 0, Ppad = 0;
  int pnib = 0, pw = 0;                    // pnib: nibble cells (counts <= 15) instead of byte cells; pw: bytes per panel row
  uint8_t* panel = nullptr; int64_t* row_ptr_r = nullptr; int32_t* col_r = nullptr; int64_t nnz_r = 0;
  size_t cap_panel = 0, cap_rp = 0, cap_colr = 0;
  const int64_t* rp() const { return P ? row_ptr_r : row_ptr; }
  const int32_t* ci() const { return P ? col_r : col; }
  int64_t enz() const { return P ? nnz_r : nnz; }      // entries of the CSR the engine walks
  Workspace* ws = nullptr;
};

#define VT_DISPATCH(vt, ...)                                                        \
  do {                                                                              \
    switch (vt) {                                                                   \
      case VT_U16: { constexpr int VT = VT_U16; __VA_ARGS__; } break;               \
      case VT_ONE: { constexpr int VT = VT_ONE; __VA_ARGS__; } break;               \
      default: { constexpr int VT = VT_F64; __VA_ARGS__; } break;                   \
    }                                                                               \
  } while (0)

template <class T> static T* dalloc(size_t n) {
  T* p = nullptr;
  CK(cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)));
  return p;
}

// Small device scratch (flags, per-column vectors, scan temporaries, staging of the host-side all-reduces): cudaMalloc / cudaFree
// calls in the load path showed up as 10-400 ms spikes of the step time, so these buffers come from a pool of retired blocks
// (power-of-two size classes, at most 512 MB held; tmc_cache_clear() releases them).
namespace pool {
static std::mutex mu;
static std::vector<std::pair<void*, int>> owned;      // every pooled pointer with its size class
static std::vector<void*> free_list[48];
static size_t held = 0;
static int cls_of_bytes(size_t bytes) { int c = 8; while (((size_t)1 << c) < bytes) ++c; return c; }
static void* take(size_t bytes) {
  const int c = cls_of_bytes(std::max<size_t>(bytes, 1));
  {
    std::lock_guard<std::mutex> lk(mu);
    if (!free_list[c].empty()) { void* p = free_list[c].back(); free_list[c].pop_back(); held -= (size_t)1 << c; return p; }
  }
  void* p = nullptr;
  CK(cudaMalloc(&p, (size_t)1 << c));
  std::lock_guard<std::mutex> lk(mu);
  owned.push_back({p, c});
  return p;
}
static void give(void* p) {
  if (!p) return;
  std::unique_lock<std::mutex> lk(mu);
  for (size_t i = 0; i < owned.size(); ++i)
    if (owned[i].first == p) {
      const int c = owned[i].second;
      if (held + ((size_t)1 << c) <= ((size_t)512 << 20)) { free_list[c].push_back(p); held += (size_t)1 << c; return; }
      owned[i] = owned.back(); owned.pop_back();
      break;
    }
  lk.unlock();
  cudaFree(p);
}
static void release_all() {
  std::lock_guard<std::mutex> lk(mu);
  for (auto& fl : free_list) { for (void* p : fl) cudaFree(p); fl.clear(); }
  // blocks still handed out stay registered; forget the freed ones
  std::vector<std::pair<void*, int>> keep;
  held = 0;
  owned.swap(keep);
}
}  // namespace pool
template <class T> static T* palloc(size_t n) { return static_cast<T*>(pool::take(std::max<size_t>(n, 1) * sizeof(T))); }
static void pfree(void* p) { pool::give(p); }

struct Workspace {
  Ctx c{};
  int K = 0, max_active = 0, use_t = 0, tcap = 0, device = 0, vt = 0;
  int32_t* rowid_scratch = nullptr;
  int64_t m = 0, cap_m = 0, cap_nnz = 0, cap_cols = 0, tn = 0;
  Report* h_rep = nullptr; Activation* h_act = nullptr; Activation* d_act = nullptr;
  int* h_err = nullptr; unsigned long long* h_stats = nullptr;
  cudaStream_t stream = nullptr;
  void* cub_tmp = nullptr; size_t cub_tmp_bytes = 0;
  std::vector<void*> allocs;
  template <class T> T* A(size_t n) { T* p = dalloc<T>(n); allocs.push_back(p); return p; }
  ~Workspace() {
    for (void* p : allocs) cudaFree(p);
    if (cub_tmp) cudaFree(cub_tmp);
    if (h_rep) cudaFreeHost(h_rep);
    if (h_act) cudaFreeHost(h_act);
    if (h_err) cudaFreeHost(h_err);
    if (h_stats) cudaFreeHost(h_stats);
    if (stream) cudaStreamDestroy(stream);
  }
};

static int pick_device(int device) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) throw TmcError(TMC_ERR_NO_DEVICE, "no CUDA device: libtmc has no CPU fallback");
  if (device < 0) { if (cudaGetDevice(&device) != cudaSuccess) device = 0; }
  if (device >= n) throw TmcError(TMC_ERR_INVALID, "device ordinal out of range");
  CK(cudaSetDevice(device));
  return device;
}

// One retired workspace is kept for re-use: cudaMalloc / cudaFree of the multi-GB arenas costs far more than a
// whole tree build on c2 (measured: ~280 ms of a 560 ms step), and a host that clusters matrix after matrix
// (or bench.py's steps) should not pay it every call.  tmc_cache_clear() releases it.
static std::mutex g_cache_mu;
static Workspace* g_ws_cache = nullptr;

static void bind_workspace(Workspace* w, tmc_matrix* b) {
  Ctx& c = w->c;
  c.row_ptr = b->rp(); c.col = b->ci(); c.val = (b->vt == VT_F64) ? (const void*)b->val : (const void*)b->valc;
  c.vt = b->vt; c.rrow = b->rrow; c.w2col = b->w2col; c.invw = b->invw;
  c.panel = b->panel; c.P = b->P; c.Ppad = b->Ppad; c.pnib = b->pnib; c.pw = b->pw; c.row_ptr0 = b->row_ptr;
  c.n_rows = (int)b->n_rows; c.n_cols = (int)b->n_cols; c.row_offset = b->row_offset;
  c.signed_mode = b->has_neg ? 1 : 0;
  const int64_t mpos = b->replicated ? (b->own_hi - b->own_lo) + b->n_rows : b->n_rows;
  c.vstride = (mpos + 15) / 16 * 16;
  c.ystride = (b->n_cols + 511) / 512 * 512;        // >= Ppad: the panel kernels read y in whole 512-column steps
  w->m = mpos;
}

static Workspace* get_workspace(tmc_matrix* b, const tmc_opts& o) {
  // positions: own rows, plus (replicated mode) room for every row that may be handed to this rank as a private subtree
  const int64_t m = b->replicated ? (b->own_hi - b->own_lo) + b->n_rows : b->n_rows;
  int K = std::max(2, (int)std::min<int64_t>(o.max_lanczos > 0 ? o.max_lanczos : 320, b->n_rows_global));
  int max_active = (int)std::min<int64_t>(std::max<int64_t>(1, b->n_rows_global / 2), o.max_active > 0 ? o.max_active : 2048);
  max_active = std::min(max_active, 16384);      // k_prepare scans up to 16 slots per thread
  {     // the per-slot feature vectors (Y arena) stay under 8 GB: wide matrices get fewer slots
    const int64_t ystride = (b->n_cols + 511) / 512 * 512;
    max_active = (int)std::max<int64_t>(std::min<int64_t>(max_active, 64), std::min<int64_t>(max_active, ((int64_t)8 << 30) / (ystride * 8)));
  }
  if (b->replicated && g_nccl.world > 1) max_active = std::max(max_active, 2);      // one shared + one private slot at least
  const int use_t = (o.scatter_mode == 0);
  if (b->ws && b->ws->K == K && b->ws->max_active == max_active && b->ws->use_t == use_t && b->ws->vt == b->vt) return b->ws;
  delete b->ws; b->ws = nullptr;
  {
    std::lock_guard<std::mutex> lk(g_cache_mu);
    Workspace* q = g_ws_cache;
    if (q && q->device == b->device && q->K == K && q->max_active == max_active && q->use_t == use_t && q->vt == b->vt && q->cap_m >= m &&
        q->cap_nnz >= b->enz() && q->cap_cols >= b->n_cols && (!use_t || q->tn >= b->enz() + (b->replicated ? b->enz() / std::max(1, g_nccl.world) + b->enz() / 16 + 65536 : 0) + 4096)) {
      g_ws_cache = nullptr;
      bind_workspace(q, b);
      b->ws = q;
      return q;
    }
  }
  Workspace* w = new Workspace();
  try {
    w->K = K; w->max_active = max_active; w->use_t = use_t; w->device = b->device; w->vt = b->vt;
    w->cap_m = m; w->cap_nnz = b->enz(); w->cap_cols = b->n_cols; w->tn = 0;
    CK(cudaStreamCreateWithFlags(&w->stream, cudaStreamNonBlocking));
    Ctx& c = w->c;
    bind_workspace(w, b);
    c.perm = w->A<int>(m); c.perm_tmp = w->A<int>(m);
    c.d = w->A<double>(m); c.sdeg = w->A<double>(m); c.xin = w->A<double>(m); c.w = w->A<double>(m); c.u2 = w->A<double>(m);
    c.label = w->A<uint8_t>(m);
    c.pacc = w->A<double>(m);
    c.d2 = w->A<double>(m);
    c.V = w->A<double>((size_t)c.vstride * (K + 2));
    c.Y = w->A<double>((size_t)c.ystride * max_active);
    c.st = w->A<Slot>(max_active); c.rep = w->A<Report>(max_active); c.max_active = max_active;
    c.redA = w->A<double>((size_t)(K + 2) * max_active); c.redB = w->A<double>((size_t)(K + 2 + R_NSCAL) * max_active);
    c.alpha = w->A<double>((size_t)(K + 2) * max_active); c.beta = w->A<double>((size_t)(K + 2) * max_active);
    c.zvec = w->A<double>((size_t)(K + 2) * max_active); c.scratch = w->A<double>((size_t)2 * (K + 2) * max_active);
    c.K = K;
    c.list_offs = w->A<int>(4 * ((size_t)max_active + 1));
    c.chunks = w->A<Chunk>((size_t)m / 8 + max_active + 8);
    c.n_chunks = w->A<int>(1);
    c.vchunks = w->A<Chunk>((size_t)m / 32 + max_active + 8);
    c.n_vchunks = w->A<int>(1);
    c.gchunks = w->A<Chunk>((size_t)m / 256 + max_active + 8);      // k_gather_smem: at least 256 positions per chunk
    c.n_gchunks = w->A<int>(1);
    c.err = w->A<int>(1);
    c.stats = w->A<unsigned long long>(4);
    c.certmin = w->A<unsigned long long>(max_active);
    c.ws = w->A<double>(m); c.ws_tmp = w->A<double>(m); c.zvec2 = w->A<double>((size_t)(K + 2) * max_active);
    c.km = w->A<double>((size_t)KM_N * max_active);
    w->d_act = w->A<Activation>(max_active);
    c.use_t = use_t;
    if (c.use_t) {      // transposed per-node blocks, double buffered (children are written to the other buffer)
      const int64_t own_nnz_est = b->replicated ? b->enz() / std::max(1, g_nccl.world) + b->enz() / 16 + 65536 : 0;
      const size_t tn = (size_t)b->enz() + (size_t)own_nnz_est + 4096;      // replicated: own block + every handed-off block
      const size_t vsz = b->vt == VT_F64 ? 8 : (b->vt == VT_U16 ? 2 : 0);
      w->tn = (int64_t)tn;
      for (int q = 0; q < 2; ++q) { c.tcol[q] = w->A<int32_t>(tn); c.tpos[q] = w->A<int32_t>(tn); c.tval[q] = w->A<char>(tn * vsz + 64); }
      w->rowid_scratch = w->A<int32_t>(tn);
      w->tcap = 32768 + 2 * max_active + 64;      // every slot can add up to two partial chunks
      c.tchunks = w->A<TChunk>(w->tcap); c.n_tchunks = w->A<int>(1);
      c.tile_cnt = w->A<int>(w->tcap); c.tile_off = w->A<long long>(w->tcap);
      c.newpos = w->A<int>(m);
      c.tb_cnt = w->A<int>(b->n_cols + 1); c.tb_colstart = w->A<long long>(b->n_cols + 2);
    }
    CK(cudaMemset(c.redB, 0, sizeof(double) * (size_t)(K + 2 + R_NSCAL) * max_active));
    CK(cudaMallocHost(&w->h_rep, sizeof(Report) * max_active));
    CK(cudaMallocHost(&w->h_act, sizeof(Activation) * max_active));
    CK(cudaMallocHost(&w->h_err, sizeof(int)));
    CK(cudaMallocHost(&w->h_stats, 4 * sizeof(unsigned long long)));
  } catch (...) { delete w; throw; }
  b->ws = w;
  return w;
}

static void retire_workspace(Workspace* w) {
  if (!w) return;
  std::lock_guard<std::mutex> lk(g_cache_mu);
  if (!g_ws_cache || (g_ws_cache->cap_nnz <= w->cap_nnz && g_ws_cache->cap_m <= w->cap_m)) { delete g_ws_cache; g_ws_cache = w; }
  else delete w;
}

static int err_to_status(int e, std::string* msg) {
  switch (e) {
    case 0: return TMC_OK;
    case 1: *msg = "column index out of range"; return TMC_ERR_INVALID;
    case 4: *msg = "zero-norm row (reference: NaN row -> SVDLIBC hang, LoadMatrix.hs:189-193)"; return TMC_ERR_ZERO_ROW;
    case 5: *msg = "NaN/Inf in matrix values"; return TMC_ERR_NAN;
    case 6: *msg = "negative matrix entry (engine requires B >= 0)"; return TMC_ERR_NEGATIVE;
    case 7: *msg = "zero / non-finite degree in a node"; return TMC_ERR_ZERO_ROW;
    case 9: *msg = "peer-memory all-reduce timed out waiting for a rank (TMC_P2P)"; return TMC_ERR_COMM;
    default: *msg = "device error flag " + std::to_string(e); return TMC_ERR_CUDA;
  }
}

// load-time collectives are needed when this rank holds (or is preparing) only a row block of the matrix
static bool sharded_coll(const tmc_matrix* b) { return g_nccl.comm && g_nccl.world > 1 && (!b->replicated || b->own_prepare); }
static int64_t blk_row(const tmc_matrix* b, int r) { return b->n_rows * r / g_nccl.world; }
// all-gather by one broadcast per owner: rank r owns bytes [off(r), off(r + 1)) of `base` (same buffer layout on every rank)
template <class OFF> static void gather_blocks(void* base, OFF off, const char* what) {
  for (int r = 0; r < g_nccl.world; ++r) {
    const int64_t o0 = off(r), o1 = off(r + 1);
    if (o1 <= o0) continue;
    char* p = reinterpret_cast<char*>(base) + o0;
    if (g_nccl.Broadcast(p, p, (size_t)(o1 - o0), NCCL_UINT8, r, g_nccl.comm, 0) != 0)
      throw TmcError(TMC_ERR_COMM, std::string("ncclBroadcast(") + what + ") failed");
  }
}
// replicated mode: fetch the raw blocks of the other ranks (only the fallbacks that walk the caller's CSR need them)
static void ensure_raw_complete(tmc_matrix* b) {
  if (b->raw_complete) return;
  gather_blocks(b->col, [&](int r) { return b->blk_e[r] * (int64_t)sizeof(int32_t); }, "raw column indices");
  gather_blocks(b->val, [&](int r) { return b->blk_e[r] * (int64_t)sizeof(double); }, "raw values");
  CK(cudaDeviceSynchronize());
  b->raw_complete = true;
}

// Row-block inputs under a communicator: a device error code raised on one rank (zero-norm row, column out of range, ...)
// must be seen by every rank, or the others would go on and wait forever in the first collective of the tree build.
static int agree_error(const tmc_matrix* b, int e) {
  if (!sharded_coll(b)) return e;
  int* d = palloc<int>(1);
  cudaMemcpy(d, &e, sizeof(int), cudaMemcpyHostToDevice);
  const int r = g_nccl.AllReduce(d, d, 1, NCCL_INT32, NCCL_MAX, g_nccl.comm, 0);
  cudaMemcpy(&e, d, sizeof(int), cudaMemcpyDeviceToHost); pfree(d);
  if (r != 0) throw TmcError(TMC_ERR_COMM, "ncclAllReduce(error code) failed");
  return e;
}

// normalise the device values in place: optional idf (b1ToB2), then row L2 (b2ToB)
static void normalise_dev(tmc_matrix* b, int normalize, double* norm_out_dev, double* idf_out_dev, bool do_l2) {
  int* d_err = palloc<int>(1);
  int* df = nullptr; double* idf = nullptr;
  try {
    CK(cudaMemset(d_err, 0, sizeof(int)));
    const int grid = 148 * 8;
    if (normalize) {
      df = palloc<int>(b->n_cols); idf = idf_out_dev ? idf_out_dev : palloc<double>(b->n_cols);
      CK(cudaMemset(df, 0, sizeof(int) * b->n_cols));
      k_df_count<<<grid, 256>>>(b->col, b->val, b->nnz, (int)b->n_cols, df, d_err);
      if (sharded_coll(b)) {      // df over all ranks' row blocks
        int r = g_nccl.AllReduce(df, df, (size_t)b->n_cols, NCCL_INT32, NCCL_SUM, g_nccl.comm, 0);
        if (r != 0) throw TmcError(TMC_ERR_COMM, "ncclAllReduce(df) failed");
      }
      k_idf<<<(int)((b->n_cols + 255) / 256), 256>>>(df, (int)b->n_cols, (double)b->n_rows_global, idf);
      k_idf_scale<<<grid, 256>>>(b->col, b->val, b->nnz, idf);
    }
    if (do_l2) k_row_l2<<<grid, 256>>>(b->row_ptr, b->col, b->val, (int)b->n_rows, (int)b->n_cols, norm_out_dev, d_err);
    CK(cudaGetLastError());
    int e = 0;
    CK(cudaMemcpy(&e, d_err, sizeof(int), cudaMemcpyDeviceToHost));
    pfree(d_err); d_err = nullptr;
    if (df) pfree(df);
    if (idf && !idf_out_dev) pfree(idf);
    df = nullptr; idf = nullptr;
    e = agree_error(b, e);
    std::string msg;
    int st = err_to_status(e, &msg);
    if (st != TMC_OK) throw TmcError(st, msg);
  } catch (...) {
    if (d_err) pfree(d_err);
    if (df) pfree(df);
    if (idf && !idf_out_dev) pfree(idf);
    throw;
  }
}

// Retired factor buffers of the last matrix (same reason as the workspace cache: cudaFree of GB-sized buffers takes
// hundreds of ms).  Released by tmc_cache_clear().
struct FactorBufs { void* valc = nullptr; double* rrow = nullptr; double* w2col = nullptr; double* invw = nullptr;
                    size_t cap_nnz = 0, cap_rows = 0, cap_cols = 0; int device = -1; };
static FactorBufs g_fb_cache;
static void factor_bufs_release(FactorBufs& f) {
  if (f.valc) cudaFree(f.valc);
  if (f.rrow) cudaFree(f.rrow);
  if (f.w2col) cudaFree(f.w2col);
  if (f.invw) cudaFree(f.invw);
  f = FactorBufs();
}
static void factor_bufs_acquire(tmc_matrix* b, bool need_valc) {
  std::lock_guard<std::mutex> lk(g_cache_mu);
  FactorBufs& f = g_fb_cache;
  const bool fits = f.device == b->device && f.cap_rows >= (size_t)b->n_rows && f.cap_cols >= (size_t)b->n_cols &&
                    (!need_valc || (f.valc && f.cap_nnz >= (size_t)b->nnz));
  if (fits) {
    b->rrow = f.rrow; b->w2col = f.w2col; b->invw = f.invw; b->valc = need_valc ? f.valc : nullptr;
    b->fb_cap_nnz = f.cap_nnz; b->fb_valc_spare = need_valc ? nullptr : f.valc;
    f = FactorBufs();
    return;
  }
  b->rrow = dalloc<double>(b->n_rows); b->w2col = dalloc<double>(b->n_cols); b->invw = dalloc<double>(b->n_cols);
  if (need_valc) { b->valc = dalloc<unsigned short>(b->nnz + 64); b->fb_cap_nnz = b->nnz; }
}
static void factor_bufs_retire(tmc_matrix* b) {
  if (!b->rrow) return;
  std::lock_guard<std::mutex> lk(g_cache_mu);
  factor_bufs_release(g_fb_cache);
  g_fb_cache.valc = b->valc ? b->valc : b->fb_valc_spare; g_fb_cache.rrow = b->rrow; g_fb_cache.w2col = b->w2col; g_fb_cache.invw = b->invw;
  g_fb_cache.cap_nnz = g_fb_cache.valc ? b->fb_cap_nnz : 0; g_fb_cache.cap_rows = b->n_rows; g_fb_cache.cap_cols = b->n_cols;
  g_fb_cache.device = b->device;
  b->valc = nullptr; b->rrow = nullptr; b->w2col = nullptr; b->invw = nullptr; b->fb_valc_spare = nullptr;
}

// Panel buffers of the last matrix, kept for the same reason.
struct PanelBufs { uint8_t* panel = nullptr; int64_t* row_ptr_r = nullptr; int32_t* col_r = nullptr;
                   size_t cap_panel = 0, cap_rp = 0, cap_colr = 0; int device = -1; };
static PanelBufs g_pb_cache;
static void panel_bufs_release(PanelBufs& f) {
  if (f.panel) cudaFree(f.panel);
  if (f.row_ptr_r) cudaFree(f.row_ptr_r);
  if (f.col_r) cudaFree(f.col_r);
  f = PanelBufs();
}
static void panel_bufs_retire(tmc_matrix* b) {
  if (!b->panel && !b->row_ptr_r && !b->col_r) return;
  std::lock_guard<std::mutex> lk(g_cache_mu);
  panel_bufs_release(g_pb_cache);
  g_pb_cache.panel = b->panel; g_pb_cache.row_ptr_r = b->row_ptr_r; g_pb_cache.col_r = b->col_r;
  g_pb_cache.cap_panel = b->cap_panel; g_pb_cache.cap_rp = b->cap_rp; g_pb_cache.cap_colr = b->cap_colr; g_pb_cache.device = b->device;
  b->panel = nullptr; b->row_ptr_r = nullptr; b->col_r = nullptr; b->P = 0; b->Ppad = 0; b->pnib = 0; b->pw = 0;
}
// take what fits from the cache, allocate the rest
static void panel_bufs_take(tmc_matrix* b, size_t need_panel, size_t need_rp, size_t need_colr) {
  std::lock_guard<std::mutex> lk(g_cache_mu);
  PanelBufs& f = g_pb_cache;
  if (f.device != b->device) panel_bufs_release(f);
  if (need_panel && !b->panel) {
    if (f.panel && f.cap_panel >= need_panel) { b->panel = f.panel; b->cap_panel = f.cap_panel; f.panel = nullptr; f.cap_panel = 0; }
    else { if (f.panel) { cudaFree(f.panel); f.panel = nullptr; f.cap_panel = 0; } b->panel = dalloc<uint8_t>(need_panel); b->cap_panel = need_panel; }
  }
  if (need_rp && !b->row_ptr_r) {
    if (f.row_ptr_r && f.cap_rp >= need_rp) { b->row_ptr_r = f.row_ptr_r; b->cap_rp = f.cap_rp; f.row_ptr_r = nullptr; f.cap_rp = 0; }
    else { if (f.row_ptr_r) { cudaFree(f.row_ptr_r); f.row_ptr_r = nullptr; f.cap_rp = 0; } b->row_ptr_r = dalloc<int64_t>(need_rp); b->cap_rp = need_rp; }
  }
  if (need_colr && !b->col_r) {
    if (f.col_r && f.cap_colr >= need_colr) { b->col_r = f.col_r; b->cap_colr = f.cap_colr; f.col_r = nullptr; f.cap_colr = 0; }
    else { if (f.col_r) { cudaFree(f.col_r); f.col_r = nullptr; f.cap_colr = 0; } b->col_r = dalloc<int32_t>(need_colr); b->cap_colr = need_colr; }
  }
  f.device = b->device;
}

// The device copies of the caller's host CSR (tmc_matrix_upload / tmc_make_tree): 20 GB for c3, allocated and freed once per call
// -- cudaFree of buffers that size takes 100-400 ms and made the end-to-end time jump between 1.12 and 1.55 s.  One retired set
// is kept (the larger of each array), released by tmc_cache_clear().
struct RawBufs { void* p[3] = {nullptr, nullptr, nullptr}; size_t cap[3] = {0, 0, 0}; int device = -1; };
static RawBufs g_raw_cache;
static void raw_bufs_release() { for (int i = 0; i < 3; ++i) { if (g_raw_cache.p[i]) cudaFree(g_raw_cache.p[i]); } g_raw_cache = RawBufs(); }
template <class T> static T* raw_take(int which, size_t n, int device, size_t* cap_out) {
  const size_t need = std::max<size_t>(n, 1) * sizeof(T);
  {
    std::lock_guard<std::mutex> lk(g_cache_mu);
    if (g_raw_cache.device == device && g_raw_cache.p[which] && g_raw_cache.cap[which] >= need) {
      T* p = static_cast<T*>(g_raw_cache.p[which]);
      *cap_out = g_raw_cache.cap[which];
      g_raw_cache.p[which] = nullptr; g_raw_cache.cap[which] = 0;
      return p;
    }
  }
  *cap_out = need;
  return dalloc<T>(n);
}
static void raw_retire(int which, void* p, size_t cap, int device) {
  if (!p) return;
  std::lock_guard<std::mutex> lk(g_cache_mu);
  if (g_raw_cache.device != device) { for (int i = 0; i < 3; ++i) { if (g_raw_cache.p[i]) cudaFree(g_raw_cache.p[i]); } g_raw_cache = RawBufs(); g_raw_cache.device = device; }
  if (!g_raw_cache.p[which] || g_raw_cache.cap[which] < cap) {
    if (g_raw_cache.p[which]) cudaFree(g_raw_cache.p[which]);
    g_raw_cache.p[which] = p; g_raw_cache.cap[which] = cap;
  } else cudaFree(p);
}

// Dense panel (DESIGN.md "dense panel"): columns present in at least `density` of the cells (TMC_PANEL_DENSITY, default 0.07)
// are renumbered to [0, P) and stored as one byte per (row, column); the other entries (and counts > 255) form the residual
// CSR, written to new arrays -- the caller's arrays are never modified.  On return (true) b->P/Ppad/panel/row_ptr_r/col_r/nnz_r
// are set, valc holds the residual values and w2col/invw follow the new column order.  Returns false when no panel pays off.
static bool build_panel(tmc_matrix* b, const int* df_dev, const double* idf_dev, int* d_flags, int nib) {
  // rows this rank prepares: all of them, or (replicated mode, own_prepare) its own block -- the blocks are gathered below
  const int64_t rl = b->own_prepare ? b->own_lo : 0, rh = b->own_prepare ? b->own_hi : b->n_rows;
  const int nr = (int)(rh - rl);
  if (getenv("TMC_NO_PANEL") || b->n_cols < 64) return false;      // rank-agreed conditions only: a collective follows
  const double density = getenv("TMC_PANEL_DENSITY") ? atof(getenv("TMC_PANEL_DENSITY")) : 0.07;
  if (!(density > 0.0)) return false;
  const int n = (int)b->n_cols;
  std::vector<int> df(n);
  CK(cudaMemcpy(df.data(), df_dev, sizeof(int) * n, cudaMemcpyDeviceToHost));
  const double thr = std::max(1.0, density * (double)b->n_rows_global);
  std::vector<int> cand;
  long long panel_nnz = 0, tot_nnz = 0;
  for (int j = 0; j < n; ++j) { tot_nnz += df[j]; if ((double)df[j] >= thr) cand.push_back(j); }
  // memory bound: the panel may take at most a quarter of the free memory; keep the most frequent columns
  size_t free_b = 0, total_b = 0;
  CK(cudaMemGetInfo(&free_b, &total_b));
  {
    std::lock_guard<std::mutex> lk(g_cache_mu);
    if (g_pb_cache.device == b->device) free_b += g_pb_cache.cap_panel;
  }
  const size_t max_cols_local = std::min<size_t>(65536, (free_b / 4) * (nib ? 2 : 1) / (size_t)std::max<int64_t>(1, b->n_rows) / 512 * 512);
  size_t max_cols = max_cols_local;
  if (g_nccl.comm && g_nccl.world > 1) {          // every rank must pick the same columns
    int* d = palloc<int>(1); int h = (int)max_cols_local;
    cudaMemcpy(d, &h, sizeof(int), cudaMemcpyHostToDevice);
    int r = g_nccl.AllReduce(d, d, 1, NCCL_INT32, NCCL_MIN, g_nccl.comm, 0);
    cudaMemcpy(&h, d, sizeof(int), cudaMemcpyDeviceToHost); pfree(d);
    if (r != 0) throw TmcError(TMC_ERR_COMM, "ncclAllReduce(panel width) failed");
    max_cols = (size_t)h;
  }
  if (cand.size() > max_cols) {
    std::sort(cand.begin(), cand.end(), [&](int a, int c2) { return df[a] != df[c2] ? df[a] > df[c2] : a < c2; });
    cand.resize(max_cols);
    std::sort(cand.begin(), cand.end());
  }
  for (int j : cand) panel_nnz += df[j];
  const int P = (int)cand.size();
  // worth it only when a real share of the entries moves (df counts are global, so every rank decides alike)
  if (P < 8 || (double)panel_nnz < 0.10 * (double)tot_nnz) return false;
  const int Ppad = (P + 511) / 512 * 512;
  const int pw = nib ? Ppad / 2 : Ppad;
  const double cap = nib ? 15.0 : 255.0;
  std::vector<int> old2new(n), new2old(n);
  {
    std::vector<char> in(n, 0);
    for (int j : cand) in[j] = 1;
    int a = 0, r = P;
    for (int j = 0; j < n; ++j) { const int id = in[j] ? a++ : r++; old2new[j] = id; new2old[id] = j; }
  }
  int* maps = palloc<int>(2 * (size_t)n);
  long long* cnt = nullptr; void* tmp = nullptr;
  try {
    CK(cudaMemcpy(maps, old2new.data(), sizeof(int) * n, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(maps + n, new2old.data(), sizeof(int) * n, cudaMemcpyHostToDevice));
    panel_bufs_take(b, 0, (size_t)b->n_rows + 1, 0);
    cnt = reinterpret_cast<long long*>(b->row_ptr_r);
    const int grid = 148 * 8;
    k_panel_count<<<grid, 256>>>(b->row_ptr + rl, b->col, b->val, nr, maps, P, cap, cnt + rl);
    if (b->own_prepare) {       // residual entries per row of every block, then the closing zero of the scan
      gather_blocks(cnt, [&](int r) { return blk_row(b, r) * (int64_t)sizeof(long long); }, "residual row counts");
      CK(cudaMemset(cnt + b->n_rows, 0, sizeof(long long)));
    }
    size_t bytes = 0;
    CK(cub::DeviceScan::ExclusiveSum(nullptr, bytes, cnt, cnt, (int)b->n_rows + 1));
    tmp = pool::take(std::max<size_t>(bytes, 16));
    CK(cub::DeviceScan::ExclusiveSum(tmp, bytes, cnt, cnt, (int)b->n_rows + 1));
    long long nnz_r = 0;
    CK(cudaMemcpy(&nnz_r, cnt + b->n_rows, sizeof(long long), cudaMemcpyDeviceToHost));
    panel_bufs_take(b, (size_t)b->n_rows * pw, 0, (size_t)nnz_r + 64);
    CK(cudaMemsetAsync(b->panel + (size_t)rl * pw, 0, (size_t)nr * pw));
    k_panel_fill<<<grid, 256>>>(b->row_ptr + rl, b->col, b->val, nr, maps, P, pw, nib, cap, cnt + rl, b->col_r,
                                b->vt == VT_U16 ? reinterpret_cast<unsigned short*>(b->valc) : nullptr, b->panel + (size_t)rl * pw, d_flags);
    k_col_factors_perm<<<(n + 255) / 256, 256>>>(idf_dev, maps + n, n, b->w2col, b->invw);
    CK(cudaGetLastError());
    int f = 0;
    CK(cudaMemcpy(&f, d_flags, sizeof(int), cudaMemcpyDeviceToHost));
    pfree(maps); maps = nullptr; pfree(tmp); tmp = nullptr;
    if (g_nccl.comm && g_nccl.world > 1) {
      int* d = palloc<int>(1); int h = (f & 32) ? 1 : 0;
      cudaMemcpy(d, &h, sizeof(int), cudaMemcpyHostToDevice);
      int r = g_nccl.AllReduce(d, d, 1, NCCL_INT32, NCCL_SUM, g_nccl.comm, 0);
      cudaMemcpy(&h, d, sizeof(int), cudaMemcpyDeviceToHost); pfree(d);
      if (r != 0) throw TmcError(TMC_ERR_COMM, "ncclAllReduce(panel flag) failed");
      if (h) f |= 32;
    }
    if (f & 32) return false;           // duplicated (row, column) entries: the plain CSR path adds them up, a byte cannot
    if (b->own_prepare) {
      // exchange the prepared blocks: panel rows and the residual entries (the scan above gave every rank the global offsets)
      std::vector<long long> rb(g_nccl.world + 1);
      for (int r = 0; r <= g_nccl.world; ++r) CK(cudaMemcpy(&rb[r], cnt + blk_row(b, r), sizeof(long long), cudaMemcpyDeviceToHost));
      gather_blocks(b->panel, [&](int r) { return blk_row(b, r) * (int64_t)pw; }, "panel rows");
      gather_blocks(b->col_r, [&](int r) { return (int64_t)rb[r] * (int64_t)sizeof(int32_t); }, "residual columns");
      if (b->vt == VT_U16) gather_blocks(b->valc, [&](int r) { return (int64_t)rb[r] * 2; }, "residual values");
    }
    b->P = P; b->Ppad = Ppad; b->pnib = nib; b->pw = pw; b->nnz_r = nnz_r;
    return true;
  } catch (...) { if (maps) pfree(maps); if (tmp) pfree(tmp); throw; }
}

// Decide the storage mode and compute what the engine needs from the device-resident input values.
//  * all values are integers in [0, 65535] (raw counts; the usual make-tree input when the engine applies tf-idf itself):
//    keep A as uint16 (or nothing when every value is 1) and the factors w_j = idf_j (or 1), r_i = 1/||(A diag w)_i||;
//    the caller's fp64 array is left untouched.
//  * otherwise: materialise B in place as before (idf scaling if asked, row L2).
// EXPERIMENT, off by default (TMC_RENUMBER=1): renumber the columns of the device CSR in place by descending global
// frequency.  It helped a non-pipelined prototype (+10 %) but measured 4 % SLOWER in the product kernels (rows lose
// their ascending column order, so neighbouring lanes no longer hit neighbouring sectors of y).
static void renumber_columns(tmc_matrix* b) {
  if (!getenv("TMC_RENUMBER") || b->nnz == 0) return;
  const int n = (int)b->n_cols;
  int* buf = dalloc<int>(4 * (size_t)n + 2);
  void* tmp = nullptr;
  try {
    int* cnt = buf; int* cnt_out = buf + n; int* ids = buf + 2 * n; int* ids_out = buf + 3 * n; int* flags = buf + 4 * n;
    CK(cudaMemset(buf, 0, sizeof(int) * (4 * (size_t)n + 2)));
    k_col_hist<<<148 * 8, 256>>>(b->col, b->nnz, n, cnt, flags);
    if (g_nccl.comm && g_nccl.world > 1 && !b->replicated) {
      int r = g_nccl.AllReduce(cnt, cnt, (size_t)n, NCCL_INT32, NCCL_SUM, g_nccl.comm, 0);
      if (r != 0) throw TmcError(TMC_ERR_COMM, "ncclAllReduce(column counts) failed");
    }
    int f = 0; CK(cudaMemcpy(&f, flags, sizeof(int), cudaMemcpyDeviceToHost));
    if (f) throw TmcError(TMC_ERR_INVALID, "column index out of range");
    k_iota_i32<<<(n + 255) / 256, 256>>>(ids, n);
    size_t bytes = 0;
    CK(cub::DeviceRadixSort::SortPairsDescending(nullptr, bytes, cnt, cnt_out, ids, ids_out, n));
    CK(cudaMalloc(&tmp, std::max<size_t>(bytes, 16)));
    CK(cub::DeviceRadixSort::SortPairsDescending(tmp, bytes, cnt, cnt_out, ids, ids_out, n));
    k_rank_from_order<<<(n + 255) / 256, 256>>>(ids_out, n, ids);          // ids := rank of every old column id
    k_remap_cols<<<148 * 8, 256>>>(b->col, b->nnz, ids);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
  } catch (...) { cudaFree(buf); if (tmp) cudaFree(tmp); throw; }
  cudaFree(buf); cudaFree(tmp);
}

static void prepare_matrix(tmc_matrix* b, int normalize, bool do_l2 = true) {
  if (b->n_rows_global < 1) throw TmcError(TMC_ERR_INVALID, "empty matrix (reference: error, LoadMatrix.hs:255)");
  if (b->n_cols < 1) throw TmcError(TMC_ERR_ZERO_ROW, "matrix without features: every row has zero norm");
  const int world = g_nccl.comm ? g_nccl.world : 1, rank = g_nccl.comm ? g_nccl.rank : 0;
  // replicated mode: prepare the own row block only and gather the prepared blocks (TMC_REPL_PREPARE=1: the first version, every
  // rank prepares the whole matrix); the fallbacks below (no compact storage / no panel) fetch the raw blocks and do just that
  b->own_prepare = b->replicated && world > 1 && do_l2 && getenv("TMC_REPL_PREPARE") == nullptr && (int)b->blk_e.size() == world + 1;
  if (!b->own_prepare) ensure_raw_complete(b);
  renumber_columns(b);
  const bool stage_dbg = getenv("TMC_DEBUG") && atoi(getenv("TMC_DEBUG")) >= 3;
  double stage_t = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
  auto stage = [&](const char* what) {      // TMC_DEBUG=3: where the load time goes (synchronises: not for timing runs)
    if (!stage_dbg) return;
    cudaDeviceSynchronize();
    const double t = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
    fprintf(stderr, "[tmc]   prepare: %-28s %8.2f ms\n", what, t - stage_t);
    stage_t = t;
  };
  int* d_flags = palloc<int>(6);          // [0] value flags, [1] error code, [2..5] two 64-bit counters of small values
  int* df = nullptr; double* idf = nullptr;
  double* wrec = nullptr; double* counts = nullptr; double* borrowed_val = nullptr;
  auto restore_val = [&]() {        // a temporary count array standing in for a borrowed value array
    if (borrowed_val) { if (b->val && b->val != borrowed_val) cudaFree(b->val); b->val = borrowed_val; borrowed_val = nullptr; }
  };
  try {
    CK(cudaMemset(d_flags, 0, 6 * sizeof(int)));
    const int grid = 148 * 8;
    if (!b->owns) {          // adopted device arrays (a host CSR was checked in validate_csr)
      int bad = 0;
      k_check_row_ptr<<<(unsigned)(b->n_rows / 256 + 1), 256>>>(b->row_ptr, b->n_rows, b->nnz, d_flags);
      CK(cudaMemcpy(&bad, d_flags, sizeof(int), cudaMemcpyDeviceToHost));
      CK(cudaMemset(d_flags, 0, sizeof(int)));
      if (agree_error(b, bad)) throw TmcError(TMC_ERR_INVALID, "row_ptr is not a non-decreasing sequence from 0 to nnz");
    }
    // the block this rank walks: rows [rl, rh), raw entries [e0, e1)
    int64_t rl = 0, rh = b->n_rows, e0 = 0, e1 = b->nnz;
    auto set_range = [&]() {
      rl = b->own_prepare ? b->own_lo : 0; rh = b->own_prepare ? b->own_hi : b->n_rows;
      e0 = b->own_prepare ? b->blk_e[rank] : 0; e1 = b->own_prepare ? b->blk_e[rank + 1] : b->nnz;
    };
    set_range();
    unsigned long long* d_small = reinterpret_cast<unsigned long long*>(d_flags + 2);
    k_scan_values<<<grid, 256>>>(b->col + e0, b->val + e0, e1 - e0, (int)b->n_cols, d_flags, d_small);
    int f[2] = {0, 0};
    unsigned long long n_small[2] = {0, 0};
    CK(cudaMemcpy(f, d_flags, 2 * sizeof(int), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(n_small, d_small, sizeof(n_small), cudaMemcpyDeviceToHost));
    stage("scan values");
    if (sharded_coll(b)) {
      // row-block inputs: every rank must take the same decision (storage mode, errors), or the others would wait forever
      int* bits = palloc<int>(5); int hb[5];
      for (int i = 0; i < 5; ++i) hb[i] = (f[0] >> i) & 1;
      cudaMemcpy(bits, hb, sizeof(hb), cudaMemcpyHostToDevice);
      int r = g_nccl.AllReduce(bits, bits, 5, NCCL_INT32, NCCL_SUM, g_nccl.comm, 0);
      cudaMemcpy(hb, bits, sizeof(hb), cudaMemcpyDeviceToHost); pfree(bits);
      if (r != 0) throw TmcError(TMC_ERR_COMM, "ncclAllReduce(value flags) failed");
      f[0] = 0; for (int i = 0; i < 5; ++i) if (hb[i]) f[0] |= 1 << i;
      std::vector<double> ns = {(double)n_small[0], (double)n_small[1]};     // exact below 2^53
      double* dn = palloc<double>(2);
      cudaMemcpy(dn, ns.data(), 16, cudaMemcpyHostToDevice);
      r = g_nccl.AllReduce(dn, dn, 2, NCCL_FLOAT64, NCCL_SUM, g_nccl.comm, 0);
      cudaMemcpy(ns.data(), dn, 16, cudaMemcpyDeviceToHost); pfree(dn);
      if (r != 0) throw TmcError(TMC_ERR_COMM, "ncclAllReduce(small-value counts) failed");
      n_small[0] = (unsigned long long)ns[0]; n_small[1] = (unsigned long long)ns[1];
    }
    if (f[0] & 1) throw TmcError(TMC_ERR_NAN, "NaN/Inf in matrix values");
    if (f[0] & 4) throw TmcError(TMC_ERR_INVALID, "column index out of range");
    b->has_neg = (f[0] & 2) != 0;        // signed input (the reference takes |B| for the degrees only): see Ctx::signed_mode
    bool small_int = !(f[0] & 8) && getenv("TMC_NO_FACTORED") == nullptr;
    if (!small_int && !b->has_neg && !normalize && b->nnz > 0 && getenv("TMC_NO_FACTORED") == nullptr && getenv("TMC_NO_RECOVER") == nullptr) {
      // B2 = A diag(w) handed over already scaled (Cluster.hs:147 receives the tf-idf matrix): recover A and w if that is exact
      wrec = palloc<double>(b->n_cols);
      k_fill_f64<<<(unsigned)((b->n_cols + 255) / 256), 256>>>(wrec, b->n_cols, INFINITY);
      k_colmin<<<grid, 256>>>(b->col + e0, b->val + e0, e1 - e0, reinterpret_cast<unsigned long long*>(wrec));
      const bool sharded = sharded_coll(b);
      if (sharded && g_nccl.AllReduce(wrec, wrec, (size_t)b->n_cols, NCCL_FLOAT64, NCCL_MIN, g_nccl.comm, 0) != 0)
        throw TmcError(TMC_ERR_COMM, "ncclAllReduce(column minima) failed");
      k_colmin_fix<<<(int)((b->n_cols + 255) / 256), 256>>>(wrec, (int)b->n_cols);
      CK(cudaMemset(d_flags + 2, 0, 4 * sizeof(int)));
      int* d_rf = palloc<int>(1);
      CK(cudaMemset(d_rf, 0, sizeof(int)));
      k_factor_check<<<grid, 256>>>(b->col + e0, b->val + e0, e1 - e0, wrec, nullptr, d_rf, d_small);
      int rf = 0;
      CK(cudaMemcpy(&rf, d_rf, sizeof(int), cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(n_small, d_small, sizeof(n_small), cudaMemcpyDeviceToHost));
      if (sharded) {
        std::vector<double> ns = {(double)n_small[0], (double)n_small[1], (rf & 8) ? 1.0 : 0.0, (rf & 16) ? 1.0 : 0.0};
        double* dn = palloc<double>(4);
        cudaMemcpy(dn, ns.data(), 32, cudaMemcpyHostToDevice);
        int r = g_nccl.AllReduce(dn, dn, 4, NCCL_FLOAT64, NCCL_SUM, g_nccl.comm, 0);
        cudaMemcpy(ns.data(), dn, 32, cudaMemcpyDeviceToHost); pfree(dn);
        if (r != 0) { pfree(d_rf); throw TmcError(TMC_ERR_COMM, "ncclAllReduce(factor check) failed"); }
        n_small[0] = (unsigned long long)ns[0]; n_small[1] = (unsigned long long)ns[1];
        rf = (ns[2] > 0 ? 8 : 0) | (ns[3] > 0 ? 16 : 0);
      }
      if (!(rf & 8)) {
        // exact: replace the values by the counts (in place when the engine owns the array, else in a temporary)
        counts = b->owns ? b->val : dalloc<double>(b->nnz);
        k_factor_check<<<grid, 256>>>(b->col + e0, b->val + e0, e1 - e0, wrec, counts + e0, d_rf, nullptr);
        borrowed_val = b->owns ? nullptr : b->val;
        b->val = counts;
        small_int = true;
        f[0] = (f[0] & ~(8 | 16)) | (rf & 16);
      } else { pfree(wrec); wrec = nullptr; }
      pfree(d_rf);
    }
    if (!small_int) {
      pfree(d_flags); d_flags = nullptr;
      b->vt = VT_F64;
      if (b->own_prepare) { ensure_raw_complete(b); b->own_prepare = false; }      // B itself is materialised: every rank does all of it
      normalise_dev(b, normalize, nullptr, nullptr, do_l2);
      return;
    }
    b->vt = (f[0] & 16) ? VT_U16 : VT_ONE;
    if (wrec) { idf = wrec; wrec = nullptr; }        // the recovered column weights take the place of idf
    const bool want_panel = getenv("TMC_NO_PANEL") == nullptr;
    if (normalize || want_panel) {       // column frequencies: idf, and the choice of the dense panel
      df = palloc<int>(b->n_cols);
      CK(cudaMemset(df, 0, sizeof(int) * b->n_cols));
      k_df_count<<<grid, 256>>>(b->col + e0, b->val + e0, e1 - e0, (int)b->n_cols, df, d_flags + 1);
      if (sharded_coll(b)) {
        int r = g_nccl.AllReduce(df, df, (size_t)b->n_cols, NCCL_INT32, NCCL_SUM, g_nccl.comm, 0);
        if (r != 0) throw TmcError(TMC_ERR_COMM, "ncclAllReduce(df) failed");
      }
    }
    if (normalize) {
      idf = palloc<double>(b->n_cols);
      k_idf<<<(int)((b->n_cols + 255) / 256), 256>>>(df, (int)b->n_cols, (double)b->n_rows_global, idf);
    }
    stage("document frequencies, idf");
    factor_bufs_acquire(b, b->vt == VT_U16);
    stage("factor buffers");
    auto row_factors = [&]() {
      if (do_l2) k_row_factor<<<grid, 256>>>(b->row_ptr + rl, b->col, b->val, idf, (int)(rh - rl), b->rrow + rl, d_flags + 1);
      else k_fill_f64<<<(unsigned)((b->n_rows + 255) / 256), 256>>>(b->rrow, b->n_rows, 1.0);      // truncated SVD of B2 itself: no row factor
    };
    row_factors();
    stage("row factors");
    bool have_panel = false;
    if (want_panel) {
      // cell format: nibbles when (nearly) everything that fits a byte also fits a nibble (UMI counts), else bytes
      int nib = (double)n_small[0] >= 0.9 * (double)n_small[1];
      if (const char* e = getenv("TMC_PANEL_BITS")) nib = atoi(e) == 4;
      have_panel = build_panel(b, df, idf, d_flags, nib);
      if (!have_panel) panel_bufs_retire(b);
    }
    stage("panel + residual");
    if (have_panel && b->own_prepare)
      gather_blocks(b->rrow, [&](int r) { return blk_row(b, r) * (int64_t)sizeof(double); }, "row factors");
    if (!have_panel) {
      if (b->own_prepare) {        // the engine will walk the caller's CSR: fetch the other blocks, finish on every rank alike
        ensure_raw_complete(b); b->own_prepare = false;
        set_range();
        row_factors();
      }
      k_col_factors<<<(int)((b->n_cols + 255) / 256), 256>>>(idf, (int)b->n_cols, b->w2col, b->invw);
      if (b->vt == VT_U16) k_to_u16<<<grid, 256>>>(b->val, reinterpret_cast<unsigned short*>(b->valc), b->nnz);
    }
    CK(cudaGetLastError());
    CK(cudaMemcpy(f, d_flags, 2 * sizeof(int), cudaMemcpyDeviceToHost));
    f[1] = agree_error(b, f[1]);
    b->own_prepare = false;
    stage("gather / finish");
    if (have_panel && b->owns && b->col) { raw_retire(1, b->col, b->raw_cap[1], b->device); b->col = nullptr; }    // only the residual is walked from here on
    if (b->owns && b->val) { raw_retire(2, b->val, b->raw_cap[2], b->device); b->val = nullptr; }       // the fp64 copy is not needed any more
    restore_val();
    pfree(d_flags); d_flags = nullptr;
    if (df) pfree(df);
    if (idf) pfree(idf);
    df = nullptr; idf = nullptr;
    if (f[1]) { std::string msg; int st = err_to_status(f[1], &msg); throw TmcError(st, msg); }
  } catch (...) {
    b->own_prepare = false;
    restore_val();
    if (d_flags) pfree(d_flags);
    if (df) pfree(df);
    if (idf) pfree(idf);
    if (wrec) pfree(wrec);
    throw;
  }
}

// host-side consumers of the values (the entry points that return host arrays) read doubles / int32 only
static void require_default_types(const tmc_csr* a, const char* who) {
  if (a && (a->val_type != TMC_VAL_F64 || a->idx_type != TMC_IDX_I32))
    throw TmcError(TMC_ERR_UNSUPPORTED, std::string(who) + " wants double values and int32 column indices (tmc_csr.val_type / idx_type = 0)");
}
static size_t val_elem_size(int t) { return t == TMC_VAL_F64 ? 8 : t == TMC_VAL_F32 ? 4 : t == TMC_VAL_U16 ? 2 : 1; }

// Host -> device copies of the caller's column indices / values [k0, k0 + cnt), widened on the device to int32 / fp64 when the
// host arrays use the compact element types of tmc_csr (the PCIe link carries the compact bytes; chunks go through two small
// staging buffers on two streams so that the widening of one chunk overlaps the copy of the next).
struct Staging { void* buf[2] = {nullptr, nullptr}; cudaStream_t st[2] = {nullptr, nullptr}; int device = -1; };
static Staging g_staging;
static const size_t kStageBytes = (size_t)64 << 20;
static void staging_release() {
  for (int i = 0; i < 2; ++i) { if (g_staging.buf[i]) cudaFree(g_staging.buf[i]); if (g_staging.st[i]) cudaStreamDestroy(g_staging.st[i]); }
  g_staging = Staging();
}
template <class SRC, class DST>
static void h2d_widen(DST* dst_dev, const SRC* src_host, int64_t cnt, int device) {
  if (cnt <= 0) return;
  std::lock_guard<std::mutex> lk(g_cache_mu);
  if (g_staging.device != device) {
    staging_release();
    for (int i = 0; i < 2; ++i) { CK(cudaMalloc(&g_staging.buf[i], kStageBytes)); CK(cudaStreamCreateWithFlags(&g_staging.st[i], cudaStreamNonBlocking)); }
    g_staging.device = device;
  }
  const int64_t per = (int64_t)(kStageBytes / sizeof(SRC));
  int q = 0;
  for (int64_t o = 0; o < cnt; o += per, q ^= 1) {
    const int64_t c = std::min(per, cnt - o);
    CK(cudaMemcpyAsync(g_staging.buf[q], src_host + o, sizeof(SRC) * c, cudaMemcpyHostToDevice, g_staging.st[q]));
    k_widen<SRC, DST><<<148 * 4, 256, 0, g_staging.st[q]>>>(reinterpret_cast<const SRC*>(g_staging.buf[q]), dst_dev + o, c);
  }
  CK(cudaStreamSynchronize(g_staging.st[0])); CK(cudaStreamSynchronize(g_staging.st[1]));
  CK(cudaGetLastError());
}
static void h2d_cols(int32_t* dst_dev, const tmc_csr* a, int64_t k0, int64_t cnt, int device) {
  if (cnt <= 0) return;
  if (a->idx_type == TMC_IDX_U16) h2d_widen(dst_dev, reinterpret_cast<const uint16_t*>(a->col_idx) + k0, cnt, device);
  else CK(cudaMemcpy(dst_dev, a->col_idx + k0, sizeof(int32_t) * cnt, cudaMemcpyHostToDevice));
}
static void h2d_vals(double* dst_dev, const tmc_csr* a, int64_t k0, int64_t cnt, int device) {
  if (cnt <= 0) return;
  switch (a->val_type) {
    case TMC_VAL_F32: h2d_widen(dst_dev, reinterpret_cast<const float*>(a->val) + k0, cnt, device); break;
    case TMC_VAL_U16: h2d_widen(dst_dev, reinterpret_cast<const uint16_t*>(a->val) + k0, cnt, device); break;
    case TMC_VAL_U8: h2d_widen(dst_dev, reinterpret_cast<const uint8_t*>(a->val) + k0, cnt, device); break;
    default: CK(cudaMemcpy(dst_dev, a->val + k0, sizeof(double) * cnt, cudaMemcpyHostToDevice));
  }
}

static void validate_csr(const tmc_csr* a) {
  if (!a || !a->row_ptr || (a->nnz > 0 && (!a->col_idx || !a->val))) throw TmcError(TMC_ERR_INVALID, "null CSR pointer");
  if (a->val_type < TMC_VAL_F64 || a->val_type > TMC_VAL_U8 || a->idx_type < TMC_IDX_I32 || a->idx_type > TMC_IDX_U16)
    throw TmcError(TMC_ERR_INVALID, "unknown tmc_csr.val_type / idx_type");
  if (a->idx_type == TMC_IDX_U16 && a->n_cols > 65536) throw TmcError(TMC_ERR_INVALID, "uint16 column indices need n_cols <= 65536");
  if (a->n_rows < 0 || a->n_cols < 0 || a->nnz < 0) throw TmcError(TMC_ERR_INVALID, "negative dimension");
  if (a->n_rows >= (int64_t)1 << 31 || a->n_cols >= (int64_t)1 << 31) throw TmcError(TMC_ERR_INVALID, "dimension exceeds int32");
  if (a->row_ptr[0] != 0 || a->row_ptr[a->n_rows] != a->nnz) throw TmcError(TMC_ERR_INVALID, "row_ptr does not span [0, nnz]");
  for (int64_t i = 0; i < a->n_rows; ++i)
    if (a->row_ptr[i + 1] < a->row_ptr[i]) throw TmcError(TMC_ERR_INVALID, "row_ptr is not non-decreasing");
  if (a->n_rows_global != 0 && (a->row_offset < 0 || a->row_offset + a->n_rows > a->n_rows_global))
    throw TmcError(TMC_ERR_INVALID, "row block outside [0, n_rows_global)");
}

// Host CSR -> device.  A full matrix handed to every rank of a communicator is sliced into this rank's
// contiguous row block; a caller-supplied block (n_rows_global > n_rows) is taken as is.
static tmc_matrix* upload(const tmc_csr* a, int device) {
  validate_csr(a);
  tmc_matrix* b = new tmc_matrix();
  try {
    b->device = pick_device(device);
    const int world = g_nccl.comm ? g_nccl.world : 1, rank = g_nccl.comm ? g_nccl.rank : 0;
    const bool is_block = a->n_rows_global != 0 && a->n_rows_global != a->n_rows;
    const bool replicate = !is_block && world > 1 && getenv("TMC_NO_HANDOFF") == nullptr;
    int64_t lo = 0, hi = a->n_rows;
    b->n_rows_global = is_block ? a->n_rows_global : a->n_rows;
    b->row_offset = is_block ? a->row_offset : 0;
    if (!is_block && world > 1) { lo = a->n_rows * rank / world; hi = a->n_rows * (rank + 1) / world; b->row_offset = lo; }
    if (replicate) {
      // every rank uploads only its own row block over PCIe; the blocks are then exchanged over NVLink (ncclBroadcast per
      // owner) so that every rank ends up with the whole matrix
      b->replicated = true; b->own_lo = lo; b->own_hi = hi; b->row_offset = 0;
      b->n_rows = a->n_rows; b->n_cols = a->n_cols; b->nnz = a->nnz; b->owns = true;
      b->row_ptr = raw_take<int64_t>(0, b->n_rows + 1, b->device, &b->raw_cap[0]); b->col = raw_take<int32_t>(1, b->nnz, b->device, &b->raw_cap[1]);
      b->val = raw_take<double>(2, b->nnz, b->device, &b->raw_cap[2]);
      CK(cudaMemcpy(b->row_ptr, a->row_ptr, sizeof(int64_t) * (b->n_rows + 1), cudaMemcpyHostToDevice));
      const int64_t k0 = a->row_ptr[lo], k1 = a->row_ptr[hi];
      h2d_cols(b->col + k0, a, k0, k1 - k0, b->device);
      h2d_vals(b->val + k0, a, k0, k1 - k0, b->device);
      // the other ranks' raw blocks stay where they are: prepare_matrix works on the own block and the ranks exchange the
      // prepared blocks over NVLink (ensure_raw_complete fetches the raw ones only for the fallbacks)
      b->raw_complete = false;
      b->blk_e.resize(world + 1);
      for (int r = 0; r <= world; ++r) b->blk_e[r] = a->row_ptr[a->n_rows * r / world];
      return b;
    }
    const int64_t k0 = a->row_ptr[lo], k1 = a->row_ptr[hi];
    b->n_rows = hi - lo; b->n_cols = a->n_cols; b->nnz = k1 - k0; b->owns = true;
    b->row_ptr = raw_take<int64_t>(0, b->n_rows + 1, b->device, &b->raw_cap[0]); b->col = raw_take<int32_t>(1, b->nnz, b->device, &b->raw_cap[1]);
    b->val = raw_take<double>(2, b->nnz, b->device, &b->raw_cap[2]);
    if (k0 == 0) {
      CK(cudaMemcpy(b->row_ptr, a->row_ptr + lo, sizeof(int64_t) * (b->n_rows + 1), cudaMemcpyHostToDevice));
    } else {
      std::vector<int64_t> rp(b->n_rows + 1);
      for (int64_t i = 0; i <= b->n_rows; ++i) rp[i] = a->row_ptr[lo + i] - k0;
      CK(cudaMemcpy(b->row_ptr, rp.data(), sizeof(int64_t) * (b->n_rows + 1), cudaMemcpyHostToDevice));
    }
    h2d_cols(b->col, a, k0, b->nnz, b->device);
    h2d_vals(b->val, a, k0, b->nnz, b->device);
  } catch (...) { tmc_matrix_free(b); throw; }
  return b;
}

// ----------------------------------------------------------------------------- the scheduler
struct HostNode { int begin, end, gsize, parent, left, right, iters; double q, theta; long long tb = 0, te = 0; int tpar = 0;
                  int priv = 0;     // the whole node lives on this rank (handed-off subtree)
                  int hid = -1;     // >= 0: root of the hid-th handed-off subtree (same numbering on every rank)
                  int owner = -1;      // [begin,end) local, gsize global
                  int warm = 0; };     // the parent left a start vector on this node's positions (Ctx::ws)

struct Engine {
  tmc_matrix* b; Workspace* w; tmc_opts o; Ctx c; cudaStream_t s;
  std::vector<int> hstate, slot_node;
  std::priority_queue<int, std::vector<int>, std::greater<int>> free_slots;   // lowest index first: keeps the y arena compact
  std::priority_queue<int, std::vector<int>, std::greater<int>> free_priv;    // replicated mode: slots of handed-off (private) nodes
  std::vector<HostNode> nodes;
  std::deque<int> pending, pending_priv;
  int hi_slot = 0;
  // replicated mode ("deeper subtrees as per-GPU tasks")
  int act_nev = 0;      // > 0: the next activation runs in truncated-SVD mode
  bool rmode = false; int n_shared_slots = 0; int handoff_rows = 0;
  size_t gs_smem = 0;           // dynamic shared memory of k_gather_smem
  bool gs_chunk_fixed = false;
  int priv_top = 0; long long t_top = 0;
  std::vector<int> handoff_list; std::vector<long long> owner_load; int n_handoffs = 0;
  int warm_child_min = 512;
  std::vector<int> small_list; int small_rows = 0;      // nodes left to the direct solver (tmc_small.cuh), its size threshold (0 = off)
  int64_t n_sweeps = 0, n_launch = 0;
  double alg_vec_bytes = 0;     // non-matrix part of the spmv_pair algorithmic bytes
  double alg_panel_bytes = 0;   // panel bytes of the pairs: one byte per (row, panel column) and half
  double build_ms = 0;
  int tg_variant = getenv("TMC_TG") ? atoi(getenv("TMC_TG")) : 0;
  int tp_nr = getenv("TMC_TP_NR") ? atoi(getenv("TMC_TP_NR")) : 8;          // rows per register set of the nibble panel kernels (experiments)
  int gp_nr = getenv("TMC_GP_NR") ? atoi(getenv("TMC_GP_NR")) : 4;
  int tp_narrow = getenv("TMC_TP_NARROW") ? atoi(getenv("TMC_TP_NARROW")) : 0;      // 8 | 16: k_tpanel_narrow<NR>
  int gu_lpr = getenv("TMC_GU_LPR") ? atoi(getenv("TMC_GU_LPR")) : 32;      // lanes per row of the residual gather (32 | 16 | 8)
  // TMC_DEBUG=2: per-phase device time (CUDA events between the launches of a sweep), printed after the tree build
  bool phase_prof = getenv("TMC_DEBUG") && atoi(getenv("TMC_DEBUG")) >= 2;
  enum { PH_PREPARE, PH_ZEROY, PH_TGATHER, PH_AR_Y, PH_GATHER, PH_YNORM, PH_DOTS, PH_AR_COEF, PH_UPDATE, PH_FINISH, PH_ADVANCE,
         PH_PARTITION, PH_TPART, PH_COMMIT_D2H, PH_N };
  double phase_ms[PH_N] = {0};
  double host_ms[4] = {0, 0, 0, 0};     // launch (host side of a sweep), sync wait, harvest+handoff, activate
  static double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
  std::vector<cudaEvent_t> ph_pool; std::vector<std::pair<cudaEvent_t, int>> ph_marks; size_t ph_used = 0;
  void mark(int phase) {        // the time since the previous mark is charged to `phase`
    if (!phase_prof) return;
    if (ph_used == ph_pool.size()) { cudaEvent_t e; CK(cudaEventCreate(&e)); ph_pool.push_back(e); }
    cudaEvent_t e = ph_pool[ph_used++];
    CK(cudaEventRecord(e, s));
    ph_marks.push_back({e, phase});
  }
  void mark_flush() {
    if (!phase_prof) return;
    for (size_t i = 1; i < ph_marks.size(); ++i) {
      float ms = 0; cudaEventElapsedTime(&ms, ph_marks[i - 1].first, ph_marks[i].first);
      phase_ms[ph_marks[i].second] += ms;
    }
    ph_marks.clear(); ph_used = 0;
  }
  void phase_report() {
    if (!phase_prof) return;
    static const char* nm[PH_N] = {"prepare", "zero_y", "tgather/scatter", "allreduce(Y)", "gather", "ynorm", "dots", "allreduce(coef+scal)",
                                   "update", "finish", "advance", "partition", "tpart(re-pack)", "commit+D2H"};
    double tot = 0; for (int i = 0; i < PH_N; ++i) tot += phase_ms[i];
    fprintf(stderr, "[tmc] rank %d device time by phase (sum %.1f ms over %lld sweeps):\n", c.rank, tot, (long long)n_sweeps);
    for (int i = 0; i < PH_N; ++i) fprintf(stderr, "[tmc]   %-22s %9.2f ms  %5.1f%%\n", nm[i], phase_ms[i], 100.0 * phase_ms[i] / std::max(tot, 1e-9));
    fprintf(stderr, "[tmc]   host: launching %.1f ms, waiting in sync %.1f ms, harvest+hand-off %.1f ms, activate %.1f ms\n", host_ms[0], host_ms[1],
            host_ms[2], host_ms[3]);
  }
  bool profile = false;
  bool use_p2p = false;      // the per-sweep all-reduces go through the peer-memory windows (tmc_p2p.cuh) instead of NCCL
  std::vector<cudaEvent_t> evs;

  Engine(tmc_matrix* b_, const tmc_opts& o_) : b(b_), o(o_) {
    pick_device(b->device);
    w = get_workspace(b, o);
    c = w->c; s = w->stream;
    c.tol = o.tol > 0 ? o.tol : 1e-10;
    if (getenv("TMC_TOL") && atof(getenv("TMC_TOL")) > 0) c.tol = atof(getenv("TMC_TOL"));      // experiments: overrides opts.tol
    c.min_modularity = o.min_modularity; c.min_size = std::max(1, o.min_size);
    c.max_restarts = o.max_restarts >= 0 ? o.max_restarts : 8; c.seed = o.seed;
    c.xs_min_rows = getenv("TMC_XS_ROWS") ? atoi(getenv("TMC_XS_ROWS")) : 8192;
    // EXPERIMENTS, off by default (measured on the c3 tree, profiles/r2_residual_smem_experiments.md: no gain over the kernels
    // they replace).  TMC_GS=1: residual gather of large nodes with y[P, n) in shared memory, when that vector fits one CTA;
    // TMC_TS=1 (needs TMC_GS=1): their transposed product from the CSR rows with shared-memory atomics.  TMC_GS_ROWS /
    // TMC_GS_CHUNK tune the node-size threshold and the chunk.
    {
      const size_t need = sizeof(double) * (size_t)(c.n_cols - c.P + 2);
      int dev_max = 0; cudaDeviceGetAttribute(&dev_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, b->device);
      c.gs_on = c.use_t && (getenv("TMC_GS") ? atoi(getenv("TMC_GS")) != 0 : 0) && need + 1024 <= (size_t)dev_max;
      c.gs_min_rows = getenv("TMC_GS_ROWS") ? atoi(getenv("TMC_GS_ROWS")) : 4096;
      gs_chunk_fixed = getenv("TMC_GS_CHUNK") != nullptr;      // else chosen per sweep (Engine::sweep)
      c.gs_chunk_rows = gs_chunk_fixed ? std::max(256, atoi(getenv("TMC_GS_CHUNK"))) : 1024;
      gs_smem = need;
      c.ts_on = c.gs_on && getenv("TMC_TS") && atoi(getenv("TMC_TS")) != 0;      // experimental: off unless asked for
      if (c.gs_on) {
        VT_DISPATCH(c.vt, CK(cudaFuncSetAttribute(k_gather_smem<VT, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need)));
        VT_DISPATCH(c.vt, CK(cudaFuncSetAttribute(k_tscatter_smem<VT, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need)));
      }
    }
    // sign-certified stop (tmc_opts.sign_stop: 0 = on, -1 = off; TMC_NO_CERT=1 switches it off too)
    // KMeansGroup on one eigenvector runs inside the sweeps (ST_KM); labels are not signs, so nothing can be sign-certified
    c.kmeans = (o.eigen_group == 1 && o.num_eigen <= 1) ? 1 : 0;
    c.cert_on = (o.sign_stop >= 0 && getenv("TMC_NO_CERT") == nullptr && !c.signed_mode && !c.kmeans) ? 1 : 0;
    c.cert_min_rows = getenv("TMC_CERT_MIN_ROWS") ? atoi(getenv("TMC_CERT_MIN_ROWS")) : 64;
    c.cert_safe = getenv("TMC_CERT_SAFE") ? atof(getenv("TMC_CERT_SAFE")) : 4.0;
    // warm start of a child's Lanczos run from its parent's Krylov space (tmc_opts.warm_start: 0 = on, -1 = off; TMC_NO_WARM=1
    // switches it off too).  The default (deflated, unsigned) solver path only.
    c.warm_on = (o.warm_start >= 0 && getenv("TMC_NO_WARM") == nullptr && !c.signed_mode && !c.kmeans) ? 1 : 0;
    c.warm_min_rows = getenv("TMC_WARM_MIN_ROWS") ? atoi(getenv("TMC_WARM_MIN_ROWS")) : 32;
    c.canon_sign = (getenv("TMC_NO_CANON") == nullptr && !c.kmeans) ? 1 : 0;
    c.cert_warm_scale = getenv("TMC_CERT_WARM_SCALE") ? atoi(getenv("TMC_CERT_WARM_SCALE")) : 1;
    c.cert_s0 = getenv("TMC_CERT_S0") ? atof(getenv("TMC_CERT_S0")) : 2e-4;
    c.warm_mix = getenv("TMC_WARM_MIX") ? atof(getenv("TMC_WARM_MIX")) : 0.0;
    c.cert_kfac = getenv("TMC_CERT_KFAC") ? atof(getenv("TMC_CERT_KFAC")) : 2.0;
    // only nodes of at least this many cells start warm: the sign certificate of a small node is too easy to pass for a start
    // vector that favours u2 (DESIGN.md "warm start"; measured: 1 wrong split in ~50 000 small certified nodes without it)
    warm_child_min = getenv("TMC_WARM_CHILD_ROWS") ? atoi(getenv("TMC_WARM_CHILD_ROWS")) : 512;
    // a node with fewer cells than that cannot have a child that starts warm: no second Ritz vector, no start vector left behind
    if (!getenv("TMC_WARM_MIN_ROWS")) c.warm_min_rows = std::max(c.warm_min_rows, warm_child_min + 1);
    // direct solve of small local nodes (tmc_opts.small_rows: 0 = default 64, -1 = off, else the threshold, at most 64)
    small_rows = o.small_rows < 0 ? 0 : (o.small_rows == 0 ? TMC_SMALL_MAX : std::min(o.small_rows, TMC_SMALL_MAX));
    if (getenv("TMC_SMALL_ROWS")) small_rows = std::max(0, std::min(atoi(getenv("TMC_SMALL_ROWS")), TMC_SMALL_MAX));
    if (c.signed_mode || c.kmeans) small_rows = 0;
    c.rank = g_nccl.comm ? g_nccl.rank : 0; c.world = g_nccl.comm ? g_nccl.world : 1;
    hstate.assign(w->max_active, ST_IDLE); slot_node.assign(w->max_active, -1);
    rmode = b->replicated && c.world > 1 && c.use_t;
    // replicated mode needs at least one private slot, or handed-off subtrees could never be placed (get_workspace keeps
    // max_active >= 2 under a communicator)
    n_shared_slots = rmode ? std::max(1, std::min(256, w->max_active / 2)) : w->max_active;
    if (rmode && n_shared_slots >= w->max_active) throw TmcError(TMC_ERR_INVALID, "internal: no private slot in replicated mode");
    for (int i = 0; i < n_shared_slots; ++i) free_slots.push(i);
    for (int i = n_shared_slots; i < w->max_active; ++i) free_priv.push(i);
    handoff_rows = getenv("TMC_HANDOFF_ROWS") ? atoi(getenv("TMC_HANDOFF_ROWS")) : 8192;
    owner_load.assign(c.world, 0);
    c.collectives = c.world > 1;
    if (c.collectives && p2p::enabled()) {      // collective: every rank builds its engine at the same point of the call
      const size_t want = std::max((size_t)c.ystride, (size_t)(c.K + 2 + R_NSCAL)) * (size_t)n_shared_slots;
      use_p2p = p2p::g_p2p.ensure(want, s);
      if (getenv("TMC_DEBUG")) fprintf(stderr, "[tmc] rank %d: peer-memory all-reduce %s (window %zu doubles x 2)\n", c.rank,
                                       use_p2p ? "on" : "unavailable, using NCCL", p2p::g_p2p.cap);
    }
    profile = o.profile != 0;
    CK(cudaMemsetAsync(c.st, 0, sizeof(Slot) * w->max_active, s));
    CK(cudaMemsetAsync(c.err, 0, sizeof(int), s));
    CK(cudaMemsetAsync(c.stats, 0, 4 * sizeof(unsigned long long), s));
  }
  ~Engine() { for (auto e : evs) cudaEventDestroy(e); for (auto e : ph_pool) cudaEventDestroy(e); }

  void init_perm(const int64_t* rows, int64_t n_sel) {
    const int m = (int)b->n_rows;
    if (!rows) {
      const int cnt = rmode ? (int)(b->own_hi - b->own_lo) : m, off = rmode ? (int)b->own_lo : 0;
      if (cnt > 0) k_iota<<<(cnt + 255) / 256, 256, 0, s>>>(c.perm, cnt, off);
    } else {
      std::vector<int> h(n_sel);
      for (int64_t i = 0; i < n_sel; ++i) {
        if (rows[i] < 0 || rows[i] >= m) throw TmcError(TMC_ERR_INVALID, "row id out of range");
        if (i && rows[i] <= rows[i - 1]) throw TmcError(TMC_ERR_INVALID, "row ids must be ascending and unique");
        h[i] = (int)rows[i];
      }
      CK(cudaMemcpyAsync(c.perm, h.data(), sizeof(int) * n_sel, cudaMemcpyHostToDevice, s));
      CK(cudaStreamSynchronize(s));
    }
  }

  // column-sorted copy of the rows at positions [begin,end) into buffer `par` at offset 0 (root / per-node API)
  void build_tblock(HostNode& nd, long long tb = 0) {
    if (!c.use_t) return;
    const int rows = nd.end - nd.begin;
    const int grid = std::max(1, std::min(148 * 8, (rows + 7) / 8));
    CK(cudaMemsetAsync(c.tb_cnt, 0, sizeof(int) * (b->n_cols + 1), s));
    k_tb_count<<<grid, 256, 0, s>>>(c, nd.begin, nd.end);
    k_tb_scan<<<1, 1024, 0, s>>>(c);
    VT_DISPATCH(c.vt, k_tb_fill<VT><<<grid, 256, 0, s>>>(c, nd.begin, nd.end, tb, 0));
    n_launch += 3;
    long long tot = 0;
    CK(cudaMemcpyAsync(&tot, c.tb_colstart + b->n_cols, sizeof(long long), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    if (tb + tot + 8 > w->tn) throw TmcError(TMC_ERR_NOMEM, "internal: transposed buffer too small for the handed-off subtrees");
    nd.tb = tb; nd.te = tb + tot; nd.tpar = 0;
  }

  // Root block when perm is the identity over all local rows: stable radix sort of the CSR entries by column
  // (cub::DeviceRadixSort, load-time plumbing) so that positions ascend inside every column; the other buffer and
  // tval[1] serve as scratch.  Falls back to the atomic-cursor build for > 2^31 - 1 entries.
  void build_tblock_sorted(HostNode& nd) {
    if (!c.use_t) return;
    const auto w0 = std::chrono::steady_clock::now();
    struct T_ { Engine* e; std::chrono::steady_clock::time_point t; ~T_() { e->build_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t).count(); } } timer_{this, w0};
    const int row_lo = rmode ? (int)b->own_lo : 0, row_hi = rmode ? (int)b->own_hi : (int)b->n_rows;
    long long e0 = 0, e1 = b->enz();
    if (rmode) {
      int64_t pe[2];
      CK(cudaMemcpy(&pe[0], b->rp() + row_lo, sizeof(int64_t), cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(&pe[1], b->rp() + row_hi, sizeof(int64_t), cudaMemcpyDeviceToHost));
      e0 = pe[0]; e1 = pe[1];
    }
    const long long n = e1 - e0;
    if (n >= (1ll << 31) - 1 || n == 0) { build_tblock(nd); return; }
    int32_t* keysA = c.tcol[1]; int32_t* keysB = c.tcol[0];
    int32_t* valsA = c.tpos[1]; int32_t* valsB = c.tpos[0];
    int32_t* rowid = w->rowid_scratch;
    CK(cudaMemcpyAsync(keysA, b->ci() + e0, sizeof(int32_t) * n, cudaMemcpyDeviceToDevice, s));
    k_tb_iota<<<148 * 8, 256, 0, s>>>(valsA, n);
    k_tb_rowid<<<148 * 8, 256, 0, s>>>(b->rp(), row_lo, row_hi, (int64_t)e0, rowid);
    cub::DoubleBuffer<int32_t> dk(keysA, keysB), dv(valsA, valsB);
    int end_bit = 1; while ((1ll << end_bit) < b->n_cols) ++end_bit;
    size_t tmp_bytes = 0;
    CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, dk, dv, (int)n, 0, end_bit, s));
    if (tmp_bytes > w->cub_tmp_bytes) {       // grown once, kept with the workspace (cudaMalloc/cudaFree here cost up to 400 ms)
      if (w->cub_tmp) cudaFree(w->cub_tmp);
      w->cub_tmp = nullptr; w->cub_tmp_bytes = 0;
      CK(cudaMalloc(&w->cub_tmp, tmp_bytes + (tmp_bytes >> 2) + 1024));
      w->cub_tmp_bytes = tmp_bytes + (tmp_bytes >> 2) + 1024;
    }
    size_t avail = w->cub_tmp_bytes;
    CK(cub::DeviceRadixSort::SortPairs(w->cub_tmp, avail, dk, dv, (int)n, 0, end_bit, s));
    if (dk.Current() != keysB) CK(cudaMemcpyAsync(keysB, dk.Current(), sizeof(int32_t) * n, cudaMemcpyDeviceToDevice, s));
    if (dv.Current() != valsB) CK(cudaMemcpyAsync(valsB, dv.Current(), sizeof(int32_t) * n, cudaMemcpyDeviceToDevice, s));
    VT_DISPATCH(c.vt, k_tb_finish_sorted<VT><<<148 * 8, 256, 0, s>>>(rowid, c.val, (int64_t)e0, valsB, c.tval[0], n));
    CK(cudaStreamSynchronize(s));
    n_launch += 6;
    nd.tb = 0; nd.te = n; nd.tpar = 0;
  }

  int activate_pending(int stop_after) {
    int n = 0;
    auto act = [&](std::deque<int>& q, std::priority_queue<int, std::vector<int>, std::greater<int>>& fs, int priv) {
      while (!q.empty() && !fs.empty()) {
        int node = q.front(); q.pop_front();
        int slot = fs.top(); fs.pop();
        w->h_act[n++] = Activation{slot, nodes[node].begin, nodes[node].end, nodes[node].gsize, node, stop_after, nodes[node].tpar, priv,
                                   nodes[node].tb, nodes[node].te, act_nev, nodes[node].warm};
        hstate[slot] = ST_DEG; slot_node[slot] = node;
        hi_slot = std::max(hi_slot, slot + 1);
      }
    };
    act(pending, free_slots, 0);
    act(pending_priv, free_priv, 1);
    if (n) {
      CK(cudaMemcpyAsync(w->d_act, w->h_act, sizeof(Activation) * n, cudaMemcpyHostToDevice, s));
      k_activate<<<(n + 127) / 128, 128, 0, s>>>(c, w->d_act, n); n_launch++;
    }
    return n;
  }

  cudaEvent_t ev() { cudaEvent_t e; CK(cudaEventCreate(&e)); evs.push_back(e); CK(cudaEventRecord(e, s)); return e; }

  struct ProfPair { cudaEvent_t a, b, c; };
  std::vector<ProfPair> prof_pairs;

  void nccl_check(int r, const char* what) {
    if (r != 0) throw TmcError(TMC_ERR_COMM, std::string(what) + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?"));
  }
  void allreduce_f64(double* p, size_t n) {
    if (c.world <= 1 || !n) return;
    // the same decision on every rank: n and the window capacity are rank-identical
    if (use_p2p && n <= p2p::g_p2p.cap) { p2p::g_p2p.allreduce(p, n, s, c.err); n_launch += 2; return; }
    nccl_check(g_nccl.AllReduce(p, p, n, NCCL_FLOAT64, NCCL_SUM, g_nccl.comm, s), "ncclAllReduce(f64)");
  }
  void allreduce_host_i32(std::vector<int>& v) {       // small host-side exchanges at the end of the tree build
    if (c.world <= 1 || v.empty()) return;
    int* d = palloc<int>(v.size());
    try {
      CK(cudaMemcpyAsync(d, v.data(), sizeof(int) * v.size(), cudaMemcpyHostToDevice, s));
      nccl_check(g_nccl.AllReduce(d, d, v.size(), NCCL_INT32, NCCL_SUM, g_nccl.comm, s), "ncclAllReduce(i32)");
      CK(cudaMemcpyAsync(v.data(), d, sizeof(int) * v.size(), cudaMemcpyDeviceToHost, s));
      CK(cudaStreamSynchronize(s));
    } catch (...) { pfree(d); throw; }
    pfree(d);
  }
  void allreduce_host_f64(std::vector<double>& v) {
    if (c.world <= 1 || v.empty()) return;
    double* d = palloc<double>(v.size());
    try {
      CK(cudaMemcpyAsync(d, v.data(), sizeof(double) * v.size(), cudaMemcpyHostToDevice, s));
      nccl_check(g_nccl.AllReduce(d, d, v.size(), NCCL_FLOAT64, NCCL_SUM, g_nccl.comm, s), "ncclAllReduce(f64)");
      CK(cudaMemcpyAsync(v.data(), d, sizeof(double) * v.size(), cudaMemcpyDeviceToHost, s));
      CK(cudaStreamSynchronize(s));
    } catch (...) { pfree(d); throw; }
    pfree(d);
  }

  // One sweep: every live slot advances by one step of its lifecycle
  // (degree pass | start-vector orthogonalisation | Lanczos step | modularity + split).
  void sweep(bool only_spmv = false) {
    int n_sc = 0, n_ga = 0, n_mod = 0, n_or = 0, n_deg = 0, live_n = 0; int64_t rows = 0;
    // collectives are issued from the SHARED slots' states only (identical on every rank); private slots never communicate
    int sh_sc = 0, sh_or = 0, sh_live = 0, hi_shared = 0, n_lan = 0, sh_lan = 0, n_km = 0;
    for (int sl = 0; sl < hi_slot; ++sl) {
      int st = hstate[sl];
      if (st < ST_DEG || st > ST_MOD) continue;
      live_n++;
      if (st == ST_LAN) n_lan++;
      if (st == ST_KM) n_km++;
      if (sl < n_shared_slots) {
        sh_live++; hi_shared = sl + 1;
        if (st == ST_LAN) sh_lan++;
        if (st == ST_DEG || st == ST_LAN || st == ST_MOD) sh_sc++;
        if (st == ST_ORTH0 || st == ST_LAN) sh_or++;
      }
      const HostNode& nd = nodes[slot_node[sl]];
      int len = nd.end - nd.begin;
      rows += len;
      if (st == ST_DEG || st == ST_LAN || st == ST_MOD) n_sc++;
      if (st == ST_DEG || st == ST_LAN) { n_ga++; alg_vec_bytes += 16.0 * (len + 1) + 32.0 * len + 16.0 * (double)b->n_cols / c.world;
                                          alg_panel_bytes += 2.0 * (double)len * c.P * (c.pnib ? 0.5 : 1.0); }
      if (st == ST_MOD) n_mod++;
      if (st == ST_DEG) n_deg++;
      if (st == ST_ORTH0 || st == ST_LAN) n_or++;
    }
    if (!live_n) return;
    const double sweep_t0 = phase_prof ? now_ms() : 0.0;
    // SpMV chunks: one row per warp while that still gives <= ~16k CTAs; vector chunks: up to 256 positions
    int R = 8;
    if (rows > 8 * 16384) R = (int)std::min<int64_t>(TMC_MAX_CHUNK, ((rows / 16384) + 7) / 8 * 8);
    int RV = 32;
    if (rows > 32 * 1184) RV = (int)std::min<int64_t>(TMC_MAX_CHUNK, ((rows / 1184) + 31) / 32 * 32);
    c.chunk_rows = R; c.vchunk_rows = RV; c.n_slots = hi_slot;
    int nch = 0, nvch = 0, ngch = 0;
    if (c.gs_on && !gs_chunk_fixed) {
      // one CTA per SM at a time (the vector takes most of the shared memory): size the chunks of the large nodes so that their
      // CTAs fill whole waves of 148, four waves when there is enough work
      int64_t big_rows = 0;
      for (int sl = 0; sl < hi_slot; ++sl) {
        int st = hstate[sl];
        if (st < ST_DEG || st > ST_MOD) continue;
        const int ll = nodes[slot_node[sl]].end - nodes[slot_node[sl]].begin;
        if (ll >= c.gs_min_rows) big_rows += ll;
      }
      const int64_t per = (big_rows + 148 * 4 - 1) / (148 * 4);
      c.gs_chunk_rows = (int)std::min<int64_t>(4096, std::max<int64_t>(256, (per + 31) / 32 * 32));
    }
    for (int sl = 0; sl < hi_slot; ++sl) {
      int st = hstate[sl];
      if (st < ST_DEG || st > ST_MOD) continue;
      const HostNode& nd = nodes[slot_node[sl]];
      const int ll = nd.end - nd.begin;
      const bool big = c.gs_on && ll >= c.gs_min_rows;      // gathered by k_gather_smem (same rule as k_prepare)
      if (!big) nch += (ll + R - 1) / R;
      else if (st == ST_DEG || st == ST_LAN || (st == ST_MOD && c.ts_on)) ngch += (ll + c.gs_chunk_rows - 1) / c.gs_chunk_rows;
      nvch += (ll + RV - 1) / RV;
    }
    const int ny = std::max(1, (int)std::min<int64_t>(64, c.ystride / 8192));
    // chunks of the transposed blocks
    int ntch = 0;
    if (c.use_t) {
      long long tot = 0;
      for (int sl = 0; sl < hi_slot; ++sl) {
        int st = hstate[sl];
        if (st == ST_DEG || st == ST_LAN || st == ST_MOD) tot += nodes[slot_node[sl]].te - nodes[slot_node[sl]].tb;
      }
      const long long TC = 2048ll * std::max<long long>(1, (tot + 2048ll * 32768 - 1) / (2048ll * 32768));
      c.tchunk = (int)TC;
      for (int sl = 0; sl < hi_slot; ++sl) {
        int st = hstate[sl];
        if (st != ST_DEG && st != ST_LAN && st != ST_MOD) continue;
        const HostNode& nd = nodes[slot_node[sl]];
        ntch += (int)((nd.te - (nd.tb & ~7ll) + TC - 1) / TC);
      }
      if (ntch > w->tcap) throw TmcError(TMC_ERR_CUDA, "internal: transposed chunk list overflow");
    }
    mark(PH_COMMIT_D2H);
    k_prepare<<<1, 1024, 0, s>>>(c);
    k_prepare_fill<<<std::max(1, std::min(148, (std::max(std::max(std::max(nch, ngch), nvch), ntch) + 255) / 256)), 256, 0, s>>>(c);
    n_launch += 2;
    if (c.vt != VT_F64 && n_deg && nvch) { k_deg_xin<<<nvch, 256, 0, s>>>(c); n_launch++; }
    mark(PH_PREPARE);
    cudaEvent_t e0 = nullptr, e1 = nullptr, e2 = nullptr;
    if (n_sc) {
      k_zero_y<<<dim3(hi_slot, ny), 256, 0, s>>>(c); n_launch++;
      mark(PH_ZEROY);
      if (profile) e0 = ev();
      if (c.use_t) {
        if (c.ts_on && ngch) { VT_DISPATCH(c.vt, k_tscatter_smem<VT, 4><<<ngch, 1024, gs_smem, s>>>(c)); n_launch++; }
        if (ntch) {
          // large nodes (long sorted column runs) read x in the strided mapping, small ones in the plain one
          int n_big = 0, n_small = 0;
          for (int sl = 0; sl < hi_slot; ++sl) {
            int st = hstate[sl];
            if (st != ST_DEG && st != ST_LAN && st != ST_MOD) continue;
            const HostNode& nd = nodes[slot_node[sl]];
            if (nd.end - nd.begin >= c.xs_min_rows) n_big++; else n_small++;
          }
          VT_DISPATCH(c.vt,
            if (tg_variant == 1) k_tgather_t<4, 1, VT><<<ntch, 256, 0, s>>>(c, 0);
            else if (tg_variant == 2) k_tgather_t<4, 0, VT><<<ntch, 256, 0, s>>>(c, 0);
            else {
              if (n_big) {
                k_tgather_t<4, 1, VT><<<ntch, 256, 0, s>>>(c, n_small ? 1 : 0);
                if (n_small) n_launch++;
              }
              if (n_small) k_tgather_t<4, 0, VT><<<ntch, 256, 0, s>>>(c, n_big ? 2 : 0);
            });
          n_launch++;
        }
      }
      else if (nch) { VT_DISPATCH(c.vt, k_scatter<VT><<<nch, 256, 0, s>>>(c)); n_launch++; }
      if (c.P && nvch) {          // dense panel half of y = A^T x: slabs of equal width (a multiple of 512 columns, at most 4096)
        const int nslab = (c.Ppad + 4095) / 4096;
        // up to 512 rows per CTA (one atomicAdd per column and CTA), fewer when that would leave SMs idle (small matrices, tree tails)
        const int G = std::max(1, std::min(512 / RV, nvch * nslab / 296));
        const int wslab = ((c.Ppad / 512 + nslab - 1) / nslab);      // warps per CTA
        if (c.pnib && tp_narrow) {      // experiment: 4-byte lanes, 256 columns per warp, 8 warps (2048 columns) per CTA
          const int nsl = (c.Ppad + 2047) / 2048;
          const int Gn = std::max(1, std::min(512 / RV, nvch * nsl / 740));
          if (tp_narrow == 16) k_tpanel_narrow<16><<<dim3((nvch + Gn - 1) / Gn, nsl), 256, 0, s>>>(c, Gn);
          else if (tp_narrow == 4) k_tpanel_narrow<4><<<dim3((nvch + Gn - 1) / Gn, nsl), 256, 0, s>>>(c, Gn);
          else k_tpanel_narrow<8><<<dim3((nvch + Gn - 1) / Gn, nsl), 256, 0, s>>>(c, Gn);
        }
        else if (c.pnib && tp_nr == 4) k_tpanel<4, 1><<<dim3((nvch + G - 1) / G, nslab), 32 * wslab, 0, s>>>(c, G);
        else if (c.pnib && tp_nr == 16) k_tpanel<16, 1><<<dim3((nvch + G - 1) / G, nslab), 32 * wslab, 0, s>>>(c, G);
        else if (c.pnib) k_tpanel<8, 1><<<dim3((nvch + G - 1) / G, nslab), 32 * wslab, 0, s>>>(c, G);
        else k_tpanel<4, 0><<<dim3((nvch + G - 1) / G, nslab), 32 * wslab, 0, s>>>(c, G);
        n_launch++;
      }
      if (profile) e1 = ev();
      mark(PH_TGATHER);
      if (c.collectives && sh_sc) { allreduce_f64(c.Y, (size_t)c.ystride * hi_shared); mark(PH_AR_Y); }
    }
    if (n_ga && (nch || ngch)) {
      if (c.P && nvch) {      // dense panel half of u = A y -> pacc
        if (c.pnib && gp_nr == 2) k_gather_panel<2, 1><<<nvch, 256, 0, s>>>(c);
        else if (c.pnib && gp_nr == 8) k_gather_panel<8, 1><<<nvch, 256, 0, s>>>(c);
        else if (c.pnib) k_gather_panel<4, 1><<<nvch, 256, 0, s>>>(c); else k_gather_panel<4, 0><<<nvch, 256, 0, s>>>(c);
        n_launch++;
      }
      if (ngch) { VT_DISPATCH(c.vt, k_gather_smem<VT, 4><<<ngch, 1024, gs_smem, s>>>(c)); n_launch++; }      // large nodes: y[P, n) in shared memory
      if (nch) {
        VT_DISPATCH(c.vt,
          if (gu_lpr == 16) k_gather_u<VT, 4, 16><<<nch, 256, 0, s>>>(c);
          else if (gu_lpr == 8) k_gather_u<VT, 4, 8><<<nch, 256, 0, s>>>(c);
          else k_gather_u<VT, 4, 32><<<nch, 256, 0, s>>>(c));
        n_launch++;
      }
    }
    mark(PH_GATHER);
    if (profile && n_sc) { e2 = ev(); prof_pairs.push_back({e0, e1, e2}); }
    if (only_spmv) { CK(cudaGetLastError()); return; }
    if (n_mod) { k_ynorm<<<dim3(hi_slot, ny), 256, 0, s>>>(c); n_launch++; mark(PH_YNORM); }
    const size_t upd_smem = sizeof(double) * ((size_t)c.K + 2 + 8 * (size_t)RV);
    // kernels depend on this rank's rows, collectives only on the (rank-identical) slot states.
    // Three all-reduces per sweep: Y above, the pass-0 coefficients, and the pass-1 coefficients + every scalar.
    if (n_or || sh_or) {
      if (nvch && n_or) { k_dots<<<nvch, 256, 0, s>>>(c, 0); n_launch++; }
      mark(PH_DOTS);
      if (c.collectives && sh_or) { allreduce_f64(c.redA, (size_t)(c.K + 2) * hi_shared); mark(PH_AR_COEF); }
      if (nvch && n_or) { k_update<<<nvch, 256, upd_smem, s>>>(c, 0); n_launch++; }
      mark(PH_UPDATE);
      if (nvch && n_or) { k_dots<<<nvch, 256, 0, s>>>(c, 1); n_launch++; }
      mark(PH_DOTS);
    }
    if (c.collectives && sh_live) { allreduce_f64(c.redB, (size_t)(c.K + 2 + R_NSCAL) * hi_shared); mark(PH_AR_COEF); }
    if (n_or && nvch) { k_update<<<nvch, 256, upd_smem, s>>>(c, 1); n_launch++; mark(PH_UPDATE); }
    k_finish<<<(hi_slot + TMC_FINISH_WARPS - 1) / TMC_FINISH_WARPS, 32 * TMC_FINISH_WARPS, sizeof(double) * TMC_FINISH_WARPS * 5 * (c.K + 2), s>>>(c); n_launch++;
    mark(PH_FINISH);
    if (c.cert_on && n_lan) {        // sign-certified stop: min |x_i| of the Ritz vectors k_finish asked about
      if (nvch) { k_cert_probe<<<nvch, 256, 0, s>>>(c); n_launch++; }
      if (c.collectives && sh_lan)   // a shared node's cells are spread over the ranks (issued from rank-identical state only)
        nccl_check(g_nccl.AllReduce(c.certmin, c.certmin, (size_t)hi_shared, NCCL_FLOAT64, NCCL_MIN, g_nccl.comm, s), "ncclAllReduce(min |x|)");
    }
    if (n_km && nvch) { k_km_assign<<<nvch, 256, 0, s>>>(c); n_launch++; }      // KMeansGroup: one Lloyd round of the slots in ST_KM
    if (nvch) { k_advance<<<nvch, 256, 0, s>>>(c); n_launch++; mark(PH_ADVANCE); }
    if (n_mod) {                                                             // local: every rank splits its own cells
      k_partition<<<hi_slot, 1024, 0, s>>>(c); n_launch++;
      mark(PH_PARTITION);
      if (c.use_t && ntch) {       // ... and the transposed blocks of the splitting nodes (stable, into the other buffer)
        k_tpart_count<<<ntch, 256, 0, s>>>(c); k_tpart_scan<<<hi_slot, 1024, 0, s>>>(c);
        VT_DISPATCH(c.vt, k_tpart_scatter<VT><<<ntch, 256, 0, s>>>(c));
        n_launch += 3;
        mark(PH_TPART);
      }
    }
    k_commit<<<(hi_slot + 127) / 128, 128, 0, s>>>(c, hi_slot); n_launch++;
    CK(cudaMemcpyAsync(w->h_rep, c.rep, sizeof(Report) * hi_slot, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(w->h_err, c.err, sizeof(int), cudaMemcpyDeviceToHost, s));
    mark(PH_COMMIT_D2H);
    const double tsync0 = phase_prof ? now_ms() : 0.0;
    CK(cudaStreamSynchronize(s));
    if (phase_prof) { const double t1 = now_ms(); host_ms[1] += t1 - tsync0; host_ms[0] += tsync0 - sweep_t0; }
    CK(cudaGetLastError());
    mark_flush();
    n_sweeps++;
    if (*w->h_err) { std::string msg; int st = err_to_status(*w->h_err, &msg); throw TmcError(st, msg); }
  }

  // harvest reports: update host mirror, record finished nodes, enqueue children
  void harvest(bool build_tree) {
    for (int sl = 0; sl < hi_slot; ++sl) {
      if (hstate[sl] < ST_DEG || hstate[sl] > ST_MOD) continue;
      const Report& r = w->h_rep[sl];
      hstate[sl] = r.state;
      if (r.state != ST_DONE) continue;
      int node = slot_node[sl];
      HostNode& nd = nodes[node];
      nd.q = r.q; nd.theta = r.theta; nd.iters = r.iters;
      if (build_tree && r.split) {
        const int b0 = nd.begin, e0 = nd.end, gs = nd.gsize, n0l = r.n0_local, n0g = r.n0;
        int li = (int)nodes.size(), ri = li + 1;
        nodes[node].left = li; nodes[node].right = ri;
        const long long ptb = nd.tb, pte = nd.te, t0 = r.tnnz0; const int cpar = nd.tpar ^ 1; const int ppriv = nd.priv;
        nodes.push_back(HostNode{b0, b0 + n0l, n0g, node, -1, -1, 0, 0.0, 0.0, ptb, ptb + t0, cpar, ppriv});
        nodes.push_back(HostNode{b0 + n0l, e0, gs - n0g, node, -1, -1, 0, 0.0, 0.0, ptb + t0, pte, cpar, ppriv});
        for (int ch : {li, ri}) {
          if (nodes[ch].gsize < 2) continue;                       // single cell: a leaf, nothing to compute
          nodes[ch].warm = (c.warm_on && r.have_z2 && nodes[ch].gsize >= warm_child_min) ? 1 : 0;
          // small and entirely on this rank: left to the direct solver after the sweeps (its positions are not touched again)
          if (small_rows > 0 && nodes[ch].gsize <= small_rows && (c.world == 1 || ppriv)) { small_list.push_back(ch); continue; }
          if (ppriv) pending_priv.push_back(ch);
          else if (rmode && nodes[ch].gsize < handoff_rows) handoff_list.push_back(ch);
          else pending.push_back(ch);
        }
      }
      hstate[sl] = ST_IDLE; slot_node[sl] = -1; if (sl < n_shared_slots) free_slots.push(sl); else free_priv.push(sl);
    }
    while (hi_slot > 0 && hstate[hi_slot - 1] == ST_IDLE) hi_slot--;
  }

  void take_slot0() { if (!free_slots.empty() && free_slots.top() == 0) free_slots.pop(); }

  // Hand small subtrees to single ranks (replicated mode).  Every rank reaches this with the same list.  The cells of a
  // node are gathered (rank order = ascending row id) with two small int32 all-reduces, the owner appends them to its
  // private position area, builds the node's column-sorted block from its resident copy of the matrix and from then on
  // works on the node alone; the other ranks only remember who owns it.
  void do_handoffs() {
    if (handoff_list.empty()) return;
    const int nh = (int)handoff_list.size(), G = c.world, me = c.rank;
    std::vector<int> lsz((size_t)G * nh, 0);
    long long total = 0;
    for (int i = 0; i < nh; ++i) { const HostNode& nd = nodes[handoff_list[i]]; lsz[(size_t)me * nh + i] = nd.end - nd.begin; total += nd.gsize; }
    allreduce_host_i32(lsz);
    CK(cudaMemsetAsync(c.perm_tmp, 0, sizeof(int) * total, s));
    std::vector<long long> off(nh + 1, 0);
    for (int i = 0; i < nh; ++i) off[i + 1] = off[i] + nodes[handoff_list[i]].gsize;
    for (int i = 0; i < nh; ++i) {
      const HostNode& nd = nodes[handoff_list[i]];
      long long o = off[i];
      for (int r = 0; r < me; ++r) o += lsz[(size_t)r * nh + i];
      if (nd.end > nd.begin) CK(cudaMemcpyAsync(c.perm_tmp + o, c.perm + nd.begin, sizeof(int) * (nd.end - nd.begin), cudaMemcpyDeviceToDevice, s));
    }
    nccl_check(g_nccl.AllReduce(c.perm_tmp, c.perm_tmp, (size_t)total, NCCL_INT32, NCCL_SUM, g_nccl.comm, s), "ncclAllReduce(hand-off ids)");
    for (int i = 0; i < nh; ++i) {
      HostNode& nd = nodes[handoff_list[i]];
      int owner = 0;
      for (int r = 1; r < G; ++r) if (owner_load[r] < owner_load[owner]) owner = r;
      owner_load[owner] += nd.gsize;
      nd.hid = n_handoffs++; nd.owner = owner;
      nd.warm = 0;                  // the cells move to new positions; the parent's start vector does not travel with them
      if (owner != me) { nd.begin = nd.end = 0; continue; }
      if ((long long)priv_top + nd.gsize > w->m) throw TmcError(TMC_ERR_NOMEM, "internal: private position area exhausted");
      CK(cudaMemcpyAsync(c.perm + priv_top, c.perm_tmp + off[i], sizeof(int) * nd.gsize, cudaMemcpyDeviceToDevice, s));
      nd.begin = priv_top; nd.end = priv_top + nd.gsize; nd.priv = 1;
      priv_top += nd.gsize;
      build_tblock(nd, t_top);
      t_top = (nd.te + 7) & ~7ll;
      pending_priv.push_back(handoff_list[i]);
    }
    CK(cudaStreamSynchronize(s));
    handoff_list.clear();
  }
  // Direct solve of the small nodes collected by harvest(): one CTA per node builds its Gram matrix and its whole subtree
  // (k_small_subtree); the pre-order records come back in one copy and become ordinary host nodes.
  void solve_small() {
    if (small_list.empty()) return;
    const int n = (int)small_list.size();
    std::vector<SmallNode> sn(n);
    long long nrec = 0;
    for (int i = 0; i < n; ++i) {
      const HostNode& nd = nodes[small_list[i]];
      sn[i] = SmallNode{nd.begin, nd.end - nd.begin, nrec};
      nrec += 2ll * (nd.end - nd.begin) - 1;
    }
    const size_t smem = small_smem_bytes();
    // scratch lists of the residual part: room for 4 x the average 64-row node (clusters of large cells have long rows), between
    // 32 k and 256 k entries per CTA (larger nodes take the kernel's fallback path); at most 2 GB in all
    const int64_t avg_row = b->n_rows > 0 ? b->enz() / b->n_rows + 1 : 1;
    const int cap = (int)std::min<int64_t>(262144, std::max<int64_t>(32768, avg_row * 256));
    const size_t stride = small_scratch_stride((int)b->n_cols, cap);
    int grid = std::min(n, 148 * 3);
    grid = (int)std::max<int64_t>(1, std::min<int64_t>(grid, ((int64_t)2 << 30) / (int64_t)stride));
    SmallNode* d_sn = nullptr; SmallRec* d_rec = nullptr; char* scr = nullptr; unsigned long long* d_prof = nullptr;
    std::vector<SmallRec> recs((size_t)nrec);
    try {
      d_sn = palloc<SmallNode>(n); d_rec = palloc<SmallRec>((size_t)nrec); scr = palloc<char>((size_t)grid * stride);
      if (phase_prof) { d_prof = palloc<unsigned long long>(4); CK(cudaMemsetAsync(d_prof, 0, 4 * sizeof(unsigned long long), s)); }
      CK(cudaMemcpyAsync(d_sn, sn.data(), sizeof(SmallNode) * n, cudaMemcpyHostToDevice, s));
      CK(cudaMemsetAsync(scr, 0, (size_t)grid * stride, s));
      VT_DISPATCH(c.vt,
        CK(cudaFuncSetAttribute(k_small_subtree<VT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_small_subtree<VT><<<grid, 256, smem, s>>>(c, d_sn, n, d_rec, scr, cap, d_prof));
      n_launch++;
      if (d_prof) {
        unsigned long long hp[4];
        CK(cudaMemcpyAsync(hp, d_prof, sizeof(hp), cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        const double tot = (double)(hp[0] + hp[1] + hp[2]) + 1e-9;
        fprintf(stderr, "[tmc] rank %d: k_small_subtree over %d nodes, %d CTAs: residual Gram %.1f %%, panel Gram %.1f %%, subtree solves %.1f %% of the CTA cycles\n",
                c.rank, n, grid, 100.0 * hp[0] / tot, 100.0 * hp[1] / tot, 100.0 * hp[2] / tot);
      }
      CK(cudaMemcpyAsync(recs.data(), d_rec, sizeof(SmallRec) * (size_t)nrec, cudaMemcpyDeviceToHost, s));
      CK(cudaMemcpyAsync(w->h_err, c.err, sizeof(int), cudaMemcpyDeviceToHost, s));
      CK(cudaStreamSynchronize(s));
      CK(cudaGetLastError());
    } catch (...) { if (d_sn) pfree(d_sn); if (d_rec) pfree(d_rec); if (scr) pfree(scr); if (d_prof) pfree(d_prof); throw; }
    pfree(d_sn); pfree(d_rec); pfree(scr); if (d_prof) pfree(d_prof);
    if (*w->h_err) { std::string msg; int st = err_to_status(*w->h_err, &msg); throw TmcError(st, msg); }
    std::vector<int> stack;
    for (int i = 0; i < n; ++i) {
      long long pos = sn[i].rec0;
      stack.clear(); stack.push_back(small_list[i]);
      while (!stack.empty()) {
        const int u = stack.back(); stack.pop_back();
        const SmallRec& r = recs[(size_t)pos++];
        nodes[u].q = r.q; nodes[u].theta = r.theta; nodes[u].iters = 0;
        if (r.leaf) continue;
        const int lsize = recs[(size_t)pos].size;      // pre-order: the left child's record comes next
        const int b0 = nodes[u].begin, e0 = nodes[u].end, ppriv = nodes[u].priv;
        const int li = (int)nodes.size(), ri = li + 1;
        nodes[u].left = li; nodes[u].right = ri;
        nodes.push_back(HostNode{b0, b0 + lsize, lsize, u, -1, -1, 0, 0.0, 0.0, 0, 0, 0, ppriv});
        nodes.push_back(HostNode{b0 + lsize, e0, e0 - b0 - lsize, u, -1, -1, 0, 0.0, 0.0, 0, 0, 0, ppriv});
        stack.push_back(ri); stack.push_back(li);
      }
    }
    small_list.clear();
  }
  int live() const { int n = 0; for (int sl = 0; sl < hi_slot; ++sl) n += (hstate[sl] >= ST_DEG && hstate[sl] <= ST_MOD); return n; }

  void run_tree() {
    const int m = rmode ? (int)(b->own_hi - b->own_lo) : (int)b->n_rows;
    nodes.clear(); pending.clear(); pending_priv.clear(); small_list.clear();
    nodes.push_back(HostNode{0, m, (int)b->n_rows_global, -1, -1, -1, 0, 0.0, 0.0});
    if (b->n_rows_global >= 2) { build_tblock_sorted(nodes[0]); pending.push_back(0); }      // run_tree starts from the identity perm
    priv_top = m; t_top = (nodes[0].te + 7) & ~7ll;
    for (;;) {
      { const double t0 = phase_prof ? now_ms() : 0.0; activate_pending(0); if (phase_prof) host_ms[3] += now_ms() - t0; }
      if (rmode && c.collectives) {
        // the shared top of the tree is finished on every rank at the same sweep: from here on no collective is issued
        bool shared_left = !pending.empty();
        for (int sl = 0; sl < std::min(hi_slot, n_shared_slots) && !shared_left; ++sl) shared_left = hstate[sl] >= ST_DEG && hstate[sl] <= ST_MOD;
        if (!shared_left) c.collectives = 0;
      }
      if (!live()) {
        if (!pending.empty() || !pending_priv.empty() || !handoff_list.empty())
          throw TmcError(TMC_ERR_CUDA, "internal: nodes left in the queue with no slot to run them");
        break;
      }
      sweep();
      const double t0 = phase_prof ? now_ms() : 0.0;
      harvest(true);
      if (rmode) do_handoffs();
      if (phase_prof) host_ms[2] += now_ms() - t0;
    }
    const double t0 = phase_prof ? now_ms() : 0.0;
    solve_small();
    if (phase_prof) fprintf(stderr, "[tmc] rank %d: direct solve of the small nodes %.2f ms\n", c.rank, now_ms() - t0);
  }

  // run slot 0 over rows [0,n) until it reports DONE
  void run_single(int n, int stop_after, bool identity = false) {
    nodes.clear(); pending.clear();
    nodes.push_back(HostNode{0, n, n, -1, -1, -1, 0, 0.0, 0.0});
    if (identity && n == (int)b->n_rows) build_tblock_sorted(nodes[0]); else build_tblock(nodes[0]);
    pending.push_back(0);
    activate_pending(stop_after);
    const int64_t cap = (int64_t)(w->K + 8) * (c.max_restarts + 2) + 16;     // Lanczos always terminates within this
    for (int64_t it = 0; live(); ++it) {
      if (it > cap) throw TmcError(TMC_ERR_CUDA, "internal: node did not finish within the sweep cap");
      sweep(); harvest(false);
    }
  }
};

// ----------------------------------------------------------------------------- C ABI
template <class F> static int guarded(F&& f) {
  try { f(); g_err.clear(); return TMC_OK; }
  catch (const TmcError& e) { return fail(e.code, e.msg); }
  catch (const std::bad_alloc&) { return fail(TMC_ERR_NOMEM, "host allocation failed"); }
  catch (const std::exception& e) { return fail(TMC_ERR_INVALID, e.what()); }
}

extern "C" {

void tmc_opts_default(tmc_opts* o) {
  if (!o) return;
  memset(o, 0, sizeof(*o));
  o->eigen_group = 0; o->num_eigen = 1; o->min_modularity = 0.0; o->min_size = 1; o->normalize = 0; o->num_runs = 0;
  o->seed = 0x2545F4914F6CDD1Dull; o->tol = 1e-10; o->max_lanczos = 320; o->max_restarts = 8; o->max_active = 2048; o->device = -1;
}
const char* tmc_last_error(void) { return g_err.c_str(); }
const char* tmc_version(void) { return "libtmc 0.1 (sm_100a)"; }
int tmc_device_count(void) { int n = 0; if (cudaGetDeviceCount(&n) != cudaSuccess) return 0; return n; }

// the per-node exports and the sweep engine know the default variant only; tmc_make_tree* dispatch the others (tmc_variants.cuh)
static void check_opts(const tmc_opts* o, bool allow_kmeans1 = false) {
  if (!o) throw TmcError(TMC_ERR_INVALID, "null opts");
  if (o->eigen_group != 0 && !(allow_kmeans1 && o->eigen_group == 1 && o->num_eigen <= 1))
    throw TmcError(TMC_ERR_UNSUPPORTED, "eigen_group KMeansGroup: only through tmc_make_tree / tmc_make_tree_dev");
  if (o->num_runs > 0) throw TmcError(TMC_ERR_UNSUPPORTED, "permutation test (--numRuns): only through tmc_make_tree* / tmc_permutation_test");
}
static void check_variant_opts(const tmc_opts* o) {
  if (!o) throw TmcError(TMC_ERR_INVALID, "null opts");
  if (o->eigen_group != 0 && o->eigen_group != 1) throw TmcError(TMC_ERR_INVALID, "eigen_group must be 0 (SignGroup) or 1 (KMeansGroup)");
  if (o->eigen_group == 1 && o->num_eigen < 1) throw TmcError(TMC_ERR_INVALID, "num_eigen must be >= 1");
  if (o->num_runs < 0) throw TmcError(TMC_ERR_INVALID, "num_runs must be >= 0");
}

int tmc_tfidf(const tmc_csr* b1, double* val_out, double* idf_out) {
  return guarded([&] {
    if (!val_out) throw TmcError(TMC_ERR_INVALID, "null val_out");
    require_default_types(b1, "tmc_tfidf");
    if (g_nccl.comm && g_nccl.world > 1) throw TmcError(TMC_ERR_UNSUPPORTED, "host-buffer tmc_tfidf is single-rank; use tmc_matrix_upload(normalize=1)");
    tmc_matrix* b = upload(b1, -1);
    double* idf_dev = nullptr;
    try {
      idf_dev = dalloc<double>(b->n_cols);
      normalise_dev(b, 1, nullptr, idf_dev, false);
      if (b->nnz) CK(cudaMemcpy(val_out, b->val, sizeof(double) * b->nnz, cudaMemcpyDeviceToHost));
      if (idf_out) CK(cudaMemcpy(idf_out, idf_dev, sizeof(double) * b->n_cols, cudaMemcpyDeviceToHost));
    } catch (...) { if (idf_dev) cudaFree(idf_dev); tmc_matrix_free(b); throw; }
    cudaFree(idf_dev); tmc_matrix_free(b);
  });
}

int tmc_row_l2(const tmc_csr* b2, double* val_out, double* norm_out) {
  return guarded([&] {
    if (!val_out) throw TmcError(TMC_ERR_INVALID, "null val_out");
    require_default_types(b2, "tmc_row_l2");
    if (g_nccl.comm && g_nccl.world > 1) throw TmcError(TMC_ERR_UNSUPPORTED, "host-buffer tmc_row_l2 is single-rank; use tmc_matrix_upload");
    tmc_matrix* b = upload(b2, -1);
    double* nd = nullptr;
    try {
      nd = dalloc<double>(b->n_rows);
      normalise_dev(b, 0, nd, nullptr, true);
      if (b->nnz) CK(cudaMemcpy(val_out, b->val, sizeof(double) * b->nnz, cudaMemcpyDeviceToHost));
      if (norm_out) CK(cudaMemcpy(norm_out, nd, sizeof(double) * b->n_rows, cudaMemcpyDeviceToHost));
    } catch (...) { if (nd) cudaFree(nd); tmc_matrix_free(b); throw; }
    cudaFree(nd); tmc_matrix_free(b);
  });
}

int tmc_matrix_upload(const tmc_csr* b2, int normalize, int device, tmc_matrix** out) {
  return guarded([&] {
    if (!out) throw TmcError(TMC_ERR_INVALID, "null out");
    const bool dbg = getenv("TMC_DEBUG") != nullptr;
    auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double w0 = now();
    tmc_matrix* b = upload(b2, device);
    const double w1 = now();
    try { prepare_matrix(b, normalize); } catch (...) { tmc_matrix_free(b); throw; }
    if (dbg) { cudaDeviceSynchronize(); fprintf(stderr, "[tmc] tmc_matrix_upload: host -> device %.2f ms, prepare %.2f ms\n", w1 - w0, now() - w1); }
    *out = b;
  });
}

int tmc_matrix_adopt_dev(int64_t n_rows, int64_t n_cols, int64_t nnz, const int64_t* row_ptr_dev, const int32_t* col_dev,
                         double* val_dev, int64_t row_offset, int64_t n_rows_global, int normalize, int device, tmc_matrix** out) {
  return guarded([&] {
    if (!out || !row_ptr_dev || !col_dev || !val_dev) throw TmcError(TMC_ERR_INVALID, "null pointer");
    if (n_rows >= (int64_t)1 << 31 || n_cols >= (int64_t)1 << 31 || n_rows_global >= (int64_t)1 << 31) throw TmcError(TMC_ERR_INVALID, "dimension exceeds int32");
    tmc_matrix* b = new tmc_matrix();
    try {
      b->device = pick_device(device);
      b->n_rows = n_rows; b->n_cols = n_cols; b->nnz = nnz; b->owns = false;
      b->n_rows_global = n_rows_global > 0 ? n_rows_global : n_rows; b->row_offset = n_rows_global > 0 ? row_offset : 0;
      if (b->row_offset < 0 || b->row_offset + n_rows > b->n_rows_global) throw TmcError(TMC_ERR_INVALID, "row block outside [0, n_rows_global)");
      if (g_nccl.comm && g_nccl.world > 1 && b->n_rows_global == n_rows && getenv("TMC_NO_HANDOFF") == nullptr) {
        b->replicated = true;      // the whole matrix is resident on every rank
        b->own_lo = n_rows * g_nccl.rank / g_nccl.world; b->own_hi = n_rows * (g_nccl.rank + 1) / g_nccl.world;
        b->blk_e.resize(g_nccl.world + 1);
        for (int r = 0; r <= g_nccl.world; ++r)
          CK(cudaMemcpy(&b->blk_e[r], row_ptr_dev + n_rows * r / g_nccl.world, sizeof(int64_t), cudaMemcpyDeviceToHost));
      }
      b->row_ptr = const_cast<int64_t*>(row_ptr_dev); b->col = const_cast<int32_t*>(col_dev); b->val = val_dev;
      const double w0 = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
      prepare_matrix(b, normalize);
      if (getenv("TMC_DEBUG")) {
        cudaDeviceSynchronize();
        fprintf(stderr, "[tmc] tmc_matrix_adopt_dev: prepare %.2f ms\n", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count() - w0);
      }
    } catch (...) { tmc_matrix_free(b); throw; }
    *out = b;
  });
}

int tmc_matrix_info(const tmc_matrix* b, int64_t* info, int n) {
  if (!b || !info || n < 0) return fail(TMC_ERR_INVALID, "null pointer");
  const int64_t v[7] = {(int64_t)b->vt, (int64_t)b->P, (int64_t)b->pw, b->enz(), b->nnz, b->n_rows, (int64_t)(b->P ? (b->pnib ? 4 : 8) : 0)};
  for (int i = 0; i < n && i < 7; ++i) info[i] = v[i];
  return TMC_OK;
}

void tmc_cache_clear(void) {
  std::lock_guard<std::mutex> lk(g_cache_mu);
  delete g_ws_cache; g_ws_cache = nullptr;
  staging_release();
  raw_bufs_release();
  pool::release_all();
  factor_bufs_release(g_fb_cache);
  panel_bufs_release(g_pb_cache);
}

void tmc_matrix_free(tmc_matrix* b) {
  if (!b) return;
  retire_workspace(b->ws); b->ws = nullptr;
  factor_bufs_retire(b);
  panel_bufs_retire(b);
  if (b->owns) { raw_retire(0, b->row_ptr, b->raw_cap[0], b->device); raw_retire(1, b->col, b->raw_cap[1], b->device); raw_retire(2, b->val, b->raw_cap[2], b->device); }
  delete b;
}

}  // extern "C"

// Under a communicator the host-composed variants (KMeansGroup, the permutation test) run as REPLICAS: the matrix is resident on
// every rank (replicated mode), so every rank builds the same result on its own GPU with the communicator hidden, and the ranks
// then check that they agree (agree_hash).  A rank that only holds a row block cannot do that: TMC_ERR_UNSUPPORTED.
struct LocalReplica {
  ncclComm_t comm; int rank, world; bool active;
  explicit LocalReplica(const tmc_matrix* b) : comm(g_nccl.comm), rank(g_nccl.rank), world(g_nccl.world), active(g_nccl.comm && g_nccl.world > 1) {
    if (!active) return;
    if (!b || !b->replicated) {
      active = false;
      throw TmcError(TMC_ERR_UNSUPPORTED, "the non-default variants need the whole matrix on every rank (pass the full matrix, not a row block)");
    }
    g_nccl.comm = nullptr; g_nccl.rank = 0; g_nccl.world = 1;
  }
  bool was_active() const { return comm && world > 1; }
  void restore() { if (active) { g_nccl.comm = comm; g_nccl.rank = rank; g_nccl.world = world; active = false; } }
  ~LocalReplica() { restore(); }
};
static uint64_t fnv1a(uint64_t h, const void* p, size_t n) {
  const unsigned char* c = static_cast<const unsigned char*>(p);
  for (size_t i = 0; i < n; ++i) { h ^= c[i]; h *= 0x100000001B3ull; }
  return h;
}
// every rank must hold the same 64-bit digest (two int32 MIN / MAX all-reduces); else TMC_ERR_COMM
static void agree_hash(uint64_t h, const char* what) {
  if (!(g_nccl.comm && g_nccl.world > 1)) return;
  int v[4] = {(int)(h & 0x7fffffff), (int)((h >> 31) & 0x7fffffff), -(int)(h & 0x7fffffff), -(int)((h >> 31) & 0x7fffffff)};
  int* d = palloc<int>(4);
  cudaMemcpy(d, v, sizeof(v), cudaMemcpyHostToDevice);
  const int r = g_nccl.AllReduce(d, d, 4, NCCL_INT32, NCCL_MIN, g_nccl.comm, 0);
  int o[4]; cudaMemcpy(o, d, sizeof(o), cudaMemcpyDeviceToHost); pfree(d);
  if (r != 0) throw TmcError(TMC_ERR_COMM, "ncclAllReduce(replica digest) failed");
  if (o[0] != -o[2] || o[1] != -o[3]) throw TmcError(TMC_ERR_COMM, std::string("the ranks built different results for ") + what);
}
#include "tmc_variants.cuh"      // TreeStorage + the host-composed variants
extern "C" {

// ---- A10, batched: the permutation test of every split of a finished tree through the sweep engine.
// Same convention as variants::permutation_pvals (labels in ascending row order, cumulative Fisher-Yates on splitmix64 seeded from
// (seed, smallest row id, size), p = #{Q_shuffled >= Q_observed - 1e-12} / num_runs), but instead of one tmc_modularity call per
// node and run -- each a degree pass plus a product on a freshly built block -- the splits of one tree depth (they are disjoint)
// sit in one slot each: one degree sweep per depth, then ONE modularity sweep per run for all of them (CSR scatter + dense panel,
// no column-sorted blocks to build).  The shuffles run on the host between sweeps.  TMC_PERM_COMPOSED=1: the per-node version.
// `ge` (one count per tree vertex) receives #{runs evaluated by THIS rank with Q_shuffled >= Q_observed - 1e-12}: under a
// communicator the runs are dealt round-robin to the ranks (every rank advances every shuffle on the host, the stream is
// cumulative, but only evaluates its own runs on its GPU); the caller sums the counts over the ranks.
static void permutation_counts_batched(tmc_matrix* b, const tmc_tree& t, int num_runs, uint64_t seed, std::vector<double>& ge_out, int rank, int world) {
  ge_out.assign((size_t)t.n_nodes, 0.0);
  if (num_runs < 1 || t.n_nodes < 2) return;
  std::vector<int> depth(t.n_nodes, 0); int maxd = 0;
  for (int64_t u = 1; u < t.n_nodes; ++u) { depth[u] = depth[t.parent[u]] + 1; maxd = std::max(maxd, depth[u]); }      // pre-order: parent < child
  std::vector<std::vector<int64_t>> by_depth(maxd + 1);
  for (int64_t u = 0; u < t.n_nodes; ++u) if (t.left[u] >= 0) by_depth[depth[u]].push_back(u);
  tmc_opts o; tmc_opts_default(&o);
  o.scatter_mode = 1;          // y = B^T x straight from the CSR rows: nothing to build per node
  Engine eng(b, o);
  eng.c.cert_on = 0; eng.c.collectives = 0;
  const int S = eng.w->max_active;
  uint8_t* d_lab = nullptr; int* d_p2s = nullptr;
  std::vector<std::pair<int64_t, uint8_t>> cells;
  try {
    for (auto& level : by_depth) {
      for (size_t k0 = 0; k0 < level.size(); k0 += (size_t)S) {
        const int ns = (int)std::min<size_t>((size_t)S, level.size() - k0);
        // the chunk's cells: per node in ascending row order, nodes side by side in one position range
        std::vector<int> hperm, p2s; std::vector<uint8_t> lab; std::vector<int> nbeg(ns + 1, 0);
        std::vector<variants::SplitMix64> rng; rng.reserve(ns);
        for (int i = 0; i < ns; ++i) {
          const int64_t u = level[k0 + i], beg = t.item_begin[u], end = t.item_end[u], cut = t.item_begin[t.right[u]];
          cells.clear();
          for (int64_t k = beg; k < end; ++k) cells.push_back({t.perm[k], (uint8_t)(k >= cut ? 1 : 0)});
          std::sort(cells.begin(), cells.end());
          for (auto& cpair : cells) { hperm.push_back((int)cpair.first); lab.push_back(cpair.second); p2s.push_back(i); }
          nbeg[i + 1] = (int)hperm.size();
          rng.emplace_back(seed ^ ((uint64_t)cells[0].first * 0x9E3779B97F4A7C15ull + (uint64_t)(end - beg)));
        }
        const int m = (int)hperm.size();
        if (m > eng.w->m) throw TmcError(TMC_ERR_CUDA, "internal: permutation chunk larger than the position arena");
        if (d_lab) { pfree(d_lab); pfree(d_p2s); }
        d_lab = palloc<uint8_t>(m); d_p2s = palloc<int>(m);
        CK(cudaMemcpyAsync(eng.c.perm, hperm.data(), sizeof(int) * m, cudaMemcpyHostToDevice, eng.s));
        CK(cudaMemcpyAsync(d_p2s, p2s.data(), sizeof(int) * m, cudaMemcpyHostToDevice, eng.s));
        // one slot per node, degree pass (stop when the slot would enter the start-vector state)
        eng.nodes.clear();
        for (int i = 0; i < ns; ++i) {
          const int n = nbeg[i + 1] - nbeg[i];
          eng.nodes.push_back(HostNode{nbeg[i], nbeg[i + 1], n, -1, -1, -1, 0, 0.0, 0.0});
          eng.w->h_act[i] = Activation{i, nbeg[i], nbeg[i + 1], n, i, ST_ORTH0, 0, 0, 0, 0, 0, 0};
          eng.hstate[i] = ST_DEG; eng.slot_node[i] = i;
        }
        eng.hi_slot = ns;
        CK(cudaMemcpyAsync(eng.w->d_act, eng.w->h_act, sizeof(Activation) * ns, cudaMemcpyHostToDevice, eng.s));
        k_activate<<<(ns + 127) / 128, 128, 0, eng.s>>>(eng.c, eng.w->d_act, ns);
        auto run_sweeps = [&](int cap) {       // until every slot of the chunk reports DONE
          for (int it = 0;; ++it) {
            bool any = false;
            for (int sl = 0; sl < ns; ++sl) any |= eng.hstate[sl] >= ST_DEG && eng.hstate[sl] <= ST_MOD;
            if (!any) break;
            if (it > cap) throw TmcError(TMC_ERR_CUDA, "internal: permutation sweep did not finish");
            eng.sweep();
            for (int sl = 0; sl < ns; ++sl) {
              if (eng.hstate[sl] < ST_DEG || eng.hstate[sl] > ST_MOD) continue;
              const Report& r = eng.w->h_rep[sl];
              eng.hstate[sl] = r.state == ST_DONE ? ST_IDLE : r.state;
              if (r.state == ST_DONE) eng.nodes[sl].q = r.q;
            }
          }
        };
        run_sweeps(8);
        std::vector<double> q_obs(ns, 0.0); std::vector<int64_t> ge(ns, 0);
        std::vector<uint8_t> cur = lab;
        for (int r = 0; r <= num_runs; ++r) {
          if (r > 0)
            for (int i = 0; i < ns; ++i) {       // cumulative Fisher-Yates on the node's labels (ascending row order)
              uint8_t* l = cur.data() + nbeg[i]; const int64_t n = nbeg[i + 1] - nbeg[i];
              for (int64_t a = n - 1; a > 0; --a) { const int64_t bidx = (int64_t)(rng[i].next() % (uint64_t)(a + 1)); std::swap(l[a], l[bidx]); }
            }
          if (r > 0 && world > 1 && (r % world) != rank) continue;       // another rank evaluates this run
          CK(cudaMemcpyAsync(d_lab, cur.data(), m, cudaMemcpyHostToDevice, eng.s));
          k_force_mod_batch<<<(ns + 127) / 128, 128, 0, eng.s>>>(eng.c, ns);
          k_set_labels_batch<<<(m + 255) / 256, 256, 0, eng.s>>>(eng.c, d_lab, d_p2s, m);
          for (int sl = 0; sl < ns; ++sl) eng.hstate[sl] = ST_MOD;
          eng.hi_slot = ns;
          run_sweeps(4);
          for (int i = 0; i < ns; ++i) {
            if (r == 0) q_obs[i] = eng.nodes[i].q;
            else if (eng.nodes[i].q >= q_obs[i] - 1e-12) ++ge[i];
          }
        }
        for (int i = 0; i < ns; ++i) ge_out[level[k0 + i]] = (double)ge[i];
      }
    }
  } catch (...) { if (d_lab) { pfree(d_lab); pfree(d_p2s); } throw; }
  if (d_lab) { pfree(d_lab); pfree(d_p2s); }
}
// Every rank ends up with the same p-values: single rank, or (communicator, matrix resident on every rank) the runs split over
// the ranks and the counts summed.  TMC_PERM_COMPOSED=1: the per-node composition, every rank as a replica.
static void permutation_pvals_dispatch(tmc_matrix* b, const tmc_tree& t, int num_runs, uint64_t seed, double* p) {
  LocalReplica local(b);
  const int rank = local.was_active() ? local.rank : 0, world = local.was_active() ? local.world : 1;
  if (getenv("TMC_PERM_COMPOSED")) {
    variants::permutation_pvals(b, t, num_runs, seed, p);
    local.restore();
  } else {
    std::vector<double> ge;
    permutation_counts_batched(b, t, num_runs, seed, ge, rank, world);
    local.restore();
    if (world > 1 && !ge.empty()) {
      double* d = palloc<double>(ge.size());
      cudaMemcpy(d, ge.data(), sizeof(double) * ge.size(), cudaMemcpyHostToDevice);
      const int r = g_nccl.AllReduce(d, d, ge.size(), NCCL_FLOAT64, NCCL_SUM, g_nccl.comm, 0);
      cudaMemcpy(ge.data(), d, sizeof(double) * ge.size(), cudaMemcpyDeviceToHost); pfree(d);
      if (r != 0) throw TmcError(TMC_ERR_COMM, "ncclAllReduce(permutation counts) failed");
    }
    for (int64_t u = 0; u < t.n_nodes; ++u)
      p[u] = (t.left[u] >= 0 && num_runs > 0) ? ge[(size_t)u] / (double)num_runs : std::numeric_limits<double>::quiet_NaN();
  }
  agree_hash(fnv1a(0xcbf29ce484222325ull, p, sizeof(double) * (size_t)t.n_nodes), "the permutation test");
}

// the non-default variants around a finished / instead of the sweep-engine tree
static int finish_variants(tmc_matrix* b, const tmc_opts* opts, tmc_tree** out, int st) {
  if (st != TMC_OK || opts->num_runs < 1) return st;
  TreeStorage* ts = reinterpret_cast<TreeStorage*>(*out);
  st = guarded([&] { permutation_pvals_dispatch(b, ts->t, opts->num_runs, opts->seed, ts->p_val.data()); });
  if (st != TMC_OK) { std::string keep = g_err; tmc_tree_free(*out); *out = nullptr; g_err = keep; }
  return st;
}
static int make_tree_kmeans(tmc_matrix* b, const variants::HostB* hb, const tmc_opts* opts, tmc_tree** out) {
  return guarded([&] {
    if (!b || !out) throw TmcError(TMC_ERR_INVALID, "null pointer");
    LocalReplica local(b);
    std::unique_ptr<TreeStorage> ts(variants::kmeans_tree(b, hb, *opts));
    local.restore();
    uint64_t h = fnv1a(0xcbf29ce484222325ull, ts->left.data(), sizeof(int64_t) * ts->left.size());
    h = fnv1a(h, ts->perm.data(), sizeof(int64_t) * ts->perm.size());
    agree_hash(h, "the KMeansGroup tree");
    ts->bind();
    *out = &ts.release()->t;
  });
}

static int make_tree_sweeps(tmc_matrix* b, const tmc_opts* opts_in, tmc_tree** out);
int tmc_make_tree_dev(tmc_matrix* b, const tmc_opts* opts, tmc_tree** out) {
  int st = guarded([&] { check_variant_opts(opts); });
  if (st != TMC_OK) return st;
  // KMeansGroup on one eigenvector runs inside the batched sweeps (the 2-means as Lloyd rounds, one per sweep); on several
  // eigenvectors it is composed per node (tmc_variants.cuh).  TMC_KMEANS_COMPOSED=1: the per-node composition for num_eigen = 1 too.
  const bool km_sweeps = opts->eigen_group == 1 && opts->num_eigen <= 1 && getenv("TMC_KMEANS_COMPOSED") == nullptr;
  if (opts->eigen_group == 1 && !km_sweeps) st = make_tree_kmeans(b, nullptr, opts, out);
  else if (km_sweeps) {
    st = guarded([&] {       // single rank, or every rank as a checked replica of the resident matrix
      LocalReplica local(b);
      int r = make_tree_sweeps(b, opts, out);
      if (r != TMC_OK) throw TmcError(r, g_err);
      local.restore();
      const TreeStorage* ts = reinterpret_cast<const TreeStorage*>(*out);
      uint64_t h = fnv1a(0xcbf29ce484222325ull, ts->left.data(), sizeof(int64_t) * ts->left.size());
      agree_hash(fnv1a(h, ts->perm.data(), sizeof(int64_t) * ts->perm.size()), "the KMeansGroup tree");
    });
  } else st = make_tree_sweeps(b, opts, out);
  return finish_variants(b, opts, out, st);
}

static int make_tree_sweeps(tmc_matrix* b, const tmc_opts* opts_in, tmc_tree** out) {
  return guarded([&] {
    if (!b || !out) throw TmcError(TMC_ERR_INVALID, "null pointer");
    tmc_opts sweep_opts = *opts_in; sweep_opts.num_runs = 0;      // the test runs afterwards, on the finished tree
    const tmc_opts* opts = &sweep_opts;
    check_opts(opts, /*allow_kmeans1=*/true);
    if (b->n_rows_global < 1) throw TmcError(TMC_ERR_INVALID, "empty matrix (reference: error, LoadMatrix.hs:255)");
    Engine eng(b, *opts);
    struct Ev { cudaEvent_t e = nullptr; ~Ev() { if (e) cudaEventDestroy(e); } } ev0, ev1;
    CK(cudaEventCreate(&ev0.e)); CK(cudaEventCreate(&ev1.e));
    cudaEvent_t t0 = ev0.e, t1 = ev1.e;
    CK(cudaEventRecord(t0, eng.s));
    const bool dbg = getenv("TMC_DEBUG") != nullptr;
    auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double w0 = now();
    eng.init_perm(nullptr, 0);
    eng.run_tree();
    CK(cudaEventRecord(t1, eng.s));
    eng.phase_report();
    if (dbg) fprintf(stderr, "[tmc] rank %d: run_tree wall %.2f ms (block build %.2f ms), sweeps %lld, nodes %zu, hand-offs %d\n", eng.c.rank, now() - w0, eng.build_ms,
                     (long long)eng.n_sweeps, eng.nodes.size(), eng.n_handoffs);
    const double w_asm = now();
    const int m = eng.rmode ? (int)(b->own_hi - b->own_lo) : (int)b->n_rows;
    const int64_t M = b->n_rows_global;
    const int world = eng.c.world, rank = eng.c.rank;
    const int used = eng.rmode ? eng.priv_top : m;
    std::vector<int> hperm(std::max(used, 1));
    CK(cudaMemcpyAsync(hperm.data(), eng.c.perm, sizeof(int) * used, cudaMemcpyDeviceToHost, eng.s));
    CK(cudaMemcpyAsync(eng.w->h_stats, eng.c.stats, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, eng.s));
    CK(cudaStreamSynchronize(eng.s));
    float ms = 0; CK(cudaEventElapsedTime(&ms, t0, t1));
    // pre-order renumbering, children [label 0, label 1]  (birch-beer treeToGraph, Cluster.hs:171-181).
    // Replicated mode: the subtrees handed to single ranks are serialised by their owners (pre-order records) and
    // exchanged with two small all-reduces; every rank then splices them into the shared top of the tree.
    struct Rec { int leaf, size, iters; double q, theta; };
    const int nh = eng.n_handoffs;
    std::vector<int> hroot(nh, -1);
    for (size_t u = 0; u < eng.nodes.size(); ++u) if (eng.nodes[u].hid >= 0) hroot[eng.nodes[u].hid] = (int)u;
    std::vector<int> hcnt(nh, 0);
    std::vector<std::vector<Rec>> hrec(nh);
    for (int h = 0; h < nh; ++h) {
      const int ur = hroot[h];
      if (eng.nodes[ur].owner != rank) continue;
      std::vector<int> st2{ur};
      while (!st2.empty()) {
        const int u = st2.back(); st2.pop_back();
        const HostNode& nd = eng.nodes[u];
        hrec[h].push_back(Rec{nd.left < 0 ? 1 : 0, nd.gsize, nd.iters, nd.q, nd.theta});
        if (nd.left >= 0) { st2.push_back(nd.right); st2.push_back(nd.left); }
      }
      hcnt[h] = (int)hrec[h].size();
    }
    if (nh) eng.allreduce_host_i32(hcnt);
    std::vector<long long> hoff(nh + 1, 0);
    for (int h = 0; h < nh; ++h) hoff[h + 1] = hoff[h] + hcnt[h];
    std::vector<int> ri((size_t)3 * hoff[nh], 0); std::vector<double> rd((size_t)2 * hoff[nh], 0.0);
    for (int h = 0; h < nh; ++h)
      for (size_t k = 0; k < hrec[h].size(); ++k) {
        const size_t o = (size_t)hoff[h] + k; const Rec& r = hrec[h][k];
        ri[3 * o] = r.leaf; ri[3 * o + 1] = r.size; ri[3 * o + 2] = r.iters; rd[2 * o] = r.q; rd[2 * o + 1] = r.theta;
      }
    if (nh && hoff[nh]) { eng.allreduce_host_i32(ri); eng.allreduce_host_f64(rd); }

    std::unique_ptr<TreeStorage> ts_owner(new TreeStorage());      // released to the caller only when everything below succeeded
    TreeStorage* ts = ts_owner.get();
    struct Item { int local; long long rec; int64_t parent; int64_t gbeg; };     // local >= 0: engine node; else record index
    {     // a saturated tree has 2 M vertices: no re-allocation while the nine arrays grow
      const size_t expect = eng.nodes.size() + (size_t)(nh ? hoff[nh] : 0);
      for (auto* v : {&ts->left, &ts->right, &ts->parent, &ts->item_begin, &ts->item_end}) v->reserve(expect);
      ts->q.reserve(expect); ts->theta.reserve(expect); ts->iters.reserve(expect);
    }
    std::vector<Item> stack; stack.push_back(Item{0, -1, -1, 0});
    std::vector<std::pair<int, int64_t>> shared_leaves;      // (engine node, global begin) of leaves whose cells are spread over ranks
    shared_leaves.reserve(eng.nodes.size() / 2 + 1);
    std::vector<std::pair<int, int64_t>> handed;             // (hid, global begin)
    while (!stack.empty()) {
      const Item it = stack.back(); stack.pop_back();
      const int64_t id = (int64_t)ts->left.size();
      ts->left.push_back(-1); ts->right.push_back(-1); ts->parent.push_back(it.parent);
      if (it.parent >= 0) { if (ts->left[it.parent] < 0) ts->left[it.parent] = id; else ts->right[it.parent] = id; }
      if (it.local >= 0 && eng.nodes[it.local].hid < 0) {
        const HostNode& nd = eng.nodes[it.local];
        ts->item_begin.push_back(it.gbeg); ts->item_end.push_back(it.gbeg + nd.gsize);
        ts->q.push_back(nd.q); ts->theta.push_back(nd.theta); ts->iters.push_back(nd.iters);
        if (nd.left >= 0) {
          // the nodes were created sweep by sweep, the walk is depth first: 2 M random accesses into 160 MB for a saturated tree
          __builtin_prefetch(&eng.nodes[nd.right]);
          if (eng.nodes[nd.left].left >= 0) { __builtin_prefetch(&eng.nodes[eng.nodes[nd.left].left]); __builtin_prefetch(&eng.nodes[eng.nodes[nd.left].right]); }
          stack.push_back(Item{nd.right, -1, id, it.gbeg + eng.nodes[nd.left].gsize});
          stack.push_back(Item{nd.left, -1, id, it.gbeg});
        } else shared_leaves.push_back({it.local, it.gbeg});
      } else {
        // a handed-off subtree (its root arrives as an engine node, its descendants as records in pre-order)
        long long k = it.rec;
        if (it.local >= 0) { const int h = eng.nodes[it.local].hid; k = hoff[h]; handed.push_back({h, it.gbeg}); }
        const int leaf = ri[3 * k], size = ri[3 * k + 1];
        ts->item_begin.push_back(it.gbeg); ts->item_end.push_back(it.gbeg + size);
        ts->q.push_back(rd[2 * k]); ts->theta.push_back(rd[2 * k + 1]); ts->iters.push_back(ri[3 * k + 2]);
        if (!leaf) {
          // pre-order: the left child is record k+1; the right child follows the whole left subtree
          long long j = k + 1, need = 1;
          while (need > 0) { need += ri[3 * j] ? -1 : 1; ++j; }
          const int lsize = ri[3 * (k + 1) + 1];
          stack.push_back(Item{-1, j, id, it.gbeg + lsize});
          stack.push_back(Item{-1, k + 1, id, it.gbeg});
        }
      }
    }
    const size_t nn = ts->left.size();
    // global permutation: inside a shared leaf, rank r's cells follow those of ranks < r (row blocks ascend with the rank,
    // so every leaf stays ascending in original row id); a handed-off subtree is written by its owner in one piece
    ts->perm.assign(M, 0);
    if (world == 1) {
      for (int i = 0; i < m; ++i) ts->perm[i] = hperm[i];
    } else {
      const size_t nl = shared_leaves.size();
      std::vector<int> lsz((size_t)world * nl, 0);
      for (size_t i = 0; i < nl; ++i) { const HostNode& nd = eng.nodes[shared_leaves[i].first]; lsz[(size_t)rank * nl + i] = nd.end - nd.begin; }
      if (nl) eng.allreduce_host_i32(lsz);
      std::vector<int> gp(M, 0);
      for (size_t i = 0; i < nl; ++i) {
        const HostNode& nd = eng.nodes[shared_leaves[i].first];
        int64_t off = shared_leaves[i].second;
        for (int r = 0; r < rank; ++r) off += lsz[(size_t)r * nl + i];
        for (int p = nd.begin; p < nd.end; ++p) gp[off + (p - nd.begin)] = (int)(b->row_offset + hperm[p]);
      }
      for (auto& hg : handed) {
        const HostNode& nd = eng.nodes[hroot[hg.first]];
        if (nd.owner != rank) continue;
        for (int p = nd.begin; p < nd.end; ++p) gp[hg.second + (p - nd.begin)] = hperm[p];
      }
      eng.allreduce_host_i32(gp);
      for (int64_t i = 0; i < M; ++i) ts->perm[i] = gp[i];
    }
    if (dbg) fprintf(stderr, "[tmc] rank %d: tree assembly + exchange after the engine %.2f ms (includes waiting for the slowest rank)\n", rank, now() - w_asm);
    tmc_tree& t = ts->t;
    t.n_nodes = (int64_t)nn; t.n_rows = M;
    ts->bind();
    t.n_sweeps = eng.n_sweeps; t.spmv_pairs_nnz = (int64_t)(eng.w->h_stats[0] + eng.w->h_stats[1]); t.kernel_launches = eng.n_launch;
    t.engine_ms = ms; t.spmv_ms = 0.0;
    if (eng.profile) {
      double tot = 0; for (auto& p : eng.prof_pairs) { float a = 0; cudaEventElapsedTime(&a, p.a, p.c); tot += a; }
      t.spmv_ms = tot;
    }
    const double bv = b->vt == VT_F64 ? 8.0 : (b->vt == VT_U16 ? 2.0 : 0.0);      // b_v of BASELINE.md section 4 as actually stored
    // stored bytes the pairs have to move: (value + column index) per CSR entry and half, plus one byte per panel cell and half
    t.spmv_alg_bytes = 2.0 * (bv + 4.0) * (double)(b->P ? (int64_t)eng.w->h_stats[2] : t.spmv_pairs_nnz) + eng.alg_panel_bytes + eng.alg_vec_bytes;
    t.value_bytes = (int32_t)bv;
    *out = &ts_owner.release()->t;
  });
}

int tmc_make_tree(const tmc_csr* b2, const tmc_opts* opts, tmc_tree** out) {
  tmc_matrix* b = nullptr;
  if (!opts) return fail(TMC_ERR_INVALID, "null opts");
  const bool dbg = getenv("TMC_DEBUG") != nullptr;
  auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  const double w0 = now();
  int st = tmc_matrix_upload(b2, opts->normalize, opts->device, &b);
  if (st != TMC_OK) return st;
  const double w1 = now();
  struct Dbg { bool on; double w0, w1; decltype(now)& clk; ~Dbg() { if (on) fprintf(stderr, "[tmc] tmc_make_tree: upload + prepare %.2f ms, tree + free %.2f ms\n", w1 - w0, clk() - w1); } } dbg_{dbg, w0, w1, now};
  if (opts->eigen_group == 1 && opts->num_eigen > 1) {
    // KMeansGroup on several eigenvectors assembles C_S per node from the host copy of B (tmc_variants.cuh)
    st = guarded([&] { check_variant_opts(opts); require_default_types(b2, "KMeansGroup with num_eigen > 1"); });
    variants::HostB hb;
    if (st == TMC_OK) st = guarded([&] { variants::host_b_from(b2, opts->normalize, hb); });
    if (st == TMC_OK) st = finish_variants(b, opts, out, make_tree_kmeans(b, &hb, opts, out));
  } else {
    st = tmc_make_tree_dev(b, opts, out);
  }
  std::string keep = g_err;
  tmc_matrix_free(b);
  g_err = keep;
  return st;
}

/* == HSD.hierarchicalSpectralCluster ... (Left (sparseToHMat mat))  (Cluster.hs:157-167, --dense) */
int tmc_make_tree_dense(const double* mat, int64_t n_rows, int64_t n_cols, const tmc_opts* opts, tmc_tree** out) {
  variants::DenseCsr d;
  int st = guarded([&] { variants::dense_to_csr(mat, n_rows, n_cols, d); });
  if (st != TMC_OK) return st;
  return tmc_make_tree(&d.view, opts, out);
}

/* == the permutation test of hierarchicalSpectralCluster's `runs` argument, for every split of a finished tree */
int tmc_permutation_test(tmc_matrix* b, const tmc_tree* tree, int num_runs, uint64_t seed, double* p_out) {
  return guarded([&] {
    if (!b || !tree || !p_out) throw TmcError(TMC_ERR_INVALID, "null pointer");
    if (num_runs < 0) throw TmcError(TMC_ERR_INVALID, "num_runs must be >= 0");
    if (tree->n_rows != b->n_rows) throw TmcError(TMC_ERR_INVALID, "tree and matrix have different numbers of cells");
    permutation_pvals_dispatch(b, *tree, num_runs, seed, p_out);
  });
}

void tmc_tree_free(tmc_tree* t) {
  if (!t) return;
  delete reinterpret_cast<TreeStorage*>(t);   // tmc_tree is the first member
}

// ---- fine-grained exports, each through the same kernels as the tree build
static int64_t sel_count(tmc_matrix* b, const int64_t* rows, int64_t n_sel) {
  if (!b) throw TmcError(TMC_ERR_INVALID, "null matrix");
  if (g_nccl.comm && g_nccl.world > 1) throw TmcError(TMC_ERR_UNSUPPORTED, "per-node exports are single-rank (use tmc_make_tree* under a communicator)");
  int64_t n = rows ? n_sel : b->n_rows;
  if (n < 1 || n > b->n_rows) throw TmcError(TMC_ERR_INVALID, "bad row selection size");
  return n;
}

int tmc_degree(tmc_matrix* b, const int64_t* rows, int64_t n_sel, double* d_out) {
  return guarded([&] {
    int64_t n = sel_count(b, rows, n_sel);
    if (!d_out) throw TmcError(TMC_ERR_INVALID, "null d_out");
    tmc_opts o; tmc_opts_default(&o);
    Engine eng(b, o);
    eng.init_perm(rows, n);
    eng.run_single((int)n, ST_ORTH0);
    CK(cudaMemcpyAsync(d_out, eng.c.d, sizeof(double) * n, cudaMemcpyDeviceToHost, eng.s));
    CK(cudaStreamSynchronize(eng.s));
  });
}

int tmc_second_left(tmc_matrix* b, const int64_t* rows, int64_t n_sel, const tmc_opts* opts, double* u2_out,
                    double* theta_out, int32_t* iters_out) {
  return guarded([&] {
    int64_t n = sel_count(b, rows, n_sel);
    check_opts(opts);
    if (n < 2) throw TmcError(TMC_ERR_INVALID, "second_left needs at least 2 rows");
    Engine eng(b, *opts);
    eng.init_perm(rows, n);
    eng.run_single((int)n, ST_MOD);
    if (u2_out) CK(cudaMemcpyAsync(u2_out, eng.c.u2, sizeof(double) * n, cudaMemcpyDeviceToHost, eng.s));
    CK(cudaStreamSynchronize(eng.s));
    if (theta_out) *theta_out = eng.nodes[0].theta;
    if (iters_out) *iters_out = eng.nodes[0].iters;
  });
}

int tmc_modularity(tmc_matrix* b, const int64_t* rows, int64_t n_sel, const uint8_t* labels, double* q_out) {
  return guarded([&] {
    int64_t n = sel_count(b, rows, n_sel);
    if (!labels || !q_out) throw TmcError(TMC_ERR_INVALID, "null pointer");
    tmc_opts o; tmc_opts_default(&o);
    Engine eng(b, o);
    eng.init_perm(rows, n);
    eng.run_single((int)n, ST_ORTH0);                     // degree pass: d, sum d
    uint8_t* dl = dalloc<uint8_t>(n);
    try {
      CK(cudaMemcpyAsync(dl, labels, n, cudaMemcpyHostToDevice, eng.s));
      // re-open slot 0 directly in the modularity state with the caller's labels
      { HostNode keep = eng.nodes[0]; keep.left = keep.right = -1; eng.nodes.clear(); eng.nodes.push_back(keep); }   // keeps the T-block
      eng.hstate[0] = ST_MOD; eng.slot_node[0] = 0; eng.hi_slot = 1;
      eng.take_slot0();
      k_force_state<<<1, 1, 0, eng.s>>>(eng.c, 0, ST_MOD, 0, ST_DONE);
      k_set_labels<<<(int)((n + 255) / 256), 256, 0, eng.s>>>(eng.c, dl, (int)n, 0);
      while (eng.live()) { eng.sweep(); eng.harvest(false); }
      *q_out = eng.nodes[0].q;
    } catch (...) { cudaFree(dl); throw; }
    cudaFree(dl);
  });
}

int tmc_partition(const int64_t* rows, int64_t n_sel, const uint8_t* labels, int64_t* rows_out, int64_t* n_left) {
  return guarded([&] {
    if (!rows || !labels || !rows_out || n_sel < 0) throw TmcError(TMC_ERR_INVALID, "null pointer");
    pick_device(-1);
    const int n = (int)n_sel;
    std::vector<int> h(n); int n1 = 0;
    for (int i = 0; i < n; ++i) { h[i] = (int)rows[i]; n1 += labels[i] ? 1 : 0; }
    Ctx c{};
    int* perm = dalloc<int>(n); int* tmp = dalloc<int>(n); uint8_t* lab = dalloc<uint8_t>(n); Slot* st = dalloc<Slot>(1);
    try {
      c.perm = perm; c.perm_tmp = tmp; c.label = lab; c.st = st;
      std::vector<uint8_t> l01(n); for (int i = 0; i < n; ++i) l01[i] = labels[i] ? 1 : 0;
      CK(cudaMemcpy(perm, h.data(), sizeof(int) * n, cudaMemcpyHostToDevice));
      CK(cudaMemcpy(lab, l01.data(), n, cudaMemcpyHostToDevice));
      k_force_partition<<<1, 1>>>(c, 0, 0, n);
      k_partition<<<1, 1024>>>(c);
      CK(cudaGetLastError());
      CK(cudaMemcpy(h.data(), perm, sizeof(int) * n, cudaMemcpyDeviceToHost));
    } catch (...) { cudaFree(perm); cudaFree(tmp); cudaFree(lab); cudaFree(st); throw; }
    cudaFree(perm); cudaFree(tmp); cudaFree(lab); cudaFree(st);
    for (int i = 0; i < n; ++i) rows_out[i] = h[i];
    if (n_left) *n_left = n - n1;
  });
}

// set slot 0 up over rows [0,n) in the Lanczos state with xin = D^{-1/2} x  (degree pass first)
static void setup_pair(Engine& eng, const int64_t* rows, int64_t n, const double* x_dev) {
  eng.init_perm(rows, n);
  eng.run_single((int)n, ST_ORTH0, rows == nullptr);
  { HostNode keep = eng.nodes[0]; keep.left = keep.right = -1; eng.nodes.clear(); eng.nodes.push_back(keep); }   // keeps the T-block
  eng.hstate[0] = ST_LAN; eng.slot_node[0] = 0; eng.hi_slot = 1;
  eng.take_slot0();
  k_force_state<<<1, 1, 0, eng.s>>>(eng.c, 0, ST_LAN, 1, 0);
  k_set_xin_scaled<<<(int)((n + 255) / 256), 256, 0, eng.s>>>(eng.c, x_dev, (int)n);
}

int tmc_spmv_pair(tmc_matrix* b, const int64_t* rows, int64_t n_sel, const double* x, double* x_out) {
  return guarded([&] {
    int64_t n = sel_count(b, rows, n_sel);
    if (!x || !x_out) throw TmcError(TMC_ERR_INVALID, "null pointer");
    tmc_opts o; tmc_opts_default(&o);
    Engine eng(b, o);
    double* xd = dalloc<double>(n);
    try {
      CK(cudaMemcpy(xd, x, sizeof(double) * n, cudaMemcpyHostToDevice));
      setup_pair(eng, rows, n, xd);
      eng.sweep(true);
      CK(cudaMemcpyAsync(x_out, eng.c.w, sizeof(double) * n, cudaMemcpyDeviceToHost, eng.s));
      CK(cudaStreamSynchronize(eng.s));
      CK(cudaGetLastError());
    } catch (...) { cudaFree(xd); throw; }
    cudaFree(xd);
  });
}

int tmc_bench_spmv_pair(tmc_matrix* b, int reps, double* scatter_ms, double* gather_ms) {
  return guarded([&] {
    int64_t n = sel_count(b, nullptr, 0);
    if (reps < 1) reps = 1;
    tmc_opts o; tmc_opts_default(&o);
    Engine eng(b, o);
    std::vector<double> hx(n);
    for (int64_t i = 0; i < n; ++i) hx[i] = 1.0 / (1.0 + (double)(i % 7));
    double* xd = dalloc<double>(n);
    try {
      CK(cudaMemcpy(xd, hx.data(), sizeof(double) * n, cudaMemcpyHostToDevice));
      setup_pair(eng, nullptr, n, xd);
      eng.profile = true;
      for (int r = 0; r < reps + 2; ++r) eng.sweep(true);
      CK(cudaStreamSynchronize(eng.s));
      double sc = 0, ga = 0;
      for (int r = 2; r < reps + 2; ++r) {
        float a = 0, g = 0;
        CK(cudaEventElapsedTime(&a, eng.prof_pairs[r].a, eng.prof_pairs[r].b));
        CK(cudaEventElapsedTime(&g, eng.prof_pairs[r].b, eng.prof_pairs[r].c));
        sc += a; ga += g;
      }
      if (scatter_ms) *scatter_ms = sc / reps;
      if (gather_ms) *gather_ms = ga / reps;
    } catch (...) { cudaFree(xd); throw; }
    cudaFree(xd);
  });
}

// ---- truncated SVD through the same SpMV pair (LSA = `secondLeft 1 N . unB2 . b1ToB2 . B1`,
// src/TooManyCells/Matrix/Preprocess.hs:560-571; `sparseSvd` idiom :512-520)
int tmc_left_singular(const tmc_csr* a, int normalize, int n_vec, const tmc_opts* opts, double* u_out, double* sigma_out, int* converged_out) {
  return guarded([&] {
    check_opts(opts);
    if (!a || !u_out || n_vec < 1) throw TmcError(TMC_ERR_INVALID, "bad arguments");
    require_default_types(a, "tmc_left_singular");
    if (g_nccl.comm && g_nccl.world > 1) throw TmcError(TMC_ERR_UNSUPPORTED, "tmc_left_singular is single-rank");
    if (n_vec > a->n_rows || n_vec > a->n_cols) throw TmcError(TMC_ERR_INVALID, "more singular vectors requested than min(rows, cols)");
    tmc_matrix* b = upload(a, opts->device);
    double *zm = nullptr, *uo = nullptr, *ev = nullptr; int* cv = nullptr;
    try {
      prepare_matrix(b, normalize, /*do_l2=*/false);
      tmc_opts o = *opts;
      o.max_lanczos = (int)std::min<int64_t>(std::max(o.max_lanczos, 8 * n_vec + 64), std::min(a->n_rows, a->n_cols) + 1);
      o.max_active = 1;
      Engine eng(b, o);
      const int K2 = eng.w->K + 2;
      zm = dalloc<double>((size_t)n_vec * K2); uo = dalloc<double>((size_t)a->n_rows * n_vec); ev = dalloc<double>(n_vec); cv = dalloc<int>(1);
      CK(cudaMemset(cv, 0, sizeof(int)));
      eng.c.zmulti = zm; eng.c.uout = uo; eng.c.ev_out = ev; eng.c.svd_conv = cv;
      eng.act_nev = n_vec;
      eng.init_perm(nullptr, 0);
      eng.run_single((int)a->n_rows, 0, true);
      CK(cudaMemcpyAsync(u_out, uo, sizeof(double) * (size_t)a->n_rows * n_vec, cudaMemcpyDeviceToHost, eng.s));
      std::vector<double> hev(n_vec); int hcv = 0;
      CK(cudaMemcpyAsync(hev.data(), ev, sizeof(double) * n_vec, cudaMemcpyDeviceToHost, eng.s));
      CK(cudaMemcpyAsync(&hcv, cv, sizeof(int), cudaMemcpyDeviceToHost, eng.s));
      CK(cudaStreamSynchronize(eng.s));
      if (sigma_out) for (int i = 0; i < n_vec; ++i) sigma_out[i] = sqrt(std::max(hev[i], 0.0));
      if (converged_out) *converged_out = hcv;
    } catch (...) { cudaFree(zm); cudaFree(uo); cudaFree(ev); cudaFree(cv); tmc_matrix_free(b); throw; }
    cudaFree(zm); cudaFree(uo); cudaFree(ev); cudaFree(cv); tmc_matrix_free(b);
  });
}

// ---- multi-GPU plumbing: NCCL communicator owned by the library (one process per GPU)
int tmc_comm_unique_id(uint8_t id_out[128]) {
  return guarded([&] {
    if (!g_nccl.load()) throw TmcError(TMC_ERR_COMM, "libnccl.so.2 not found (dlopen)");
    ncclUniqueId id;
    int r = g_nccl.GetUniqueId(&id);
    if (r != 0) throw TmcError(TMC_ERR_COMM, "ncclGetUniqueId failed");
    memcpy(id_out, id.internal, 128);
  });
}
int tmc_comm_init(const uint8_t id_in[128], int rank, int world, int device) {
  return guarded([&] {
    if (!g_nccl.load()) throw TmcError(TMC_ERR_COMM, "libnccl.so.2 not found (dlopen)");
    if (g_nccl.comm) throw TmcError(TMC_ERR_COMM, "communicator already initialised");
    pick_device(device);
    ncclUniqueId id; memcpy(id.internal, id_in, 128);
    ncclComm_t comm = nullptr;
    int r = g_nccl.CommInitRank(&comm, world, id, rank);
    if (r != 0) throw TmcError(TMC_ERR_COMM, std::string("ncclCommInitRank: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?"));
    g_nccl.comm = comm; g_nccl.rank = rank; g_nccl.world = world;
  });
}
/* Collective check of the peer-memory all-reduce against ncclAllReduce on n_elems doubles (rank-dependent pseudo-random data):
 * max |p2p - nccl| over the vector (0 expected up to summation order: NCCL's order is not specified, so ~1e-16 relative), and the
 * average device time per call of either path over `reps` calls.  p2p_us = -1 when the windows could not be set up. */
int tmc_comm_selftest(int64_t n_elems, int reps, double* max_abs_diff, double* p2p_us, double* nccl_us) {
  return guarded([&] {
    if (!g_nccl.comm || g_nccl.world < 2) throw TmcError(TMC_ERR_COMM, "tmc_comm_selftest needs a communicator with >= 2 ranks");
    if (n_elems < 1 || reps < 1 || !max_abs_diff || !p2p_us || !nccl_us) throw TmcError(TMC_ERR_INVALID, "bad arguments");
    cudaStream_t s; CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    const size_t n = (size_t)n_elems;
    std::vector<double> h(n), r1(n), r2(n);
    uint64_t z = 0x9E3779B97F4A7C15ull * (uint64_t)(g_nccl.rank + 1);
    for (size_t i = 0; i < n; ++i) { z = z * 6364136223846793005ull + 1442695040888963407ull; h[i] = (double)(z >> 11) / 9007199254740992.0 - 0.5; }
    double* a = dalloc<double>(n); double* b = dalloc<double>(n); int* err = dalloc<int>(1);
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    try {
      CK(cudaMemsetAsync(err, 0, sizeof(int), s));
      const bool ok = p2p::g_p2p.ensure(n, s) && n <= p2p::g_p2p.cap;
      *p2p_us = -1.0; *max_abs_diff = 0.0;
      auto fill = [&](double* d) { CK(cudaMemcpyAsync(d, h.data(), sizeof(double) * n, cudaMemcpyHostToDevice, s)); };
      fill(a);
      if (g_nccl.AllReduce(a, a, n, NCCL_FLOAT64, NCCL_SUM, g_nccl.comm, s) != 0) throw TmcError(TMC_ERR_COMM, "ncclAllReduce failed");
      CK(cudaMemcpyAsync(r1.data(), a, sizeof(double) * n, cudaMemcpyDeviceToHost, s));
      if (ok) {
        fill(b);
        p2p::g_p2p.allreduce(b, n, s, err);
        CK(cudaMemcpyAsync(r2.data(), b, sizeof(double) * n, cudaMemcpyDeviceToHost, s));
      }
      CK(cudaStreamSynchronize(s));
      if (ok) {
        int herr = 0; CK(cudaMemcpy(&herr, err, sizeof(int), cudaMemcpyDeviceToHost));
        if (herr) throw TmcError(TMC_ERR_COMM, "peer-memory all-reduce timed out");
        for (size_t i = 0; i < n; ++i) *max_abs_diff = std::max(*max_abs_diff, std::fabs(r1[i] - r2[i]));
      }
      float ms = 0;
      CK(cudaEventRecord(e0, s));
      for (int k = 0; k < reps; ++k)
        if (g_nccl.AllReduce(a, a, n, NCCL_FLOAT64, NCCL_SUM, g_nccl.comm, s) != 0) throw TmcError(TMC_ERR_COMM, "ncclAllReduce failed");
      CK(cudaEventRecord(e1, s)); CK(cudaStreamSynchronize(s)); CK(cudaEventElapsedTime(&ms, e0, e1));
      *nccl_us = 1e3 * ms / reps;
      if (ok) {
        CK(cudaEventRecord(e0, s));
        for (int k = 0; k < reps; ++k) p2p::g_p2p.allreduce(b, n, s, err);
        CK(cudaEventRecord(e1, s)); CK(cudaStreamSynchronize(s)); CK(cudaEventElapsedTime(&ms, e0, e1));
        *p2p_us = 1e3 * ms / reps;
      }
    } catch (...) { cudaFree(a); cudaFree(b); cudaFree(err); cudaEventDestroy(e0); cudaEventDestroy(e1); cudaStreamDestroy(s); throw; }
    cudaFree(a); cudaFree(b); cudaFree(err); cudaEventDestroy(e0); cudaEventDestroy(e1); cudaStreamDestroy(s);
  });
}

void tmc_comm_destroy(void) {
  if (g_nccl.comm) p2p::g_p2p.destroy(0);
  if (g_nccl.comm && g_nccl.CommDestroy) g_nccl.CommDestroy(g_nccl.comm);
  g_nccl.comm = nullptr; g_nccl.rank = 0; g_nccl.world = 1;
}

}  // extern "C"

#include "tmc_ingest.cuh"
