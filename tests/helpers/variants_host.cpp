// CPU harness for too-many-cells_b200/csrc/tmc_variants.cuh (test infrastructure): the header is compiled unchanged against
// stand-ins of the six libtmc exports it composes; the stand-ins forward to callbacks that tests/test_host.py implements with
// the oracle.  What this covers is the host logic of the variants (recursion, pre-order layout, k-means, label shuffles,
// p-values, dense -> CSR), not the GPU kernels behind the real exports — those have their own -m gpu parity tests.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <memory>
#include <string>
#include <vector>

#include "../../include/tmc.h"

struct TmcError { int code; std::string msg; TmcError(int c, std::string m) : code(c), msg(std::move(m)) {} };
static thread_local std::string g_err;
static struct { void* comm = nullptr; int world = 1; } g_nccl;
enum { VT_F64 = 0, VT_U16 = 1, VT_ONE = 2 };
struct tmc_matrix { int64_t n_rows, n_rows_global, n_cols; int vt; };

typedef int (*cb_modularity)(const int64_t*, int64_t, const uint8_t*, double*);
typedef int (*cb_second_left)(const int64_t*, int64_t, double*, double*);
typedef int (*cb_degree)(const int64_t*, int64_t, double*);
typedef int (*cb_left_singular)(const tmc_csr*, int, int, double*, double*);
typedef int (*cb_values)(const tmc_csr*, double*);
static cb_modularity f_mod; static cb_second_left f_sl; static cb_degree f_deg; static cb_left_singular f_ls; static cb_values f_tfidf, f_l2;

extern "C" {
int tmc_modularity(tmc_matrix*, const int64_t* rows, int64_t n, const uint8_t* lab, double* q) { return f_mod(rows, n, lab, q); }
int tmc_second_left(tmc_matrix*, const int64_t* rows, int64_t n, const tmc_opts*, double* u2, double* th, int32_t* it) {
  *it = 0; return f_sl(rows, n, u2, th);
}
int tmc_degree(tmc_matrix*, const int64_t* rows, int64_t n, double* d) { return f_deg(rows, n, d); }
int tmc_left_singular(const tmc_csr* a, int normalize, int n_vec, const tmc_opts*, double* u, double* sig, int* conv) {
  *conv = 1; return f_ls(a, normalize, n_vec, u, sig);
}
int tmc_tfidf(const tmc_csr* a, double* v, double*) { return f_tfidf(a, v); }
int tmc_row_l2(const tmc_csr* a, double* v, double*) { return f_l2(a, v); }
}

#include "../../too-many-cells_b200/csrc/tmc_variants.cuh"

template <class F> static int guarded(F&& f) {
  try { f(); return 0; } catch (const TmcError& e) { g_err = e.msg; return e.code; }
}

extern "C" {
void vh_set_callbacks(cb_modularity a, cb_second_left b, cb_degree c, cb_left_singular d, cb_values e, cb_values f) {
  f_mod = a; f_sl = b; f_deg = c; f_ls = d; f_tfidf = e; f_l2 = f;
}
const char* vh_last_error() { return g_err.c_str(); }
// KMeansGroup tree (+ optional permutation test), exactly the sequence tmc_make_tree runs for eigen_group = 1
int vh_kmeans_tree(const tmc_csr* a, const tmc_opts* o, tmc_tree** out) {
  return guarded([&] {
    tmc_matrix b{a->n_rows, a->n_rows, a->n_cols, VT_F64};
    variants::HostB hb;
    if (o->num_eigen > 1) variants::host_b_from(a, o->normalize, hb);
    TreeStorage* ts = variants::kmeans_tree(&b, o->num_eigen > 1 ? &hb : nullptr, *o);
    ts->bind();
    if (o->num_runs > 0) variants::permutation_pvals(&b, ts->t, o->num_runs, o->seed, ts->p_val.data());
    *out = &ts->t;
  });
}
int vh_permutation_test(int64_t n_rows, const tmc_tree* t, int num_runs, uint64_t seed, double* p) {
  return guarded([&] { tmc_matrix b{n_rows, n_rows, 0, VT_F64}; variants::permutation_pvals(&b, *t, num_runs, seed, p); });
}
void vh_tree_free(tmc_tree* t) { delete reinterpret_cast<TreeStorage*>(t); }
int vh_kmeans2(const double* pts, int64_t n, int e, uint8_t* lab_out) {
  std::vector<uint8_t> lab; variants::kmeans2_labels(pts, n, e, lab); memcpy(lab_out, lab.data(), n); return 0;
}
// dense -> CSR: returns nnz, fills caller arrays sized n_rows+1 / n_rows*n_cols
int64_t vh_dense_to_csr(const double* mat, int64_t n_rows, int64_t n_cols, int64_t* rp, int32_t* ci, double* v) {
  variants::DenseCsr d;
  if (guarded([&] { variants::dense_to_csr(mat, n_rows, n_cols, d); }) != 0) return -1;
  memcpy(rp, d.rp.data(), sizeof(int64_t) * (n_rows + 1));
  if (!d.v.empty()) { memcpy(ci, d.ci.data(), sizeof(int32_t) * d.ci.size()); memcpy(v, d.v.data(), sizeof(double) * d.v.size()); }
  return (int64_t)d.v.size();
}
uint64_t vh_splitmix(uint64_t seed, int n_skip) { variants::SplitMix64 r(seed); uint64_t x = 0; for (int i = 0; i <= n_skip; ++i) x = r.next(); return x; }
}
