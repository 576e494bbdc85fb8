// SURVEY 8(f) N4 / 8(a) A10 — the non-default variants of the engine call, behind the same C ABI:
//
//   * KMeansGroup / --num-eigen   (src/TooManyCells/Program/Options.hs:172-173; spectralClusterK in spectral-clustering):
//     labels from 2-means on the rows of eigenvectors 2..e+1 of C = D^{-1/2} B_S instead of the sign of u2
//   * --numRuns permutation test  (Options.hs:174; `_pVal`, MakeTree/Types.hs:124-130, printed at Print.hs:124-126):
//     labels of a split shuffled num_runs times, Q recomputed, empirical p
//   * --dense                     (MakeTree/Cluster.hs:157-167, `HSD.hierarchicalSpectralCluster ... . sparseToHMat`)
//
// First version: composed on the host from the per-node exports above (tmc_second_left, tmc_left_singular, tmc_degree,
// tmc_modularity), i.e. every product with the matrix runs on the GPU through the same kernels as the default tree, but
// nodes are visited one at a time like the reference's own recursion, not in the batched sweeps of Engine::run_tree.
// Upstream bodies are un-vendored ([UPSTREAM-RECALL]); the conventions fixed here are restated in
// oracle/tmc_oracle.py (`kmeans2_labels`, `permutation_p_values`) and are what the parity tests pin:
//   k-means: Lloyd, k = 2, start = the two rows with the smallest / largest first coordinate, ties go to cluster 1,
//            at most 100 rounds; label 1 = the cluster started from the largest first coordinate.  A node of n cells uses
//            min(num_eigen, n/4 - 1) >= 1 eigenvectors.
//   p-value: p = #{runs : Q_shuffled >= Q_observed - 1e-12} / num_runs; shuffle = Fisher-Yates over the node's labels in
//            ascending row order, driven by splitmix64 seeded with seed ^ (min_row * 0x9E3779B97F4A7C15 + size), so the
//            stream of a node does not depend on how the tree numbers it.
//
// This header is plain host C++ over the C ABI: tests/helpers/variants_host.cpp compiles it against oracle-backed stand-ins
// of the six exports it calls, so the composition logic is covered by the CPU test suite (tests/test_host.py).
#pragma once

// the library-owned storage behind a tmc_tree (tmc_tree is the first member: tmc_tree_free casts back)
struct TreeStorage {
  tmc_tree t;
  std::vector<int64_t> left, right, parent, item_begin, item_end, perm;
  std::vector<double> q, theta, p_val; std::vector<int32_t> iters;
  void bind() {      // point the C view at the vectors; p_val = NaN ("Nothing") until a permutation test fills it
    p_val.assign(left.size(), std::numeric_limits<double>::quiet_NaN());
    t.left = left.data(); t.right = right.data(); t.parent = parent.data(); t.q = q.data(); t.theta = theta.data(); t.p_val = p_val.data();
    t.iters = iters.data(); t.item_begin = item_begin.data(); t.item_end = item_end.data(); t.perm = perm.data();
  }
};

namespace variants {

struct SplitMix64 {
  uint64_t s;
  explicit SplitMix64(uint64_t seed) : s(seed) {}
  uint64_t next() {
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
  }
};

static void shuffle_labels(std::vector<uint8_t>& lab, SplitMix64& rng) {
  for (int64_t i = (int64_t)lab.size() - 1; i > 0; --i) {
    const int64_t j = (int64_t)(rng.next() % (uint64_t)(i + 1));
    std::swap(lab[i], lab[j]);
  }
}

static inline void must(int st) { if (st != TMC_OK) throw TmcError(st, g_err); }

// (under a communicator the callers hide it and every rank runs the variant as a replica: LocalReplica in tmc_engine.cu)
static void single_rank_only(const char* what) {
  if (g_nccl.comm && g_nccl.world > 1) throw TmcError(TMC_ERR_UNSUPPORTED, std::string(what) + " is single-rank in this version");
}

// ---- A10: permutation test of every split of a finished tree; p[u] = NaN for leaves
static void permutation_pvals(tmc_matrix* b, const tmc_tree& t, int num_runs, uint64_t seed, double* p) {
  single_rank_only("the permutation test (--numRuns)");
  for (int64_t u = 0; u < t.n_nodes; ++u) p[u] = std::numeric_limits<double>::quiet_NaN();
  if (num_runs < 1) return;
  std::vector<std::pair<int64_t, uint8_t>> cells;
  std::vector<int64_t> rows; std::vector<uint8_t> lab0, lab;
  for (int64_t u = 0; u < t.n_nodes; ++u) {
    if (t.left[u] < 0) continue;
    const int64_t beg = t.item_begin[u], end = t.item_end[u], cut = t.item_begin[t.right[u]], n = end - beg;
    cells.clear();
    for (int64_t k = beg; k < end; ++k) cells.push_back({t.perm[k], (uint8_t)(k >= cut ? 1 : 0)});
    std::sort(cells.begin(), cells.end());
    rows.resize(n); lab0.resize(n);
    for (int64_t i = 0; i < n; ++i) { rows[i] = cells[i].first; lab0[i] = cells[i].second; }
    double q_obs = 0;
    must(tmc_modularity(b, rows.data(), n, lab0.data(), &q_obs));      // same code path as the shuffled runs
    SplitMix64 rng(seed ^ ((uint64_t)rows[0] * 0x9E3779B97F4A7C15ull + (uint64_t)n));
    int64_t ge = 0;
    lab = lab0;
    for (int r = 0; r < num_runs; ++r) {
      shuffle_labels(lab, rng);
      double qp = 0;
      must(tmc_modularity(b, rows.data(), n, lab.data(), &qp));
      if (qp >= q_obs - 1e-12) ++ge;
    }
    p[u] = (double)ge / (double)num_runs;
  }
}

// ---- KMeansGroup: Lloyd's 2-means on n points of dimension e (row-major); see the conventions at the top
static void kmeans2_labels(const double* pts, int64_t n, int e, std::vector<uint8_t>& lab) {
  lab.assign(n, 0);
  int64_t lo = 0, hi = 0;
  for (int64_t i = 1; i < n; ++i) {
    if (pts[i * e] < pts[lo * e]) lo = i;
    if (pts[i * e] > pts[hi * e]) hi = i;
  }
  std::vector<double> c0(pts + lo * e, pts + lo * e + e), c1(pts + hi * e, pts + hi * e + e), s0(e), s1(e);
  if (pts[lo * e] == pts[hi * e]) { lab.assign(n, 1); return; }      // no spread in u2: one cluster, the node becomes a leaf
  for (int round = 0; round < 100; ++round) {
    bool changed = round == 0;
    int64_t n1 = 0;
    std::fill(s0.begin(), s0.end(), 0.0); std::fill(s1.begin(), s1.end(), 0.0);
    for (int64_t i = 0; i < n; ++i) {
      double d0 = 0, d1 = 0;
      for (int k = 0; k < e; ++k) { const double x = pts[i * e + k]; d0 += (x - c0[k]) * (x - c0[k]); d1 += (x - c1[k]) * (x - c1[k]); }
      const uint8_t l = d1 <= d0 ? 1 : 0;
      if (l != lab[i]) { lab[i] = l; changed = true; }
      if (l) { ++n1; for (int k = 0; k < e; ++k) s1[k] += pts[i * e + k]; }
      else   {       for (int k = 0; k < e; ++k) s0[k] += pts[i * e + k]; }
    }
    if (!changed || n1 == 0 || n1 == n) break;
    for (int k = 0; k < e; ++k) { c0[k] = s0[k] / (double)(n - n1); c1[k] = s1[k] / (double)n1; }
  }
}

// B = b2ToB(B2) on the host: only needed for num_eigen > 1, where C_S is assembled per node for tmc_left_singular
struct HostB { const tmc_csr* a = nullptr; std::vector<double> val; };

static void host_b_from(const tmc_csr* a, int normalize, HostB& hb) {
  hb.a = a; hb.val.resize((size_t)a->nnz);
  tmc_csr cur = *a;
  std::vector<double> b2;
  if (normalize) { b2.resize((size_t)a->nnz); must(tmc_tfidf(a, b2.data(), nullptr)); cur.val = b2.data(); }
  must(tmc_row_l2(&cur, hb.val.data(), nullptr));
}

// eigen-coordinates of the node's cells: n x e_eff, columns = singular vectors 2..e_eff+1 of C_S
static int node_coordinates(tmc_matrix* b, const HostB* hb, const std::vector<int64_t>& rows, int num_eigen, const tmc_opts& o0,
                            std::vector<double>& pts, double* theta, int32_t* iters) {
  // the truncated-SVD slot mode wants Krylov room: a node of n cells uses at most n/4 - 1 eigenvectors (and at least one)
  const int64_t n = (int64_t)rows.size();
  int e = (int)std::max<int64_t>(1, std::min<int64_t>(num_eigen, n / 4 - 1));
  if (hb) e = (int)std::max<int64_t>(1, std::min<int64_t>(e, hb->a->n_cols / 4 - 1));
  if (e <= 1 || !hb) {
    pts.resize(n);
    must(tmc_second_left(b, rows.data(), n, &o0, pts.data(), theta, iters));
    return 1;
  }
  std::vector<double> d(n);
  must(tmc_degree(b, rows.data(), n, d.data()));
  const tmc_csr* a = hb->a;
  std::vector<int64_t> rp(n + 1, 0);
  for (int64_t i = 0; i < n; ++i) rp[i + 1] = rp[i] + (a->row_ptr[rows[i] + 1] - a->row_ptr[rows[i]]);
  std::vector<int32_t> ci((size_t)rp[n]); std::vector<double> cv((size_t)rp[n]);
  for (int64_t i = 0; i < n; ++i) {
    const double s = 1.0 / std::sqrt(d[i]);
    int64_t o = rp[i];
    for (int64_t k = a->row_ptr[rows[i]]; k < a->row_ptr[rows[i] + 1]; ++k, ++o) { ci[o] = a->col_idx[k]; cv[o] = hb->val[k] * s; }
  }
  tmc_csr cs{n, a->n_cols, rp[n], rp.data(), ci.data(), cv.data(), 0, 0};
  std::vector<double> u((size_t)n * (e + 1)), sig(e + 1);
  int conv = 0;
  must(tmc_left_singular(&cs, 0, e + 1, &o0, u.data(), sig.data(), &conv));
  pts.resize((size_t)n * e);
  for (int64_t i = 0; i < n; ++i) for (int k = 0; k < e; ++k) pts[i * e + k] = u[i * (e + 1) + 1 + k];
  *theta = sig[1] * sig[1]; *iters = 0;
  return e;
}

// hierarchicalSpectralCluster KMeansGroup ...: the same `go` as the default engine (split iff Q > minMod, both labels occur,
// rows > 1, both sides >= minSize; children [label 0, label 1]; stable row order), one node at a time
static TreeStorage* kmeans_tree(tmc_matrix* b, const HostB* hb, const tmc_opts& opts) {
  single_rank_only("eigen_group KMeansGroup");
  if (opts.num_eigen > 1 && !hb)
    throw TmcError(TMC_ERR_UNSUPPORTED, "KMeansGroup with num_eigen > 1 needs the host matrix: call tmc_make_tree, not tmc_make_tree_dev");
  if (b->n_rows_global < 1) throw TmcError(TMC_ERR_INVALID, "empty matrix (reference: error, LoadMatrix.hs:255)");
  tmc_opts o0 = opts; o0.eigen_group = 0; o0.num_runs = 0;
  const auto w0 = std::chrono::steady_clock::now();
  std::unique_ptr<TreeStorage> ts(new TreeStorage());
  struct Pend { std::vector<int64_t> rows; int64_t parent; };
  std::vector<Pend> stack;
  { Pend root; root.rows.resize(b->n_rows); for (int64_t i = 0; i < b->n_rows; ++i) root.rows[i] = i; root.parent = -1; stack.push_back(std::move(root)); }
  std::vector<double> pts; std::vector<uint8_t> lab;
  const int64_t min_size = std::max(1, opts.min_size);
  while (!stack.empty()) {
    Pend nd = std::move(stack.back()); stack.pop_back();
    const int64_t id = (int64_t)ts->left.size(), n = (int64_t)nd.rows.size();
    ts->left.push_back(-1); ts->right.push_back(-1); ts->parent.push_back(nd.parent);
    if (nd.parent >= 0) { if (ts->left[nd.parent] < 0) ts->left[nd.parent] = id; else ts->right[nd.parent] = id; }
    const int64_t beg = (int64_t)ts->perm.size();
    ts->item_begin.push_back(beg); ts->item_end.push_back(beg + n);
    double q = 0, theta = 0; int32_t iters = 0;
    bool split = false;
    if (n >= 2) {
      const int e = node_coordinates(b, hb, nd.rows, opts.num_eigen, o0, pts, &theta, &iters);
      kmeans2_labels(pts.data(), n, e, lab);
      must(tmc_modularity(b, nd.rows.data(), n, lab.data(), &q));
      int64_t n1 = 0; for (uint8_t l : lab) n1 += l;
      split = q > opts.min_modularity && n1 >= min_size && n - n1 >= min_size;
    }
    ts->q.push_back(q); ts->theta.push_back(theta); ts->iters.push_back(iters);
    if (!split) { ts->perm.insert(ts->perm.end(), nd.rows.begin(), nd.rows.end()); continue; }
    Pend l, r; l.parent = r.parent = id;
    for (int64_t i = 0; i < n; ++i) (lab[i] ? r.rows : l.rows).push_back(nd.rows[i]);
    stack.push_back(std::move(r)); stack.push_back(std::move(l));      // pre-order: label 0 first
  }
  tmc_tree& t = ts->t;
  memset(&t, 0, sizeof(t));
  t.n_nodes = (int64_t)ts->left.size(); t.n_rows = b->n_rows;
  t.engine_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - w0).count();
  t.value_bytes = b->vt == VT_F64 ? 8 : (b->vt == VT_U16 ? 2 : 0);
  return ts.release();
}

// dense cells x features (row-major) -> CSR of the non-zero entries (sparseToHMat's inverse: zeros carry no weight in any
// product of the path, so the dense and the sparse algorithm see the same B)
struct DenseCsr { std::vector<int64_t> rp; std::vector<int32_t> ci; std::vector<double> v; tmc_csr view; };
static void dense_to_csr(const double* mat, int64_t n_rows, int64_t n_cols, DenseCsr& out) {
  if (!mat || n_rows < 0 || n_cols < 0) throw TmcError(TMC_ERR_INVALID, "null / negative dense matrix");
  if (n_rows >= (int64_t)1 << 31 || n_cols >= (int64_t)1 << 31) throw TmcError(TMC_ERR_INVALID, "dimension exceeds int32");
  out.rp.assign(n_rows + 1, 0);
  for (int64_t i = 0; i < n_rows; ++i) {
    const double* row = mat + i * n_cols;
    for (int64_t j = 0; j < n_cols; ++j) if (row[j] != 0.0) { out.ci.push_back((int32_t)j); out.v.push_back(row[j]); }   // NaN != 0: kept, rejected later
    out.rp[i + 1] = (int64_t)out.v.size();
  }
  out.view = tmc_csr{n_rows, n_cols, (int64_t)out.v.size(), out.rp.data(), out.ci.data(), out.v.data(), 0, 0};
}

}  // namespace variants
