/* CPU port of the reference's make-tree hot path in plain C — TEST INFRASTRUCTURE / CPU BASELINE ONLY.
 *
 * Nothing in the product links or calls this file; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs do.  PARITY UNPINNED against the Haskell reference (see tmc_oracle.py):
 * the reference's arithmetic lives in un-vendored Hackage packages, so this restates the published algorithm
 * the way the reference executes it:
 *
 *   go items B            hierarchical-spectral-clustering, called at MakeTree/Cluster.hs:147-156,169
 *     B_S  = rows of B    the node's matrix is MATERIALISED (extractRow / fromRowsL), one node at a time, one core
 *     d    = B_S (B_S^T 1)                                 bToD          (spectral-clustering)
 *     C    = D^{-1/2} B_S                                  bdToC
 *     u2   = second left singular vector of C              secondLeft -> sparseSvd -> SVDLIBC las2:
 *            single-vector Lanczos, two sparse mat-vecs per step (opb), convergence at kappa (1e-6 in las2)
 *     Q    = getBModularity labels B_S                     modularity 0.2.1.1
 *
 * Differences from las2 that do not change the result: the trivial pair (sigma_1 = 1, u_1 = D^{1/2}1) is
 * deflated analytically instead of being found as the first Ritz pair, and re-orthogonalisation is full rather
 * than selective (both make this port FASTER than las2 per node, i.e. a conservative baseline).
 * Single-threaded by default: the reference recursion is sequential and SVDLIBC is single-threaded; a pthread mode
 * (tmc_port_set_threads) exists only to report a more generous all-cores baseline next to the faithful one.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { int64_t m, n, nnz; int64_t* ptr; int32_t* col; double* val; } csr_t;

typedef struct {
  int64_t n_nodes, cap;
  int64_t *left, *right, *begin, *end;    /* items of node i: perm[begin[i] .. end[i]) */
  double *q, *theta; int32_t *iters;
} tree_t;

static double dot(const double* a, const double* b, int64_t n) { double s = 0; for (int64_t i = 0; i < n; ++i) s += a[i] * b[i]; return s; }

/* ---- optional thread pool (pthreads; this image's gcc has no libgomp).  1 thread by default: the reference's recursion is
 * sequential and SVDLIBC single-threaded, and with 1 thread every loop below runs exactly as written above (bit-identical
 * results).  tmc_port_set_threads(k > 1) spreads the two mat-vecs and the re-orthogonalisation over k threads: "what if the
 * reference's path used every host core", a second and more generous baseline that bench.py reports next to the faithful one
 * (sums are re-associated, so results agree to rounding only). */
#include <pthread.h>
typedef void (*job_fn)(int tid, int nt, void* arg);
static int g_threads = 1;
static pthread_t* g_pool = NULL;
static pthread_barrier_t g_bar;
static job_fn g_job; static void* g_arg; static volatile int g_quit = 0;
static void* pool_worker(void* p) {
  const int tid = (int)(intptr_t)p;
  for (;;) { pthread_barrier_wait(&g_bar); if (g_quit) break; g_job(tid, g_threads, g_arg); pthread_barrier_wait(&g_bar); }
  return NULL;
}
static void run_par(job_fn f, void* arg) { g_job = f; g_arg = arg; pthread_barrier_wait(&g_bar); f(0, g_threads, arg); pthread_barrier_wait(&g_bar); }
void tmc_port_set_threads(int k) {
  if (k < 1) k = 1;
  if (k == g_threads) return;
  if (g_threads > 1) { g_quit = 1; pthread_barrier_wait(&g_bar); for (int i = 1; i < g_threads; ++i) pthread_join(g_pool[i], NULL); pthread_barrier_destroy(&g_bar); free(g_pool); g_pool = NULL; g_quit = 0; }
  g_threads = k;
  if (k > 1) { pthread_barrier_init(&g_bar, NULL, (unsigned)k); g_pool = (pthread_t*)calloc((size_t)k, sizeof(pthread_t));
    for (int i = 1; i < k; ++i) pthread_create(&g_pool[i], NULL, pool_worker, (void*)(intptr_t)i); }
}
int tmc_port_get_threads(void) { return g_threads; }
#define PAR_MIN_NNZ 200000        /* smaller loops stay serial: two barrier crossings cost more than they do */

/* ---- symmetric tridiagonal: largest eigenvalue by Sturm bisection, eigenvector by twisted factorisation */
static int sturm(const double* a, const double* b, int j, double x) {
  int c = 0; double d = a[0] - x; if (fabs(d) < 1e-290) d = -1e-290; if (d < 0) c++;
  for (int i = 1; i < j; ++i) { d = (a[i] - x) - b[i] * b[i] / d; if (fabs(d) < 1e-290) d = -1e-290; if (d < 0) c++; }
  return c;
}
static double top_eig(const double* a, const double* b, int j) {
  double lo = a[0], hi = a[0];
  for (int i = 0; i < j; ++i) {
    double r = (i > 0 ? fabs(b[i]) : 0) + (i + 1 < j ? fabs(b[i + 1]) : 0);
    if (a[i] - r < lo) lo = a[i] - r;
    if (a[i] + r > hi) hi = a[i] + r;
  }
  double span = fmax(fabs(lo), fabs(hi)); lo -= 1e-14 * span + 1e-300; hi += 1e-14 * span + 1e-300;
  for (int it = 0; it < 200; ++it) { double mid = 0.5 * (lo + hi); if (mid <= lo || mid >= hi) break; if (sturm(a, b, j, mid) >= j) hi = mid; else lo = mid; }
  return 0.5 * (lo + hi);
}
static void twisted(const double* a, const double* b, int j, double th, double* dp, double* dm, double* z) {
  if (j == 1) { z[0] = 1; return; }
  const double piv = 1e-290;
  dp[0] = a[0] - th;
  for (int i = 1; i < j; ++i) { double p = dp[i - 1]; if (fabs(p) < piv) p = p < 0 ? -piv : piv; dp[i - 1] = p; dp[i] = (a[i] - th) - b[i] * b[i] / p; }
  dm[j - 1] = a[j - 1] - th;
  for (int i = j - 2; i >= 0; --i) { double p = dm[i + 1]; if (fabs(p) < piv) p = p < 0 ? -piv : piv; dm[i + 1] = p; dm[i] = (a[i] - th) - b[i + 1] * b[i + 1] / p; }
  int r = 0; double best = INFINITY;
  for (int i = 0; i < j; ++i) { double g = fabs(dp[i] + dm[i] - (a[i] - th)); if (g < best) { best = g; r = i; } }
  z[r] = 1;
  for (int i = r - 1; i >= 0; --i) { double p = dp[i]; if (fabs(p) < piv) p = p < 0 ? -piv : piv; z[i] = -b[i + 1] / p * z[i + 1]; }
  for (int i = r + 1; i < j; ++i) { double p = dm[i]; if (fabs(p) < piv) p = p < 0 ? -piv : piv; z[i] = -b[i] / p * z[i - 1]; }
  double nn = sqrt(dot(z, z, j)); for (int i = 0; i < j; ++i) z[i] /= nn;
}

/* ---- one node: materialise B_S, degree, Lanczos for u2, labels, Q.  Returns Lanczos steps. */
typedef struct { double* y; double* w; double* d; double* s; double* u1; double* V; double* u2; double *al, *be, *z, *dp, *dm, *h; int kmax; } work_t;

typedef struct { const csr_t* a; const double* in; double* out; double* part; } mv_arg;
static double* g_ypart = NULL; static int64_t g_ypart_n = 0;             /* nthreads private copies of y for the scatter */
static void at_x_job(int tid, int nt, void* p) {                          /* row block tid scatters into its own copy of y */
  mv_arg* g = (mv_arg*)p; const csr_t* a = g->a;
  double* yt = g->part + (size_t)tid * a->n;
  memset(yt, 0, sizeof(double) * a->n);
  const int64_t lo = a->m * tid / nt, hi = a->m * (tid + 1) / nt;
  for (int64_t i = lo; i < hi; ++i) { double xi = g->in[i]; if (xi == 0) continue; for (int64_t k = a->ptr[i]; k < a->ptr[i + 1]; ++k) yt[a->col[k]] += a->val[k] * xi; }
}
static void at_x_sum_job(int tid, int nt, void* p) {
  mv_arg* g = (mv_arg*)p; const int64_t n = g->a->n, lo = n * tid / nt, hi = n * (tid + 1) / nt;
  for (int64_t j = lo; j < hi; ++j) { double s = 0; for (int q = 0; q < nt; ++q) s += g->part[(size_t)q * n + j]; g->out[j] = s; }
}
static void a_y_job(int tid, int nt, void* p) {
  mv_arg* g = (mv_arg*)p; const csr_t* a = g->a; const int64_t lo = a->m * tid / nt, hi = a->m * (tid + 1) / nt;
  for (int64_t i = lo; i < hi; ++i) { double s = 0; for (int64_t k = a->ptr[i]; k < a->ptr[i + 1]; ++k) s += a->val[k] * g->in[a->col[k]]; g->out[i] = s; }
}
static void at_x(const csr_t* a, const double* x, double* y) {          /* y = A^T x */
  if (g_threads > 1 && a->nnz >= PAR_MIN_NNZ && g_ypart && g_ypart_n >= a->n) {
    mv_arg g = {a, x, y, g_ypart}; run_par(at_x_job, &g); run_par(at_x_sum_job, &g); return;
  }
  memset(y, 0, sizeof(double) * a->n);
  for (int64_t i = 0; i < a->m; ++i) { double xi = x[i]; if (xi == 0) continue; for (int64_t k = a->ptr[i]; k < a->ptr[i + 1]; ++k) y[a->col[k]] += a->val[k] * xi; }
}
static void a_y(const csr_t* a, const double* y, double* x) {           /* x = A y */
  if (g_threads > 1 && a->nnz >= PAR_MIN_NNZ) { mv_arg g = {a, y, x, NULL}; run_par(a_y_job, &g); return; }
  for (int64_t i = 0; i < a->m; ++i) { double s = 0; for (int64_t k = a->ptr[i]; k < a->ptr[i + 1]; ++k) s += a->val[k] * y[a->col[k]]; x[i] = s; }
}
/* re-orthogonalisation in two parallel regions: h = V^T w (per-thread partial sums), then w -= V h */
typedef struct { const double* V; double* w; double* h; double* hpart; int64_t m; int j; } cgs_arg;
static void cgs_dots_job(int tid, int nt, void* p) {
  cgs_arg* g = (cgs_arg*)p; const int64_t lo = g->m * tid / nt, hi = g->m * (tid + 1) / nt;
  for (int l = 0; l <= g->j; ++l) { const double* vl = g->V + (int64_t)l * g->m; double s = 0; for (int64_t i = lo; i < hi; ++i) s += vl[i] * g->w[i]; g->hpart[(size_t)tid * (g->j + 1) + l] = s; }
}
static void cgs_update_job(int tid, int nt, void* p) {
  cgs_arg* g = (cgs_arg*)p; const int64_t lo = g->m * tid / nt, hi = g->m * (tid + 1) / nt;
  for (int l = 0; l <= g->j; ++l) { const double hl = g->h[l]; const double* vl = g->V + (int64_t)l * g->m; for (int64_t i = lo; i < hi; ++i) g->w[i] -= hl * vl[i]; }
}
static uint64_t mix(uint64_t x) { x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull; x = (x ^ (x >> 27)) * 0x94D049BB133111EBull; return x ^ (x >> 31); }

static int node_u2(const csr_t* bs, const int64_t* rows, work_t* wk, double tol, double* theta_out) {
  const int64_t m = bs->m;
  double* ones = wk->w;
  for (int64_t i = 0; i < m; ++i) ones[i] = 1.0;
  at_x(bs, ones, wk->y); a_y(bs, wk->y, wk->d);                          /* d = B (B^T 1) */
  double sumd = 0; for (int64_t i = 0; i < m; ++i) sumd += wk->d[i];
  /* C = D^{-1/2} B is materialised like bdToC does */
  csr_t c = *bs; c.val = (double*)malloc(sizeof(double) * (bs->nnz > 0 ? bs->nnz : 1));
  for (int64_t i = 0; i < m; ++i) { wk->s[i] = 1.0 / sqrt(wk->d[i]); wk->u1[i] = sqrt(wk->d[i] / sumd);
    for (int64_t k = bs->ptr[i]; k < bs->ptr[i + 1]; ++k) c.val[k] = bs->val[k] * wk->s[i]; }
  double* V = wk->V;                                                     /* V[0] = u1, V[l] = q_l */
  memcpy(V, wk->u1, sizeof(double) * m);
  double* q = V + m;
  for (int64_t i = 0; i < m; ++i) q[i] = (double)(mix(0x2545F4914F6CDD1Dull + (uint64_t)(rows[i] + 1) * 0x9E3779B97F4A7C15ull) >> 11) * (2.0 / 9007199254740992.0) - 1.0;
  double h0 = dot(V, q, m); for (int64_t i = 0; i < m; ++i) q[i] -= h0 * V[i];
  double nq = sqrt(dot(q, q, m)); for (int64_t i = 0; i < m; ++i) q[i] /= nq;
  int j = 1, steps = 0; double theta = 0;
  for (;;) {
    double* qj = V + (int64_t)j * m; double* w = wk->w;
    at_x(&c, qj, wk->y); a_y(&c, wk->y, w); steps++;                     /* w = C C^T q_j : the two mat-vecs of las2's opb */
    double alpha = 0;
    for (int pass = 0; pass < 2; ++pass) {                               /* full re-orthogonalisation, twice */
      if (g_threads > 1 && m * (int64_t)(j + 1) >= PAR_MIN_NNZ) {
        double* hpart = (double*)malloc(sizeof(double) * (size_t)g_threads * (j + 1));
        cgs_arg g = {V, w, wk->h, hpart, m, j};
        run_par(cgs_dots_job, &g);
        for (int l = 0; l <= j; ++l) { double s2 = 0; for (int q = 0; q < g_threads; ++q) s2 += hpart[(size_t)q * (j + 1) + l]; wk->h[l] = s2; }
        run_par(cgs_update_job, &g);
        free(hpart);
      } else {
        for (int l = 0; l <= j; ++l) wk->h[l] = dot(V + (int64_t)l * m, w, m);
        for (int l = 0; l <= j; ++l) { const double hl = wk->h[l]; const double* vl = V + (int64_t)l * m; for (int64_t i = 0; i < m; ++i) w[i] -= hl * vl[i]; }
      }
      alpha += wk->h[j];
    }
    double beta = sqrt(dot(w, w, m));
    wk->al[j - 1] = alpha; wk->be[j] = beta; if (j == 1) wk->be[0] = 0;
    theta = top_eig(wk->al, wk->be, j);
    twisted(wk->al, wk->be, j, theta, wk->dp, wk->dm, wk->z);
    double resid = beta * fabs(wk->z[j - 1]);
    if (j >= m - 1 || beta <= 1e-13 || resid <= tol * theta || resid <= 1e-14 || j >= wk->kmax) break;
    double* qn = V + (int64_t)(j + 1) * m; for (int64_t i = 0; i < m; ++i) qn[i] = w[i] / beta;
    j++;
  }
  for (int64_t i = 0; i < m; ++i) { double u = 0; for (int l = 1; l <= j; ++l) u += wk->z[l - 1] * V[(int64_t)l * m + i]; wk->u2[i] = u; }
  free(c.val);
  *theta_out = theta;
  return steps;
}

static double modularity(const csr_t* bs, const uint8_t* lab, work_t* wk) {
  const int64_t m = bs->m;
  double* ones = wk->w; for (int64_t i = 0; i < m; ++i) ones[i] = 1.0;
  at_x(bs, ones, wk->y); a_y(bs, wk->y, wk->d);
  double L = -(double)m; for (int64_t i = 0; i < m; ++i) L += wk->d[i];
  if (L == 0) return 0;
  double q = 0;
  for (int c = 0; c < 2; ++c) {
    double nc = 0, lc = 0;
    for (int64_t i = 0; i < m; ++i) { ones[i] = (lab[i] == c) ? 1.0 : 0.0; nc += ones[i]; lc += ones[i] * wk->d[i]; }
    at_x(bs, ones, wk->y);
    double oc = dot(wk->y, wk->y, bs->n) - nc; lc -= nc;
    q += oc / L - (lc / L) * (lc / L);
  }
  return q;
}

static void tree_push(tree_t* t, int64_t b, int64_t e) {
  if (t->n_nodes == t->cap) {
    t->cap = t->cap ? 2 * t->cap : 1024;
    t->left = realloc(t->left, 8 * t->cap); t->right = realloc(t->right, 8 * t->cap); t->begin = realloc(t->begin, 8 * t->cap);
    t->end = realloc(t->end, 8 * t->cap); t->q = realloc(t->q, 8 * t->cap); t->theta = realloc(t->theta, 8 * t->cap); t->iters = realloc(t->iters, 4 * t->cap);
  }
  int64_t i = t->n_nodes++; t->left[i] = t->right[i] = -1; t->begin[i] = b; t->end[i] = e; t->q[i] = 0; t->theta[i] = 0; t->iters[i] = 0;
}

/* Entry point.  b: row-L2-normalised B (the caller does b1ToB2 / b2ToB with numpy, as loadAllSSM does before the engine).
 * perm (m ints, out): rows grouped by node, ascending inside leaves.  Node arrays are returned through tree_t (pre-order).
 * Returns total Lanczos steps (= pairs of mat-vecs); *work_nnz accumulates nnz touched by those pairs. */
int64_t tmc_port_make_tree(int64_t m, int64_t n, const int64_t* ptr, const int32_t* col, const double* val, double min_mod, double tol,
                           int kmax, int64_t* perm, tree_t* t, double* work_nnz) {
  memset(t, 0, sizeof(*t));
  for (int64_t i = 0; i < m; ++i) perm[i] = i;
  work_t wk; wk.kmax = kmax < 2 ? 2 : kmax;
  wk.y = malloc(8 * (n + 1)); wk.w = malloc(8 * (m + 1)); wk.d = malloc(8 * (m + 1)); wk.s = malloc(8 * (m + 1)); wk.u1 = malloc(8 * (m + 1));
  wk.u2 = malloc(8 * (m + 1)); wk.al = malloc(8 * (wk.kmax + 2)); wk.be = malloc(8 * (wk.kmax + 2)); wk.z = malloc(8 * (wk.kmax + 2));
  wk.dp = malloc(8 * (wk.kmax + 2)); wk.dm = malloc(8 * (wk.kmax + 2)); wk.h = malloc(8 * (wk.kmax + 2));
  wk.V = NULL; int64_t vcap = 0;
  if (g_threads > 1) { free(g_ypart); g_ypart = (double*)malloc(sizeof(double) * (size_t)g_threads * (n + 1)); g_ypart_n = n; }
  int64_t total_steps = 0; *work_nnz = 0;
  uint8_t* lab = malloc(m + 1); int64_t* tmp = malloc(8 * (m + 1));
  /* explicit pre-order stack: push root; children are appended right after their parent is processed, left first */
  int64_t* stack = malloc(8 * (2 * m + 2)); int64_t sp = 0;
  tree_push(t, 0, m); stack[sp++] = 0;
  while (sp) {
    const int64_t node = stack[--sp];
    const int64_t b = t->begin[node], e = t->end[node], ms = e - b;
    if (ms < 2) continue;
    /* materialise B_S (what extractRow / fromRowsL do in the reference, per node) */
    csr_t bs; bs.m = ms; bs.n = n; bs.ptr = malloc(8 * (ms + 1)); bs.ptr[0] = 0;
    for (int64_t i = 0; i < ms; ++i) bs.ptr[i + 1] = bs.ptr[i] + (ptr[perm[b + i] + 1] - ptr[perm[b + i]]);
    bs.nnz = bs.ptr[ms]; bs.col = malloc(4 * (bs.nnz + 1)); bs.val = malloc(8 * (bs.nnz + 1));
    for (int64_t i = 0; i < ms; ++i) { const int64_t r = perm[b + i], k0 = ptr[r], len = ptr[r + 1] - k0;
      memcpy(bs.col + bs.ptr[i], col + k0, 4 * len); memcpy(bs.val + bs.ptr[i], val + k0, 8 * len); }
    const int64_t kdim = (ms < wk.kmax ? ms : wk.kmax) + 3;
    if (kdim * ms > vcap) { free(wk.V); vcap = kdim * ms; wk.V = malloc(8 * vcap); }
    double theta = 0;
    const int steps = node_u2(&bs, perm + b, &wk, tol, &theta);
    total_steps += steps; *work_nnz += (double)steps * (double)bs.nnz;
    int64_t n1 = 0;
    for (int64_t i = 0; i < ms; ++i) { lab[i] = wk.u2[i] >= 0 ? 1 : 0; n1 += lab[i]; }
    const double q = modularity(&bs, lab, &wk);
    t->q[node] = q; t->theta[node] = theta; t->iters[node] = steps;
    free(bs.ptr); free(bs.col); free(bs.val);
    if (q > min_mod && n1 > 0 && n1 < ms) {
      int64_t z = 0, o = ms - n1;                                         /* stable split: [label 0 | label 1] */
      for (int64_t i = 0; i < ms; ++i) { if (lab[i]) tmp[o++] = perm[b + i]; else tmp[z++] = perm[b + i]; }
      memcpy(perm + b, tmp, 8 * ms);
      const int64_t li = t->n_nodes; tree_push(t, b, b + (ms - n1)); tree_push(t, b + (ms - n1), e);
      t->left[node] = li; t->right[node] = li + 1;
      stack[sp++] = li + 1; stack[sp++] = li;                              /* left is processed first */
    }
  }
  free(wk.y); free(wk.w); free(wk.d); free(wk.s); free(wk.u1); free(wk.u2); free(wk.al); free(wk.be); free(wk.z); free(wk.dp); free(wk.dm);
  free(wk.h); free(wk.V); free(lab); free(tmp); free(stack);
  free(g_ypart); g_ypart = NULL; g_ypart_n = 0;
  return total_steps;
}

void tmc_port_tree_free(tree_t* t) { free(t->left); free(t->right); free(t->begin); free(t->end); free(t->q); free(t->theta); free(t->iters); memset(t, 0, sizeof(*t)); }
