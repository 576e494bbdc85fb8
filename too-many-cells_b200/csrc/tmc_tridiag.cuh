// Symmetric tridiagonal helpers for the batched Lanczos (K5 in SURVEY.md 7.1):
// top eigenvalue of T_j by Sturm-count multisection and its eigenvector by a
// twisted factorisation, O(j) per evaluation so a warp per tree node suffices.
// Replaces the imtqlb / imtql2 QL step of SVDLIBC las2, which
// the reference reaches through `sparseSvd` (twin idiom Matrix/Preprocess.hs:512-520).
//
// Conventions: a[0..j) diagonal (alpha_1..alpha_j), b[1..j) off-diagonals, b[i] couples
// rows i-1 and i (b[0] unused).  Compiles for host too (unit-tested on CPU, no GPU needed).
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define TMC_HD __host__ __device__ __forceinline__
#else
#define TMC_HD inline
#endif

// number of eigenvalues of T_j strictly below x
TMC_HD int tmc_sturm_count(const double* a, const double* b, int j, double x, double pivmin) {
  int cnt = 0;
  double dl = a[0] - x;
  if (fabs(dl) < pivmin) dl = -pivmin;
  if (dl < 0) cnt++;
  for (int i = 1; i < j; ++i) {
    dl = (a[i] - x) - b[i] * b[i] / dl;
    if (fabs(dl) < pivmin) dl = -pivmin;
    if (dl < 0) cnt++;
  }
  return cnt;
}

// Gershgorin bounds and pivmin
TMC_HD void tmc_gershgorin(const double* a, const double* b, int j, double* lo, double* hi, double* pivmin) {
  double l = a[0], h = a[0], bmax = 0.0;
  for (int i = 0; i < j; ++i) {
    double r = (i > 0 ? fabs(b[i]) : 0.0) + (i + 1 < j ? fabs(b[i + 1]) : 0.0);
    l = fmin(l, a[i] - r);
    h = fmax(h, a[i] + r);
    if (i > 0) bmax = fmax(bmax, b[i] * b[i]);
  }
  double span = fmax(fabs(l), fabs(h));
  *lo = l - 1e-14 * span - 1e-300;
  *hi = h + 1e-14 * span + 1e-300;
  *pivmin = fmax(2.2250738585072014e-308 * fmax(1.0, bmax) * 1e4, 1e-290);
}

// Eigenvector of T_j for eigenvalue theta by twisted factorisation (Parlett/Dhillon "getvec").
// dp, dm: scratch of j doubles each.  z (j) is normalised to unit 2-norm.
TMC_HD void tmc_twisted_vector(const double* a, const double* b, int j, double theta,
                               double* dp, double* dm, double* z, double pivmin) {
  if (j == 1) { z[0] = 1.0; return; }
  // forward pivots D+_i
  dp[0] = a[0] - theta;
  for (int i = 1; i < j; ++i) {
    double p = dp[i - 1];
    if (fabs(p) < pivmin) p = (p < 0 ? -pivmin : pivmin);
    dp[i - 1] = p;
    dp[i] = (a[i] - theta) - b[i] * b[i] / p;
  }
  // backward pivots D-_i
  dm[j - 1] = a[j - 1] - theta;
  for (int i = j - 2; i >= 0; --i) {
    double p = dm[i + 1];
    if (fabs(p) < pivmin) p = (p < 0 ? -pivmin : pivmin);
    dm[i + 1] = p;
    dm[i] = (a[i] - theta) - b[i + 1] * b[i + 1] / p;
  }
  // twist index minimising |gamma_r|
  int r = 0;
  double best = INFINITY;
  for (int i = 0; i < j; ++i) {
    double g = dp[i] + dm[i] - (a[i] - theta);
    if (fabs(g) < best) { best = fabs(g); r = i; }
  }
  z[r] = 1.0;
  for (int i = r - 1; i >= 0; --i) {
    double p = dp[i];
    if (fabs(p) < pivmin) p = (p < 0 ? -pivmin : pivmin);
    z[i] = -b[i + 1] / p * z[i + 1];
  }
  for (int i = r + 1; i < j; ++i) {
    double p = dm[i];
    if (fabs(p) < pivmin) p = (p < 0 ? -pivmin : pivmin);
    z[i] = -b[i] / p * z[i - 1];
  }
  double nn = 0.0;
  for (int i = 0; i < j; ++i) nn += z[i] * z[i];
  nn = 1.0 / sqrt(nn);
  for (int i = 0; i < j; ++i) z[i] *= nn;
}

// Serial reference of the multisection (host tests + single-lane fallback): largest eigenvalue of T_j.
TMC_HD double tmc_top_eig_serial(const double* a, const double* b, int j, double lo_hint) {
  double lo, hi, pivmin;
  tmc_gershgorin(a, b, j, &lo, &hi, &pivmin);
  if (lo_hint > lo && lo_hint < hi && tmc_sturm_count(a, b, j, lo_hint, pivmin) < j) lo = lo_hint;
  for (int it = 0; it < 200; ++it) {
    double mid = 0.5 * (lo + hi);
    if (mid <= lo || mid >= hi) break;
    if (tmc_sturm_count(a, b, j, mid, pivmin) >= j) hi = mid; else lo = mid;
  }
  return 0.5 * (lo + hi);
}

#if defined(__CUDACC__)
// Warp-cooperative 32-way multisection: every lane evaluates one Sturm count per round.
// All 32 lanes must call; returns the same value on every lane.
// rel_stop > 0: stop as soon as the bracket is narrower than rel_stop * max(|lo|, |hi|) and return its UPPER end (a rigorous upper
// bound of the eigenvalue: what the sign certificate wants for theta_2, which it needs to three digits only)
__device__ __forceinline__ double tmc_top_eig_warp(const double* a, const double* b, int j, double lo_hint, double hi_hint = 1e300, int kth = 1,
                                                   double rel_stop = 0.0) {
  // kth = 1: the largest eigenvalue; kth = i: the i-th largest (smallest x with count(x) >= j - i + 1)
  const unsigned lane = threadIdx.x & 31u;
  const int target = j - kth + 1;
  double lo, hi, pivmin;
  tmc_gershgorin(a, b, j, &lo, &hi, &pivmin);
  if (lo_hint > lo && lo_hint < hi) {
    int c = tmc_sturm_count(a, b, j, lo_hint, pivmin);
    if (c < target) lo = lo_hint;
  }
  if (hi_hint > lo && hi_hint < hi) {        // e.g. previous theta + previous residual: usually a valid, much tighter upper bound
    int c = tmc_sturm_count(a, b, j, hi_hint, pivmin);
    if (c >= target) hi = hi_hint;
  }
  for (int round = 0; round < 40; ++round) {
    double w = (hi - lo) / 33.0;
    double x = lo + w * (double)(lane + 1);
    bool degenerate = !(x > lo && x < hi);
    int c = degenerate ? -1 : tmc_sturm_count(a, b, j, x, pivmin);
    unsigned above = __ballot_sync(0xffffffffu, (!degenerate) && c >= target);   // x is an upper bound
    unsigned valid = __ballot_sync(0xffffffffu, !degenerate);
    if (valid == 0u) break;
    double nlo = lo, nhi = hi;
    if (above) {
      int first = __ffs(above) - 1;             // smallest x that bounds from above
      nhi = __shfl_sync(0xffffffffu, x, first);
      // largest valid lane below `first` is a lower bound
      unsigned below = valid & ((1u << first) - 1u);
      if (below) { int last = 31 - __clz(below); nlo = __shfl_sync(0xffffffffu, x, last); }
    } else {
      int last = 31 - __clz(valid);
      nlo = __shfl_sync(0xffffffffu, x, last);
    }
    if (nlo == lo && nhi == hi) break;
    lo = nlo; hi = nhi;
    if (hi - lo <= fmax(4.0e-16, rel_stop) * fmax(fabs(lo), fabs(hi))) break;
  }
  return rel_stop > 0.0 ? hi : 0.5 * (lo + hi);
}
#endif
