// Host-compiled shim around the product's host/device tridiagonal helpers so they
// can be unit-tested on CPU (no GPU) against LAPACK.
#include "../../too-many-cells_b200/csrc/tmc_tridiag.cuh"
extern "C" {
double th_top_eig(const double* a, const double* b, int j, double lo_hint) { return tmc_top_eig_serial(a, b, j, lo_hint); }
void th_vector(const double* a, const double* b, int j, double theta, double* dp, double* dm, double* z) {
  double lo, hi, piv; tmc_gershgorin(a, b, j, &lo, &hi, &piv);
  tmc_twisted_vector(a, b, j, theta, dp, dm, z, piv);
}
int th_sturm(const double* a, const double* b, int j, double x) {
  double lo, hi, piv; tmc_gershgorin(a, b, j, &lo, &hi, &piv);
  return tmc_sturm_count(a, b, j, x, piv);
}
}
