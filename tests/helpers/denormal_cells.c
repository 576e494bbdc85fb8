/* Host check of the arithmetic the dense-panel kernels rely on (tmc_kernels.cuh, "dense panel halves"): a stored count b used
 * as the denormal double with bit pattern (hi = 0, lo = b) -- or, for nibble k, lo = b << 4k -- multiplied inside an FMA with the
 * operand pre-scaled by 2^960 (and 16^-k), accumulated, and scaled back by 2^114, gives bit for bit the same sum as converting
 * the count to double first.  IEEE-754 fma is the same on the host and on the device.  Prints the number of mismatches. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

static double from_bits(uint64_t u) { double d; memcpy(&d, &u, 8); return d; }
static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static uint64_t rnd(void) { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return rng_state; }
static double rnd_operand(void) {             /* magnitudes from 1e-12 to 1e12, both signs */
  double m = (double)(rnd() >> 11) / 9007199254740992.0 + 0.5;
  int e = (int)(rnd() % 81) - 40;
  return ((rnd() & 1) ? -1.0 : 1.0) * ldexp(m, e);
}

int main(void) {
  const double UP = 0x1p960, DOWN = 0x1p114;
  long bad = 0;
  for (int trial = 0; trial < 2000; ++trial) {
    double ref = 0.0, acc_b = 0.0, acc_n[8] = {0, 0, 0, 0, 0, 0, 0, 0}, ref_n[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 400; ++i) {
      const double x = rnd_operand();
      const unsigned b = (rnd() % 4 == 0) ? (unsigned)(rnd() % 256) : (unsigned)(rnd() % 4);      /* byte cell */
      ref = fma((double)b, x, ref);
      acc_b = fma(from_bits((uint64_t)b), x * UP, acc_b);
      const unsigned w = (unsigned)rnd();                                                          /* eight nibble cells */
      for (int k = 0; k < 8; ++k) {
        const unsigned c = (w >> (4 * k)) & 15u;
        ref_n[k] = fma((double)c, x, ref_n[k]);
        /* T side: the accumulator carries 16^k; gather side: the operand is pre-scaled by 16^-k -- both exact powers of two */
        acc_n[k] = fma(from_bits((uint64_t)(w & (15u << (4 * k)))), (x * UP) * ldexp(1.0, -4 * k), acc_n[k]);
      }
    }
    if (acc_b * DOWN != ref) bad++;
    for (int k = 0; k < 8; ++k) if (acc_n[k] * DOWN != ref_n[k]) bad++;
    /* T-side variant: unscaled operand, accumulator divided by 16^k at the end */
    double acc_t[8] = {0, 0, 0, 0, 0, 0, 0, 0}, ref_t[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 400; ++i) {
      const double x = rnd_operand();
      const unsigned w = (unsigned)rnd();
      for (int k = 0; k < 8; ++k) {
        ref_t[k] = fma((double)((w >> (4 * k)) & 15u), x, ref_t[k]);
        acc_t[k] = fma(from_bits((uint64_t)(w & (15u << (4 * k)))), x * UP, acc_t[k]);
      }
    }
    for (int k = 0; k < 8; ++k) if (acc_t[k] * (DOWN / (double)(1u << (4 * k))) != ref_t[k]) bad++;
  }
  printf("%ld\n", bad);
  return bad != 0;
}
