// fastmath.h — fp64 exp and reciprocal for the per-spot likelihood, built for short dependency chains.
//
// The spot loop is bound by instruction issue and by the latency of dependent fp64 chains (16-32 warps per SM,
// one DFMA every ~8 cycles per warp).  CUDA's exp(double) is ~60 issued instructions with a 14-deep Horner chain
// plus special-case handling that never triggers for the arguments of this model.  exp_fast evaluates the same
// range reduction (x = n ln2 + r, |r| <= ln2/2) and a degree-13 Taylor polynomial in Estrin form (depth 5 instead
// of 13); faithful to < 2 ulp on [-708, 709] (tests/test_fastmath.py checks 4M arguments against libm), inf / 0 /
// NaN outside like exp().  The header also compiles for the host (CSP_HD) so the CPU tests can pin it.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

#include "philox.h"  // CSP_HD

namespace csp {

CSP_HD int hi_word(double x) {
#ifdef __CUDA_ARCH__
  return __double2hiint(x);
#else
  int64_t b;
  std::memcpy(&b, &x, 8);
  return (int)(b >> 32);
#endif
}
CSP_HD int lo_word(double x) {
#ifdef __CUDA_ARCH__
  return __double2loint(x);
#else
  int64_t b;
  std::memcpy(&b, &x, 8);
  return (int)(uint32_t)b;
#endif
}
CSP_HD double from_words(int hi, int lo) {
#ifdef __CUDA_ARCH__
  return __hiloint2double(hi, lo);
#else
  const int64_t b = ((int64_t)hi << 32) | (uint32_t)lo;
  double x;
  std::memcpy(&x, &b, 8);
  return x;
#endif
}

CSP_HD double exp_fast(double x) {
  // outside the range where 2^n is a normal number: the library routine (gradual underflow, inf, NaN)
  if (!(fabs(x) < 708.0)) return exp(x);
  const double kMagic = 6755399441055744.0;  // 1.5 * 2^52: adding it rounds to the nearest integer, held in the low word
  const double t = fma(x, 1.4426950408889634074, kMagic);
  const int n = lo_word(t);
  const double nf = t - kMagic;
  double r = fma(nf, -6.93147180369123816490e-01, x);  // ln2 = hi + lo, hi with 32 trailing zero bits
  r = fma(nf, -1.90821492927058770002e-10, r);
  // exp(r) = 1 + r + r^2 (c2 + c3 r + ... + c13 r^11), Estrin
  const double r2 = r * r, r4 = r2 * r2, r8 = r4 * r4;
  const double q01 = fma(r, 1.0 / 6.0, 0.5);
  const double q23 = fma(r, 1.0 / 120.0, 1.0 / 24.0);
  const double q45 = fma(r, 1.0 / 5040.0, 1.0 / 720.0);
  const double q67 = fma(r, 1.0 / 362880.0, 1.0 / 40320.0);
  const double q89 = fma(r, 1.0 / 39916800.0, 1.0 / 3628800.0);
  const double qab = fma(r, 1.0 / 6227020800.0, 1.0 / 479001600.0);
  const double q0123 = fma(r2, q23, q01);
  const double q4567 = fma(r2, q67, q45);
  const double q89ab = fma(r2, qab, q89);
  const double qlo = fma(r4, q4567, q0123);
  const double q = fma(r8, q89ab, qlo);
  const double p = fma(r2, q, r) + 1.0;  // in (0.70, 1.42)
  return from_words(hi_word(p) + (n << 20), lo_word(p));
}

// 1 / x for a normal, positive x that is not close to the overflow / underflow thresholds (the caller's x is a
// probability-like quantity in (1e-300, 2]): hardware seed + two Newton steps, <= 1 ulp, no special-case branches
CSP_HD double rcp_fast(double x) {
#ifdef __CUDA_ARCH__
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  return fma(r, e, r);
#else
  return 1.0 / x;
#endif
}

// sum of logs as the log of a running product: mant in [1, 2) and an integer exponent, renormalised after every
// factor (5 integer instructions instead of a ~100-instruction log per factor).  Factors must be normal positive
// numbers; log_value() = log(mant) + expo ln2 once at the end.
struct LogProd {
  double mant;
  int expo;
  CSP_HD void init() { mant = 1.0; expo = 0; }
  CSP_HD void mul(double f) {
    const double p = mant * f;
    const int hi = hi_word(p);
    expo += (hi >> 20) - 1023;
    mant = from_words((hi & 0x000fffff) | 0x3ff00000, lo_word(p));
  }
  CSP_HD double log_value() const { return log(mant) + (double)expo * 0.693147180559945309417; }
};

}  // namespace csp
