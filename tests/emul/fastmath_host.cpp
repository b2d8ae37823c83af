// TEST INFRASTRUCTURE: host build of csrc/fastmath.h (the same source the kernels compile) for accuracy tests.
#include "fastmath.h"
extern "C" void fm_exp(const double *x, double *y, long n) {
  for (long i = 0; i < n; ++i) y[i] = csp::exp_fast(x[i]);
}
extern "C" double fm_logprod(const double *f, long n) {
  csp::LogProd lp;
  lp.init();
  for (long i = 0; i < n; ++i) lp.mul(f[i]);
  return lp.log_value();
}
