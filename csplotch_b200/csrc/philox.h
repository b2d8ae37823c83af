// Philox4x32-10 keyed stream of the sampler.
//   key     = (seed, chain_id)
//   counter = (index, iteration, purpose << 24 | sub, gene_id)
// The stream is a function of (gene, chain, iteration, purpose, element) only, so a draw never
// depends on which GPU, CTA or thread produced it, nor on how genes were sharded.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define CSP_HD __host__ __device__ __forceinline__
#else
#define CSP_HD inline
#endif

namespace csp {

enum Purpose : uint32_t {
  PURP_INIT = 1,         // uniform(-R,R) initial point; iteration field = attempt
  PURP_MOMENTUM = 2,     // momentum refresh of a transition
  PURP_SS_MOMENTUM = 3,  // momentum draws inside the step-size search; sub = attempt
  PURP_DIRECTION = 4,    // tree doubling direction; index = depth
  PURP_ACCEPT = 5,       // top-level biased progressive sampling; index = depth
  PURP_RESERVOIR = 6     // in-subtree multinomial (reservoir) sampling; index = leapfrog serial
};

CSP_HD void philox4x32_10(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                          uint32_t out[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    const uint32_t n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    const uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// 53-bit uniform in the open interval (0,1)
CSP_HD double u53(uint32_t hi, uint32_t lo) {
  const uint64_t x = (((uint64_t)hi << 32) | lo) >> 11;
  return ((double)x + 0.5) * (1.0 / 9007199254740992.0);
}

struct RngKey {
  uint32_t seed, chain, gene;
};

CSP_HD void rng_u2(const RngKey &r, uint32_t purpose, uint32_t sub, uint32_t iter, uint32_t idx, double &u0,
                   double &u1) {
  uint32_t o[4];
  philox4x32_10(r.seed, r.chain, idx, iter, (purpose << 24) | (sub & 0xFFFFFFu), r.gene, o);
  u0 = u53(o[0], o[1]);
  u1 = u53(o[2], o[3]);
}

CSP_HD double rng_u(const RngKey &r, uint32_t purpose, uint32_t sub, uint32_t iter, uint32_t idx) {
  double a, b;
  rng_u2(r, purpose, sub, iter, idx, a, b);
  return a;
}

// standard normal of element `elem` (Box-Muller on the pair elem>>1; lane elem&1)
CSP_HD double rng_normal(const RngKey &r, uint32_t purpose, uint32_t sub, uint32_t iter, uint32_t elem) {
  double u0, u1;
  rng_u2(r, purpose, sub, iter, elem >> 1, u0, u1);
  const double rad = sqrt(-2.0 * log(u0));
#if defined(__CUDA_ARCH__)
  double s, c;
  sincospi(2.0 * u1, &s, &c);
#else
  const double ang = 6.283185307179586476925286766559 * u1;
  const double s = sin(ang), c = cos(ang);
#endif
  return (elem & 1u) ? rad * s : rad * c;
}

// uniform(-R, R) initial value of element `elem` (pair elem>>1, lane elem&1)
CSP_HD double rng_init(const RngKey &r, uint32_t attempt, uint32_t elem, double radius) {
  double u0, u1;
  rng_u2(r, PURP_INIT, 0, attempt, elem >> 1, u0, u1);
  return radius * (2.0 * ((elem & 1u) ? u1 : u0) - 1.0);
}

}  // namespace csp
