/*
 * mrl_arith.h -- arithmetic that must be bit-identical on host (gcc) and device (nvcc).
 *
 * Part of the ABI contract of mindrl_b200.h:
 *
 *  mrl_detpowf(x, y): x**y for the two exponentiations of the PER path
 *     priority ** alpha        (PriorityReplayBufferUpdate)
 *     (p/sum * size) ** -beta  (importance weights of PriorityReplayBufferSample)
 *   Neither glibc's nor CUDA's pow is specified bit-for-bit, and a 1-ulp difference in a leaf
 *   priority changes node sums and therefore sampled indices.  This routine uses only IEEE-754
 *   double +,-,*,/ and fma in a fixed order, so gcc (-ffp-contract=off) and nvcc (-fmad=false)
 *   produce the same bits.  It is accurate to <1e-14 relative before the final rounding to
 *   float, i.e. it equals the correctly rounded (float)pow((double)x,(double)y) except for
 *   ~1e-7 of inputs, where it is off by one float ulp (checked in tests/test_oracle.py::test_detpow_equals_correctly_rounded_pow).
 *
 *  mrl_hash32 / mrl_uniform01 / mrl_uniform_below: counter-based generator used when the caller
 *   does not inject its own uniform draws (splitmix64 finaliser over seed + counter).
 *
 * Build rule for every TU that includes this file: no fast-math, no implicit fma contraction.
 */
#ifndef MRL_ARITH_H_
#define MRL_ARITH_H_

#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define MRL_HD __host__ __device__ __forceinline__
#else
#define MRL_HD static inline
#endif

#define MRL_MIN_PRIORITY 1e-7f /* non-positive p**alpha is clamped to this */

MRL_HD double mrl_ll_as_double(long long v) {
#if defined(__CUDA_ARCH__)
  return __longlong_as_double(v);
#else
  double d;
  memcpy(&d, &v, sizeof(d));
  return d;
#endif
}

MRL_HD long long mrl_double_as_ll(double d) {
#if defined(__CUDA_ARCH__)
  return __double_as_longlong(d);
#else
  long long v;
  memcpy(&v, &d, sizeof(v));
  return v;
#endif
}

MRL_HD float mrl_detpowf(float xf, float yf) {
  if (yf == 0.0f) return 1.0f;
  if (xf != xf || yf != yf) return xf + yf; /* NaN */
  if (xf < 0.0f) return (xf - xf) / (xf - xf); /* NaN: negative base is outside the contract */
  if (xf == 0.0f) return yf > 0.0f ? 0.0f : INFINITY;
  if (xf > 3.4028234663852886e38f) return yf > 0.0f ? INFINITY : 0.0f;
  if (xf == 1.0f) return 1.0f;

  const double LN2 = 0.6931471805599453094;
  const double LN2_HI = 6.93147180369123816490e-01;
  const double LN2_LO = 1.90821492927058770002e-10;
  const double INV_LN2 = 1.4426950408889634074;

  /* x = m * 2^e, m in (sqrt(1/2), sqrt(2)] ; float subnormals are normal doubles */
  double x = (double)xf;
  long long bits = mrl_double_as_ll(x);
  int e = (int)((bits >> 52) & 0x7ff) - 1023;
  double m = mrl_ll_as_double((bits & 0x000fffffffffffffLL) | 0x3ff0000000000000LL);
  if (m > 1.4142135623730951) {
    m = m * 0.5;
    e = e + 1;
  }
  /* log(m) = 2 atanh(s), s = (m-1)/(m+1), |s| <= 0.1716 */
  double f = m - 1.0;
  double s = f / (2.0 + f);
  double z = s * s;
  double p = 1.0 / 25.0;
  p = fma(p, z, 1.0 / 23.0);
  p = fma(p, z, 1.0 / 21.0);
  p = fma(p, z, 1.0 / 19.0);
  p = fma(p, z, 1.0 / 17.0);
  p = fma(p, z, 1.0 / 15.0);
  p = fma(p, z, 1.0 / 13.0);
  p = fma(p, z, 1.0 / 11.0);
  p = fma(p, z, 1.0 / 9.0);
  p = fma(p, z, 1.0 / 7.0);
  p = fma(p, z, 1.0 / 5.0);
  p = fma(p, z, 1.0 / 3.0);
  p = fma(p, z, 1.0);
  double logm = (2.0 * s) * p;
  double lg = fma((double)e, LN2, logm);
  double t = (double)yf * lg;
  if (t > 89.5) return INFINITY;
  if (t < -104.5) return 0.0f;

  /* exp(t) = 2^k * exp(r), |r| <= ln2/2 */
  double kd = floor(fma(t, INV_LN2, 0.5));
  double r = fma(-kd, LN2_HI, t);
  r = fma(-kd, LN2_LO, r);
  double q = 1.0 / 87178291200.0; /* 1/14! */
  q = fma(q, r, 1.0 / 6227020800.0);
  q = fma(q, r, 1.0 / 479001600.0);
  q = fma(q, r, 1.0 / 39916800.0);
  q = fma(q, r, 1.0 / 3628800.0);
  q = fma(q, r, 1.0 / 362880.0);
  q = fma(q, r, 1.0 / 40320.0);
  q = fma(q, r, 1.0 / 5040.0);
  q = fma(q, r, 1.0 / 720.0);
  q = fma(q, r, 1.0 / 120.0);
  q = fma(q, r, 1.0 / 24.0);
  q = fma(q, r, 1.0 / 6.0);
  q = fma(q, r, 0.5);
  q = fma(q, r, 1.0);
  q = fma(q, r, 1.0);
  long long k = (long long)kd; /* in [-151, 130] */
  double scale = mrl_ll_as_double((k + 1023LL) << 52);
  return (float)(q * scale); /* IEEE round-to-nearest-even; overflows to inf, underflows to subnormal/0 */
}

/* priority ** alpha with the reference's clamp of non-positive results */
MRL_HD float mrl_priority_pow(float priority, float alpha) {
  float p = mrl_detpowf(priority, alpha);
  if (!(p > 0.0f)) p = MRL_MIN_PRIORITY;
  return p;
}

/* importance weight before normalisation: (priority / sum * size) ** -beta, fp32 steps */
MRL_HD float mrl_is_weight(float priority, float sum, float size, float beta) {
  float prob = priority / sum;
  float scaled = prob * size;
  return mrl_detpowf(scaled, -beta);
}

MRL_HD uint32_t mrl_hash32(uint64_t seed, uint64_t ctr) {
  uint64_t zz = seed + 0x9E3779B97F4A7C15ULL * (ctr + 1ULL);
  zz = (zz ^ (zz >> 30)) * 0xBF58476D1CE4E5B9ULL;
  zz = (zz ^ (zz >> 27)) * 0x94D049BB133111EBULL;
  zz = zz ^ (zz >> 31);
  return (uint32_t)(zz >> 32);
}

/* float in [0,1) with 24 random bits */
MRL_HD float mrl_uniform01(uint64_t seed, uint64_t ctr) {
  return (float)(mrl_hash32(seed, ctr) >> 8) * (1.0f / 16777216.0f);
}

/* integer in [0, bound), bound < 2^31 (multiply-shift) */
MRL_HD int64_t mrl_uniform_below(uint64_t seed, uint64_t ctr, int64_t bound) {
  return (int64_t)(((uint64_t)mrl_hash32(seed, ctr) * (uint64_t)bound) >> 32);
}

MRL_HD uint64_t mrl_mix_seed(uint64_t seed0, uint64_t seed1) {
  return (seed0 << 32) ^ seed1 ^ 0xD1B54A32D192ED03ULL;
}

/* synthetic row filler shared by mrl_fill_rows and the oracle */
MRL_HD uint8_t mrl_fill_byte(uint64_t seed, int64_t row, int64_t off) {
  return (uint8_t)(mrl_hash32(seed ^ (uint64_t)row * 0x100000001B3ULL, (uint64_t)(off >> 2)) >>
                   (8 * (off & 3)));
}

#endif /* MRL_ARITH_H_ */
