// scan.cuh -- argument block shared by the reverse-scan kernels (scan.cu, scan_tma.cu).
#pragma once
#include "common.cuh"

namespace mrl {

enum { OP_DR = 0, OP_GAE = 1, OP_AFFINE = 2 };

struct ScanArgs {
  const float *a;       // reward (DR, GAE) / a (AFFINE)
  const float *v;       // value (GAE) / b (AFFINE)
  const float *nv;      // next_value (GAE) or NULL
  const float *last;    // last_value / x_last or NULL
  const uint8_t *done;  // or NULL
  float *out;           // returns / advantages / x
  float *ret;           // GAE returns or NULL
  int64_t T, B, E, N;   // N = B*E columns
  int64_t st, sc;       // element strides along t and along the column index
  int64_t dst, dsc;     // same for done (indexed by column / E)
  float gamma, gl;      // gl = fl(gamma*lambda)
  // != NULL: the kernel also accumulates (sum, sum of squares) of everything it stores to `out` (fp32 over a
  // lane's 8..32 consecutive steps, float64 beyond that) in a fixed order, one partial per CTA at mom[blockIdx.x] -- the moments PPO / A2C normalise with right
  // after the scan (ppo.py:361-368, a2c.py:190-192) without another pass over the array.  Honoured by the
  // TMA kernels (scan_tma.cu); the launchers report how many partials were written.
  double2 *mom;
};
constexpr int kMomentSlots = 4096;  // most CTAs a moment-producing launch may have

__device__ __forceinline__ float scan_step(float a, float b, float x) {
  return __fadd_rn(a, __fmul_rn(b, x));
}

// TMA-ring kernels (scan_tma.cu).  Return MRL_OK after launching, or kScanNotApplicable when the shape
// or the alignment does not fit (the caller then takes the register kernels of scan.cu).
constexpr int kScanNotApplicable = -1000;
// *mom_parts (when A.mom != NULL): number of partials the launch writes, 0 = this launch does not produce them
int scan_tm_strip_launch(const ScanArgs &A, int op, bool force, cudaStream_t st, int *mom_parts);
int scan_bm_row_launch(const ScanArgs &A, int op, bool force, cudaStream_t st, int *mom_parts);

}  // namespace mrl
