// scan.cu -- reverse first-order recurrences x[t] = a[t] + b[t]*x[t+1] over trajectory batches.
//
// Replaces rl_ops.DiscountedReturn (discounted_return.py:71,82; CPU loop :84-92) and the in-learner
// graph-mode `while` loops of PPO (ppo.py:320-350), MAPPO (mappo.py:478-497), V-trace
// (vtrace.py:88-93) and TD(lambda) (coma.py:489-501).
//
// Kernels (all HBM-bound, no tensor cores: nothing here is a contraction)
//   scan_seq_kernel         thread per column, t = T-1..0, 8 rows of loads in flight per thread.
//                           Same operation order as the reference loop => bit-identical to the
//                           oracle.  Coalesced for time-major data.
//   scan_tm_chunked_kernel  time-major: CTA = 32 columns x W time-chunks.  Pass 1 reduces each chunk
//                           to an affine map, the maps are combined through shared memory, pass 2
//                           replays the chunk from its carry (second read hits L2).  Gives T-parallel
//                           work when B alone cannot fill 148 SMs (4096 envs = 128 warps).
//   scan_bm_warp_kernel     batch-major: warp per environment, 128 steps per tile (float4 per lane),
//                           per-lane composite + 5-step shuffle suffix scan of affine maps; a done flag
//                           makes b = 0, which cuts the segment inside the same scan.
// The two parallel kernels re-associate fp32 (spec: <= 1e-5 relative); MRL_SCAN_SEQUENTIAL selects the
// bit-comparable kernel.
#include <stdlib.h>

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
};

template <int OP>
__device__ __forceinline__ float scan_init(const ScanArgs &A, int64_t col) {
  if (OP == OP_GAE) return 0.0f;
  return A.last ? __ldg(A.last + col) : 0.0f;
}

// one element: produces (a, b) of x = a + b*x_next and the value needed by the store
template <int OP>
__device__ __forceinline__ void scan_load(const ScanArgs &A, int64_t t, int64_t col, float &a,
                                          float &b, float &aux) {
  const int64_t i = t * A.st + col * A.sc;
  if (OP == OP_AFFINE) {
    a = __ldg(A.a + i);
    b = __ldg(A.v + i);
    aux = 0.0f;
    return;
  }
  float m = 1.0f;
  if (A.done) {
    const int64_t bcol = A.E == 1 ? col : col / A.E;
    m = __fsub_rn(1.0f, __ldg(A.done + t * A.dst + bcol * A.dsc) ? 1.0f : 0.0f);
  }
  const float r = __ldg(A.a + i);
  if (OP == OP_DR) {
    a = r;
    b = __fmul_rn(m, A.gamma);
    aux = 0.0f;
  } else {
    const float v = __ldg(A.v + i);
    float nv;
    if (A.nv) nv = __ldg(A.nv + i);
    else if (t == A.T - 1) nv = A.last ? __ldg(A.last + col) : 0.0f;
    else nv = __ldg(A.v + i + A.st);
    const float gvm = __fmul_rn(__fmul_rn(A.gamma, nv), m);
    a = __fsub_rn(__fadd_rn(r, gvm), v);
    b = __fmul_rn(A.gl, m);
    aux = v;
  }
}

__device__ __forceinline__ float scan_step(float a, float b, float x) {
  return __fadd_rn(a, __fmul_rn(b, x));
}

template <int OP>
__device__ __forceinline__ void scan_store(const ScanArgs &A, int64_t t, int64_t col, float x,
                                           float aux) {
  const int64_t i = t * A.st + col * A.sc;
  A.out[i] = x;
  if (OP == OP_GAE && A.ret) A.ret[i] = __fadd_rn(x, aux);
}

// walk [t_lo, t_hi) downwards from carry x, U rows of loads in flight
template <int OP, int U>
__device__ __forceinline__ float scan_walk(const ScanArgs &A, int64_t col, int64_t t_lo,
                                           int64_t t_hi, float x) {
  for (int64_t t0 = t_hi; t0 > t_lo; t0 -= U) {
    float a[U], b[U], aux[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t t = t0 - 1 - u;
      if (t >= t_lo) scan_load<OP>(A, t, col, a[u], b[u], aux[u]);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t t = t0 - 1 - u;
      if (t >= t_lo) {
        x = scan_step(a[u], b[u], x);
        scan_store<OP>(A, t, col, x, aux[u]);
      }
    }
  }
  return x;
}

template <int OP>
__global__ void __launch_bounds__(128) scan_seq_kernel(const __grid_constant__ ScanArgs A) {
  const int64_t col = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= A.N) return;
  scan_walk<OP, 8>(A, col, 0, A.T, scan_init<OP>(A, col));
}

template <int OP>
__global__ void __launch_bounds__(1024)
    scan_tm_chunked_kernel(const __grid_constant__ ScanArgs A, int64_t chunk) {
  __shared__ float sA[32][32];
  __shared__ float sB[32][32];
  const int lane = threadIdx.x, w = threadIdx.y, W = blockDim.y;
  const int64_t col = (int64_t)blockIdx.x * 32 + lane;
  const bool valid = col < A.N;
  int64_t t_lo = (int64_t)w * chunk, t_hi = t_lo + chunk;
  if (t_lo > A.T) t_lo = A.T;
  if (t_hi > A.T) t_hi = A.T;

  // pass 1: x[t_lo] = ca + cb * x[t_hi]
  float ca = 0.0f, cb = 1.0f;
  if (valid) {
    constexpr int U = 8;
    for (int64_t t0 = t_hi; t0 > t_lo; t0 -= U) {
      float a[U], b[U], aux[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t t = t0 - 1 - u;
        if (t >= t_lo) scan_load<OP>(A, t, col, a[u], b[u], aux[u]);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t t = t0 - 1 - u;
        if (t >= t_lo) {
          ca = scan_step(a[u], b[u], ca);
          cb = __fmul_rn(b[u], cb);
        }
      }
    }
  }
  sA[w][lane] = ca;
  sB[w][lane] = cb;
  __syncthreads();
  if (!valid) return;
  // carry into this chunk = later chunks applied to the boundary value, last chunk first
  float x = scan_init<OP>(A, col);
  for (int j = W - 1; j > w; --j) x = scan_step(sA[j][lane], sB[j][lane], x);
  // pass 2: replay the chunk (inputs are L2-resident from pass 1)
  scan_walk<OP, 8>(A, col, t_lo, t_hi, x);
}

// ---- time-major single-pass tile scan -----------------------------------------------------
// CTA = 32 columns x (TM_W warps x TM_S steps) = one [128 x 32] tile held in registers.  Tiles are
// handed out by a ticket counter in dependency order (latest time chunk first).  Each CTA reduces
// its tile to one affine map per column and publishes it (aggregate + release flag); a CTA then
// folds the aggregates of all later time chunks of its column group into its carry -- aggregates do
// not depend on any carry, so there is no serial chain across CTAs -- and replays its tile from
// registers.  One global read and one global write per element.
constexpr int TM_W = 8;    // warps per CTA
constexpr int TM_S = 16;   // steps per thread
constexpr int TM_TC = TM_W * TM_S;
constexpr int TM_MAX_CHUNKS = 64;  // time chunks folded directly; longer T uses the chunked kernel

struct TileWs {
  unsigned int *ticket;  // 1 counter
  unsigned int *flag;    // [n_chunks * n_groups]
  float2 *agg;           // [n_chunks * n_groups][32]
};

template <int OP>
__global__ void __launch_bounds__(TM_W * 32)
    scan_tm_tile_kernel(const __grid_constant__ ScanArgs A, TileWs ws, int n_chunks, int n_groups) {
  __shared__ float sA[TM_W][32];
  __shared__ float sB[TM_W][32];
  __shared__ float s_ext[32];
  __shared__ unsigned int s_ticket;
  const int lane = threadIdx.x, w = threadIdx.y;
  if (lane == 0 && w == 0) s_ticket = atomicAdd(ws.ticket, 1u);
  __syncthreads();
  const unsigned int ticket = s_ticket;
  const int chunk = n_chunks - 1 - (int)(ticket / (unsigned)n_groups);  // latest chunk first
  const int group = (int)(ticket % (unsigned)n_groups);
  const int64_t col = (int64_t)group * 32 + lane;
  const bool valid = col < A.N;
  const int64_t t_base = (int64_t)chunk * TM_TC + (int64_t)w * TM_S;

  float a[TM_S], b[TM_S], aux[TM_S];
  if (OP == OP_GAE && !A.nv) {
    // V[t+1] of step s is V[t] of step s+1: load TM_S+1 values once instead of 2*TM_S
    float vv[TM_S + 1], rr[TM_S], mm[TM_S];
#pragma unroll
    for (int s = 0; s <= TM_S; ++s) {
      const int64_t t = t_base + s;
      vv[s] = 0.0f;
      if (valid && t < A.T) vv[s] = __ldg(A.v + t * A.st + col);
      else if (valid && t == A.T && A.last) vv[s] = __ldg(A.last + col);
    }
#pragma unroll
    for (int s = 0; s < TM_S; ++s) {
      const int64_t t = t_base + s;
      rr[s] = 0.0f, mm[s] = 1.0f;
      if (valid && t < A.T) {
        rr[s] = __ldg(A.a + t * A.st + col);
        if (A.done) mm[s] = __fsub_rn(1.0f, __ldg(A.done + t * A.dst + col) ? 1.0f : 0.0f);
      }
    }
#pragma unroll
    for (int s = 0; s < TM_S; ++s) {
      const int64_t t = t_base + s;
      if (valid && t < A.T) {
        const float gvm = __fmul_rn(__fmul_rn(A.gamma, vv[s + 1]), mm[s]);
        a[s] = __fsub_rn(__fadd_rn(rr[s], gvm), vv[s]);
        b[s] = __fmul_rn(A.gl, mm[s]);
        aux[s] = vv[s];
      } else {
        a[s] = 0.0f, b[s] = 1.0f, aux[s] = 0.0f;
      }
    }
  } else {
#pragma unroll
    for (int s = 0; s < TM_S; ++s) {
      const int64_t t = t_base + s;
      if (valid && t < A.T) {
        scan_load<OP>(A, t, col, a[s], b[s], aux[s]);
      } else {
        a[s] = 0.0f, b[s] = 1.0f, aux[s] = 0.0f;
      }
    }
  }
  float ca = 0.0f, cb = 1.0f;  // x[t_base] = ca + cb * x[t_base + TM_S]
#pragma unroll
  for (int s = TM_S - 1; s >= 0; --s) {
    ca = scan_step(a[s], b[s], ca);
    cb = __fmul_rn(b[s], cb);
  }
  sA[w][lane] = ca;
  sB[w][lane] = cb;
  __syncthreads();

  const int tile = chunk * n_groups + group;
  if (w == 0) {
    // publish this tile's aggregate: x[tile start] = ta + tb * x[tile end]
    float ta = 0.0f, tb = 1.0f;
#pragma unroll
    for (int j = TM_W - 1; j >= 0; --j) {
      ta = scan_step(sA[j][lane], sB[j][lane], ta);
      tb = __fmul_rn(sB[j][lane], tb);
    }
    if (chunk > 0) {
      __stcg(&ws.agg[(size_t)tile * 32 + lane], make_float2(ta, tb));
      __threadfence();
      __syncwarp();
      if (lane == 0) {
        asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(ws.flag + tile), "r"(1u) : "memory");
      }
    }
    // carry from all later chunks of this column group
    float x = valid ? scan_init<OP>(A, col) : 0.0f;
    for (int c = n_chunks - 1; c > chunk; --c) {
      const int other = c * n_groups + group;
      if (lane == 0) {
        unsigned int f;
        do {
          asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(f) : "l"(ws.flag + other) : "memory");
        } while (f == 0u);
      }
      __syncwarp();
      const float2 g = __ldcg(&ws.agg[(size_t)other * 32 + lane]);
      x = scan_step(g.x, g.y, x);
    }
    s_ext[lane] = x;
  }
  __syncthreads();
  if (!valid) return;
  float x = s_ext[lane];
  for (int j = TM_W - 1; j > w; --j) x = scan_step(sA[j][lane], sB[j][lane], x);
#pragma unroll
  for (int s = TM_S - 1; s >= 0; --s) {
    const int64_t t = t_base + s;
    if (t < A.T) {
      x = scan_step(a[s], b[s], x);
      scan_store<OP>(A, t, col, x, aux[s]);
    }
  }
}

// ---- batch-major warp scan ----------------------------------------------------------------
// A tile is Q sub-tiles of 128 steps.  Every load instruction of a sub-tile is one fully coalesced
// 512-byte warp access (lane l owns steps 4l..4l+3 of the sub-tile); all Q sub-tiles of the NEXT tile
// are requested before the current tile is scanned, so a warp keeps 2*Q*(bytes per quad) in flight.
struct BmRaw {
  float r[4], v[4], nv[4];
  uint32_t d;  // 4 done bytes
};

template <int OP, bool VEC, bool HAS_NV>
__device__ __forceinline__ void bm_load(const ScanArgs &A, int64_t row, int64_t t0, BmRaw &q) {
  const int64_t i = row * A.T + t0;
  if (VEC) {
    // T % 4 == 0 and 16-byte aligned bases: whole quad is in range when t0 < T
    if (t0 < A.T) {
      const float4 r = __ldcs(reinterpret_cast<const float4 *>(A.a + i));
      q.r[0] = r.x, q.r[1] = r.y, q.r[2] = r.z, q.r[3] = r.w;
      if (OP != OP_DR) {
        const float4 v = __ldcs(reinterpret_cast<const float4 *>(A.v + i));
        q.v[0] = v.x, q.v[1] = v.y, q.v[2] = v.z, q.v[3] = v.w;
      }
      if (OP == OP_GAE && HAS_NV) {
        const float4 n = __ldcs(reinterpret_cast<const float4 *>(A.nv + i));
        q.nv[0] = n.x, q.nv[1] = n.y, q.nv[2] = n.z, q.nv[3] = n.w;
      }
      q.d = (OP != OP_AFFINE && A.done) ? __ldcs(reinterpret_cast<const uint32_t *>(A.done + i)) : 0u;
    }
  } else {
    q.d = 0u;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (t0 + k < A.T) {
        q.r[k] = __ldg(A.a + i + k);
        if (OP != OP_DR) q.v[k] = __ldg(A.v + i + k);
        if (OP == OP_GAE && HAS_NV) q.nv[k] = __ldg(A.nv + i + k);
        if (OP != OP_AFFINE && A.done) q.d |= (uint32_t)(__ldg(A.done + i + k) != 0) << (8 * k);
      }
    }
  }
}

template <int OP, bool VEC, bool HAS_NV, int Q>
__global__ void __launch_bounds__(256) scan_bm_warp_kernel(const __grid_constant__ ScanArgs A) {
  const int lane = threadIdx.x & 31;
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= A.N) return;  // whole warp leaves together
  constexpr unsigned FULL = 0xffffffffu;
  constexpr int64_t TILE = 128 * Q;

  float carry = scan_init<OP>(A, row);
  const float lastv = (OP == OP_GAE && !HAS_NV && A.last) ? __ldg(A.last + row) : 0.0f;
  float v_after = lastv;  // V at the first step after the current sub-tile

  int64_t base = ((A.T - 1) / TILE) * TILE;
  BmRaw cur[Q], nxt[Q];
#pragma unroll
  for (int q = 0; q < Q; ++q) bm_load<OP, VEC, HAS_NV>(A, row, base + 128 * q + 4 * lane, cur[q]);
  for (; base >= 0; base -= TILE) {
    if (base >= TILE) {  // request the whole next tile before touching this one
#pragma unroll
      for (int q = 0; q < Q; ++q)
        bm_load<OP, VEC, HAS_NV>(A, row, base - TILE + 128 * q + 4 * lane, nxt[q]);
    }
#pragma unroll
    for (int q = Q - 1; q >= 0; --q) {
      const int64_t t0 = base + 128 * q + 4 * lane;
      if (base + 128 * q >= A.T) continue;  // sub-tile entirely past the end (warp-uniform)
      const BmRaw &c = cur[q];
      float a[4], b[4], aux[4];
      float v_next_lane = 0.0f;
      if (OP == OP_GAE && !HAS_NV) {
        v_next_lane = __shfl_down_sync(FULL, c.v[0], 1);
        if (lane == 31) v_next_lane = v_after;
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int64_t t = t0 + k;
        if (t < A.T) {
          if (OP == OP_AFFINE) {
            a[k] = c.r[k];
            b[k] = c.v[k];
            aux[k] = 0.0f;
          } else {
            const float m = __fsub_rn(1.0f, ((c.d >> (8 * k)) & 0xffu) ? 1.0f : 0.0f);
            if (OP == OP_DR) {
              a[k] = c.r[k];
              b[k] = __fmul_rn(m, A.gamma);
              aux[k] = 0.0f;
            } else {
              float nv;
              if (HAS_NV) nv = c.nv[k];
              else if (t == A.T - 1) nv = lastv;
              else nv = (k < 3) ? c.v[k < 3 ? k + 1 : 3] : v_next_lane;
              const float gvm = __fmul_rn(__fmul_rn(A.gamma, nv), m);
              a[k] = __fsub_rn(__fadd_rn(c.r[k], gvm), c.v[k]);
              b[k] = __fmul_rn(A.gl, m);
              aux[k] = c.v[k];
            }
          }
        } else {  // identity map: passes the carry through
          a[k] = 0.0f;
          b[k] = 1.0f;
          aux[k] = 0.0f;
        }
      }
      if (OP == OP_GAE && !HAS_NV) v_after = __shfl_sync(FULL, c.v[0], 0);

      // this lane's 4 steps as one affine map: x[t0] = ca + cb * x[t0+4]
      float ca = 0.0f, cb = 1.0f;
#pragma unroll
      for (int k = 3; k >= 0; --k) {
        ca = scan_step(a[k], b[k], ca);
        cb = __fmul_rn(b[k], cb);
      }
      // suffix scan over lanes (later time = higher lane): x[t0] = ca + cb * x[sub-tile end]
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const float oa = __shfl_down_sync(FULL, ca, off);
        const float ob = __shfl_down_sync(FULL, cb, off);
        if (lane + off < 32) {
          ca = scan_step(ca, cb, oa);  // ca + cb*oa
          cb = __fmul_rn(cb, ob);
        }
      }
      const float x_here = scan_step(ca, cb, carry);  // x[t0]
      float x = __shfl_down_sync(FULL, x_here, 1);    // x[t0+4] = next lane's x[t0]
      if (lane == 31) x = carry;
      float o[4], o2[4];
#pragma unroll
      for (int k = 3; k >= 0; --k) {
        x = scan_step(a[k], b[k], x);
        o[k] = x;
        o2[k] = __fadd_rn(x, aux[k]);
      }
      const int64_t i = row * A.T + t0;
      if (VEC) {
        if (t0 < A.T) {
          __stcs(reinterpret_cast<float4 *>(A.out + i), make_float4(o[0], o[1], o[2], o[3]));
          if (OP == OP_GAE && A.ret)
            __stcs(reinterpret_cast<float4 *>(A.ret + i), make_float4(o2[0], o2[1], o2[2], o2[3]));
        }
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (t0 + k < A.T) {
            A.out[i + k] = o[k];
            if (OP == OP_GAE && A.ret) A.ret[i + k] = o2[k];
          }
      }
      carry = __shfl_sync(FULL, x_here, 0);
    }
#pragma unroll
    for (int q = 0; q < Q; ++q) cur[q] = nxt[q];
  }
}

template <int OP, bool VEC, bool HAS_NV>
static void launch_bm(const ScanArgs &A, cudaStream_t st) {
  const int64_t blocks = (A.N * 32 + 255) / 256;
  if (A.T > 512) scan_bm_warp_kernel<OP, VEC, HAS_NV, 4><<<(unsigned)blocks, 256, 0, st>>>(A);
  else scan_bm_warp_kernel<OP, VEC, HAS_NV, 1><<<(unsigned)blocks, 256, 0, st>>>(A);
}

static bool aligned16(const void *p) { return (((uintptr_t)p) & 15) == 0; }

// workspace of the tile scan (ticket, flags, aggregates), grow-only, one per process
static uint8_t *g_tile_ws = nullptr;
static size_t g_tile_ws_bytes = 0;
static int g_tile_mode = -1;  // MRL_TM_KERNEL=chunked forces the two-pass kernel

static int tile_workspace(size_t need, uint8_t **p) {
  if (need > g_tile_ws_bytes) {
    if (g_tile_ws) cudaFree(g_tile_ws);
    g_tile_ws = nullptr;
    g_tile_ws_bytes = 0;
    MRL_CUDA(cudaMalloc((void **)&g_tile_ws, need));
    g_tile_ws_bytes = need;
  }
  *p = g_tile_ws;
  return MRL_OK;
}

template <int OP>
static int launch_scan(ScanArgs A, int32_t layout, int32_t mode, cudaStream_t st) {
  if (A.T <= 0 || A.N <= 0) return MRL_OK;
  if (g_tile_mode < 0) {
    const char *e = getenv("MRL_TM_KERNEL");
    g_tile_mode = (e && e[0] == 'c') ? 1 : 0;
  }
  if (layout == MRL_TIME_MAJOR) {
    A.st = A.N, A.sc = 1, A.dst = A.B, A.dsc = 1;
    bool chunked;
    if (mode == MRL_SCAN_SEQUENTIAL) chunked = false;
    else if (mode == MRL_SCAN_PARALLEL) chunked = A.T >= 2;
    else chunked = A.T >= 64 && A.N < (int64_t)kNumSM * 1024;
    const int64_t n_chunks = (A.T + TM_TC - 1) / TM_TC;
    if (chunked && g_tile_mode == 0 && n_chunks <= TM_MAX_CHUNKS) {
      const int64_t n_groups = (A.N + 31) / 32;
      const size_t tiles = (size_t)n_chunks * n_groups;
      const size_t flag_bytes = (4 + tiles * 4 + 255) / 256 * 256;
      uint8_t *wsp;
      int rc = tile_workspace(flag_bytes + tiles * 32 * sizeof(float2), &wsp);
      if (rc) return rc;
      MRL_CUDA(cudaMemsetAsync(wsp, 0, flag_bytes, st));
      TileWs ws;
      ws.ticket = reinterpret_cast<unsigned int *>(wsp);
      ws.flag = ws.ticket + 1;
      ws.agg = reinterpret_cast<float2 *>(wsp + flag_bytes);
      dim3 block(32, TM_W);
      scan_tm_tile_kernel<OP><<<(unsigned)tiles, block, 0, st>>>(A, ws, (int)n_chunks, (int)n_groups);
    } else if (chunked) {
      int W = 32;
      while (W > 1 && A.T / W < 8) W >>= 1;
      const int64_t chunk = (A.T + W - 1) / W;
      dim3 block(32, W);
      const int64_t blocks = (A.N + 31) / 32;
      scan_tm_chunked_kernel<OP><<<(unsigned)blocks, block, 0, st>>>(A, chunk);
    } else {
      const int64_t blocks = (A.N + 127) / 128;
      scan_seq_kernel<OP><<<(unsigned)blocks, 128, 0, st>>>(A);
    }
  } else {
    A.st = 1, A.sc = A.T, A.dst = 1, A.dsc = A.T;
    if (mode == MRL_SCAN_SEQUENTIAL) {
      const int64_t blocks = (A.N + 127) / 128;
      scan_seq_kernel<OP><<<(unsigned)blocks, 128, 0, st>>>(A);
    } else {
      const bool vec = (A.T % 4 == 0) && aligned16(A.a) && aligned16(A.out) &&
                       (OP == OP_DR || aligned16(A.v)) && (!A.nv || aligned16(A.nv)) &&
                       (!A.ret || aligned16(A.ret)) && (!A.done || (((uintptr_t)A.done) & 3) == 0);
      const bool has_nv = OP == OP_GAE && A.nv != nullptr;
      if (vec && has_nv) launch_bm<OP, true, true>(A, st);
      else if (vec) launch_bm<OP, true, false>(A, st);
      else if (has_nv) launch_bm<OP, false, true>(A, st);
      else launch_bm<OP, false, false>(A, st);
    }
  }
  MRL_LAUNCHED();
  return MRL_OK;
}

// grow-only device scratch for the *_host entry points
static uint8_t *g_scratch = nullptr;
static size_t g_scratch_bytes = 0;
static int scratch(size_t need, uint8_t **p) {
  if (need > g_scratch_bytes) {
    if (g_scratch) cudaFree(g_scratch);
    g_scratch = nullptr;
    g_scratch_bytes = 0;
    MRL_CUDA(cudaMalloc((void **)&g_scratch, need));
    g_scratch_bytes = need;
  }
  *p = g_scratch;
  return MRL_OK;
}
static size_t up256(size_t x) { return (x + 255) / 256 * 256; }

}  // namespace mrl

using namespace mrl;

extern "C" int mrl_discounted_return(const float *reward, const uint8_t *done,
                                     const float *last_value, float *out, int64_t T, int64_t B,
                                     int64_t E, float gamma, int32_t layout, int32_t mode,
                                     void *stream) {
  MRL_REQUIRE(T >= 0 && B >= 0 && E >= 1, "mrl_discounted_return: bad shape");
  MRL_REQUIRE(layout == MRL_TIME_MAJOR || (layout == MRL_BATCH_MAJOR && E == 1),
              "mrl_discounted_return: batch-major layout needs E == 1");
  MRL_REQUIRE(gamma >= 0.0f && gamma <= 1.0f,
              "The discounted factor should be a number in range [0, 1], but got %f.", (double)gamma);
  if (T == 0 || B == 0) return MRL_OK;
  MRL_REQUIRE(reward && out, "mrl_discounted_return: NULL reward/out");
  ScanArgs A{};
  A.a = reward, A.done = done, A.last = last_value, A.out = out;
  A.T = T, A.B = B, A.E = E, A.N = B * E, A.gamma = gamma;
  return launch_scan<OP_DR>(A, layout, mode, as_stream(stream));
}

extern "C" int mrl_gae(const float *reward, const float *value, const float *next_value,
                       const float *last_value, const uint8_t *done, float *adv, float *ret,
                       int64_t T, int64_t B, float gamma, float lambda, int32_t layout,
                       int32_t mode, void *stream) {
  MRL_REQUIRE(T >= 0 && B >= 0, "mrl_gae: bad shape");
  MRL_REQUIRE(layout == MRL_TIME_MAJOR || layout == MRL_BATCH_MAJOR, "mrl_gae: bad layout");
  if (T == 0 || B == 0) return MRL_OK;
  MRL_REQUIRE(reward && value && adv, "mrl_gae: NULL reward/value/adv");
  ScanArgs A{};
  A.a = reward, A.v = value, A.nv = next_value, A.last = last_value, A.done = done;
  A.out = adv, A.ret = ret;
  A.T = T, A.B = B, A.E = 1, A.N = B, A.gamma = gamma, A.gl = gamma * lambda;
  return launch_scan<OP_GAE>(A, layout, mode, as_stream(stream));
}

extern "C" int mrl_affine_scan(const float *a, const float *b, const float *x_last, float *out,
                               int64_t T, int64_t B, int32_t layout, int32_t mode, void *stream) {
  MRL_REQUIRE(T >= 0 && B >= 0, "mrl_affine_scan: bad shape");
  MRL_REQUIRE(layout == MRL_TIME_MAJOR || layout == MRL_BATCH_MAJOR, "mrl_affine_scan: bad layout");
  if (T == 0 || B == 0) return MRL_OK;
  MRL_REQUIRE(a && b && out, "mrl_affine_scan: NULL a/b/out");
  ScanArgs A{};
  A.a = a, A.v = b, A.last = x_last, A.out = out;
  A.T = T, A.B = B, A.E = 1, A.N = B;
  return launch_scan<OP_AFFINE>(A, layout, mode, as_stream(stream));
}

extern "C" int mrl_discounted_return_host(const float *reward, const uint8_t *done,
                                          const float *last_value, float *out, int64_t T,
                                          int64_t B, int64_t E, float gamma, int32_t layout,
                                          int32_t mode, void *stream) {
  MRL_REQUIRE(T >= 0 && B >= 0 && E >= 1, "mrl_discounted_return_host: bad shape");
  if (T == 0 || B == 0) return MRL_OK;
  cudaStream_t st = as_stream(stream);
  const size_t n = (size_t)T * B * E, nd = (size_t)T * B, nl = (size_t)B * E;
  const size_t o_r = 0, o_out = o_r + up256(n * 4), o_d = o_out + up256(n * 4),
               o_l = o_d + up256(nd), total = o_l + up256(nl * 4);
  uint8_t *s;
  int rc = scratch(total, &s);
  if (rc) return rc;
  MRL_CUDA(cudaMemcpyAsync(s + o_r, reward, n * 4, cudaMemcpyHostToDevice, st));
  if (done) MRL_CUDA(cudaMemcpyAsync(s + o_d, done, nd, cudaMemcpyHostToDevice, st));
  if (last_value) MRL_CUDA(cudaMemcpyAsync(s + o_l, last_value, nl * 4, cudaMemcpyHostToDevice, st));
  rc = mrl_discounted_return((const float *)(s + o_r), done ? s + o_d : nullptr,
                             last_value ? (const float *)(s + o_l) : nullptr,
                             (float *)(s + o_out), T, B, E, gamma, layout, mode, stream);
  if (rc) return rc;
  MRL_CUDA(cudaMemcpyAsync(out, s + o_out, n * 4, cudaMemcpyDeviceToHost, st));
  MRL_CUDA(cudaStreamSynchronize(st));
  return MRL_OK;
}

extern "C" int mrl_gae_host(const float *reward, const float *value, const float *next_value,
                            const float *last_value, const uint8_t *done, float *adv, float *ret,
                            int64_t T, int64_t B, float gamma, float lambda, int32_t layout,
                            int32_t mode, void *stream) {
  MRL_REQUIRE(T >= 0 && B >= 0, "mrl_gae_host: bad shape");
  if (T == 0 || B == 0) return MRL_OK;
  cudaStream_t st = as_stream(stream);
  const size_t n = (size_t)T * B;
  const size_t f = up256(n * 4);
  const size_t o_r = 0, o_v = f, o_nv = 2 * f, o_adv = 3 * f, o_ret = 4 * f, o_d = 5 * f,
               o_l = o_d + up256(n), total = o_l + up256((size_t)B * 4);
  uint8_t *s;
  int rc = scratch(total, &s);
  if (rc) return rc;
  MRL_CUDA(cudaMemcpyAsync(s + o_r, reward, n * 4, cudaMemcpyHostToDevice, st));
  MRL_CUDA(cudaMemcpyAsync(s + o_v, value, n * 4, cudaMemcpyHostToDevice, st));
  if (next_value) MRL_CUDA(cudaMemcpyAsync(s + o_nv, next_value, n * 4, cudaMemcpyHostToDevice, st));
  if (done) MRL_CUDA(cudaMemcpyAsync(s + o_d, done, n, cudaMemcpyHostToDevice, st));
  if (last_value)
    MRL_CUDA(cudaMemcpyAsync(s + o_l, last_value, (size_t)B * 4, cudaMemcpyHostToDevice, st));
  rc = mrl_gae((const float *)(s + o_r), (const float *)(s + o_v),
               next_value ? (const float *)(s + o_nv) : nullptr,
               last_value ? (const float *)(s + o_l) : nullptr, done ? s + o_d : nullptr,
               (float *)(s + o_adv), ret ? (float *)(s + o_ret) : nullptr, T, B, gamma, lambda,
               layout, mode, stream);
  if (rc) return rc;
  MRL_CUDA(cudaMemcpyAsync(adv, s + o_adv, n * 4, cudaMemcpyDeviceToHost, st));
  if (ret) MRL_CUDA(cudaMemcpyAsync(ret, s + o_ret, n * 4, cudaMemcpyDeviceToHost, st));
  MRL_CUDA(cudaStreamSynchronize(st));
  return MRL_OK;
}
