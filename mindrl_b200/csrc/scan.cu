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
//   scan_tm_wtile_kernel    time-major, default: a warp owns 32 columns x 32 steps in registers; tiles
//                           are handed out latest-chunk-first by a ticket; the carry comes from a
//                           decoupled look-back over flag-in-payload words.  One read, one write.
//   scan_tm_tile4_kernel    time-major, 128-column CTA tiles (aligned inputs, next_value form of GAE)
//   scan_tm_chunked_kernel  time-major two-pass fallback (unaligned next_value form)
//   scan_bm_warp_kernel     batch-major: warp per environment, 128 steps per tile (float4 per lane),
//                           per-lane composite + 5-step shuffle suffix scan of affine maps; a done flag
//                           makes b = 0, which cuts the segment inside the same scan.
// The two parallel kernels re-associate fp32 (spec: <= 1e-5 relative); MRL_SCAN_SEQUENTIAL selects the
// bit-comparable kernel.
#include <stdlib.h>

#include <map>

#include "scan.cuh"

namespace mrl {

template <int OP>
__device__ __forceinline__ float scan_init(const ScanArgs &A, int64_t col) {
  if (OP == OP_GAE) return 0.0f;
  return A.last ? __ldg(A.last + col) : 0.0f;
}

// V[T] of a column for the GAE bootstrap (value[t+1] form)
template <int OP>
__device__ __forceinline__ float scan_lastv(const ScanArgs &A, int64_t col) {
  return (OP == OP_GAE && !A.nv && A.last) ? __ldg(A.last + col) : 0.0f;
}

// One element: (a, b) of x = a + b*x_next plus the value the store needs.  (t, col) must be in range:
// every load here is unconditional and address selection is branch-free, so that an unrolled caller
// gets all its loads in flight at once (predicated loads are serialised by the compiler).
template <int OP, bool HD>
__device__ __forceinline__ void scan_load(const ScanArgs &A, int64_t t, int64_t col, float lastv,
                                          float &a, float &b, float &aux) {
  const int64_t i = t * A.st + col * A.sc;
  if (OP == OP_AFFINE) {
    a = __ldg(A.a + i);
    b = __ldg(A.v + i);
    aux = 0.0f;
    return;
  }
  float m = 1.0f;
  if (HD) {
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
    const bool boot = !A.nv && t == A.T - 1;
    const float *np = A.nv ? A.nv + i : (boot ? A.v + i : A.v + i + A.st);
    float nv = __ldg(np);
    nv = boot ? lastv : nv;
    const float gvm = __fmul_rn(__fmul_rn(A.gamma, nv), m);
    a = __fsub_rn(__fadd_rn(r, gvm), v);
    b = __fmul_rn(A.gl, m);
    aux = v;
  }
}

template <int OP>
__device__ __forceinline__ void scan_store(const ScanArgs &A, int64_t t, int64_t col, float x,
                                           float aux) {
  const int64_t i = t * A.st + col * A.sc;
  A.out[i] = x;
  if (OP == OP_GAE && A.ret) A.ret[i] = __fadd_rn(x, aux);
}

// walk [t_lo, t_hi) downwards from carry x, U rows of loads in flight
template <int OP, bool HD, int U>
__device__ __forceinline__ float scan_walk(const ScanArgs &A, int64_t col, float lastv, int64_t t_lo,
                                           int64_t t_hi, float x) {
  for (int64_t t0 = t_hi; t0 > t_lo; t0 -= U) {
    float a[U], b[U], aux[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      int64_t t = t0 - 1 - u;
      t = t < t_lo ? t_lo : t;  // clamped: the load is always legal, the result may be unused
      scan_load<OP, HD>(A, t, col, lastv, a[u], b[u], aux[u]);
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

template <int OP, bool HD>
__global__ void __launch_bounds__(128) scan_seq_kernel(const __grid_constant__ ScanArgs A) {
  const int64_t col = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= A.N) return;
  scan_walk<OP, HD, 8>(A, col, scan_lastv<OP>(A, col), 0, A.T, scan_init<OP>(A, col));
}

template <int OP, bool HD>
__global__ void __launch_bounds__(1024)
    scan_tm_chunked_kernel(const __grid_constant__ ScanArgs A, int64_t chunk) {
  __shared__ float sA[32][32];
  __shared__ float sB[32][32];
  const int lane = threadIdx.x, w = threadIdx.y, W = blockDim.y;
  const int64_t col = (int64_t)blockIdx.x * 32 + lane;
  const bool valid = col < A.N;
  const int64_t colc = valid ? col : A.N - 1;
  int64_t t_lo = (int64_t)w * chunk, t_hi = t_lo + chunk;
  if (t_lo > A.T) t_lo = A.T;
  if (t_hi > A.T) t_hi = A.T;
  const float lastv = scan_lastv<OP>(A, colc);

  // pass 1: x[t_lo] = ca + cb * x[t_hi]
  float ca = 0.0f, cb = 1.0f;
  {
    constexpr int U = 8;
    for (int64_t t0 = t_hi; t0 > t_lo; t0 -= U) {
      float a[U], b[U], aux[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        int64_t t = t0 - 1 - u;
        t = t < t_lo ? t_lo : t;
        scan_load<OP, HD>(A, t, colc, lastv, a[u], b[u], aux[u]);
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
  // pass 2: replay the chunk
  scan_walk<OP, HD, 8>(A, col, lastv, t_lo, t_hi, x);
}

// ---- time-major single-pass tile scan, 512-byte rows ---------------------------------------
// CTA = 128 columns x (T4_W warps x T4_S steps) held in registers; a lane owns 4 adjacent columns (one
// 16-byte load per array and step).  Tiles are handed out by a ticket counter in dependency order
// (latest time chunk first).  The carry comes from a decoupled
// look-back: every tile publishes its aggregate (flag 1) and later its inclusive value at the tile
// start (flag 2); a tile walks towards later chunks, composing aggregates until it meets an
// inclusive value.
constexpr int T4_W = 4;
constexpr int T4_S = 8;
constexpr int T4_TC = T4_W * T4_S;  // 32 steps
constexpr int T4_COLS = 128;

struct Tile4Ws {
  unsigned int *ticket;
  unsigned int *flag;  // [tiles] 0 none, 1 aggregate, 2 inclusive
  float2 *agg;         // [tiles][128] (a, b)
  float *incl;         // [tiles][128] x at the tile start
};

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int *p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_u32(unsigned int *p, unsigned int v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <int OP, bool HD>
__global__ void __launch_bounds__(T4_W * 32, 3)
    scan_tm_tile4_kernel(const __grid_constant__ ScanArgs A, Tile4Ws ws, int n_chunks, int n_groups) {
  __shared__ __align__(16) float sA[T4_W][T4_COLS];
  __shared__ __align__(16) float sB[T4_W][T4_COLS];
  __shared__ __align__(16) float s_ext[T4_COLS];
  __shared__ unsigned int s_ticket;
  __shared__ int s_zero;
  const int lane = threadIdx.x, w = threadIdx.y;
  if (lane == 0 && w == 0) {
    s_ticket = atomicAdd(ws.ticket, 1u);
    s_zero = 0;
  }
  __syncthreads();
  const unsigned int ticket = s_ticket;
  const int chunk = n_chunks - 1 - (int)(ticket / (unsigned)n_groups);
  const int group = (int)(ticket % (unsigned)n_groups);
  const int64_t col = (int64_t)group * T4_COLS + lane * 4;
  const bool valid = col < A.N;  // N % 4 == 0: a lane's 4 columns are valid together
  const int64_t colc = valid ? col : A.N - 4;
  const int64_t t_base = (int64_t)chunk * T4_TC + (int64_t)w * T4_S;
  const int64_t t_max = A.T - 1;

  // ---- loads: all unconditional (clamped), all in flight together -------------------------
  float4 rr[T4_S], vv[T4_S + 1], nn[T4_S];
  uint32_t dd[T4_S];
#pragma unroll
  for (int s = 0; s < T4_S; ++s) {
    int64_t t = t_base + s;
    t = t > t_max ? t_max : t;
    rr[s] = __ldcs(reinterpret_cast<const float4 *>(A.a + t * A.N + colc));
    dd[s] = HD ? __ldcs(reinterpret_cast<const uint32_t *>(A.done + t * A.N + colc)) : 0u;
  }
  if (OP != OP_DR) {
#pragma unroll
    for (int s = 0; s <= T4_S; ++s) {
      int64_t t = t_base + s;
      t = t > t_max ? t_max : t;
      if (s < T4_S || OP == OP_GAE)
        vv[s] = __ldcs(reinterpret_cast<const float4 *>(A.v + t * A.N + colc));
    }
  }
  const bool has_nv = OP == OP_GAE && A.nv != nullptr;
  if (has_nv) {
#pragma unroll
    for (int s = 0; s < T4_S; ++s) {
      int64_t t = t_base + s;
      t = t > t_max ? t_max : t;
      nn[s] = __ldcs(reinterpret_cast<const float4 *>(A.nv + t * A.N + colc));
    }
  }
  float4 lastv = make_float4(0.f, 0.f, 0.f, 0.f);
  if (OP == OP_GAE && !has_nv && A.last) lastv = __ldcg(reinterpret_cast<const float4 *>(A.last + colc));
  // Scheduling fence: ptxas otherwise sinks the first consumer between the loads, and the in-order
  // warp then sits out a full memory latency before it issues the rest.  The barrier pins the loads,
  // the never-taken branch on an opaque zero pins the consumers behind it.
  __syncthreads();
  if (*reinterpret_cast<volatile int *>(&s_zero) != 0) return;

  // ---- per-element maps ------------------------------------------------------------------
  float a[T4_S][4], b[T4_S][4];
#pragma unroll
  for (int s = 0; s < T4_S; ++s) {
    const int64_t t = t_base + s;
    const bool in = valid && t <= t_max;
    const float r4[4] = {rr[s].x, rr[s].y, rr[s].z, rr[s].w};
    float v4[4] = {0.f, 0.f, 0.f, 0.f}, n4[4] = {0.f, 0.f, 0.f, 0.f};
    if (OP != OP_DR) {
      v4[0] = vv[s].x, v4[1] = vv[s].y, v4[2] = vv[s].z, v4[3] = vv[s].w;
    }
    if (OP == OP_GAE) {
      const float4 q = has_nv ? nn[s] : ((t + 1 > t_max) ? lastv : vv[s + 1]);
      n4[0] = q.x, n4[1] = q.y, n4[2] = q.z, n4[3] = q.w;
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float aa, bb;
      if (OP == OP_AFFINE) {
        aa = r4[c];
        bb = v4[c];
      } else {
        const float m = __fsub_rn(1.0f, ((dd[s] >> (8 * c)) & 0xffu) ? 1.0f : 0.0f);
        if (OP == OP_DR) {
          aa = r4[c];
          bb = __fmul_rn(m, A.gamma);
        } else {
          const float gvm = __fmul_rn(__fmul_rn(A.gamma, n4[c]), m);
          aa = __fsub_rn(__fadd_rn(r4[c], gvm), v4[c]);
          bb = __fmul_rn(A.gl, m);
        }
      }
      a[s][c] = in ? aa : 0.0f;
      b[s][c] = in ? bb : 1.0f;
    }
  }
  float ca[4] = {0.f, 0.f, 0.f, 0.f}, cb[4] = {1.f, 1.f, 1.f, 1.f};
#pragma unroll
  for (int s = T4_S - 1; s >= 0; --s)
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      ca[c] = scan_step(a[s][c], b[s][c], ca[c]);
      cb[c] = __fmul_rn(b[s][c], cb[c]);
    }
  *reinterpret_cast<float4 *>(&sA[w][lane * 4]) = make_float4(ca[0], ca[1], ca[2], ca[3]);
  *reinterpret_cast<float4 *>(&sB[w][lane * 4]) = make_float4(cb[0], cb[1], cb[2], cb[3]);
  __syncthreads();

  const size_t tile = (size_t)chunk * n_groups + group;
  if (w == 0) {
    float ta[4] = {0.f, 0.f, 0.f, 0.f}, tb[4] = {1.f, 1.f, 1.f, 1.f};
#pragma unroll
    for (int j = T4_W - 1; j >= 0; --j) {
      const float4 ja = *reinterpret_cast<const float4 *>(&sA[j][lane * 4]);
      const float4 jb = *reinterpret_cast<const float4 *>(&sB[j][lane * 4]);
      const float ja4[4] = {ja.x, ja.y, ja.z, ja.w}, jb4[4] = {jb.x, jb.y, jb.z, jb.w};
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        ta[c] = scan_step(ja4[c], jb4[c], ta[c]);
        tb[c] = __fmul_rn(jb4[c], tb[c]);
      }
    }
    if (chunk > 0) {  // somebody looks back at this tile
      float4 *dst = reinterpret_cast<float4 *>(ws.agg + tile * T4_COLS + lane * 4);
      __stcg(dst, make_float4(ta[0], tb[0], ta[1], tb[1]));
      __stcg(dst + 1, make_float4(ta[2], tb[2], ta[3], tb[3]));
      __threadfence();
      __syncwarp();
      if (lane == 0) st_release_u32(ws.flag + tile, 1u);
    }
    // Decoupled look-back towards later chunks, one window of 32 chunks per round trip: lane j
    // polls chunk+1+j until it has at least an aggregate; the nearest chunk that already has its
    // inclusive value (or the end of the trajectory) closes the window.  x_in = P(x at that point).
    float pa[4] = {0.f, 0.f, 0.f, 0.f}, pb[4] = {1.f, 1.f, 1.f, 1.f};
    float xin[4];
    int c = chunk + 1;
    while (true) {
      const int cj = c + lane;
      unsigned int f = 2u;  // beyond the last chunk: the boundary value plays the inclusive role
      if (cj < n_chunks) {
        const unsigned int *fp = ws.flag + (size_t)cj * n_groups + group;
        do {
          f = ld_acquire_u32(fp);
        } while (f == 0u);
      }
      const unsigned int closed = __ballot_sync(0xffffffffu, f == 2u);
      const int jstop = closed ? __ffs(closed) - 1 : 32;  // aggregates c .. c+jstop-1, then inclusive
      for (int j0 = 0; j0 < jstop; j0 += 4) {
        float4 g0[4], g1[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int j = j0 + u < jstop ? j0 + u : jstop - 1;
          const float4 *src = reinterpret_cast<const float4 *>(
              ws.agg + ((size_t)(c + j) * n_groups + group) * T4_COLS + lane * 4);
          g0[u] = __ldcg(src);
          g1[u] = __ldcg(src + 1);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (j0 + u < jstop) {
            const float ga[4] = {g0[u].x, g0[u].z, g1[u].x, g1[u].z};
            const float gb[4] = {g0[u].y, g0[u].w, g1[u].y, g1[u].w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              pa[k] = scan_step(pa[k], pb[k], ga[k]);  // P o M
              pb[k] = __fmul_rn(pb[k], gb[k]);
            }
          }
        }
      }
      if (jstop < 32) {
        const int cs = c + jstop;
        float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
        if (cs >= n_chunks) {
          if (OP != OP_GAE && A.last) q = __ldg(reinterpret_cast<const float4 *>(A.last + colc));
        } else {
          q = __ldcg(reinterpret_cast<const float4 *>(
              ws.incl + ((size_t)cs * n_groups + group) * T4_COLS + lane * 4));
        }
        const float q4[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) xin[k] = scan_step(pa[k], pb[k], q4[k]);
        break;
      }
      c += 32;
    }
    if (chunk > 0) {  // inclusive value at this tile's start for the tiles before it
      float inc[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) inc[k] = scan_step(ta[k], tb[k], xin[k]);
      __stcg(reinterpret_cast<float4 *>(ws.incl + tile * T4_COLS + lane * 4),
             make_float4(inc[0], inc[1], inc[2], inc[3]));
      __threadfence();
      __syncwarp();
      if (lane == 0) st_release_u32(ws.flag + tile, 2u);
    }
    *reinterpret_cast<float4 *>(&s_ext[lane * 4]) = make_float4(xin[0], xin[1], xin[2], xin[3]);
  }
  __syncthreads();
  if (!valid) return;
  const float4 e = *reinterpret_cast<const float4 *>(&s_ext[lane * 4]);
  float x[4] = {e.x, e.y, e.z, e.w};
  for (int j = T4_W - 1; j > w; --j) {
    const float4 ja = *reinterpret_cast<const float4 *>(&sA[j][lane * 4]);
    const float4 jb = *reinterpret_cast<const float4 *>(&sB[j][lane * 4]);
    x[0] = scan_step(ja.x, jb.x, x[0]);
    x[1] = scan_step(ja.y, jb.y, x[1]);
    x[2] = scan_step(ja.z, jb.z, x[2]);
    x[3] = scan_step(ja.w, jb.w, x[3]);
  }
#pragma unroll
  for (int s = T4_S - 1; s >= 0; --s) {
    const int64_t t = t_base + s;
    if (t <= t_max) {
#pragma unroll
      for (int c = 0; c < 4; ++c) x[c] = scan_step(a[s][c], b[s][c], x[c]);
      const int64_t i = t * A.N + col;
      __stcs(reinterpret_cast<float4 *>(A.out + i), make_float4(x[0], x[1], x[2], x[3]));
      if (OP == OP_GAE && A.ret)
        __stcs(reinterpret_cast<float4 *>(A.ret + i),
               make_float4(__fadd_rn(x[0], vv[s].x), __fadd_rn(x[1], vv[s].y),
                           __fadd_rn(x[2], vv[s].z), __fadd_rn(x[3], vv[s].w)));
    }
  }
}

// ---- time-major single-pass scan, warp-owned tiles ------------------------------------------
// A warp owns a tile of 32 columns x WT_S steps in registers and never meets a block barrier after
// its loads: the four warps of a CTA take four adjacent column groups of the same time chunk (a time
// row is fetched in 512-byte pieces at about the same moment) but publish, look back and replay on
// their own.
//
// Carry exchange = decoupled look-back with flag-in-payload words (the protocol NCCL calls LL): every
// published float travels in an aligned 8-byte word {value bits, launch epoch}.  8-byte accesses are
// single-copy atomic, so a reader that sees the epoch sees the value: no __threadfence, no release /
// acquire, no separate flag round trip, no per-launch memset.  A tile publishes its aggregate map
// (a, b), looks at a window of WT_WIN later chunks in ONE round trip (aggregate and inclusive words of
// each), composes aggregates up to the nearest chunk whose inclusive value is out, publishes its own
// inclusive value, and replays its tile from registers.
constexpr int WT_S = 32;
constexpr int WT_WARPS = 4;
constexpr int WT_WIN = 8;

// Launch state of the look-back kernels lives IN the workspace and is advanced by the last CTA of every
// launch (not by the host): a launch reads `base` and `epoch` when it starts, takes tickets from `ticket`,
// and its last CTA to finish sets base <- ticket, epoch <- epoch + 1 for the next launch on the stream.
// Nothing launch-specific travels in the kernel parameters, so a captured CUDA graph replays correctly.
// (The epoch is 32 bits and words are only trusted when they carry the current epoch; a word written exactly
// 2^32 launches ago and not since could be mistaken -- every launch of a shape rewrites all of its words.)
struct LLCtl {
  unsigned long long ticket;  // tickets handed out since the workspace was created
  unsigned long long base;    // value of `ticket` when the current launch started
  unsigned int epoch;         // tag of the current launch's words
  unsigned int done;          // CTAs of the current launch that have finished
};
struct WTileWs {
  LLCtl *ctl;
  uint4 *agg;   // [tiles][32] {a, epoch, b, epoch}
  uint2 *incl;  // [tiles][32] {x, epoch}
};
// end of a look-back kernel: every thread of the CTA has issued its last ticket / word
__device__ __forceinline__ void ll_finish(LLCtl *ctl) {
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int prev;
    asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(prev) : "l"(&ctl->done) : "memory");
    if (prev == gridDim.x - 1) {  // all other CTAs read base / epoch before their own increment
      const unsigned long long t = *reinterpret_cast<volatile unsigned long long *>(&ctl->ticket);
      const unsigned int e = *reinterpret_cast<volatile unsigned int *>(&ctl->epoch);
      *reinterpret_cast<volatile unsigned long long *>(&ctl->base) = t;
      *reinterpret_cast<volatile unsigned int *>(&ctl->epoch) = e + 1u == 0u ? 1u : e + 1u;
      *reinterpret_cast<volatile unsigned int *>(&ctl->done) = 0u;
    }
  }
}

__device__ __forceinline__ uint4 ld_relaxed_v4(const uint4 *p) {
  uint4 v;
  asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p)
               : "memory");
  return v;
}
__device__ __forceinline__ uint2 ld_relaxed_v2(const uint2 *p) {
  uint2 v;
  asm volatile("ld.relaxed.gpu.global.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_v4(uint4 *p, uint4 v) {
  asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y),
               "r"(v.z), "r"(v.w)
               : "memory");
}
__device__ __forceinline__ void st_relaxed_v2(uint2 *p, uint2 v) {
  asm volatile("st.relaxed.gpu.global.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(v.x), "r"(v.y) : "memory");
}

// Everything a warp does with its tile once the loads are in registers.  INTERIOR: the tile lies
// completely inside [0,T) x [0,N), so no step or column needs masking (the common case; it halves the
// instruction count of the kernel, which is otherwise issue-bound rather than HBM-bound).
template <int OP, bool HD, bool INTERIOR>
__device__ __forceinline__ void wtile_body(const ScanArgs &A, const WTileWs &ws, unsigned int epoch,
                                           float (&a)[WT_S], float (&v)[WT_S + 1],
                                           const unsigned char (&db)[WT_S], float lastv, int chunk,
                                           int wgroup, int n_chunks, int n_wgroups, int lane,
                                           int64_t col, int64_t colc, bool valid) {
  constexpr unsigned FULL = 0xffffffffu;
  const int64_t t_base = (int64_t)chunk * WT_S;
  const int64_t t_max = A.T - 1;
  unsigned int dmask = 0u, inmask = 0xffffffffu;
  if (HD) {
#pragma unroll
    for (int s = 0; s < WT_S; ++s) dmask |= (db[s] ? 1u : 0u) << s;
  }
  if (!INTERIOR) {
    inmask = 0u;
#pragma unroll
    for (int s = 0; s < WT_S; ++s) inmask |= ((valid && t_base + s <= t_max) ? 1u : 0u) << s;
  }
  const float bgain = OP == OP_GAE ? A.gl : A.gamma;
  if (OP == OP_GAE) {
#pragma unroll
    for (int s = 0; s < WT_S; ++s) {
      const float vn = (!INTERIOR && t_base + s + 1 > t_max) ? lastv : v[s + 1];
      const float gv = __fmul_rn(A.gamma, vn);
      const float gvm = ((dmask >> s) & 1u) ? 0.0f : gv;
      a[s] = __fsub_rn(__fadd_rn(a[s], gvm), v[s]);
    }
  }
  // x_new = a + b * x with b = 0 on a done step, bgain otherwise (or v[s] for the generic scan)
  auto step = [&](int s, float x) -> float {
    if (OP == OP_AFFINE) return scan_step(a[s], v[s], x);
    const float y = scan_step(a[s], bgain, x);
    return ((dmask >> s) & 1u) ? a[s] : y;
  };
  float ca = 0.0f, cb = 1.0f;
#pragma unroll
  for (int s = WT_S - 1; s >= 0; --s) {
    if (INTERIOR || ((inmask >> s) & 1u)) {
      ca = step(s, ca);
      if (OP == OP_AFFINE) cb = __fmul_rn(v[s], cb);
      else cb = ((dmask >> s) & 1u) ? 0.0f : __fmul_rn(bgain, cb);
    }
  }

  const size_t tile = (size_t)chunk * n_wgroups + wgroup;
  if (chunk > 0)
    st_relaxed_v4(ws.agg + tile * 32 + lane,
                  make_uint4(__float_as_uint(ca), epoch, __float_as_uint(cb), epoch));

  // ---- look-back, WT_WIN later chunks per round trip --------------------------------------------
  float pa_ = 0.0f, pb_ = 1.0f, xin = 0.0f;
  int c = chunk + 1;
  bool found = false;
  while (!found) {
    if (c >= n_chunks) {  // ran off the end of the trajectory: boundary value
      float q = 0.0f;
      if (OP != OP_GAE && A.last) q = __ldcg(A.last + colc);
      xin = scan_step(pa_, pb_, q);
      break;
    }
    uint4 g[WT_WIN];
    uint2 q[WT_WIN];
    const int c0 = c;  // window = chunks c0 .. c0 + WT_WIN - 1
#pragma unroll
    for (int j = 0; j < WT_WIN; ++j) {
      const int cj = c0 + j < n_chunks ? c0 + j : n_chunks - 1;
      const size_t other = ((size_t)cj * n_wgroups + wgroup) * 32 + lane;
      g[j] = ld_relaxed_v4(ws.agg + other);
      q[j] = ld_relaxed_v2(ws.incl + other);
    }
    bool retry = false;
#pragma unroll
    for (int j = 0; j < WT_WIN; ++j) {
      if (found || retry) continue;
      if (c0 + j >= n_chunks) {  // window reaches the end: boundary value closes it
        float qq = 0.0f;
        if (OP != OP_GAE && A.last) qq = __ldcg(A.last + colc);
        xin = scan_step(pa_, pb_, qq);
        found = true;
      } else if (__all_sync(FULL, q[j].y == epoch)) {
        xin = scan_step(pa_, pb_, __uint_as_float(q[j].x));
        found = true;
      } else if (__all_sync(FULL, g[j].y == epoch && g[j].w == epoch)) {
        pa_ = scan_step(pa_, pb_, __uint_as_float(g[j].x));  // P o M
        pb_ = __fmul_rn(pb_, __uint_as_float(g[j].z));
        ++c;
      } else {
        retry = true;  // chunk c has published nothing yet: poll again from c
      }
    }
  }
  if (chunk > 0)
    st_relaxed_v2(ws.incl + tile * 32 + lane, make_uint2(__float_as_uint(scan_step(ca, cb, xin)), epoch));
  if (!INTERIOR && !valid) return;
  // ---- replay from registers ----------------------------------------------------------------
  float x = xin;
  // stepping pointers: one 64-bit add per store instead of a 64-bit multiply-add
  float *po = A.out + (t_base + WT_S - 1) * A.N + col;
  const int64_t dret = (OP == OP_GAE && A.ret) ? A.ret - A.out : 0;
  const bool want_ret = OP == OP_GAE && A.ret != nullptr;
#pragma unroll
  for (int s = WT_S - 1; s >= 0; --s) {
    if (INTERIOR || t_base + s <= t_max) {
      x = step(s, x);
      __stcs(po, x);
      if (want_ret) __stcs(po + dret, __fadd_rn(x, v[s]));
    }
    po -= A.N;
  }
}

template <int OP, bool HD>
__global__ void __launch_bounds__(WT_WARPS * 32, OP == OP_GAE ? 3 : 4)
    scan_tm_wtile_kernel(const __grid_constant__ ScanArgs A, WTileWs ws, int n_chunks, int n_wgroups,
                         int n_cgroups) {
  __shared__ unsigned int s_ticket, s_epoch;
  __shared__ int s_zero;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    const unsigned long long base = *reinterpret_cast<volatile unsigned long long *>(&ws.ctl->base);
    s_epoch = *reinterpret_cast<volatile unsigned int *>(&ws.ctl->epoch);
    s_ticket = (unsigned int)(atomicAdd(&ws.ctl->ticket, 1ull) - base);
    s_zero = 0;
  }
  __syncthreads();
  const unsigned int ticket = s_ticket, epoch = s_epoch;
  const int chunk = n_chunks - 1 - (int)(ticket / (unsigned)n_cgroups);  // latest chunk first
  int wgroup = (int)(ticket % (unsigned)n_cgroups) * WT_WARPS + w;
  const bool warp_on = wgroup < n_wgroups;
  wgroup = warp_on ? wgroup : n_wgroups - 1;  // idle warps shadow a real tile up to the barrier
  const int64_t col = (int64_t)wgroup * 32 + lane;
  const bool valid = warp_on && col < A.N;
  const int64_t colc = col < A.N ? col : A.N - 1;
  const int64_t dcol = A.E == 1 ? colc : colc / A.E;
  const int64_t t_base = (int64_t)chunk * WT_S;
  const int64_t t_max = A.T - 1;

  // ---- loads: unconditional and all in flight.  Interior tiles use plain strided pointers ------
  float a[WT_S], v[WT_S + 1];
  unsigned char db[WT_S];
  const bool interior = t_base + WT_S <= t_max;  // also covers v[WT_S]
  if (interior) {
    const float *pa = A.a + t_base * A.N + colc;
#pragma unroll
    for (int s = 0; s < WT_S; ++s) a[s] = __ldcs(pa + (int64_t)s * A.N);
    if (OP != OP_DR) {
      const float *pv = A.v + t_base * A.N + colc;
#pragma unroll
      for (int s = 0; s <= WT_S; ++s)
        if (s < WT_S || OP == OP_GAE) v[s] = __ldcs(pv + (int64_t)s * A.N);
    }
    if (HD) {
      const unsigned char *pd = A.done + t_base * A.B + dcol;
#pragma unroll
      for (int s = 0; s < WT_S; ++s) db[s] = __ldcs(pd + (int64_t)s * A.B);
    }
  } else {
#pragma unroll
    for (int s = 0; s < WT_S; ++s) {
      int64_t t = t_base + s;
      t = t > t_max ? t_max : t;
      a[s] = __ldcs(A.a + t * A.N + colc);
      if (HD) db[s] = __ldcs(A.done + t * A.B + dcol);
    }
    if (OP != OP_DR) {
#pragma unroll
      for (int s = 0; s <= WT_S; ++s) {
        int64_t t = t_base + s;
        t = t > t_max ? t_max : t;
        if (s < WT_S || OP == OP_GAE) v[s] = __ldcs(A.v + t * A.N + colc);
      }
    }
  }
  float lastv = 0.0f;  // (the next_value form of GAE is dispatched to another kernel)
  if (OP == OP_GAE && A.last) lastv = __ldcg(A.last + colc);
  // Scheduling fence (see scan_tm_tile4_kernel): the barrier pins the loads above, the never-taken
  // branch on an opaque zero pins their consumers below.
  __syncthreads();
  if (*reinterpret_cast<volatile int *>(&s_zero) != 0) return;
  if (warp_on) {
    const bool full_cols = ((int64_t)wgroup + 1) * 32 <= A.N;
    if (interior && full_cols)
      wtile_body<OP, HD, true>(A, ws, epoch, a, v, db, lastv, chunk, wgroup, n_chunks, n_wgroups, lane,
                               col, colc, valid);
    else
      wtile_body<OP, HD, false>(A, ws, epoch, a, v, db, lastv, chunk, wgroup, n_chunks, n_wgroups, lane,
                                col, colc, valid);
  }
  ll_finish(ws.ctl);
}

// ---- time-major single-pass scan, software-pipelined warp tiles (default for aligned shapes) --
// Same tiles, same look-back protocol and the same arithmetic as scan_tm_wtile_kernel, but warps are
// persistent and every tile is first staged in shared memory with cp.async (16 bytes per lane and
// instruction): a warp requests the tile of its NEXT ticket before it works on the current one, so
// its loads are in flight during the composite, the look-back and the stores instead of only at the
// start of a tile (37 us against 40 us for GAE over 4096 x 2048).  Needs T % 32 == 0, N % 32 == 0,
// E == 1.  What bounds both variants today is instruction issue, not HBM and not the look-back: ~42
// instructions per element (scalar address arithmetic for 96 accesses per lane and tile) on 12
// resident warps per SM; with the look-back and the ticket removed the kernel still takes 27 us for
// DiscountedReturn where a copy with the same access pattern takes 14 us (scratch/pattern_test.cu).
constexpr int PIPE_WARPS = 4;
struct PipeBuf {
  float a[WT_S][32];
  float v[WT_S + 1][32];
  unsigned char d[WT_S][32];
};

__device__ __forceinline__ void cp_async16(void *smem, const void *g) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(g)
               : "memory");
}
__device__ __forceinline__ void cp_async4(void *smem, const void *g) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(g)
               : "memory");
}

template <int OP, bool HD>
__global__ void __launch_bounds__(PIPE_WARPS * 32, 3)
    scan_tm_pipe_kernel(const __grid_constant__ ScanArgs A, WTileWs ws, int n_chunks, int n_wgroups) {
  extern __shared__ __align__(16) unsigned char pipe_smem[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  // (every warp reads the launch state itself, before its first ticket: the last CTA rewrites it only after
  // all CTAs -- hence all warps -- have finished)
  const unsigned long long ticket_base = *reinterpret_cast<volatile unsigned long long *>(&ws.ctl->base);
  const unsigned int epoch = *reinterpret_cast<volatile unsigned int *>(&ws.ctl->epoch);
  PipeBuf *bufs = reinterpret_cast<PipeBuf *>(pipe_smem) + 2 * w;
  const long long n_tiles = (long long)n_chunks * n_wgroups;
  constexpr unsigned FULL = 0xffffffffu;
  const int64_t t_max = A.T - 1;

  auto take = [&]() -> long long {
    unsigned long long t = 0;
    if (lane == 0) t = atomicAdd(&ws.ctl->ticket, 1ull) - ticket_base;
    t = __shfl_sync(FULL, t, 0);
    return (long long)t < n_tiles ? (long long)t : -1;
  };
  auto issue = [&](long long tk, PipeBuf &B) {
    const int chunk = n_chunks - 1 - (int)(tk / n_wgroups);  // latest chunk first
    const int wgroup = (int)(tk % n_wgroups);
    const int64_t t_base = (int64_t)chunk * WT_S;
    const int64_t col0 = (int64_t)wgroup * 32 + (lane & 7) * 4;
    const int r0 = lane >> 3;
#pragma unroll
    for (int i = 0; i < WT_S / 4; ++i) {
      const int row = r0 + 4 * i;
      cp_async16(&B.a[row][(lane & 7) * 4], A.a + (t_base + row) * A.N + col0);
      if (HD) cp_async4(&B.d[row][(lane & 7) * 4], A.done + (t_base + row) * A.B + col0);
    }
    if (OP != OP_DR) {
#pragma unroll
      for (int i = 0; i <= WT_S / 4; ++i) {
        const int row = r0 + 4 * i;
        if (row < WT_S || (OP == OP_GAE && row == WT_S)) {
          int64_t t = t_base + row;
          t = t > t_max ? t_max : t;  // row 32 of the last chunk: unused (the bootstrap takes its place)
          cp_async16(&B.v[row][(lane & 7) * 4], A.v + t * A.N + col0);
        }
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  long long cur = take();
  int cb = 0;
  if (cur >= 0) issue(cur, bufs[0]);
  while (cur >= 0) {
    const long long nxt = take();
    if (nxt >= 0) {
      issue(nxt, bufs[cb ^ 1]);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncwarp();
    const PipeBuf &B = bufs[cb];
    float a[WT_S], v[WT_S + 1];
    unsigned char db[WT_S];
#pragma unroll
    for (int s = 0; s < WT_S; ++s) {
      a[s] = B.a[s][lane];
      if (HD) db[s] = B.d[s][lane];
    }
    if (OP != OP_DR) {
#pragma unroll
      for (int s = 0; s <= WT_S; ++s)
        if (s < WT_S || OP == OP_GAE) v[s] = B.v[s][lane];
    }
    __syncwarp();  // the buffer may be refilled by the next iteration's prefetch
    const int chunk = n_chunks - 1 - (int)(cur / n_wgroups);
    const int wgroup = (int)(cur % n_wgroups);
    const int64_t col = (int64_t)wgroup * 32 + lane;
    if (chunk == n_chunks - 1) {  // holds the bootstrap step
      float lastv = 0.0f;
      if (OP == OP_GAE && A.last) lastv = __ldcg(A.last + col);
      wtile_body<OP, HD, false>(A, ws, epoch, a, v, db, lastv, chunk, wgroup, n_chunks, n_wgroups, lane, col,
                                col, true);
    } else {
      wtile_body<OP, HD, true>(A, ws, epoch, a, v, db, 0.0f, chunk, wgroup, n_chunks, n_wgroups, lane, col,
                               col, true);
    }
    cur = nxt;
    cb ^= 1;
  }
  ll_finish(ws.ctl);
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
  if (OP != OP_GAE && A.T >= 2048) scan_bm_warp_kernel<OP, VEC, HAS_NV, 8><<<(unsigned)blocks, 256, 0, st>>>(A);
  else if (A.T > 512) scan_bm_warp_kernel<OP, VEC, HAS_NV, 4><<<(unsigned)blocks, 256, 0, st>>>(A);
  else scan_bm_warp_kernel<OP, VEC, HAS_NV, 1><<<(unsigned)blocks, 256, 0, st>>>(A);
}


static bool aligned16(const void *p) { return (((uintptr_t)p) & 15) == 0; }

// Workspaces are keyed by (device, stream): two scans running at once on different streams or devices never
// share flag words, partials or scratch (process-global statics let concurrent launches corrupt
// each other's look-back words, and lived on whichever device came first).  Work on ONE stream is ordered by
// the stream.  Buffers only grow; a retired buffer is kept until exit, because a captured graph may still
// point at it.
struct StreamWs {
  uint8_t *ll = nullptr;  // look-back words: [LLCtl | agg | incl]
  size_t ll_bytes = 0;
  uint8_t *tile = nullptr;  // tile4 kernel: flags, aggregates
  size_t tile_bytes = 0;
  double2 *moments = nullptr;  // [kMomentSlots] partial (sum, sum of squares) + [1] total + ticket
  uint8_t *scratch = nullptr;  // *_host scans, V-trace intermediates
  size_t scratch_bytes = 0;
};
static std::mutex g_ws_mu;
static std::map<std::pair<int, cudaStream_t>, StreamWs> g_ws;

static StreamWs *stream_ws(cudaStream_t st) {
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> g(g_ws_mu);
  return &g_ws[std::make_pair(dev, st)];  // (std::map: references stay valid)
}
static bool capturing(cudaStream_t st) {
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cap) != cudaSuccess) {
    cudaGetLastError();
    return true;
  }
  return cap != cudaStreamCaptureStatusNone;
}
// grow-only buffer of a stream workspace; `zero`: cleared when (re)allocated.  Allocation is not a stream
// operation, so a capturing stream must find the buffer in place (run the call once before capturing).
static int ws_reserve(uint8_t **buf, size_t *have, size_t need, bool zero, cudaStream_t st, bool *fresh = nullptr) {
  if (fresh) *fresh = false;
  if (need <= *have) return MRL_OK;
  if (capturing(st)) {
    set_error("this call needs %zu bytes of workspace that are not allocated yet: run it once on this stream "
              "before capturing it in a CUDA graph", need);
    return MRL_ERR_STATE;
  }
  size_t cap = 1 << 16;
  while (cap < need) cap <<= 1;
  uint8_t *p = nullptr;
  MRL_CUDA(cudaMalloc((void **)&p, cap));
  if (zero) MRL_CUDA(cudaMemsetAsync(p, 0, cap, st));
  *buf = p, *have = cap;  // (the previous buffer stays allocated: graphs / launches in flight may use it)
  if (fresh) *fresh = true;
  return MRL_OK;
}

static int g_tile_mode = -1;  // MRL_TM_KERNEL=chunked forces the two-pass kernel
static int64_t g_host_chunks = 0;  // option "host_scan_chunks": pieces of the *_host scans, 0 = auto, 1 = no pipelining
static int g_bm_mode = 0;     // 0 auto, 1 warp-per-row register kernel only, 2 TMA row kernel whenever it fits
extern int64_t g_strip_cols, g_strip_rpu, g_strip_min_cols, g_strip_stages, g_strip_groups, g_row_ctas_per_sm, g_row_stages, g_row_spl;  // scan_tma.cu

// "tm_kernel": 0 auto, 2 two-pass chunked, 3 128-column CTA tiles, 5 warp tiles with look-back,
// 6 TMA strips whenever the alignment allows; "strip_cols": 0 auto / 16 / 32; "bm_kernel": see g_bm_mode
int set_scan_option(const char *key, int64_t value) {
  std::string k(key);
  if (k == "tm_kernel" && (value == 0 || value == 2 || value == 3 || value == 5 || value == 6)) g_tile_mode = (int)value;
  else if (k == "strip_cols" && (value == 0 || value == 16 || value == 28 || value == 32)) g_strip_cols = value;
  else if (k == "strip_rpu" && (value == 0 || value == 8 || value == 16 || value == 32)) g_strip_rpu = value;
  else if (k == "row_ctas_per_sm" && value >= 1 && value <= 12) g_row_ctas_per_sm = value;
  else if (k == "row_spl" && (value == 8 || value == 16)) g_row_spl = value;
  else if (k == "row_stages" && value >= 0 && value <= 12) g_row_stages = value;
  else if (k == "strip_groups" && value >= 0 && value <= 2) g_strip_groups = value;
  else if (k == "strip_stages" && value >= 0 && value <= 12) g_strip_stages = value;
  else if (k == "strip_min_cols" && value >= 0) g_strip_min_cols = value;
  else if (k == "bm_kernel" && value >= 0 && value <= 2) g_bm_mode = (int)value;
  else if (k == "host_scan_chunks" && value >= 0 && value <= 16) g_host_chunks = value;
  else return MRL_ERR_INVALID;
  return MRL_OK;
}

template <int OP, bool HD>
static int launch_scan_hd(ScanArgs A, int32_t layout, int32_t mode, cudaStream_t st, int *mom_parts) {
  if (layout == MRL_TIME_MAJOR) {
    A.st = A.N, A.sc = 1, A.dst = A.B, A.dsc = 1;
    bool chunked;
    if (mode == MRL_SCAN_SEQUENTIAL) chunked = false;
    else if (mode == MRL_SCAN_PARALLEL) chunked = A.T >= 2;
    else chunked = A.T >= 64 && A.N < (int64_t)sm_count() * 1024;
    if (chunked && (g_tile_mode == 0 || g_tile_mode == 6)) {
      const int rc = scan_tm_strip_launch(A, OP, g_tile_mode == 6, st, mom_parts);
      if (rc != kScanNotApplicable) {
        if (rc != MRL_OK) return rc;
        MRL_LAUNCHED();
        return MRL_OK;
      }
    }
    const bool vec4 = A.E == 1 && A.N % 4 == 0 && aligned16(A.a) && aligned16(A.out) &&
                      (OP == OP_DR || aligned16(A.v)) && (!A.nv || aligned16(A.nv)) &&
                      (!A.ret || aligned16(A.ret)) && (!A.last || aligned16(A.last)) &&
                      (!A.done || (((uintptr_t)A.done) & 3) == 0);
    const bool pipe_ok = A.E == 1 && A.T % WT_S == 0 && A.N % 32 == 0 && aligned16(A.a) &&
                         (OP == OP_DR || aligned16(A.v)) && (!A.done || (((uintptr_t)A.done) & 3) == 0);
    if (chunked && (g_tile_mode == 5 || g_tile_mode == 0 || g_tile_mode == 6) && !(OP == OP_GAE && A.nv)) {
      const bool pipe = pipe_ok && g_tile_mode != 5;
      const int64_t nch = (A.T + WT_S - 1) / WT_S;
      const int64_t n_wgroups = (A.N + 31) / 32;
      const int64_t n_cgroups = (n_wgroups + WT_WARPS - 1) / WT_WARPS;
      const size_t tiles = (size_t)nch * n_wgroups;
      const size_t agg_bytes = tiles * 32 * sizeof(uint4);
      const size_t need = 256 + agg_bytes + tiles * 32 * sizeof(uint2);
      // flag-in-payload workspace of this (device, stream): zeroed when (re)allocated, then validated by the
      // launch epoch, which the kernels advance themselves (LLCtl) -- capturable in a CUDA graph
      StreamWs *W = stream_ws(st);
      bool fresh = false;
      int rc = ws_reserve(&W->ll, &W->ll_bytes, need, true, st, &fresh);
      if (rc) return rc;
      if (fresh) {
        const LLCtl init = {0ull, 0ull, 1u, 0u};
        MRL_CUDA(cudaMemcpyAsync(W->ll, &init, sizeof(init), cudaMemcpyHostToDevice, st));
      }
      WTileWs ws;
      ws.ctl = reinterpret_cast<LLCtl *>(W->ll);
      ws.agg = reinterpret_cast<uint4 *>(W->ll + 256);
      ws.incl = reinterpret_cast<uint2 *>(W->ll + 256 + agg_bytes);
      if (pipe) {
        const size_t smem = sizeof(PipeBuf) * 2 * PIPE_WARPS;
        if (first_use_on_device((const void *)scan_tm_pipe_kernel<OP, HD>))
          MRL_CUDA(cudaFuncSetAttribute(scan_tm_pipe_kernel<OP, HD>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int64_t ctas = (int64_t)sm_count() * 3;
        if (ctas * PIPE_WARPS > (int64_t)tiles) ctas = ((int64_t)tiles + PIPE_WARPS - 1) / PIPE_WARPS;
        scan_tm_pipe_kernel<OP, HD><<<(unsigned)ctas, PIPE_WARPS * 32, smem, st>>>(A, ws, (int)nch, (int)n_wgroups);
      } else {
        scan_tm_wtile_kernel<OP, HD><<<(unsigned)(nch * n_cgroups), WT_WARPS * 32, 0, st>>>(
            A, ws, (int)nch, (int)n_wgroups, (int)n_cgroups);
      }
    } else if (chunked && (g_tile_mode == 3 || g_tile_mode == 0 || g_tile_mode == 6) && vec4) {
      const int64_t n4 = (A.T + T4_TC - 1) / T4_TC;
      const int64_t n_groups = (A.N + T4_COLS - 1) / T4_COLS;
      const size_t tiles = (size_t)n4 * n_groups;
      const size_t flag_bytes = (4 + tiles * 4 + 255) / 256 * 256;
      const size_t agg_bytes = tiles * T4_COLS * sizeof(float2);
      StreamWs *W = stream_ws(st);
      int rc = ws_reserve(&W->tile, &W->tile_bytes, flag_bytes + agg_bytes + tiles * T4_COLS * sizeof(float), false, st);
      if (rc) return rc;
      uint8_t *wsp = W->tile;
      MRL_CUDA(cudaMemsetAsync(wsp, 0, flag_bytes, st));
      Tile4Ws ws;
      ws.ticket = reinterpret_cast<unsigned int *>(wsp);
      ws.flag = ws.ticket + 1;
      ws.agg = reinterpret_cast<float2 *>(wsp + flag_bytes);
      ws.incl = reinterpret_cast<float *>(wsp + flag_bytes + agg_bytes);
      dim3 block(32, T4_W);
      scan_tm_tile4_kernel<OP, HD><<<(unsigned)tiles, block, 0, st>>>(A, ws, (int)n4, (int)n_groups);
    } else if (chunked) {
      int W = 32;
      while (W > 1 && A.T / W < 8) W >>= 1;
      const int64_t chunk = (A.T + W - 1) / W;
      dim3 block(32, W);
      const int64_t blocks = (A.N + 31) / 32;
      scan_tm_chunked_kernel<OP, HD><<<(unsigned)blocks, block, 0, st>>>(A, chunk);
    } else {
      const int64_t blocks = (A.N + 127) / 128;
      scan_seq_kernel<OP, HD><<<(unsigned)blocks, 128, 0, st>>>(A);
    }
  } else {
    A.st = 1, A.sc = A.T, A.dst = 1, A.dsc = A.T;
    if (mode == MRL_SCAN_SEQUENTIAL) {
      const int64_t blocks = (A.N + 127) / 128;
      scan_seq_kernel<OP, HD><<<(unsigned)blocks, 128, 0, st>>>(A);
    } else {
      if (g_bm_mode != 1) {
        const int rc = scan_bm_row_launch(A, OP, g_bm_mode == 2, st, mom_parts);
        if (rc != kScanNotApplicable) {
          if (rc != MRL_OK) return rc;
          MRL_LAUNCHED();
          return MRL_OK;
        }
      }
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

template <int OP>
static int launch_scan(ScanArgs A, int32_t layout, int32_t mode, cudaStream_t st, int *mom_parts = nullptr) {
  if (mom_parts) *mom_parts = 0;
  if (A.T <= 0 || A.N <= 0) return MRL_OK;
  if (g_tile_mode < 0) {
    const char *e = getenv("MRL_TM_KERNEL");
    // MRL_TM_KERNEL (same choices as option "tm_kernel"): default auto = TMA strips for wide aligned batches,
    // else warp tiles with look-back (cp.async pipeline when aligned); "s" strips whenever aligned, "w" warp
    // tiles only, "4" 128-column CTA tiles (aligned; also serves the next_value form of GAE), "c" two-pass
    g_tile_mode = 0;
    if (e && e[0] == 'c') g_tile_mode = 2;
    if (e && e[0] == '4') g_tile_mode = 3;
    if (e && e[0] == 'w') g_tile_mode = 5;
    if (e && e[0] == 's') g_tile_mode = 6;
  }
  if (OP != OP_AFFINE && A.done) return launch_scan_hd<OP, true>(A, layout, mode, st, mom_parts);
  return launch_scan_hd<OP, false>(A, layout, mode, st, mom_parts);
}

// ---- advantage normalisation ----------------------------------------------------------------
// (x - mean) / (sqrt(var) + eps) over all elements, moments in float64 and a fixed order (ppo.py:361-368,
// a2c.py:190-192).  The moments come as per-CTA partials: from the scan kernel that produced x (ScanArgs.mom:
// no extra pass) or from norm_partial_kernel; the apply pass folds them in index order in every CTA.
__global__ void __launch_bounds__(256)
    norm_partial_kernel(const float *__restrict__ x, int64_t n, double2 *__restrict__ part) {
  __shared__ double ss[8], sq[8];
  double s = 0.0, q = 0.0;
  const int64_t n4 = n / 4;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const bool al = (((uintptr_t)x) & 15) == 0;
  if (al) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
      const float4 v = __ldg(reinterpret_cast<const float4 *>(x) + i);
      s += (double)v.x + (double)v.y + (double)v.z + (double)v.w;
      q += (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z + (double)v.w * v.w;
    }
  }
  for (int64_t i = (al ? n4 * 4 : 0) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const double v = __ldg(x + i);
    s += v;
    q += v * v;
  }
#pragma unroll
  for (int off = 16; off; off >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, off);
    q += __shfl_xor_sync(0xffffffffu, q, off);
  }
  if ((threadIdx.x & 31) == 0) ss[threadIdx.x >> 5] = s, sq[threadIdx.x >> 5] = q;
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
    for (int w = 0; w < 8; ++w) a += ss[w], b += sq[w];
    part[blockIdx.x] = make_double2(a, b);
  }
}

__global__ void __launch_bounds__(256)
    norm_apply_kernel(const float *__restrict__ x, float *__restrict__ out, int64_t n, float eps,
                      const double2 *__restrict__ part, int nparts) {
  __shared__ double ra[8], rb[8];
  __shared__ float s_mean, s_inv;
  // (launched as a programmatic dependent of the scan that produced x and the partials: resident while that kernel
  // drains, released when its memory is flushed; a no-op after an ordinary launch)
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool vec = ((((uintptr_t)x) | ((uintptr_t)out)) & 15) == 0;
  const int64_t n4 = vec ? n / 4 : 0;
  // the first two quads of this thread are requested BEFORE the partials are folded: the fold (a dependent chain
  // of loads, shuffles and a barrier, ~2 us) then runs under their latency instead of in front of the stream
  constexpr int U = 2;
  float4 pre[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const int64_t i = gid + u * stride;
    pre[u] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < n4) pre[u] = __ldcs(reinterpret_cast<const float4 *>(x) + i);
  }
  {
    // 256 threads, contiguous runs of the partials, folded in index order (same in every CTA)
    const int per = (nparts + 255) / 256;
    double a = 0.0, b = 0.0;
    for (int i = threadIdx.x * per; i < nparts && i < (threadIdx.x + 1) * per; ++i) {
      const double2 v = __ldcg(part + i);
      a += v.x, b += v.y;
    }
#pragma unroll
    for (int off = 16; off; off >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, off);
      b += __shfl_xor_sync(0xffffffffu, b, off);
    }
    if ((threadIdx.x & 31) == 0) ra[threadIdx.x >> 5] = a, rb[threadIdx.x >> 5] = b;
    __syncthreads();
    if (threadIdx.x == 0) {
      a = 0.0, b = 0.0;
      for (int w = 0; w < 8; ++w) a += ra[w], b += rb[w];
      const double mean = a / (double)n;
      double var = b / (double)n - mean * mean;
      var = var > 0.0 ? var : 0.0;
      s_mean = (float)mean;
      s_inv = 1.0f / (sqrtf((float)var) + eps);
    }
  }
  __syncthreads();
  const float mean = s_mean, inv = s_inv;
  auto norm4 = [&](const float4 v) {
    return make_float4(__fmul_rn(__fsub_rn(v.x, mean), inv), __fmul_rn(__fsub_rn(v.y, mean), inv),
                       __fmul_rn(__fsub_rn(v.z, mean), inv), __fmul_rn(__fsub_rn(v.w, mean), inv));
  };
  for (int64_t i0 = gid; i0 < n4; i0 += U * stride) {
    float4 nxt[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {  // the next round's loads go out before this round's stores
      const int64_t i = i0 + (U + u) * stride;
      nxt[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i < n4) nxt[u] = __ldcs(reinterpret_cast<const float4 *>(x) + i);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = i0 + u * stride;
      if (i < n4) __stcs(reinterpret_cast<float4 *>(out) + i, norm4(pre[u]));
    }
#pragma unroll
    for (int u = 0; u < U; ++u) pre[u] = nxt[u];
  }
  for (int64_t i = n4 * 4 + gid; i < n; i += stride) out[i] = __fmul_rn(__fsub_rn(x[i], mean), inv);
}

static int moments_ws(cudaStream_t st, double2 **p) {
  StreamWs *W = stream_ws(st);
  if (!W->moments) {
    if (capturing(st)) {
      set_error("the normalisation workspace of this stream is not allocated yet: run the call once before capturing");
      return MRL_ERR_STATE;
    }
    MRL_CUDA(cudaMalloc((void **)&W->moments, sizeof(double2) * kMomentSlots));
  }
  *p = W->moments;
  return MRL_OK;
}

// x -> out with the moments taken from `nparts` partials already in `part` (nparts == 0: computed here)
static int normalize_impl(const float *x, float *out, int64_t n, float eps, double2 *part, int nparts,
                          cudaStream_t st) {
  int64_t ctas = (n + 2047) / 2048;
  const int64_t cap = (int64_t)sm_count() * 8;
  ctas = ctas > cap ? cap : ctas;
  if (nparts == 0) {
    nparts = (int)(ctas > kMomentSlots ? kMomentSlots : ctas);
    norm_partial_kernel<<<nparts, 256, 0, st>>>(x, n, part);
    MRL_LAUNCHED();
  }
  MRL_CUDA(launch_pdl(norm_apply_kernel, dim3((unsigned)ctas), dim3(256), 0, st, x, out, n, eps,
                      (const double2 *)part, nparts));
  MRL_LAUNCHED();
  return MRL_OK;
}

// grow-only device scratch of a stream (the *_host entry points, V-trace intermediates)
static int scratch(size_t need, uint8_t **p, cudaStream_t st) {
  StreamWs *W = stream_ws(st);
  const int rc = ws_reserve(&W->scratch, &W->scratch_bytes, need, false, st);
  *p = W->scratch;
  return rc;
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

extern "C" int mrl_normalize(const float *x, float *out, int64_t n, float eps, void *stream) {
  MRL_REQUIRE(n >= 0 && (n == 0 || (x && out)), "mrl_normalize: bad arguments");
  if (n == 0) return MRL_OK;
  cudaStream_t st = as_stream(stream);
  double2 *part;
  const int rc = moments_ws(st, &part);
  if (rc) return rc;
  return normalize_impl(x, out, n, eps, part, 0, st);
}

extern "C" int mrl_gae_normalized(const float *reward, const float *value, const float *next_value,
                                  const float *last_value, const uint8_t *done, float *adv, float *ret,
                                  int64_t T, int64_t B, float gamma, float lambda, float eps, int32_t layout,
                                  int32_t mode, void *stream) {
  MRL_REQUIRE(T >= 0 && B >= 0, "mrl_gae_normalized: bad shape");
  MRL_REQUIRE(layout == MRL_TIME_MAJOR || layout == MRL_BATCH_MAJOR, "mrl_gae_normalized: bad layout");
  if (T == 0 || B == 0) return MRL_OK;
  MRL_REQUIRE(reward && value && adv, "mrl_gae_normalized: NULL reward/value/adv");
  cudaStream_t st = as_stream(stream);
  ScanArgs A{};
  A.a = reward, A.v = value, A.nv = next_value, A.last = last_value, A.done = done;
  A.out = adv, A.ret = ret;
  A.T = T, A.B = B, A.E = 1, A.N = B, A.gamma = gamma, A.gl = gamma * lambda;
  int rc = moments_ws(st, &A.mom);
  if (rc) return rc;
  int parts = 0;
  rc = launch_scan<OP_GAE>(A, layout, mode, st, &parts);
  if (rc) return rc;
  return normalize_impl(adv, adv, T * B, eps, A.mom, parts, st);
}

extern "C" int mrl_discounted_return_normalized(const float *reward, const uint8_t *done, const float *last_value,
                                                float *out, int64_t T, int64_t B, int64_t E, float gamma,
                                                float eps, int32_t layout, int32_t mode, void *stream) {
  MRL_REQUIRE(T >= 0 && B >= 0 && E >= 1, "mrl_discounted_return_normalized: bad shape");
  MRL_REQUIRE(layout == MRL_TIME_MAJOR || (layout == MRL_BATCH_MAJOR && E == 1),
              "mrl_discounted_return_normalized: batch-major layout needs E == 1");
  MRL_REQUIRE(gamma >= 0.0f && gamma <= 1.0f,
              "The discounted factor should be a number in range [0, 1], but got %f.", (double)gamma);
  if (T == 0 || B == 0) return MRL_OK;
  MRL_REQUIRE(reward && out, "mrl_discounted_return_normalized: NULL reward/out");
  cudaStream_t st = as_stream(stream);
  ScanArgs A{};
  A.a = reward, A.done = done, A.last = last_value, A.out = out;
  A.T = T, A.B = B, A.E = E, A.N = B * E, A.gamma = gamma;
  int rc = moments_ws(st, &A.mom);
  if (rc) return rc;
  int parts = 0;
  rc = launch_scan<OP_DR>(A, layout, mode, st, &parts);
  if (rc) return rc;
  return normalize_impl(out, out, T * B * E, eps, A.mom, parts, st);
}

namespace mrl {
// ---- V-trace (IMPALA) -----------------------------------------------------------------------
// get_vs_and_advantages (mindspore_rl/algorithm/impala/vtrace.py:48-108), time-major [T, B]:
//   rho = min(exp(log_rho), clip_rho); c = min(rho, clip_c); pg_rho = min(rho, clip_pg)          (:79-82)
//   delta_t = rho_t * (r_t + discount_t * V_{t+1} - V_t),  V_T = bootstrap                       (:84-86)
//   acc_t = delta_t + discount_t * c_t * acc_{t+1} * mask_t,  acc_T = 0                          (:88-93)
//   vs_t = acc_t + V_t                                                                          (:99)
//   pg_adv_t = pg_rho_t * (r_t + discount_t * vs_{t+1} - V_t),  vs_T = bootstrap                 (:103-105)
// The reference runs ~12 element-wise ops and a Python loop over T; here it is either ONE kernel (thread per
// column walking t = T-1..0 in the reference's operation order with 4 steps of loads in flight: 28 bytes per
// element, everything else in registers -- vs_{t+1} is simply the previous step's result) or, for long and
// narrow trajectories, prep -> parallel affine scan -> post (three passes instead of twelve).
struct VtraceArgs {
  const float *log_rhos, *discounts, *rewards, *values, *bootstrap, *masks;  // masks may be NULL (= 1)
  float clip_rho, clip_pg, clip_c;                                           // +inf = no clipping
  float *vs, *pg;
  int64_t T, B;
};

__global__ void __launch_bounds__(128) vtrace_seq_kernel(const __grid_constant__ VtraceArgs V) {
  const int64_t col = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= V.B) return;
  constexpr int U = 4;
  const float boot = __ldg(V.bootstrap + col);
  float acc = 0.0f;      // vs_minus_v_xs of the step after
  float v_next = boot;   // V_{t+1}
  float vs_next = boot;  // vs_{t+1}
  for (int64_t t0 = V.T; t0 > 0; t0 -= U) {
    float lr[U], dc[U], rw[U], vl[U], mk[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      int64_t t = t0 - 1 - u;
      t = t < 0 ? 0 : t;  // clamped: the load is always legal, the result may be unused
      const int64_t i = t * V.B + col;
      lr[u] = __ldg(V.log_rhos + i), dc[u] = __ldg(V.discounts + i), rw[u] = __ldg(V.rewards + i);
      vl[u] = __ldg(V.values + i), mk[u] = V.masks ? __ldg(V.masks + i) : 1.0f;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t t = t0 - 1 - u;
      if (t >= 0) {
        const float rho = fminf(expf(lr[u]), V.clip_rho);
        const float c = fminf(rho, V.clip_c), pg_rho = fminf(rho, V.clip_pg);
        const float delta = __fmul_rn(rho, __fsub_rn(__fadd_rn(rw[u], __fmul_rn(dc[u], v_next)), vl[u]));
        acc = __fadd_rn(delta, __fmul_rn(__fmul_rn(__fmul_rn(dc[u], c), acc), mk[u]));
        const float vs = __fadd_rn(acc, vl[u]);
        const int64_t i = t * V.B + col;
        V.vs[i] = vs;
        V.pg[i] = __fmul_rn(pg_rho, __fsub_rn(__fadd_rn(rw[u], __fmul_rn(dc[u], vs_next)), vl[u]));
        v_next = vl[u], vs_next = vs;
      }
    }
  }
}

// a = delta, b = discount * c * mask for the affine scan
__global__ void __launch_bounds__(256) vtrace_prep_kernel(const __grid_constant__ VtraceArgs V, float *__restrict__ a,
                                                          float *__restrict__ b) {
  const int64_t n = V.T * V.B, stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const bool lastrow = i >= n - V.B;
    const float v1 = lastrow ? __ldg(V.bootstrap + (i - (n - V.B))) : __ldg(V.values + i + V.B);
    const float rho = fminf(expf(__ldg(V.log_rhos + i)), V.clip_rho);
    const float c = fminf(rho, V.clip_c);
    const float dc = __ldg(V.discounts + i);
    a[i] = __fmul_rn(rho, __fsub_rn(__fadd_rn(__ldg(V.rewards + i), __fmul_rn(dc, v1)), __ldg(V.values + i)));
    b[i] = __fmul_rn(__fmul_rn(dc, c), V.masks ? __ldg(V.masks + i) : 1.0f);
  }
}

// acc (in V.vs) -> vs and pg_adv; row t needs acc and V of row t+1: rows are walked by whole CTAs so that the
// in-place update of row t+1 never races the read of it (vs is written to V.vs only after pg used acc[t+1])
__global__ void __launch_bounds__(256) vtrace_post_kernel(const __grid_constant__ VtraceArgs V, const float *__restrict__ acc) {
  const int64_t n = V.T * V.B, stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const bool lastrow = i >= n - V.B;
    const float vl = __ldg(V.values + i);
    const float vs = __fadd_rn(__ldg(acc + i), vl);
    const float vs1 = lastrow ? __ldg(V.bootstrap + (i - (n - V.B)))
                              : __fadd_rn(__ldg(acc + i + V.B), __ldg(V.values + i + V.B));
    const float rho = fminf(expf(__ldg(V.log_rhos + i)), V.clip_rho);
    const float pg_rho = fminf(rho, V.clip_pg);
    V.vs[i] = vs;
    V.pg[i] = __fmul_rn(pg_rho, __fsub_rn(__fadd_rn(__ldg(V.rewards + i), __fmul_rn(__ldg(V.discounts + i), vs1)), vl));
  }
}

}  // namespace mrl
using namespace mrl;

extern "C" int mrl_vtrace(const float *log_rhos, const float *discounts, const float *rewards, const float *values,
                          const float *bootstrap_value, const float *masks, float clip_rho_threshold,
                          float clip_pg_threshold, float clip_cs_threshold, float *vs, float *pg_advantages,
                          int64_t T, int64_t B, int32_t mode, void *stream) {
  MRL_REQUIRE(T >= 0 && B >= 0, "mrl_vtrace: bad shape");
  if (T == 0 || B == 0) return MRL_OK;
  MRL_REQUIRE(log_rhos && discounts && rewards && values && bootstrap_value && vs && pg_advantages,
              "mrl_vtrace: NULL argument");
  MRL_REQUIRE(vs != pg_advantages, "mrl_vtrace: vs and pg_advantages must be distinct arrays");
  cudaStream_t st = as_stream(stream);
  VtraceArgs V;
  V.log_rhos = log_rhos, V.discounts = discounts, V.rewards = rewards, V.values = values;
  V.bootstrap = bootstrap_value, V.masks = masks;
  V.clip_rho = clip_rho_threshold, V.clip_pg = clip_pg_threshold, V.clip_c = clip_cs_threshold;
  V.vs = vs, V.pg = pg_advantages, V.T = T, V.B = B;
  const bool seq = mode == MRL_SCAN_SEQUENTIAL ||
                   (mode == MRL_SCAN_AUTO && (B >= 32768 || T < 64 || T * B < (1 << 18)));
  if (seq) {
    vtrace_seq_kernel<<<(unsigned)((B + 127) / 128), 128, 0, st>>>(V);
    MRL_LAUNCHED();
    return MRL_OK;
  }
  const size_t f = ((size_t)T * B * 4 + 255) / 256 * 256;
  uint8_t *sc;
  int rc = scratch(3 * f, &sc, st);
  if (rc) return rc;
  float *a = (float *)sc, *b = (float *)(sc + f), *acc = (float *)(sc + 2 * f);
  int64_t ctas = (T * B + 1023) / 1024;
  const int64_t cap = (int64_t)sm_count() * 8;
  ctas = ctas > cap ? cap : ctas;
  vtrace_prep_kernel<<<(unsigned)ctas, 256, 0, st>>>(V, a, b);
  MRL_LAUNCHED();
  ScanArgs A{};
  A.a = a, A.v = b, A.last = nullptr, A.out = acc;
  A.T = T, A.B = B, A.E = 1, A.N = B;
  rc = launch_scan<OP_AFFINE>(A, MRL_TIME_MAJOR, MRL_SCAN_PARALLEL, st);
  if (rc) return rc;
  vtrace_post_kernel<<<(unsigned)ctas, 256, 0, st>>>(V, acc);
  MRL_LAUNCHED();
  return MRL_OK;
}

// ---- TD(lambda) targets of COMA ---------------------------------------------------------------
// build_td_lambda_targets (mindspore_rl/algorithm/coma/coma.py:489-501), batch-major:
//   ret[:, t] = (td_lambda*gamma) * ret[:, t+1]
//               + mask[:, t] * (rewards[:, t] + ((1-td_lambda)*gamma) * target_qs[:, t+1] * (1 - terminated[:, t]))
// t = max_t-1 .. 0 on a zero-initialised ret shaped like target_qs [B, Tq, A] (Tq > max_t); rewards, terminated,
// mask are [B, Tr, RA] with RA = 1 (broadcast over the A agents) or A.  Episodes x agents is a few hundred
// rows of <= a few hundred steps: one thread per (episode, agent) row walks it in the reference's operation
// order (bit-identical), adjacent threads = adjacent agents (coalesced).
namespace mrl {
__global__ void __launch_bounds__(128)
    td_lambda_kernel(const float *__restrict__ rewards, const float *__restrict__ terminated,
                     const float *__restrict__ mask, const float *__restrict__ q, float *__restrict__ ret,
                     int64_t B, int64_t Tq, int64_t Tr, int64_t A, int64_t RA, int64_t max_t, float lg, float bg) {
  const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= B * A) return;
  const int64_t b = row / A, a = row - b * A;
  const float *qr = q + b * Tq * A + a;
  float *rr = ret + b * Tq * A + a;
  const int64_t ra = RA == 1 ? 0 : a;
  for (int64_t t = max_t; t < Tq; ++t) rr[t * A] = 0.0f;  // zeros_like(target_qs) beyond the recurrence
  float x = 0.0f;
  for (int64_t t = max_t - 1; t >= 0; --t) {
    const int64_t i = (b * Tr + t) * RA + ra;
    const float boot = __fmul_rn(__fmul_rn(bg, qr[(t + 1) * A]), __fsub_rn(1.0f, terminated[i]));
    x = __fadd_rn(__fmul_rn(lg, x), __fmul_rn(mask[i], __fadd_rn(rewards[i], boot)));
    rr[t * A] = x;
  }
}
}  // namespace mrl

extern "C" int mrl_td_lambda(const float *rewards, const float *terminated, const float *mask,
                             const float *target_qs, float *ret, int64_t B, int64_t Tq, int64_t Tr, int64_t A,
                             int64_t RA, int64_t max_t, float lambda_gamma, float bootstrap_gamma, void *stream) {
  MRL_REQUIRE(B >= 0 && Tq >= 1 && Tr >= 0 && A >= 1 && (RA == 1 || RA == A), "mrl_td_lambda: bad shape");
  MRL_REQUIRE(max_t >= 0 && max_t < Tq && max_t <= Tr, "mrl_td_lambda: need max_t < Tq and max_t <= Tr");
  if (B == 0) return MRL_OK;
  MRL_REQUIRE(rewards && terminated && mask && target_qs && ret, "mrl_td_lambda: NULL argument");
  td_lambda_kernel<<<(unsigned)((B * A + 127) / 128), 128, 0, as_stream(stream)>>>(
      rewards, terminated, mask, target_qs, ret, B, Tq, Tr, A, RA, max_t, lambda_gamma, bootstrap_gamma);
  MRL_LAUNCHED();
  return MRL_OK;
}

// ---- host arrays in, host arrays out: chunked along the batch axis and pipelined ------------------------
// The scans are independent per column, so a [T,B] / [B,T] problem that lives in host memory is cut into
// K pieces of the batch axis: while piece c is scanned, piece c+1 comes in and piece c-1 goes out -- both
// PCIe directions busy at once instead of copy-in, scan, copy-out in a row (C4: 142 MB of PCIe traffic, the
// scan itself is 1 % of the call).  Pieces are dense on the device (2-D copies re-pitch them), so every kernel
// of this file applies unchanged.  Pinned host memory overlaps; pageable memory degrades to staged copies.
struct HostScan {
  const float *a, *v, *nv, *last;
  const uint8_t *done;
  float *out, *ret;
  int64_t T, B, E;
  int32_t layout;
};

template <typename Launch>
static int scan_host_pipelined(const HostScan &H, cudaStream_t st, Launch launch) {
  const int64_t T = H.T, B = H.B, E = H.E;
  const size_t bytes_in = (size_t)T * B * E * 4 * (1 + (H.v ? 1 : 0) + (H.nv ? 1 : 0));
  // measured at 4096 x 2048 GAE (scratch/time_host_scan.py): 1 piece 2.79 ms, 2: 2.26, 4: 1.96, 8: 2.13,
  // 16: 2.42 -- every piece is 5-7 DMA submissions, so pieces of ~16 MB of input
  int64_t K = g_host_chunks > 0 ? g_host_chunks : (int64_t)(bytes_in / (16u << 20));
  K = K < 1 ? 1 : (K > 16 ? 16 : K);
  int64_t Bc = ((B + K - 1) / K + 31) / 32 * 32;  // columns (rows) per piece
  K = (B + Bc - 1) / Bc;
  const size_t f = up256((size_t)T * Bc * E * 4), fd = up256((size_t)T * Bc), fl = up256((size_t)Bc * E * 4);
  const size_t o_a = 0, o_v = f, o_nv = 2 * f, o_out = 3 * f, o_ret = 4 * f, o_d = 5 * f, o_l = o_d + fd;
  const size_t piece = o_l + fl;
  // one host scan at a time per process: the copy streams and events below are shared (the call is
  // synchronous anyway)
  static std::mutex host_mu;
  std::lock_guard<std::mutex> host_guard(host_mu);
  uint8_t *s;
  int rc = scratch(piece * (size_t)K, &s, st);
  if (rc) return rc;
  // copy streams and events of the CURRENT device (streams and events belong to the device they were created on)
  struct HostPipe {
    cudaStream_t s_in = nullptr, s_out = nullptr;
    cudaEvent_t ev_in[16], ev_k[16], ev_start = nullptr, ev_done = nullptr;
  };
  static HostPipe pipes[64];
  int dev = 0;
  MRL_CUDA(cudaGetDevice(&dev));
  MRL_REQUIRE(dev >= 0 && dev < 64, "device index out of range");
  HostPipe &P = pipes[dev];
  cudaStream_t &s_in = P.s_in, &s_out = P.s_out;
  cudaEvent_t(&ev_in)[16] = P.ev_in, (&ev_k)[16] = P.ev_k;
  cudaEvent_t &ev_start = P.ev_start, &ev_done = P.ev_done;
  if (!s_in) {
    MRL_CUDA(cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking));
    MRL_CUDA(cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking));
    for (int i = 0; i < 16; ++i) {
      MRL_CUDA(cudaEventCreateWithFlags(&ev_in[i], cudaEventDisableTiming));
      MRL_CUDA(cudaEventCreateWithFlags(&ev_k[i], cudaEventDisableTiming));
    }
    MRL_CUDA(cudaEventCreateWithFlags(&ev_start, cudaEventDisableTiming));
    MRL_CUDA(cudaEventCreateWithFlags(&ev_done, cudaEventDisableTiming));
  }
  const bool tm = H.layout == MRL_TIME_MAJOR;
  // copy one array of a piece: time-major pieces are column blocks (re-pitched), batch-major pieces are rows
  auto copy = [&](void *dst, const void *src, int64_t c0, int64_t bc, size_t esz, bool to_dev, cudaStream_t cs) {
    const cudaMemcpyKind kind = to_dev ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost;
    if (tm) {
      const size_t wide = (size_t)B * esz, narrow = (size_t)bc * esz;
      if (to_dev) return cudaMemcpy2DAsync(dst, narrow, (const uint8_t *)src + (size_t)c0 * esz, wide, narrow, (size_t)T, kind, cs);
      return cudaMemcpy2DAsync((uint8_t *)dst + (size_t)c0 * esz, wide, src, narrow, narrow, (size_t)T, kind, cs);
    }
    const size_t bytes = (size_t)bc * T * esz, off = (size_t)c0 * T * esz;
    if (to_dev) return cudaMemcpyAsync(dst, (const uint8_t *)src + off, bytes, kind, cs);
    return cudaMemcpyAsync((uint8_t *)dst + off, src, bytes, kind, cs);
  };
  MRL_CUDA(cudaEventRecord(ev_start, st));
  MRL_CUDA(cudaStreamWaitEvent(s_in, ev_start, 0));
  MRL_CUDA(cudaStreamWaitEvent(s_out, ev_start, 0));
  for (int64_t c = 0; c < K; ++c) {
    const int64_t c0 = c * Bc, bc = (c0 + Bc <= B) ? Bc : B - c0;
    uint8_t *p = s + piece * (size_t)c;
    MRL_CUDA(copy(p + o_a, H.a, c0, bc, 4 * (size_t)E, true, s_in));
    if (H.v) MRL_CUDA(copy(p + o_v, H.v, c0, bc, 4 * (size_t)E, true, s_in));
    if (H.nv) MRL_CUDA(copy(p + o_nv, H.nv, c0, bc, 4 * (size_t)E, true, s_in));
    if (H.done) MRL_CUDA(copy(p + o_d, H.done, c0, bc, 1, true, s_in));
    if (H.last)
      MRL_CUDA(cudaMemcpyAsync(p + o_l, H.last + c0 * E, (size_t)bc * E * 4, cudaMemcpyHostToDevice, s_in));
    MRL_CUDA(cudaEventRecord(ev_in[c], s_in));
    MRL_CUDA(cudaStreamWaitEvent(st, ev_in[c], 0));
    HostScan D = H;
    D.a = (const float *)(p + o_a), D.v = H.v ? (const float *)(p + o_v) : nullptr;
    D.nv = H.nv ? (const float *)(p + o_nv) : nullptr, D.last = H.last ? (const float *)(p + o_l) : nullptr;
    D.done = H.done ? p + o_d : nullptr;
    D.out = (float *)(p + o_out), D.ret = H.ret ? (float *)(p + o_ret) : nullptr;
    D.B = bc;
    rc = launch(D);
    if (rc) return rc;
    MRL_CUDA(cudaEventRecord(ev_k[c], st));
    MRL_CUDA(cudaStreamWaitEvent(s_out, ev_k[c], 0));
    MRL_CUDA(copy(H.out, p + o_out, c0, bc, 4 * (size_t)E, false, s_out));
    if (H.ret) MRL_CUDA(copy(H.ret, p + o_ret, c0, bc, 4 * (size_t)E, false, s_out));
  }
  MRL_CUDA(cudaEventRecord(ev_done, s_out));
  MRL_CUDA(cudaStreamWaitEvent(st, ev_done, 0));
  MRL_CUDA(cudaStreamSynchronize(st));
  return MRL_OK;
}

extern "C" int mrl_discounted_return_host(const float *reward, const uint8_t *done,
                                          const float *last_value, float *out, int64_t T,
                                          int64_t B, int64_t E, float gamma, int32_t layout,
                                          int32_t mode, void *stream) {
  MRL_REQUIRE(T >= 0 && B >= 0 && E >= 1, "mrl_discounted_return_host: bad shape");
  MRL_REQUIRE(layout == MRL_TIME_MAJOR || layout == MRL_BATCH_MAJOR, "mrl_discounted_return_host: bad layout");
  if (T == 0 || B == 0) return MRL_OK;
  MRL_REQUIRE(reward && out, "mrl_discounted_return_host: NULL reward/out");
  HostScan H{reward, nullptr, nullptr, last_value, done, out, nullptr, T, B, E, layout};
  return scan_host_pipelined(H, as_stream(stream), [&](const HostScan &D) {
    return mrl_discounted_return(D.a, D.done, D.last, D.out, D.T, D.B, D.E, gamma, layout, mode, stream);
  });
}

extern "C" int mrl_gae_host(const float *reward, const float *value, const float *next_value,
                            const float *last_value, const uint8_t *done, float *adv, float *ret,
                            int64_t T, int64_t B, float gamma, float lambda, int32_t layout,
                            int32_t mode, void *stream) {
  MRL_REQUIRE(T >= 0 && B >= 0, "mrl_gae_host: bad shape");
  MRL_REQUIRE(layout == MRL_TIME_MAJOR || layout == MRL_BATCH_MAJOR, "mrl_gae_host: bad layout");
  if (T == 0 || B == 0) return MRL_OK;
  MRL_REQUIRE(reward && value && adv, "mrl_gae_host: NULL reward/value/adv");
  HostScan H{reward, value, next_value, last_value, done, adv, ret, T, B, 1, layout};
  return scan_host_pipelined(H, as_stream(stream), [&](const HostScan &D) {
    return mrl_gae(D.a, D.v, D.nv, D.last, D.done, D.out, D.ret, D.T, D.B, gamma, lambda, layout, mode, stream);
  });
}
