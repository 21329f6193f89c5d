// scan_tma.cu -- reverse scans as persistent CTAs fed by the TMA engine (sm_100a).
//
// Same recurrences as scan.cu (x[t] = a[t] + b[t]*x[t+1]; DiscountedReturn discounted_return.py:84-92,
// PPO ppo.py:320-350, MAPPO mappo.py:478-497, V-trace vtrace.py:88-93), different machine mapping:
//
//   * one elected thread of a producer warp keeps a ring of shared-memory stages full with
//     cp.async.bulk[.tensor] (mbarrier complete_tx), so the bytes in flight per SM do not depend on
//     how many registers or warps the arithmetic needs and no thread computes a global address;
//   * 8 consumer warps take a stage out of shared memory (conflict-free), build the affine map of
//     their 8-step (time-major) / 128-step (batch-major) piece, exchange the 8 / 16 maps of the tile
//     through shared memory (ONE named barrier per tile), chain them from the CTA's running carry and
//     replay their piece;
//   * results are staged in shared memory and leave through cp.async.bulk stores (bulk groups), one
//     per warp, so a warp never waits for another warp's output.
//
// A CTA walks its strip (time-major: 32 or 16 columns, all T) or its environments (batch-major: whole
// rows) from the last step to the first, so the carry never leaves the CTA: no look-back, no flags,
// no workspace.  Parallel kernels re-associate fp32 (spec: 1e-5 relative).
#include <cuda.h>

#include "scan.cuh"

namespace mrl {
namespace {

// ---- PTX helpers ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(s32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int x, int y, uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::
          "r"(s32(dst)),
      "l"(map), "r"(x), "r"(y), "r"(s32(bar))
      : "memory");
}
__device__ __forceinline__ void prefetch_tensormap(const CUtensorMap *map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap *map, int x, int y, const void *src) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(map),
               "r"(x), "r"(y), "r"(s32(src))
               : "memory");
}
__device__ __forceinline__ void bulk_load(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
                   "r"(s32(dst)),
               "l"(src), "r"(bytes), "r"(s32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_store(void *dst, const void *src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(s32(src)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// A consumer warp has read its part of a stage with ordinary (generic-proxy) loads and hands the stage
// back to the producer, whose refill is an async-proxy write: the proxy fence orders the two (without it
// the refill was observed to land under loads still in flight: bit-flips of whole tiles, timing dependent).
__device__ __forceinline__ void release_stage(uint64_t *empty_bar, int lane) {
  fence_async_smem();
  __syncwarp();
  if (lane == 0) mbar_arrive(empty_bar);
}
template <int ID = 1>
__device__ __forceinline__ void consumer_barrier() { asm volatile("bar.sync %0, 256;" ::"n"(ID) : "memory"); }

constexpr int kWarps = 8;  // consumer warps; warp kWarps is the producer
constexpr int kThreads = (kWarps + 1) * 32;

// =============================================================================================
// time-major strips
// =============================================================================================
// a unit = COLS adjacent lanes of a warp x RPU consecutive steps

// The parallel kernels re-associate fp32 anyway (spec: 1e-5 relative), so the maps are composed and
// replayed with fused multiply-adds: half the dependent latency and one instruction per step.  The
// per-element (a, b) keep the reference's operation order.
__device__ __forceinline__ float fstep(float a, float b, float x) { return __fmaf_rn(b, x, a); }

constexpr int kMaxStages = 12;

template <int OP, bool HD, int COLS, int RPU, int NG>
struct StripCfg {
  static constexpr int kThreads = (NG * kWarps + 1) * 32;  // NG consumer groups of 8 warps + the producer warp
  static constexpr int kUPW = 32 / COLS;                  // units per warp
  static constexpr int kUnits = kWarps * kUPW;            // units per tile
  static constexpr int kRows = kUnits * RPU;              // steps per tile (64 .. 256)
  static constexpr int kVRows = OP == OP_GAE ? kRows + 1 : (OP == OP_AFFINE ? kRows : 0);
  static constexpr int kOffV = kRows * COLS * 4;
  // done bytes per row in shared memory.  A TMA box starts on a 16-byte boundary and is a multiple of 16 bytes wide:
  // a 28-column strip's flags start at byte 28k, so it loads the 48 bytes from the boundary below (an unaligned start
  // is an illegal instruction, found the hard way) and reads them at an offset of 0, 4, 8 or 12
  static constexpr int kDW = COLS % 16 == 0 ? COLS : (COLS + 30) / 16 * 16;
  static constexpr int kOffD = kOffV + (kVRows * COLS * 4 + 127) / 128 * 128;
  static constexpr int kStage = kOffD + (HD ? kRows * kDW : 0);
  static constexpr int kTx = kRows * COLS * 4 + kVRows * COLS * 4 + (HD ? kRows * kDW : 0);
  static constexpr int kNOut = OP == OP_GAE ? 2 : 1;
  static constexpr int kWarpRows = RPU * kUPW;            // rows one warp produces per tile
  static constexpr int kWarpOut = kWarpRows * COLS;       // floats per output array
  static constexpr int kCompBytes = NG * 2 * kUnits * COLS * 8 + 2 * COLS * 4;  // maps (per group, double-buffered) + carries
  static constexpr int kFixed = kCompBytes + 2 * 8 * 8;
  static constexpr int kOut1 = NG * kWarps * kNOut * kWarpOut * 4;  // one output staging buffer
  static constexpr int kBarBytes = (2 * kMaxStages + 2) * 8;
  // ring depth and output double-buffering are chosen at launch from the shared memory a CTA may take
  // (all of the SM when the strips fit one per SM, half otherwise)
  static constexpr int smem_bytes(int stages, int obufs) {
    return stages * kStage + obufs * kOut1 + kCompBytes + kBarBytes;
  }
};

// NG == 2: two consumer groups take alternate tiles of the strip, so that the loads and the map building
// of tile k+1 run while tile k is chained, replayed and stored (a lone group of 8 warps is latency-bound
// on its own dependent chain).  The running carry passes between the groups through shared memory and
// one mbarrier per tile parity.
// MOM: the kernel also accumulates the moments of `out` (ScanArgs.mom).  A template parameter, not a runtime
// flag: the predicate alone in the replay loop cost the moment-less kernels 3 us of 25 (measured, round 2).
template <int OP, bool HD, int COLS, int RPU, int NG, bool MOM>
__global__ void __launch_bounds__((NG * kWarps + 1) * 32)
    scan_tm_strip_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_v,
                         const __grid_constant__ CUtensorMap tm_v1, const __grid_constant__ CUtensorMap tm_d,
                         const __grid_constant__ CUtensorMap tm_out, const __grid_constant__ CUtensorMap tm_ret,
                         const __grid_constant__ ScanArgs A, int n_tiles, int n_stages, int n_obufs) {
  using C = StripCfg<OP, HD, COLS, RPU, NG>;
  extern __shared__ __align__(1024) unsigned char strip_raw[];  // (no integer casts: keeps the shared address space)
  unsigned char *stages = strip_raw;
  float *ostage = reinterpret_cast<float *>(stages + n_stages * C::kStage);
  float2 *comp = reinterpret_cast<float2 *>(reinterpret_cast<unsigned char *>(ostage) + n_obufs * C::kOut1);
  uint64_t *full = reinterpret_cast<uint64_t *>(reinterpret_cast<unsigned char *>(comp) + C::kCompBytes);
  uint64_t *empty = full + kMaxStages;
  uint64_t *cbar = empty + kMaxStages;  // [2]: carry into the tiles of each parity is published
  float *carry = reinterpret_cast<float *>(comp + NG * 2 * C::kUnits * COLS);  // [2][COLS]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int col0 = blockIdx.x * COLS;
  const int T = (int)A.T;
  if (MOM) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");  // the apply pass may become resident

  if (threadIdx.x == 0) {
    for (int s = 0; s < n_stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], kWarps);
    }
    mbar_init(&cbar[0], 1);
    mbar_init(&cbar[1], 1);
    fence_async_smem();
  } else if (threadIdx.x == NG * kWarps * 32) {
    // the descriptors live in the kernel's parameter block: have them in the TMA unit's cache before the first
    // load is issued (their fetch is otherwise the first thing every CTA's pipeline waits for)
    prefetch_tensormap(&tm_a);
    if (OP != OP_DR) prefetch_tensormap(&tm_v);
    if (OP == OP_GAE) prefetch_tensormap(&tm_v1);
    if (HD) prefetch_tensormap(&tm_d);
  } else if (threadIdx.x == 32) {
    prefetch_tensormap(&tm_out);
    if (OP == OP_GAE) prefetch_tensormap(&tm_ret);
  }
  __syncthreads();

  if (warp == NG * kWarps) {  // ---- producer ----
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 1;  // parity of the previous use of the stage
#pragma unroll 1
      for (int k = 0; k < n_tiles; ++k) {
        if (k >= n_stages) mbar_wait(&empty[s], ph);
        unsigned char *sb = stages + s * C::kStage;
        const int row0 = (n_tiles - 1 - k) * C::kRows;  // the first tile may reach past T: zero-filled
        mbar_expect_tx(&full[s], C::kTx);
        tma_load_2d(sb, &tm_a, col0, row0, &full[s]);
        if (OP != OP_DR) tma_load_2d(sb + C::kOffV, &tm_v, col0, row0, &full[s]);
        // V of the first step after the tile (a box may have at most 256 rows, so it travels alone)
        if (OP == OP_GAE) tma_load_2d(sb + C::kOffV + C::kRows * COLS * 4, &tm_v1, col0, row0 + C::kRows, &full[s]);
        if (HD) tma_load_2d(sb + C::kOffD, &tm_d, col0 & ~15, row0, &full[s]);
        if (++s == n_stages) s = 0, ph ^= 1u;
      }
    }
    return;
  }

  // ---- consumers ----
  // (COLS = 28: lanes 28..31 shadow lanes 0..3 -- same loads, same values to the same shared-memory words)
  const int c = lane % COLS, h = (lane / COLS) % C::kUPW;
  const int grp = warp / kWarps, wl = warp % kWarps;
  const int unit = wl * C::kUPW + h;
  const int col = col0 + c;
  float X = 0.0f, lastv = 0.0f;
  if (col < A.N && A.last) {
    const float l = __ldg(A.last + col);
    if (OP == OP_GAE) lastv = l;
    else X = l;
  }
  const bool want_ret = OP == OP_GAE && A.ret != nullptr;
  const bool colok = col < A.N && lane < C::kUPW * COLS;
  double msum = 0.0, msq = 0.0;  // (MOM) moments of `out` for the normalisation that follows
  float *my_out = ostage + warp * (n_obufs * C::kNOut * C::kWarpOut);
  int s = grp;  // (n_stages >= 2 >= NG)
  uint32_t ph = 0;

#pragma unroll 1
  for (int k = grp; k < n_tiles; k += NG) {
    mbar_wait(&full[s], ph);
    const unsigned char *sb = stages + s * C::kStage;
    const int e0 = unit * RPU * COLS + c;  // first element of this lane's column piece
    float ea[RPU], eb[RPU], aux[RPU];
    {
      float r[RPU], vv[RPU + 1];
      unsigned char dd[RPU];
      const float *sa = reinterpret_cast<const float *>(sb) + e0;
#pragma unroll
      for (int i = 0; i < RPU; ++i) r[i] = sa[i * COLS];
      if (OP != OP_DR) {
        const float *sv = reinterpret_cast<const float *>(sb + C::kOffV) + e0;
#pragma unroll
        for (int i = 0; i < RPU + (OP == OP_GAE ? 1 : 0); ++i) vv[i] = sv[i * COLS];
      }
      if (HD) {
        const unsigned char *sd = sb + C::kOffD + unit * RPU * C::kDW + (col0 & 15) + c;
#pragma unroll
        for (int i = 0; i < RPU; ++i) dd[i] = sd[i * C::kDW];
      }
      release_stage(&empty[s], lane);
      const int g0 = (n_tiles - 1 - k) * C::kRows + unit * RPU;  // global step of this lane's first row
      if (OP == OP_GAE && k == 0) {  // row T (zero-filled by the DMA) is the bootstrap value
#pragma unroll
        for (int i = 0; i < RPU; ++i)
          if (g0 + i + 1 == T) vv[i + 1] = lastv;
      }
#pragma unroll
      for (int i = 0; i < RPU; ++i) {
        if (OP == OP_AFFINE) {
          ea[i] = r[i], eb[i] = vv[i], aux[i] = 0.0f;
        } else {
          const float m = HD ? (dd[i] ? 0.0f : 1.0f) : 1.0f;
          if (OP == OP_DR) {
            ea[i] = r[i], eb[i] = __fmul_rn(m, A.gamma), aux[i] = 0.0f;
          } else {
            const float gvm = __fmul_rn(__fmul_rn(A.gamma, vv[i + 1]), m);
            ea[i] = __fsub_rn(__fadd_rn(r[i], gvm), vv[i]);
            eb[i] = __fmul_rn(A.gl, m);
            aux[i] = vv[i];
          }
        }
      }
      if (k == 0) {  // steps past T in the first tile: identity maps
#pragma unroll
        for (int i = 0; i < RPU; ++i)
          if (g0 + i >= T) ea[i] = 0.0f, eb[i] = 1.0f;
      }
    }
    // this piece as one affine map x[first step] = ca + cb * x[first step after the piece]
    float ca = 0.0f, cb = 1.0f;
#pragma unroll
    for (int i = RPU - 1; i >= 0; --i) {
      ca = fstep(ea[i], eb[i], ca);
      cb = __fmul_rn(eb[i], cb);
    }
    float2 *cbuf = comp + (grp * 2 + ((k / NG) & 1)) * (C::kUnits * COLS);
    cbuf[unit * COLS + c] = make_float2(ca, cb);
    if (NG == 1 || grp == 0) consumer_barrier<1>();
    else consumer_barrier<2>();
    if (NG == 2 && k > 0) {  // the other group has chained tile k-1
      mbar_wait(&cbar[k & 1], (uint32_t)(((k >> 1) - (1 - (k & 1))) & 1));
      X = carry[(k & 1) * COLS + c];
    }
    float x = X, xin = X;
#pragma unroll
    for (int u = C::kUnits - 1; u >= 0; --u) {
      if (u == unit) xin = x;
      const float2 m = cbuf[u * COLS + c];
      x = fstep(m.x, m.y, x);
    }
    X = x;
    if (NG == 2 && wl == 0 && k + 1 < n_tiles) {
      if (h == 0) carry[((k + 1) & 1) * COLS + c] = x;
      __syncwarp();
      if (lane == 0) mbar_arrive(&cbar[(k + 1) & 1]);
    }

    // replay, stage, store
    float *oa = my_out + (n_obufs == 2 ? ((k / NG) & 1) : 0) * (C::kNOut * C::kWarpOut);
    if (lane == 0) {  // the last store issued from this buffer has been read
      if (n_obufs == 2) bulk_wait_read<1>();
      else bulk_wait_read<0>();
    }
    __syncwarp();
    x = xin;
    float tsum = 0.0f, tsq = 0.0f;
#pragma unroll
    for (int i = RPU - 1; i >= 0; --i) {
      x = fstep(ea[i], eb[i], x);
      oa[(h * RPU + i) * COLS + c] = x;
      if (want_ret) oa[C::kWarpOut + (h * RPU + i) * COLS + c] = __fadd_rn(x, aux[i]);
      if (MOM && colok && (k > 0 || (n_tiles - 1) * C::kRows + unit * RPU + i < T)) {
        tsum = __fadd_rn(tsum, x);  // fp32 over this lane's RPU steps of the tile, float64 across tiles:
        tsq = __fmaf_rn(x, x, tsq);  // per-element float64 cost the kernel 4 us of 24
      }
    }
    if (MOM) {
      msum += (double)tsum;
      msq += (double)tsq;
    }
    fence_async_smem();
    __syncwarp();
    if (lane == 0) {
      const int row = (n_tiles - 1 - k) * C::kRows + wl * C::kWarpRows;
      if (row < T) {  // rows past T are clipped by the tensor map
        tma_store_2d(&tm_out, col0, row, oa);
        if (want_ret) tma_store_2d(&tm_ret, col0, row, oa + C::kWarpOut);
      }
      bulk_commit();
    }
    s += NG;
    if (s >= n_stages) s -= n_stages, ph ^= 1u;
  }
  if (MOM && NG == 1) {  // fixed-order fold: lanes, then the 8 warps in order -> one partial per CTA
#pragma unroll
    for (int off = 16; off; off >>= 1) {
      msum += __shfl_xor_sync(0xffffffffu, msum, off);
      msq += __shfl_xor_sync(0xffffffffu, msq, off);
    }
    double2 *red = reinterpret_cast<double2 *>(comp);
    consumer_barrier<1>();  // every warp has read its last maps
    if (lane == 0) red[warp] = make_double2(msum, msq);
    consumer_barrier<1>();
    if (warp == 0 && lane == 0) {
      double a = 0.0, b = 0.0;
      for (int w = 0; w < kWarps; ++w) a += red[w].x, b += red[w].y;
      A.mom[blockIdx.x] = make_double2(a, b);
    }
  }
  if (lane == 0) bulk_wait_all();
}

}  // namespace
int64_t g_strip_stages = 0;  // 0 = as many as fit; else an upper bound (for experiments)
int64_t g_strip_groups = 0;  // consumer groups per CTA: 0 = auto, 1, 2
int64_t g_row_ctas_per_sm = 4;  // batch-major rows: resident CTAs per SM the ring depth is sized for
int64_t g_row_stages = 0;       // 0 = from g_row_ctas_per_sm
namespace {

// ---- host: tensor maps ------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
    else
      cudaGetLastError();
  }
  return fn;
}

// A tensor map is a pure function of (base, element size, rows, cols, box): learners call the scans on the same
// trajectory buffers every iteration, and six cuTensorMapEncodeTiled calls are ~2 us of host time in front of
// a 14 us kernel (they sit between the caller's event and the launch when the GPU is idle).  Small direct-mapped
// cache, per thread (no locking; a stale entry can only be hit by identical arguments, which encode identically).
struct MapKey {
  const void *base;
  int64_t rows, cols;
  int32_t box_cols, box_rows, is_u8;
  bool operator==(const MapKey &o) const {
    return base == o.base && rows == o.rows && cols == o.cols && box_cols == o.box_cols && box_rows == o.box_rows &&
           is_u8 == o.is_u8;
  }
};
struct MapSlot {
  MapKey key;
  CUtensorMap map;
  bool valid = false;
};
constexpr int kMapSlots = 64;

bool make_map_uncached(CUtensorMap *m, const void *base, bool is_u8, int64_t rows, int64_t cols, int box_cols,
                       int box_rows);

// row-major [rows, cols] array of `esize`-byte elements, box = box_rows x box_cols
bool make_map(CUtensorMap *m, const void *base, bool is_u8, int64_t rows, int64_t cols, int box_cols,
              int box_rows) {
  static thread_local MapSlot cache[kMapSlots];
  const MapKey key{base, rows, cols, box_cols, box_rows, is_u8 ? 1 : 0};
  uint64_t h = reinterpret_cast<uintptr_t>(base) >> 4;
  h ^= (uint64_t)rows * 0x9E3779B97F4A7C15ull ^ (uint64_t)cols * 0xC2B2AE3D27D4EB4Full ^
       ((uint64_t)box_rows << 20) ^ ((uint64_t)box_cols << 8) ^ (uint64_t)key.is_u8;
  h ^= h >> 29;
  MapSlot &sl = cache[h % kMapSlots];
  if (sl.valid && sl.key == key) {
    *m = sl.map;
    return true;
  }
  if (!make_map_uncached(m, base, is_u8, rows, cols, box_cols, box_rows)) return false;
  sl.key = key, sl.map = *m, sl.valid = true;
  return true;
}

bool make_map_uncached(CUtensorMap *m, const void *base, bool is_u8, int64_t rows, int64_t cols, int box_cols,
                       int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  const int esize = is_u8 ? 1 : 4;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * esize};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  const CUresult r = fn(m, is_u8 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                        const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

bool al16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

template <int OP, bool HD, int COLS, int RPU, int NG, bool MOM = false>
int strip_launch(ScanArgs A, cudaStream_t st, int *mom_parts) {
  using C = StripCfg<OP, HD, COLS, RPU, NG>;
  static_assert(C::kRows <= 256 && C::smem_bytes(2, 1) <= (NG == 1 && RPU <= 16 ? 111 : 225) * 1024, "tile too large");
  CUtensorMap ma, mv, mv1, md, mo, mr;
  if (!make_map(&ma, A.a, false, A.T, A.N, COLS, C::kRows)) return kScanNotApplicable;
  mv = ma, mv1 = ma, md = ma;
  if (OP != OP_DR && !make_map(&mv, A.v, false, A.T, A.N, COLS, C::kRows)) return kScanNotApplicable;
  if (OP == OP_GAE && !make_map(&mv1, A.v, false, A.T, A.N, COLS, 1)) return kScanNotApplicable;
  if (HD && !make_map(&md, A.done, true, A.T, A.N, C::kDW, C::kRows)) return kScanNotApplicable;
  if (!make_map(&mo, A.out, false, A.T, A.N, COLS, C::kWarpRows)) return kScanNotApplicable;
  mr = mo;
  if (OP == OP_GAE && A.ret && !make_map(&mr, A.ret, false, A.T, A.N, COLS, C::kWarpRows))
    return kScanNotApplicable;
  const int64_t n_strips = (A.N + COLS - 1) / COLS;
  if (MOM && n_strips > kMomentSlots) return strip_launch<OP, HD, COLS, RPU, NG, false>(A, st, mom_parts);
  if (!MOM) A.mom = nullptr;
  if (first_use_on_device((const void *)scan_tm_strip_kernel<OP, HD, COLS, RPU, NG, MOM>))
    MRL_CUDA(cudaFuncSetAttribute(scan_tm_strip_kernel<OP, HD, COLS, RPU, NG, MOM>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
  const int n_tiles = (int)((A.T + C::kRows - 1) / C::kRows);
  if (RPU > 16 && n_strips > sm_count()) return kScanNotApplicable;  // 256-step tiles need the whole SM
  const int budget = ((NG == 2 || n_strips <= sm_count()) ? 226 : 112) * 1024 - 1024;
  int obufs = 2, stages = (budget - C::smem_bytes(0, 2)) / C::kStage;
  if (stages < 3) obufs = 1, stages = (budget - C::smem_bytes(0, 1)) / C::kStage;
  stages = stages > kMaxStages ? kMaxStages : stages;
  stages = stages > n_tiles ? (n_tiles < 2 ? 2 : n_tiles) : stages;
  if (g_strip_stages > 0 && g_strip_stages < stages) stages = (int)(g_strip_stages < 2 ? 2 : g_strip_stages);
  scan_tm_strip_kernel<OP, HD, COLS, RPU, NG, MOM><<<(unsigned)n_strips, C::kThreads, C::smem_bytes(stages, obufs), st>>>(
      ma, mv, mv1, md, mo, mr, A, n_tiles, stages, obufs);
  if (mom_parts) *mom_parts = A.mom ? (int)n_strips : 0;
  return MRL_OK;
}

template <int OP, bool HD>
int strip_launch_cols(const ScanArgs &A, int cols, int rpu, cudaStream_t st, int *mp) {
  // moments (ScanArgs.mom) are produced by the default shapes of the two scans that get normalised afterwards
  // (GAE: ppo.py:361-368; DiscountedReturn: a2c.py:190-192); any other configuration leaves them to a separate pass
  if (A.mom && OP != OP_AFFINE && cols == 32 && g_strip_groups != 2) {
    if (rpu == 32) {
      const int rc = strip_launch<OP, HD, 32, 32, 1, OP != OP_AFFINE>(A, st, mp);
      if (rc != kScanNotApplicable) return rc;
      rpu = 16;
    }
    if (rpu == 16) return strip_launch<OP, HD, 32, 16, 1, OP != OP_AFFINE>(A, st, mp);
  }
  // two consumer groups per CTA: opt-in (measured equal to one group at 4096 x 2048: once the ring is
  // deep enough the strips are bound by HBM, not by the consumers' dependent chain)
  const bool two = g_strip_groups == 2;
  if (cols == 32 && rpu == 16 && two) return strip_launch<OP, HD, 32, 16, 2>(A, st, mp);
  if (cols == 32 && rpu == 32) {
    const int rc = strip_launch<OP, HD, 32, 32, 1>(A, st, mp);
    if (rc != kScanNotApplicable) return rc;
    rpu = 16;
  }
  if (cols == 16) return rpu == 16 ? strip_launch<OP, HD, 16, 16, 1>(A, st, mp) : strip_launch<OP, HD, 16, 8, 1>(A, st, mp);
  return rpu == 16 ? strip_launch<OP, HD, 32, 16, 1>(A, st, mp) : strip_launch<OP, HD, 32, 8, 1>(A, st, mp);
}

// 28-column strips: 4096 columns make 147 strips -- one per SM on all 148 -- where 32-column strips make 128
template <int OP, bool HD>
int strip_launch_28(const ScanArgs &A, int rpu, cudaStream_t st, int *mp) {
  constexpr bool kMomOk = OP != OP_AFFINE;
  if (rpu == 32) {
    const int rc = (A.mom && kMomOk) ? strip_launch<OP, HD, 28, 32, 1, kMomOk>(A, st, mp) : strip_launch<OP, HD, 28, 32, 1>(A, st, mp);
    if (rc != kScanNotApplicable) return rc;
  }
  return (A.mom && kMomOk) ? strip_launch<OP, HD, 28, 16, 1, kMomOk>(A, st, mp) : strip_launch<OP, HD, 28, 16, 1>(A, st, mp);
}

}  // namespace

int64_t g_strip_cols = 0;     // 0 = auto, 16, 32
int64_t g_strip_rpu = 0;      // steps per lane and tile: 0 = auto, 8, 16, 32
int64_t g_strip_min_cols = 2048;  // auto: narrower batches keep the look-back kernels (time-parallel)

// Time-major [T, N], E == 1.  Needs 16-byte aligned bases and row pitches (N % 4 == 0; N % 16 == 0 with
// done flags): the tensor maps address rows by pitch.
int scan_tm_strip_launch(const ScanArgs &A, int op, bool force, cudaStream_t st, int *mom_parts) {
  if (mom_parts) *mom_parts = 0;
  if (A.E != 1 || (op == OP_GAE && A.nv)) return kScanNotApplicable;
  if (A.T >= (1LL << 31) - 256 || A.N >= (1LL << 31) - 64 || A.N % 4 != 0) return kScanNotApplicable;
  if (!al16(A.a) || !al16(A.out) || (op != OP_DR && !al16(A.v)) || (A.ret && !al16(A.ret)))
    return kScanNotApplicable;
  const bool hd = op != OP_AFFINE && A.done != nullptr;
  if (hd && (A.N % 16 != 0 || !al16(A.done))) return kScanNotApplicable;
  if (!force && (A.N < g_strip_min_cols || A.T < 64)) return kScanNotApplicable;
  int cols = (int)g_strip_cols, rpu = (int)g_strip_rpu;
  if (cols == 0) cols = 32;
  // 256-step tiles (32 steps per lane) halve the per-tile costs that are not overlapped with the data
  // movement (barrier, map chain, store hand-off): measured 19.4 -> 16.4 us for DiscountedReturn and
  // 22.6 -> 20.6 us for GAE without returns at 4096 x 2048; GAE with both outputs is HBM-bound either
  // way and keeps the deeper ring of the 128-step tiles.  (Falls back to 16 when the strips do not get an SM each.)
  if (rpu == 0) rpu = (op == OP_GAE && A.ret) ? 16 : 32;
  int *mp = mom_parts;
  if (cols == 28) {
    if (op == OP_DR) return hd ? strip_launch_28<OP_DR, true>(A, rpu, st, mp) : strip_launch_28<OP_DR, false>(A, rpu, st, mp);
    if (op == OP_GAE) return hd ? strip_launch_28<OP_GAE, true>(A, rpu, st, mp) : strip_launch_28<OP_GAE, false>(A, rpu, st, mp);
    return strip_launch_28<OP_AFFINE, false>(A, rpu, st, mp);
  }
  if (op == OP_DR)
    return hd ? strip_launch_cols<OP_DR, true>(A, cols, rpu, st, mp) : strip_launch_cols<OP_DR, false>(A, cols, rpu, st, mp);
  if (op == OP_GAE)
    return hd ? strip_launch_cols<OP_GAE, true>(A, cols, rpu, st, mp) : strip_launch_cols<OP_GAE, false>(A, cols, rpu, st, mp);
  return strip_launch_cols<OP_AFFINE, false>(A, cols, rpu, st, mp);
}

// =============================================================================================
// batch-major rows
// =============================================================================================
// A tile is one environment's chunk of up to BM_L steps (8 KB per float array, one bulk copy each).
// Warp w owns positions [256w, 256w+256) as two 128-step sub-tiles (lane l: steps 4l..4l+3 of each,
// float4 from shared memory); the 16 sub-tile maps of a tile are chained from the running carry of the
// environment.  A CTA takes environments blockIdx.x, blockIdx.x + gridDim.x, ... and walks the chunks
// of each from the last to the first.
namespace {

constexpr int BM_L = 2048;

// SPL = consecutive steps a lane owns in a tile: 8 (8 consumer warps, the default) or 16 (4 consumer warps: half the
// shuffle scans, barriers and chain links per element, 160-thread CTAs -- option "row_spl").  Measured at C4 (round 2,
// tools/rows_sweep.sh): 16 steps per lane is SLOWER, GAE 32.9 against 28.4 us and DiscountedReturn 20.7 against 19.5 us
// at any number of resident CTAs, and 4 steps per lane (16 warps) slower still (35 / 29 us): the rows are bound by the
// latency of their own dependent chain per tile, which fewer warps hide worse, not by instruction issue.
template <int OP, bool HD, bool NVF, int SPL>
struct RowCfg {
  static constexpr int kW = BM_L / (32 * SPL);  // consumer warps
  static constexpr int kThreads = (kW + 1) * 32;
  static constexpr int kOffV = BM_L * 4;
  static constexpr int kOffNV = kOffV + (OP != OP_DR ? BM_L * 4 + 128 : 0);  // +: sv[BM_L] stays in bounds
  static constexpr int kOffD = kOffNV + (NVF ? BM_L * 4 : 0);
  static constexpr int kStage = kOffD + (HD ? BM_L : 0);
  static constexpr int kNOut = OP == OP_GAE ? 2 : 1;
  static constexpr int kWarpOut = BM_L / kW;  // floats a warp produces per tile and output array
  static constexpr int kOutBytes = kW * 2 * kNOut * kWarpOut * 4;
  static constexpr int kCompBytes = 2 * 16 * 8 > 16 * 16 ? 2 * 16 * 8 : 16 * 16;  // maps [2][16], or 16 double2 partials
  static constexpr int smem_bytes(int stages) {
    return stages * kStage + kOutBytes + kCompBytes + 2 * kMaxStages * 8;
  }
};

template <int NT>
__device__ __forceinline__ void row_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory"); }

template <int OP, bool HD, bool NVF, bool MOM, int SPL>
__global__ void __launch_bounds__((BM_L / (32 * SPL) + 1) * 32)
    scan_bm_row_kernel(const __grid_constant__ ScanArgs A, int n_chunks, int n_stages) {
  using C = RowCfg<OP, HD, NVF, SPL>;
  constexpr int W = C::kW;
  extern __shared__ __align__(1024) unsigned char row_raw[];
  unsigned char *stages = row_raw;
  float *ostage = reinterpret_cast<float *>(stages + n_stages * C::kStage);
  float2 *comp = reinterpret_cast<float2 *>(reinterpret_cast<unsigned char *>(ostage) + C::kOutBytes);
  uint64_t *full = reinterpret_cast<uint64_t *>(reinterpret_cast<unsigned char *>(comp) + C::kCompBytes);
  uint64_t *empty = full + kMaxStages;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t T = A.T;
  const int64_t G = gridDim.x;
  if (MOM) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");  // the apply pass may become resident
  const int64_t n_rows = A.N > (int64_t)blockIdx.x ? (A.N - blockIdx.x + G - 1) / G : 0;
  const int64_t n_items = n_rows * n_chunks;

  if (threadIdx.x == 0) {
    for (int s = 0; s < n_stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], W);
    }
    fence_async_smem();
  }
  __syncthreads();

  if (warp == W) {  // ---- producer ----
    if (lane == 0) {
      int64_t row = blockIdx.x;
      int c = 0;
      int s = 0;
      uint32_t ph = 1;  // parity of the previous use of the stage
#pragma unroll 1
      for (int64_t k = 0; k < n_items; ++k) {
        if (k >= n_stages) mbar_wait(&empty[s], ph);
        unsigned char *sb = stages + s * C::kStage;
        const int64_t t_end = T - (int64_t)c * BM_L;
        const int64_t t_begin = t_end > BM_L ? t_end - BM_L : 0;
        const uint32_t len = (uint32_t)(t_end - t_begin);
        const int64_t off = row * T + t_begin;
        mbar_expect_tx(&full[s], len * (4u + (OP != OP_DR ? 4u : 0u) + (NVF ? 4u : 0u) + (HD ? 1u : 0u)));
        bulk_load(sb, A.a + off, len * 4u, &full[s]);
        if (OP != OP_DR) bulk_load(sb + C::kOffV, A.v + off, len * 4u, &full[s]);
        if (NVF) bulk_load(sb + C::kOffNV, A.nv + off, len * 4u, &full[s]);
        if (HD) bulk_load(sb + C::kOffD, A.done + off, len, &full[s]);
        if (++c == n_chunks) c = 0, row += G;
        if (++s == n_stages) s = 0, ph ^= 1u;
      }
    }
    return;
  }

  // ---- consumers ----
  // warp w owns positions [32*SPL*w, 32*SPL*(w+1)) of the tile, lane l the SPL consecutive steps from 32*SPL*w + SPL*l
  constexpr unsigned FULL = 0xffffffffu;
  const bool want_ret = OP == OP_GAE && A.ret != nullptr;
  double msum = 0.0, msq = 0.0;  // (MOM) moments of `out` for the normalisation that follows
  const bool use_last = A.last != nullptr && !(OP == OP_GAE && NVF);
  float *my_out = ostage + warp * (2 * C::kNOut * C::kWarpOut);
  float X = 0.0f, v_after = 0.0f;
  float last_next = (use_last && n_rows > 0) ? __ldg(A.last + blockIdx.x) : 0.0f;
  int64_t row = blockIdx.x;
  int c = 0;
  const int p0 = warp * (32 * SPL) + SPL * lane;
  int s = 0;
  uint32_t ph = 0;

#pragma unroll 1
  for (int64_t k = 0; k < n_items; ++k) {
    const int64_t t_end = T - (int64_t)c * BM_L;
    const int64_t t_begin = t_end > BM_L ? t_end - BM_L : 0;
    const int len = (int)(t_end - t_begin);
    if (c == 0) {  // a new environment: boundary values were requested one environment ago
      X = OP == OP_GAE ? 0.0f : last_next;
      v_after = OP == OP_GAE ? last_next : 0.0f;
      last_next = (use_last && row + G < A.N) ? __ldg(A.last + row + G) : 0.0f;
    }
    mbar_wait(&full[s], ph);
    const unsigned char *sb = stages + s * C::kStage;
    float ea[SPL], eb[SPL], aux[SPL];
    float v_first = 0.0f;
    {
      float rr[SPL], vq[SPL + 1], nq[SPL];
      uint32_t dd[SPL / 4];
      float edge = 0.0f;
#pragma unroll
      for (int q = 0; q < SPL / 4; ++q) {
        const float4 r4 = *reinterpret_cast<const float4 *>(sb + 4 * p0 + 16 * q);
        rr[4 * q] = r4.x, rr[4 * q + 1] = r4.y, rr[4 * q + 2] = r4.z, rr[4 * q + 3] = r4.w;
        if (OP != OP_DR) {
          const float4 v4 = *reinterpret_cast<const float4 *>(sb + C::kOffV + 4 * p0 + 16 * q);
          vq[4 * q] = v4.x, vq[4 * q + 1] = v4.y, vq[4 * q + 2] = v4.z, vq[4 * q + 3] = v4.w;
        }
        if (NVF) {
          const float4 n4 = *reinterpret_cast<const float4 *>(sb + C::kOffNV + 4 * p0 + 16 * q);
          nq[4 * q] = n4.x, nq[4 * q + 1] = n4.y, nq[4 * q + 2] = n4.z, nq[4 * q + 3] = n4.w;
        }
        dd[q] = HD ? *reinterpret_cast<const uint32_t *>(sb + C::kOffD + p0 + 4 * q) : 0u;
      }
      if (OP == OP_GAE && !NVF) {
        if (lane == 31) edge = *reinterpret_cast<const float *>(sb + C::kOffV + 4 * (p0 + SPL));
        v_first = *reinterpret_cast<const float *>(sb + C::kOffV);
      }
      release_stage(&empty[s], lane);
      const bool whole = len == BM_L;  // (block-uniform) no step of the tile lies past the chunk
      if (OP == OP_GAE && !NVF) {
        const float nb = __shfl_down_sync(FULL, vq[0], 1);
        vq[SPL] = lane == 31 ? edge : nb;
        if (p0 + SPL >= len) vq[SPL] = v_after;  // the chunk ends inside or right after this lane's run
        if (!whole) {
#pragma unroll
          for (int q = 1; q < SPL; ++q)
            if (p0 + q >= len) vq[q] = v_after;  // only read as V[t+1] of the last valid step
        }
      }
#pragma unroll
      for (int q = 0; q < SPL; ++q) {
        if (OP == OP_AFFINE) {
          ea[q] = rr[q], eb[q] = vq[q], aux[q] = 0.0f;
        } else {
          const bool dn = HD && ((dd[q >> 2] >> (8 * (q & 3))) & 0xffu) != 0u;
          if (OP == OP_DR) {
            ea[q] = rr[q], eb[q] = dn ? 0.0f : A.gamma, aux[q] = 0.0f;
          } else {
            const float gv = __fmul_rn(A.gamma, NVF ? nq[q] : vq[q + 1]);
            ea[q] = __fsub_rn(__fadd_rn(rr[q], dn ? 0.0f : gv), vq[q]);
            eb[q] = dn ? 0.0f : A.gl;
            aux[q] = vq[q];
          }
        }
      }
      if (!whole) {
#pragma unroll
        for (int q = 0; q < SPL; ++q)
          if (p0 + q >= len) ea[q] = 0.0f, eb[q] = 1.0f;  // past the chunk: identity map
      }
    }
    // per lane: SPL steps as one map; then a suffix scan over the lanes (later time = higher lane)
    float ca = 0.0f, cb = 1.0f;
#pragma unroll
    for (int q = SPL - 1; q >= 0; --q) {
      ca = fstep(ea[q], eb[q], ca);
      cb = __fmul_rn(eb[q], cb);
    }
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      float oa_ = __shfl_down_sync(FULL, ca, off);
      float ob_ = __shfl_down_sync(FULL, cb, off);
      if (lane + off >= 32) oa_ = 0.0f, ob_ = 1.0f;  // identity beyond the last lane
      ca = fstep(ca, cb, oa_);
      cb = __fmul_rn(cb, ob_);
    }
    float2 *cbuf = comp + (k & 1) * 16;
    if (lane == 0) cbuf[warp] = make_float2(ca, cb);
    row_barrier<32 * W>();
    float x = X, xin = X;
#pragma unroll
    for (int u = W - 1; u >= 0; --u) {
      if (u == warp) xin = x;
      const float2 m = cbuf[u];
      x = fstep(m.x, m.y, x);
    }
    X = x;
    if (OP == OP_GAE && !NVF) v_after = v_first;

    float *oa = my_out + (k & 1) * (C::kNOut * C::kWarpOut);
    if (lane == 0) bulk_wait_read<1>();
    __syncwarp();
    {
      const float x_here = fstep(ca, cb, xin);            // x at this lane's first step
      float xe = __shfl_down_sync(FULL, x_here, 1);       // x right after this lane's run
      if (lane == 31) xe = xin;
      float o[SPL], o2[SPL];
#pragma unroll
      for (int q = SPL - 1; q >= 0; --q) {
        xe = fstep(ea[q], eb[q], xe);
        o[q] = xe;
        o2[q] = __fadd_rn(xe, aux[q]);
      }
      if (MOM) {  // fp32 over the lane's steps, float64 across tiles
        float tsum = 0.0f, tsq = 0.0f;
#pragma unroll
        for (int q = 0; q < SPL; ++q)
          if (p0 + q < len) {
            tsum = __fadd_rn(tsum, o[q]);
            tsq = __fmaf_rn(o[q], o[q], tsq);
          }
        msum += (double)tsum;
        msq += (double)tsq;
      }
      float4 *dst = reinterpret_cast<float4 *>(oa + SPL * lane);
#pragma unroll
      for (int q = 0; q < SPL / 4; ++q) dst[q] = make_float4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
      if (want_ret) {
        float4 *dst2 = reinterpret_cast<float4 *>(oa + C::kWarpOut + SPL * lane);
#pragma unroll
        for (int q = 0; q < SPL / 4; ++q) dst2[q] = make_float4(o2[4 * q], o2[4 * q + 1], o2[4 * q + 2], o2[4 * q + 3]);
      }
    }
    fence_async_smem();
    __syncwarp();
    if (lane == 0) {
      int n = len - warp * C::kWarpOut;
      n = n > C::kWarpOut ? C::kWarpOut : n;
      if (n > 0) {
        const int64_t off = row * T + t_begin + warp * C::kWarpOut;
        bulk_store(A.out + off, oa, (uint32_t)n * 4u);
        if (want_ret) bulk_store(A.ret + off, oa + C::kWarpOut, (uint32_t)n * 4u);
      }
      bulk_commit();
    }
    if (++c == n_chunks) c = 0, row += G;
    if (++s == n_stages) s = 0, ph ^= 1u;
  }
  if (MOM) {  // fixed-order fold: lanes, then the warps in order -> one partial per CTA
#pragma unroll
    for (int off = 16; off; off >>= 1) {
      msum += __shfl_xor_sync(FULL, msum, off);
      msq += __shfl_xor_sync(FULL, msq, off);
    }
    double2 *red = reinterpret_cast<double2 *>(comp);
    row_barrier<32 * W>();  // every warp has read its last maps
    if (lane == 0) red[warp] = make_double2(msum, msq);
    row_barrier<32 * W>();
    if (warp == 0 && lane == 0) {
      double a = 0.0, b = 0.0;
      for (int w = 0; w < W; ++w) a += red[w].x, b += red[w].y;
      A.mom[blockIdx.x] = make_double2(a, b);
    }
  }
  if (lane == 0) bulk_wait_all();
}

template <int OP, bool HD, bool NVF, bool MOM, int SPL>
int row_launch_m(ScanArgs A, cudaStream_t st, int *mom_parts) {
  using C = RowCfg<OP, HD, NVF, SPL>;
  if (!MOM) A.mom = nullptr;
  if (first_use_on_device((const void *)scan_bm_row_kernel<OP, HD, NVF, MOM, SPL>))
    MRL_CUDA(cudaFuncSetAttribute(scan_bm_row_kernel<OP, HD, NVF, MOM, SPL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  C::smem_bytes(kMaxStages) < 226 * 1024 ? C::smem_bytes(kMaxStages) : 226 * 1024));
  const int n_chunks = (int)((A.T + BM_L - 1) / BM_L);
  // ring depth: the SM's shared memory split between g_row_ctas_per_sm resident CTAs
  const int per_sm = (int)g_row_ctas_per_sm;
  int stages = ((227 * 1024) / per_sm - 1024 - C::smem_bytes(0)) / C::kStage;
  stages = stages > kMaxStages ? kMaxStages : (stages < 2 ? 2 : stages);
  if (g_row_stages > 0) stages = (int)(g_row_stages < 2 ? 2 : g_row_stages);
  const int smem = C::smem_bytes(stages);
  // as many CTAs as fit at once, trimmed so that every CTA gets the same number of rows
  int fit = (227 * 1024) / (smem + 1024);
  const int by_threads = 2048 / C::kThreads;
  fit = fit > by_threads ? by_threads : (fit < 1 ? 1 : fit);
  const int64_t max_ctas = (int64_t)sm_count() * fit;
  const int64_t per = (A.N + max_ctas - 1) / max_ctas;
  const int64_t ctas = (A.N + per - 1) / per;
  scan_bm_row_kernel<OP, HD, NVF, MOM, SPL><<<(unsigned)ctas, C::kThreads, smem, st>>>(A, n_chunks, stages);
  if (mom_parts) *mom_parts = MOM ? (int)ctas : 0;  // (ctas <= 12 * SMs < kMomentSlots)
  return MRL_OK;
}

}  // namespace
int64_t g_row_spl = 8;  // option "row_spl": steps per lane of the batch-major rows (8 or 16)
namespace {

template <int OP, bool HD, bool NVF>
int row_launch(const ScanArgs &A, cudaStream_t st, int *mom_parts) {
  constexpr bool kMomOk = OP != OP_AFFINE;
  if (g_row_spl == 16) {
    if (A.mom && kMomOk) return row_launch_m<OP, HD, NVF, kMomOk, 16>(A, st, mom_parts);
    return row_launch_m<OP, HD, NVF, false, 16>(A, st, mom_parts);
  }

  if (A.mom && kMomOk) return row_launch_m<OP, HD, NVF, kMomOk, 8>(A, st, mom_parts);
  return row_launch_m<OP, HD, NVF, false, 8>(A, st, mom_parts);
}

}  // namespace

int64_t g_row_min_t = 512;

// Batch-major [N, T].  Bulk copies need 16-byte aligned row starts: T % 4 == 0 (T % 16 == 0 with done).
int scan_bm_row_launch(const ScanArgs &A, int op, bool force, cudaStream_t st, int *mom_parts) {
  if (mom_parts) *mom_parts = 0;
  int *mp = mom_parts;
  if (A.T % 4 != 0 || A.T >= (1LL << 40)) return kScanNotApplicable;
  if (!al16(A.a) || !al16(A.out) || (op != OP_DR && !al16(A.v)) || (A.ret && !al16(A.ret)) ||
      (A.nv && !al16(A.nv)))
    return kScanNotApplicable;
  const bool hd = op != OP_AFFINE && A.done != nullptr;
  if (hd && (A.T % 16 != 0 || !al16(A.done))) return kScanNotApplicable;
  if (!force && (A.T < g_row_min_t || A.N * A.T < (1 << 20))) return kScanNotApplicable;
  if (op == OP_DR) return hd ? row_launch<OP_DR, true, false>(A, st, mp) : row_launch<OP_DR, false, false>(A, st, mp);
  if (op == OP_AFFINE) return row_launch<OP_AFFINE, false, false>(A, st, mp);
  if (A.nv) return hd ? row_launch<OP_GAE, true, true>(A, st, mp) : row_launch<OP_GAE, false, true>(A, st, mp);
  return hd ? row_launch<OP_GAE, true, false>(A, st, mp) : row_launch<OP_GAE, false, false>(A, st, mp);
}

}  // namespace mrl
