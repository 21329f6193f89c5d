// copy.cu -- row movers of the experience data path (sm_100a).
//
//   copy_seg_kernel      ring append / get_item: contiguous runs of rows, widest aligned unit
//                        (128-bit when the column allows), 4 units in flight per thread.
//   gather_ldg_kernel    index-driven row gather through registers (small rows / fallback).
//   gather_bulk_kernel   index-driven row gather through the TMA engine: one thread per CTA drives
//                        a STAGES-deep shared-memory ring with cp.async.bulk global->shared
//                        (mbarrier complete_tx) and cp.async.bulk shared->global (bulk groups).
//                        SASS: UBLKCP.  No register traffic, no per-byte instructions.
//   fill_rows_kernel     synthetic rows for benches (slot id in bytes 0..7, hash after).
//
// Replaces the copy halves of BufferAppend / BufferGetItem / BufferSample and
// PriorityReplayBufferSample (uniform_replay_buffer.py:114,128,139; priority_replay_buffer.py:106).
#include <stdarg.h>
#include <stdlib.h>

#include "bulk.cuh"

namespace mrl {

// ---- segment copy ------------------------------------------------------------------------
template <class V>
__device__ __forceinline__ void copy_units(const uint8_t *__restrict__ src,
                                           uint8_t *__restrict__ dst, int64_t bytes) {
  const V *s = reinterpret_cast<const V *>(src);
  V *d = reinterpret_cast<V *>(dst);
  const int64_t n = bytes / (int64_t)sizeof(V);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < n; i += 4 * stride) {
    V a = s[i], b = s[i + stride], c = s[i + 2 * stride], e = s[i + 3 * stride];
    d[i] = a;
    d[i + stride] = b;
    d[i + 2 * stride] = c;
    d[i + 3 * stride] = e;
  }
  for (; i < n; i += stride) d[i] = s[i];
}

__global__ void __launch_bounds__(256)
    copy_seg_kernel(ColTable tab, int64_t ring_row, int64_t io_row, int64_t nrows, int to_ring) {
  pdl_prologue();  // (the DQN loop's insert -> sample pair: the launch latencies of the two overlap)
  const int c = blockIdx.y;
  const int64_t rb = tab.row_bytes[c];
  uint8_t *ring = tab.col[c] + ring_row * rb;
  uint8_t *io = tab.io[c] + io_row * rb;
  const uint8_t *src = to_ring ? io : ring;
  uint8_t *dst = to_ring ? ring : io;
  const int64_t bytes = nrows * rb;
  switch (tab.vec[c]) {
    case 16: copy_units<uint4>(src, dst, bytes); break;
    case 8: copy_units<uint2>(src, dst, bytes); break;
    case 4: copy_units<uint32_t>(src, dst, bytes); break;
    case 2: copy_units<uint16_t>(src, dst, bytes); break;
    default: copy_units<uint8_t>(src, dst, bytes); break;
  }
}

int copy_segments_dir(const ColTable &tab, int64_t ring_row, int64_t io_row, int64_t nrows,
                      int to_ring, cudaStream_t st) {
  if (nrows <= 0) return MRL_OK;
  if (to_ring == 2) {
    // io rows are HOST memory: a run of rows is contiguous in the source and in the ring, so the DMA
    // engine puts it in place directly (no staging copy, no kernel, no second pass over HBM)
    for (int c = 0; c < tab.ncols; ++c)
      MRL_CUDA(cudaMemcpyAsync(tab.col[c] + ring_row * tab.row_bytes[c], tab.io[c] + io_row * tab.row_bytes[c],
                               (size_t)(nrows * tab.row_bytes[c]), cudaMemcpyHostToDevice, st));
    return MRL_OK;
  }
  int64_t max_units = 1;
  for (int c = 0; c < tab.ncols; ++c) {
    int64_t u = nrows * tab.row_bytes[c] / tab.vec[c];
    if (u > max_units) max_units = u;
  }
  int64_t blocks = (max_units + 256 * 4 - 1) / (256 * 4);
  if (blocks < 1) blocks = 1;
  if (blocks > sm_count() * 8) blocks = sm_count() * 8;
  dim3 grid((unsigned)blocks, (unsigned)tab.ncols);
  MRL_CUDA(launch_pdl(copy_seg_kernel, grid, dim3(256), 0, st, tab, ring_row, io_row, nrows, to_ring));
  MRL_LAUNCHED();
  return MRL_OK;
}

int copy_segments(const ColTable &tab, int64_t dst_row, int64_t src_row, int64_t nrows,
                  cudaStream_t st) {
  return copy_segments_dir(tab, dst_row, src_row, nrows, 1, st);
}

// ---- gather through registers ------------------------------------------------------------
template <class V, class I>
__device__ __forceinline__ void gather_units(const uint8_t *__restrict__ col,
                                             uint8_t *__restrict__ out,
                                             const int64_t *__restrict__ slots, int64_t n,
                                             int64_t rb, int64_t capacity) {
  const I upr = (I)(rb / (int64_t)sizeof(V));
  const I total = (I)n * upr;
  const I stride = (I)gridDim.x * blockDim.x;
  I u = (I)blockIdx.x * blockDim.x + threadIdx.x;
  V *d = reinterpret_cast<V *>(out);
  for (; u < total; u += 4 * stride) {
    V v[4];
    I uu[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      uu[k] = u + (I)k * stride;
      if (uu[k] < total) {
        const I i = uu[k] / upr;
        const I off = uu[k] - i * upr;
        int64_t slot = __ldg(slots + i);
        slot = slot < 0 ? 0 : (slot >= capacity ? capacity - 1 : slot);
        v[k] = reinterpret_cast<const V *>(col + slot * rb)[off];
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (uu[k] < total) d[uu[k]] = v[k];
  }
}

__global__ void __launch_bounds__(256)
    gather_ldg_kernel(ColTable tab, const int64_t *__restrict__ slots, int64_t n,
                      int64_t capacity, uint32_t col_mask) {
  pdl_prologue();
  int c = blockIdx.y;
  // blockIdx.y enumerates the set bits of col_mask
  {
    uint32_t m = col_mask;
    for (int k = 0; k < c; ++k) m &= m - 1;
    c = __ffs(m) - 1;
  }
  const int64_t rb = tab.row_bytes[c];
  const bool small = n * rb < (int64_t)0x7fffffff;
  switch (tab.vec[c]) {
    case 16:
      if (small) gather_units<uint4, uint32_t>(tab.col[c], tab.io[c], slots, n, rb, capacity);
      else gather_units<uint4, uint64_t>(tab.col[c], tab.io[c], slots, n, rb, capacity);
      break;
    case 8:
      if (small) gather_units<uint2, uint32_t>(tab.col[c], tab.io[c], slots, n, rb, capacity);
      else gather_units<uint2, uint64_t>(tab.col[c], tab.io[c], slots, n, rb, capacity);
      break;
    case 4:
      if (small) gather_units<uint32_t, uint32_t>(tab.col[c], tab.io[c], slots, n, rb, capacity);
      else gather_units<uint32_t, uint64_t>(tab.col[c], tab.io[c], slots, n, rb, capacity);
      break;
    case 2:
      if (small) gather_units<uint16_t, uint32_t>(tab.col[c], tab.io[c], slots, n, rb, capacity);
      else gather_units<uint16_t, uint64_t>(tab.col[c], tab.io[c], slots, n, rb, capacity);
      break;
    default:
      if (small) gather_units<uint8_t, uint32_t>(tab.col[c], tab.io[c], slots, n, rb, capacity);
      else gather_units<uint8_t, uint64_t>(tab.col[c], tab.io[c], slots, n, rb, capacity);
      break;
  }
}

// ---- gather through the TMA engine -------------------------------------------------------
struct BulkPlan {
  int32_t ncols;
  int32_t col_id[MRL_MAX_COLS];
  int32_t chunk_bytes[MRL_MAX_COLS];
  int32_t nchunks[MRL_MAX_COLS];
  int64_t item_begin[MRL_MAX_COLS + 1];  // items of column k: [item_begin[k], item_begin[k+1])
};

struct BulkItem {
  const uint8_t *src;
  uint8_t *dst;
  uint32_t bytes;
};

__device__ __forceinline__ void bulk_decode(const ColTable &tab, const BulkPlan &plan,
                                            const int64_t *__restrict__ slots, int64_t capacity,
                                            int64_t item, BulkItem &it) {
  int k = 0;
  while (k + 1 < plan.ncols && item >= plan.item_begin[k + 1]) ++k;
  const int c = plan.col_id[k];
  const int64_t local = item - plan.item_begin[k];
  const int32_t nch = plan.nchunks[k];
  const int64_t i = local / nch;
  const int32_t j = (int32_t)(local - i * nch);
  const int64_t rb = tab.row_bytes[c];
  const int64_t off = (int64_t)j * plan.chunk_bytes[k];
  int64_t rem = rb - off;
  it.bytes = (uint32_t)(rem < plan.chunk_bytes[k] ? rem : plan.chunk_bytes[k]);
  int64_t slot = __ldg(slots + i);
  slot = slot < 0 ? 0 : (slot >= capacity ? capacity - 1 : slot);
  it.src = tab.col[c] + slot * rb + off;
  it.dst = tab.io[c] + i * rb + off;
}

template <int STAGES>
__global__ void __launch_bounds__(32)
    gather_bulk_kernel(const __grid_constant__ ColTable tab, const __grid_constant__ BulkPlan plan,
                       const int64_t *__restrict__ slots, int64_t capacity, int32_t stage_bytes) {
  extern __shared__ __align__(128) uint8_t stage_mem[];
  __shared__ __align__(8) uint64_t full_bar[STAGES];
  __shared__ BulkItem items[32];
  const int lane = threadIdx.x;

  if (lane == 0) {
    for (int s = 0; s < STAGES; ++s) mbar_init(&full_bar[s], 1);
    fence_async_smem();
  }
  pdl_prologue();  // the slots come from the sample kernel just before this one

  const int64_t total = plan.item_begin[plan.ncols];
  const int64_t first = blockIdx.x;
  const int64_t step = gridDim.x;
  const int64_t nmine = total > first ? (total - first + step - 1) / step : 0;

  // Lane 0 drives the DMA ring; the whole warp decodes the items 32 at a time (each needs slots[i], a
  // dependent load: decoded one by one in the ring loop those latencies added up to most of the
  // kernel's fixed cost -- 3.8 us against 2.0 us for an elementwise copy of the same volume).
  uint8_t *dst_of[STAGES];
  uint32_t bytes_of[STAGES];
#pragma unroll 1
  for (int64_t k = 0; k < nmine + STAGES - 1; ++k) {
    if ((k & 31) == 0 && k < nmine) {
      __syncwarp();
      if (k + lane < nmine) bulk_decode(tab, plan, slots, capacity, first + (k + lane) * step, items[lane]);
      __syncwarp();
    }
    if (lane != 0) continue;
    if (k >= STAGES - 1) {  // drain: item j has landed -> write it out
      const int64_t j = k - (STAGES - 1);
      const int s = (int)(j % STAGES);
      mbar_wait(&full_bar[s], (uint32_t)((j / STAGES) & 1));
      fence_async_smem();
#pragma unroll
      for (int q = 0; q < STAGES; ++q)
        if (q == s) bulk_s2g(dst_of[q], stage_mem + (size_t)q * stage_bytes, bytes_of[q]);
      bulk_commit();
    }
    if (k < nmine) {  // fill: stage k%STAGES was last read by the store of item k-STAGES
      const BulkItem cur = items[k & 31];
      const int s = (int)(k % STAGES);
      if (k >= STAGES) bulk_wait_read<1>();
#pragma unroll
      for (int q = 0; q < STAGES; ++q)
        if (q == s) {
          dst_of[q] = cur.dst;
          bytes_of[q] = cur.bytes;
        }
      mbar_expect_tx(&full_bar[s], cur.bytes);
      bulk_g2s(stage_mem + (size_t)s * stage_bytes, cur.src, cur.bytes, &full_bar[s]);
    }
  }
  if (lane == 0) bulk_wait_all();
}

// tunables (mrl_set_option / environment): gather_path 0 = auto, 1 = registers only, 2 = TMA ring
// whenever the column is 16-byte aligned
static int64_t env_i64(const char *name, int64_t dflt) {
  const char *e = getenv(name);
  return e ? atoll(e) : dflt;
}
int64_t g_gather_path = env_i64("MRL_GATHER_PATH", 0);
int64_t g_bulk_min_row = env_i64("MRL_BULK_MIN_ROW", 2048);
int64_t g_bulk_chunk = env_i64("MRL_BULK_CHUNK", 8192);
int64_t g_bulk_ctas_per_sm = env_i64("MRL_BULK_CTAS_PER_SM", 4);
int64_t g_bulk_stages = env_i64("MRL_BULK_STAGES", 4);

extern long long *g_debug_timers;  // prb.cu
extern bool g_no_inline_update;    // prb.cu
extern bool g_host_zero_copy;      // prb.cu
extern bool g_fused_gather;        // prb.cu
extern bool g_exact_shard_weights; // prb.cu
extern long long g_shard_timeout_ns;  // prb.cu
bool g_pdl = true;                 // programmatic dependent launch of the PER-path kernels
bool g_tree_prefetch = true;       // tree kernels pull the lines of their later round trips into L2 first

int set_copy_option(const char *key, int64_t value) {
  std::string k(key);
  if (k == "debug_timers") {
    g_debug_timers = reinterpret_cast<long long *>(value);
    return MRL_OK;
  }
  if (k == "inline_update" && (value == 0 || value == 1)) {
    g_no_inline_update = value == 0;
    return MRL_OK;
  }
  if (k == "pdl" && (value == 0 || value == 1)) {
    g_pdl = value == 1;
    return MRL_OK;
  }
  if (k == "host_zero_copy" && (value == 0 || value == 1)) {
    g_host_zero_copy = value == 1;
    return MRL_OK;
  }
  if (k == "shard_weights" && (value == 0 || value == 1)) {
    g_exact_shard_weights = value == 1;
    return MRL_OK;
  }
  if (k == "fused_gather" && (value == 0 || value == 1)) {
    g_fused_gather = value == 1;
    return MRL_OK;
  }
  if (k == "shard_timeout_ms" && value >= 1 && value <= 3600000) {
    g_shard_timeout_ns = value * 1000000ll;
    return MRL_OK;
  }
  if (k == "tree_prefetch" && (value == 0 || value == 1)) {
    g_tree_prefetch = value == 1;
    return MRL_OK;
  }
  if (k == "gather_path" && value >= 0 && value <= 2) g_gather_path = value;
  else if (k == "bulk_min_row" && value >= 16) g_bulk_min_row = value;
  else if (k == "bulk_chunk" && value >= 16 && value <= 49152 && value % 16 == 0) g_bulk_chunk = value;
  else if (k == "bulk_ctas_per_sm" && value >= 1 && value <= 16) g_bulk_ctas_per_sm = value;
  else if (k == "bulk_stages" && (value == 2 || value == 4 || value == 8 || value == 12)) g_bulk_stages = value;
  else return MRL_ERR_INVALID;
  return MRL_OK;
}

template <int STAGES>
static int launch_bulk(const ColTable &tab, const BulkPlan &plan, const int64_t *slots, int64_t capacity,
                       int32_t stage_bytes, int64_t blocks, cudaStream_t st) {
  const size_t smem = (size_t)stage_bytes * STAGES;
  // (per device: the limit is a property of the function ON a device; always the most a CTA may take, so
  // one setting serves every stage size)
  if (first_use_on_device((const void *)gather_bulk_kernel<STAGES>)) {
    cudaFuncAttributes fa;
    MRL_CUDA(cudaFuncGetAttributes(&fa, gather_bulk_kernel<STAGES>));
    MRL_CUDA(cudaFuncSetAttribute(gather_bulk_kernel<STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  227 * 1024 - (int)fa.sharedSizeBytes));
  }
  MRL_CUDA(launch_pdl(gather_bulk_kernel<STAGES>, dim3((unsigned)blocks), dim3(32), smem, st, tab, plan, slots,
                      capacity, stage_bytes));
  return MRL_OK;
}

int gather_rows(const ColTable &tab, const int64_t *slots, int64_t n, int64_t capacity,
                cudaStream_t st, uint32_t col_mask) {
  if (n <= 0) return MRL_OK;
  const int64_t bulk_min_row = g_bulk_min_row, bulk_chunk = g_bulk_chunk;
  const int64_t bulk_ctas_per_sm = g_bulk_ctas_per_sm;
  const int path = (int)g_gather_path;
  uint32_t ldg_mask = 0;
  BulkPlan plan;
  plan.ncols = 0;
  plan.item_begin[0] = 0;
  int32_t stage_bytes = 0;
  for (int c = 0; c < tab.ncols; ++c) {
    if (!(col_mask & (1u << c))) continue;
    const int64_t rb = tab.row_bytes[c];
    const bool eligible = tab.vec[c] == 16 && path != 1 && (path == 2 || rb >= bulk_min_row);
    if (!eligible) {
      ldg_mask |= 1u << c;
      continue;
    }
    const int k = plan.ncols++;
    int64_t nch = (rb + bulk_chunk - 1) / bulk_chunk;
    int64_t chunk = ((rb + nch - 1) / nch + 15) / 16 * 16;
    nch = (rb + chunk - 1) / chunk;
    plan.col_id[k] = c;
    plan.chunk_bytes[k] = (int32_t)chunk;
    plan.nchunks[k] = (int32_t)nch;
    plan.item_begin[k + 1] = plan.item_begin[k] + n * nch;
    if (chunk > stage_bytes) stage_bytes = (int32_t)chunk;
  }
  if (plan.ncols > 0) {
    const int64_t total = plan.item_begin[plan.ncols];
    int64_t blocks = sm_count() * bulk_ctas_per_sm;
    if (blocks > total) blocks = total;
    stage_bytes = (stage_bytes + 127) / 128 * 128;
    int stages = (int)g_bulk_stages;
    while (stages > 2 && (size_t)stage_bytes * stages > 200 * 1024) stages = stages == 12 ? 8 : stages / 2;
    int rc;
    if (stages == 12) rc = launch_bulk<12>(tab, plan, slots, capacity, stage_bytes, blocks, st);
    else if (stages == 8) rc = launch_bulk<8>(tab, plan, slots, capacity, stage_bytes, blocks, st);
    else if (stages == 4) rc = launch_bulk<4>(tab, plan, slots, capacity, stage_bytes, blocks, st);
    else rc = launch_bulk<2>(tab, plan, slots, capacity, stage_bytes, blocks, st);
    if (rc) return rc;
    MRL_LAUNCHED();
  }
  if (ldg_mask) {
    int ncols = __builtin_popcount(ldg_mask);
    int64_t max_units = 1;
    for (int c = 0; c < tab.ncols; ++c)
      if (ldg_mask & (1u << c)) {
        int64_t u = n * tab.row_bytes[c] / tab.vec[c];
        if (u > max_units) max_units = u;
      }
    int64_t blocks = (max_units + 256 * 4 - 1) / (256 * 4);
    if (blocks > sm_count() * 8) blocks = sm_count() * 8;
    dim3 grid((unsigned)blocks, (unsigned)ncols);
    MRL_CUDA(launch_pdl(gather_ldg_kernel, grid, dim3(256), 0, st, tab, slots, n, capacity, ldg_mask));
    MRL_LAUNCHED();
  }
  return MRL_OK;
}

// ---- synthetic fill ----------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    fill_rows_kernel(uint8_t *col, int64_t nrows, int64_t row_bytes, uint64_t seed) {
  const int64_t words_per_row = (row_bytes + 3) / 4;
  const int64_t total = nrows * words_per_row;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < total; w += stride) {
    const int64_t r = w / words_per_row;
    const int64_t wi = w - r * words_per_row;
    uint32_t v = mrl_hash32(seed ^ (uint64_t)r * 0x100000001B3ULL, (uint64_t)wi);
    if (row_bytes >= 8 && wi < 2) v = (uint32_t)((uint64_t)r >> (32 * wi));
    uint8_t *p = col + r * row_bytes + wi * 4;
    const int64_t left = row_bytes - wi * 4;
    if (left >= 4 && (((uintptr_t)p) & 3) == 0) {
      *reinterpret_cast<uint32_t *>(p) = v;
    } else {
      for (int b = 0; b < 4 && b < left; ++b) p[b] = (uint8_t)(v >> (8 * b));
    }
  }
}

// ---- swap of the two leading axes of a column: out[b][a][:] = in[a][b][:] -------------------
// MSRL.get_replay_buffer_elements(transpose=True, shape=(1, 0, 2)) (core/msrl.py:535-557): every PPO
// trainer turns its [T, env, d] columns into [env, T, d] this way before the [env, T] scans.
// Elements of 1/2/4/8 bytes: 32 x 32 tiles through shared memory (both sides coalesced).  Wider rows:
// one thread per aligned unit of the OUTPUT (coalesced stores, row-contiguous loads served by L2).
template <typename T>
__global__ void __launch_bounds__(256) transpose_tile_kernel(const T *__restrict__ in, T *__restrict__ out,
                                                            int64_t A, int64_t B) {
  __shared__ T tile[32][33];
  const int64_t tiles_b = (B + 31) / 32;
  const int64_t n_tiles = ((A + 31) / 32) * tiles_b;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    const int64_t a0 = (t / tiles_b) * 32, b0 = (t % tiles_b) * 32;
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
      const int64_t a = a0 + ty + j, b = b0 + tx;
      if (a < A && b < B) tile[ty + j][tx] = in[a * B + b];
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
      const int64_t b = b0 + ty + j, a = a0 + tx;
      if (a < A && b < B) out[b * A + a] = tile[tx][ty + j];
    }
    __syncthreads();
  }
}

template <typename V>
__global__ void __launch_bounds__(256) transpose_rows_kernel(const V *__restrict__ in, V *__restrict__ out,
                                                            int64_t A, int64_t B, int64_t U) {
  const int64_t total = A * B * U;
  for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (int64_t)gridDim.x * blockDim.x) {
    const int64_t u = o % U, ba = o / U;
    const int64_t a = ba % A, b = ba / A;
    out[o] = in[(a * B + b) * U + u];
  }
}

template <typename T>
static void launch_tile(const void *in, void *out, int64_t A, int64_t B, cudaStream_t st) {
  const int64_t n_tiles = ((A + 31) / 32) * ((B + 31) / 32);
  const int64_t grid = n_tiles < (int64_t)sm_count() * 8 ? n_tiles : (int64_t)sm_count() * 8;
  transpose_tile_kernel<T><<<(unsigned)grid, 256, 0, st>>>((const T *)in, (T *)out, A, B);
}
template <typename V>
static void launch_rows(const void *in, void *out, int64_t A, int64_t B, int64_t row_bytes, cudaStream_t st) {
  const int64_t U = row_bytes / (int64_t)sizeof(V);
  const int64_t blocks = (A * B * U + 255) / 256;
  const int64_t grid = blocks < (int64_t)sm_count() * 16 ? blocks : (int64_t)sm_count() * 16;
  transpose_rows_kernel<V><<<(unsigned)grid, 256, 0, st>>>((const V *)in, (V *)out, A, B, U);
}

}  // namespace mrl

extern "C" int mrl_set_option(const char *key, int64_t value) {
  MRL_REQUIRE(key, "mrl_set_option: NULL key");
  int rc = mrl::set_copy_option(key, value);
  if (rc) rc = mrl::set_scan_option(key, value);
  if (rc) mrl::set_error("mrl_set_option: unknown key or value out of range: %s=%lld", key, (long long)value);
  return rc;
}

extern "C" int mrl_transpose01(const void *in, void *out, int64_t A, int64_t B, int64_t row_bytes,
                               void *stream) {
  MRL_REQUIRE(A >= 0 && B >= 0 && row_bytes > 0, "mrl_transpose01: bad shape");
  if (A == 0 || B == 0) return MRL_OK;
  MRL_REQUIRE(in && out && in != out, "mrl_transpose01: need distinct input and output");
  cudaStream_t st = mrl::as_stream(stream);
  const uintptr_t al = (uintptr_t)in | (uintptr_t)out | (uintptr_t)row_bytes;
  if (row_bytes == 8 && al % 8 == 0) mrl::launch_tile<uint64_t>(in, out, A, B, st);
  else if (row_bytes == 4 && al % 4 == 0) mrl::launch_tile<uint32_t>(in, out, A, B, st);
  else if (row_bytes == 2 && al % 2 == 0) mrl::launch_tile<uint16_t>(in, out, A, B, st);
  else if (row_bytes == 1) mrl::launch_tile<uint8_t>(in, out, A, B, st);
  else if (al % 16 == 0) mrl::launch_rows<uint4>(in, out, A, B, row_bytes, st);
  else if (al % 8 == 0) mrl::launch_rows<uint2>(in, out, A, B, row_bytes, st);
  else if (al % 4 == 0) mrl::launch_rows<uint32_t>(in, out, A, B, row_bytes, st);
  else if (al % 2 == 0) mrl::launch_rows<uint16_t>(in, out, A, B, row_bytes, st);
  else mrl::launch_rows<uint8_t>(in, out, A, B, row_bytes, st);
  MRL_LAUNCHED();
  return MRL_OK;
}

extern "C" int mrl_fill_rows(void *col, int64_t nrows, int64_t row_bytes, uint64_t seed,
                             void *stream) {
  MRL_REQUIRE(col && nrows >= 0 && row_bytes > 0, "mrl_fill_rows: bad arguments");
  if (nrows == 0) return MRL_OK;
  mrl::fill_rows_kernel<<<mrl::sm_count() * 8, 256, 0, mrl::as_stream(stream)>>>(
      (uint8_t *)col, nrows, row_bytes, seed);
  MRL_LAUNCHED();
  return MRL_OK;
}
