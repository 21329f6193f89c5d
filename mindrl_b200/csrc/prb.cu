// prb.cu -- PriorityReplayBuffer ops: PriorityReplayBuffer{Create,Push,Sample,Update,Destroy}
// (mindspore_rl/core/priority_replay_buffer.py:60-138).
//
// Layout in HBM
//   ring columns   [capacity, row_bytes[i]] bytes each, row-contiguous (hidden behind the handle,
//                  as in the reference)
//   sum[2*cap2], mn[2*cap2]   fp32 segment tree in heap order (root = 1, children 2i / 2i+1,
//                  leaves at cap2 + slot, cap2 = capacity rounded up to a power of two), stored as
//                  two arrays: the descent reads only `sum`, so a 32-descendant hop is one 128-byte
//                  line.
//   tag[cap2]      uint64 (call epoch << 32 | batch position): last-writer-wins arbiter of update
//   state          {max_priority}
//
// Kernels
//   prb_sample_kernel   warp per sample.  Top <=10 levels of `sum` staged in shared memory per CTA;
//                       below that the warp hops 5 levels per step: one coalesced 128-byte load of
//                       the 32 descendants, pairwise sums rebuilt with shuffles (bit-identical to the
//                       stored inner nodes because every parent is fl(left+right)), then five
//                       shuffle-compare steps.  4 dependent L2 loads instead of 20 for 2^20 leaves.
//                       Importance weights and small-row gather are fused behind the descent.
//   prb_update_kernel   one CTA.  atomicMax tag arbitration -> winners write leaves -> level-synchronous
//                       parent recompute for the lower levels -> top 10 levels recomputed in shared
//                       memory.  Parents are always recomputed from both children (never delta
//                       atomics), so the tree equals the sequential oracle's bit for bit.
//   prb_push_kernel     one warp: row write, leaf <- max_priority, 20 sibling loads in parallel, path
//                       recomputed through shuffles.
//   prb_rebuild_kernel  range rebuild after a batched push.
#include <float.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace mrl {

struct PrbState {
  float max_priority;
  float pad[3];
};

struct PrbBuffer {
  int64_t capacity = 0, cap2 = 1;
  int32_t levels = 0;  // log2(cap2)
  float alpha = 0.f;
  int32_t ncols = 0;
  int64_t row_bytes[MRL_MAX_COLS];
  uint8_t *cols[MRL_MAX_COLS];
  float *sum = nullptr, *mn = nullptr;
  unsigned long long *tag = nullptr;
  PrbState *state = nullptr;
  int64_t size = 0, head = -1;
  uint64_t seed = 0, ctr = 0, epoch = 0;
  bool small_rows = true;  // gather fused into the sample kernel
  Staging stage;
};

static Registry<PrbBuffer> g_prb;

constexpr int kTopLevels = 11;          // levels 0..10 (2047 nodes) staged in shared memory
constexpr int kSampleWarps = 8;         // warps (= samples) per CTA of prb_sample_kernel
constexpr int64_t kFusedRowLimit = 512;  // column row bytes up to which sample() gathers in-kernel

__global__ void __launch_bounds__(256)
    prb_init_kernel(float *sum, float *mn, unsigned long long *tag, PrbState *state, int64_t cap2) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < 2 * cap2; i += stride) {
    sum[i] = 0.0f;
    mn[i] = FLT_MAX;
    if (i < cap2) tag[i] = 0ull;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) state->max_priority = 1.0f;
}

// ---- sample ------------------------------------------------------------------------------
struct SampleArgs {
  const float *sum;
  const float *mn;
  const float *uniforms;  // or NULL
  const float *gstats;    // [world][4] or NULL
  int64_t *indices;
  float *weights;
  int64_t n, cap2, capacity, size;
  int32_t levels, world;
  float beta;
  uint64_t seed, ctr;
  uint32_t gather_mask;  // columns whose rows this kernel copies itself (narrow rows)
};

__global__ void __launch_bounds__(kSampleWarps * 32)
    prb_sample_kernel(const __grid_constant__ SampleArgs S, const __grid_constant__ ColTable tab) {
  __shared__ float top[1 << kTopLevels];
  const int top_levels = S.levels < kTopLevels ? S.levels : kTopLevels;  // levels fully in smem
  // nodes [1, 2^(top_levels+1)) cover levels 0..top_levels, but only keep what fits: levels
  // 0..top_levels-1 plus (when levels < kTopLevels) the leaf level is read from global.
  const int staged = 1 << top_levels;  // nodes 1 .. staged-1  => levels 0 .. top_levels-1
  for (int i = threadIdx.x; i < staged; i += blockDim.x) top[i] = i ? __ldcg(S.sum + i) : 0.0f;
  __syncthreads();

  const int lane = threadIdx.x & 31;
  const int64_t i = (int64_t)blockIdx.x * kSampleWarps + (threadIdx.x >> 5);
  if (i >= S.n) return;
  constexpr unsigned FULL = 0xffffffffu;

  const float root_sum = staged > 1 ? top[1] : __ldcg(S.sum + 1);
  const float seg = __fdiv_rn(root_sum, (float)S.n);
  const float u = S.uniforms ? __ldg(S.uniforms + i) : mrl_uniform01(S.seed, S.ctr + (uint64_t)i);
  float mass = __fmul_rn(__fadd_rn(u, (float)i), seg);

  // levels whose children are staged: node at level l has children at level l+1 <= top_levels-1
  int64_t node = 1;
  int level = 0;
  for (; level + 1 < top_levels; ++level) {
    const float left = top[2 * node];
    if (mass <= left) {
      node = 2 * node;
    } else {
      mass = __fsub_rn(mass, left);
      node = 2 * node + 1;
    }
  }
  // warp hops: h <= 5 levels per 128-byte load
  float leaf_p = 0.0f;
  int remaining = S.levels - level;
  if (remaining == 0) leaf_p = __ldcg(S.sum + node);  // capacity 1
  while (remaining > 0) {
    const int h = remaining < 5 ? remaining : 5;
    const int64_t base = node << h;
    float p[6];
    p[0] = lane < (1 << h) ? __ldcg(S.sum + base + lane) : 0.0f;
#pragma unroll
    for (int k = 1; k <= 5; ++k) p[k] = __fadd_rn(p[k - 1], __shfl_xor_sync(FULL, p[k - 1], 1 << (k - 1)));
    int off = 0;
#pragma unroll
    for (int k = 4; k >= 0; --k) {
      if (k < h) {  // left child at this step spans 2^k leaves starting at `off`
        const float left = __shfl_sync(FULL, p[k], off);
        if (!(mass <= left)) {
          mass = __fsub_rn(mass, left);
          off += 1 << k;
        }
      }
    }
    node = base + off;
    remaining -= h;
    if (remaining == 0) leaf_p = __shfl_sync(FULL, p[0], off);
  }
  const int64_t leaf = node - S.cap2;

  if (lane == 0) {
    float w_sum = root_sum, w_min = __ldcg(S.mn + 1), w_size = (float)S.size;
    if (S.gstats) {
      w_sum = 0.0f, w_min = FLT_MAX, w_size = 0.0f;
      for (int r = 0; r < S.world; ++r) {
        w_sum = __fadd_rn(w_sum, __ldg(S.gstats + 4 * r));
        const float m = __ldg(S.gstats + 4 * r + 1);
        w_min = m < w_min ? m : w_min;
        w_size = __fadd_rn(w_size, __ldg(S.gstats + 4 * r + 2));
      }
    }
    const float max_w = mrl_is_weight(w_min, w_sum, w_size, S.beta);
    S.indices[i] = leaf;
    S.weights[i] = __fdiv_rn(mrl_is_weight(leaf_p, w_sum, w_size, S.beta), max_w);
  }
  if (S.gather_mask) {
    const int64_t slot = leaf < S.capacity ? leaf : S.capacity - 1;
    for (int c = 0; c < tab.ncols; ++c) {
      if (!((S.gather_mask >> c) & 1u)) continue;
      const int64_t rb = tab.row_bytes[c];
      const uint8_t *src = tab.col[c] + slot * rb;
      uint8_t *dst = tab.io[c] + i * rb;
      const int v = tab.vec[c];
      if (v == 16) {
        for (int64_t o = lane * 16; o < rb; o += 512)
          *reinterpret_cast<uint4 *>(dst + o) = *reinterpret_cast<const uint4 *>(src + o);
      } else if (v >= 4) {
        for (int64_t o = lane * 4; o < rb; o += 128)
          *reinterpret_cast<uint32_t *>(dst + o) = *reinterpret_cast<const uint32_t *>(src + o);
      } else {
        for (int64_t o = lane; o < rb; o += 32) dst[o] = src[o];
      }
    }
  }
}

// ---- update ------------------------------------------------------------------------------
struct UpdateArgs {
  float *sum;
  float *mn;
  unsigned long long *tag;
  PrbState *state;
  const int64_t *indices;
  const float *priorities;
  int64_t n, cap2;
  int32_t levels;
  float alpha;
  uint64_t epoch;
};

__device__ __forceinline__ void recompute_node(float *sum, float *mn, int64_t node) {
  const float a = __ldcg(sum + 2 * node), b = __ldcg(sum + 2 * node + 1);
  const float ma = __ldcg(mn + 2 * node), mb = __ldcg(mn + 2 * node + 1);
  __stcg(sum + node, __fadd_rn(a, b));
  __stcg(mn + node, ma < mb ? ma : mb);
}

__global__ void __launch_bounds__(1024) prb_update_kernel(const __grid_constant__ UpdateArgs U) {
  __shared__ float ssum[2 << kTopLevels];
  __shared__ float smin[2 << kTopLevels];
  const int tid = threadIdx.x, nthr = blockDim.x;

  // 0) arbitration: the highest batch position of a leaf wins; max over ALL priorities (also the
  //    overwritten ones) goes to max_priority, as in the sequential loop
  float local_max = 0.0f;
  for (int64_t k = tid; k < U.n; k += nthr) {
    const int64_t leaf = __ldg(U.indices + k);
    atomicMax(U.tag + leaf, (unsigned long long)((U.epoch << 32) | (uint64_t)k));
    const float p = mrl_priority_pow(__ldg(U.priorities + k), U.alpha);
    local_max = p > local_max ? p : local_max;
  }
#pragma unroll
  for (int off = 16; off; off >>= 1) {
    const float o = __shfl_xor_sync(0xffffffffu, local_max, off);
    local_max = o > local_max ? o : local_max;
  }
  if ((tid & 31) == 0 && local_max > 0.0f)
    atomicMax(reinterpret_cast<int *>(&U.state->max_priority), __float_as_int(local_max));
  __syncthreads();

  // 1) winners write their leaf
  for (int64_t k = tid; k < U.n; k += nthr) {
    const int64_t leaf = __ldg(U.indices + k);
    if (__ldcg(U.tag + leaf) == (unsigned long long)((U.epoch << 32) | (uint64_t)k)) {
      const float p = mrl_priority_pow(__ldg(U.priorities + k), U.alpha);
      __stcg(U.sum + U.cap2 + leaf, p);
      __stcg(U.mn + U.cap2 + leaf, p);
    }
  }
  __syncthreads();

  // 2) lower levels, level-synchronous (duplicate ancestors write identical values)
  const int split = U.levels < kTopLevels ? U.levels : kTopLevels;  // level held in smem as input
  for (int level = U.levels - 1; level >= split; --level) {
    const int shift = U.levels - level;
    for (int64_t k = tid; k < U.n; k += nthr) {
      const int64_t node = (U.cap2 + __ldg(U.indices + k)) >> shift;
      recompute_node(U.sum, U.mn, node);
    }
    __syncthreads();
  }

  // 3) top `split` levels rebuilt in shared memory from the 2^split nodes of level `split`
  const int width = 1 << split;
  for (int i = tid; i < width; i += nthr) {
    ssum[width + i] = __ldcg(U.sum + width + i);
    smin[width + i] = __ldcg(U.mn + width + i);
  }
  __syncthreads();
  for (int level = split - 1; level >= 0; --level) {
    const int w = 1 << level;
    for (int i = tid; i < w; i += nthr) {
      const int node = w + i;
      ssum[node] = __fadd_rn(ssum[2 * node], ssum[2 * node + 1]);
      const float ma = smin[2 * node], mb = smin[2 * node + 1];
      smin[node] = ma < mb ? ma : mb;
    }
    __syncthreads();
  }
  for (int i = tid + 1; i < width; i += nthr) {
    __stcg(U.sum + i, ssum[i]);
    __stcg(U.mn + i, smin[i]);
  }
}

// ---- update, low-latency variant (n <= kFastUpdateMax) -------------------------------------
// Same result as prb_update_kernel, organised around the number of DEPENDENT memory round trips
// (the tree is cold in L2 whenever a learner step ran in between):
//   * duplicates are arbitrated in a shared-memory hash table (leaf -> highest batch position),
//     no global tag traffic;
//   * the lower levels are rebuilt three levels per barrier: a thread loads the aligned group of 8
//     nodes below its ancestor (one 32-byte sector per array), rebuilds the 4+2+1 nodes above them
//     and stores them -- 9 levels cost 3 round trips instead of 9;
//   * every line a later round will need is prefetched into L2 in the first phase.
constexpr int kFastUpdateMax = 4096;
constexpr int kFastItems = kFastUpdateMax / 1024;

__device__ __forceinline__ uint32_t upd_hash(int64_t leaf, int bits) {
  return ((uint32_t)leaf * 2654435761u) >> (32 - bits);
}
__device__ __forceinline__ void prefetch_l2(const void *p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

__global__ void __launch_bounds__(1024)
    prb_update_fast_kernel(const __grid_constant__ UpdateArgs U, int slot_bits) {
  __shared__ float ssum[2 << kTopLevels];
  __shared__ float smin[2 << kTopLevels];
  extern __shared__ __align__(16) unsigned char upd_dyn[];
  const int nslots = 1 << slot_bits;
  unsigned long long *table = reinterpret_cast<unsigned long long *>(upd_dyn);
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int split = U.levels < kTopLevels ? U.levels : kTopLevels;

  for (int i = tid; i < nslots; i += nthr) table[i] = 0ull;
  // prefetch the inputs of the shared-memory top rebuild (level `split`)
  {
    const int width = 1 << split;
    for (int i = tid * 32; i < width; i += nthr * 32) {
      prefetch_l2(U.sum + width + i);
      prefetch_l2(U.mn + width + i);
    }
  }
  __syncthreads();

  int64_t leaf_of[kFastItems];
  float p_of[kFastItems];
  int slot_of[kFastItems];
  float local_max = 0.0f;
#pragma unroll
  for (int r = 0; r < kFastItems; ++r) {
    const int64_t k = tid + (int64_t)r * nthr;
    leaf_of[r] = -1;
    p_of[r] = 0.0f;
    slot_of[r] = 0;
    if (k < U.n) {
      const int64_t leaf = __ldg(U.indices + k);
      leaf_of[r] = leaf;
      // lines of the three-level rounds
      for (int lvl = U.levels; lvl > split; lvl -= 3) {
        const int64_t g = ((U.cap2 + leaf) >> (U.levels - lvl)) & ~(int64_t)7;
        prefetch_l2(U.sum + g);
        prefetch_l2(U.mn + g);
      }
      const float p = mrl_priority_pow(__ldg(U.priorities + k), U.alpha);
      p_of[r] = p;
      local_max = p > local_max ? p : local_max;
      const unsigned long long mine = ((unsigned long long)(leaf + 1) << 32) | (unsigned long long)k;
      uint32_t h = upd_hash(leaf, slot_bits);
      while (true) {
        unsigned long long cur = *reinterpret_cast<volatile unsigned long long *>(table + h);
        if (cur == 0ull) {
          cur = atomicCAS(table + h, 0ull, mine);
          if (cur == 0ull) break;
        }
        if ((cur >> 32) == (unsigned long long)(leaf + 1)) {
          atomicMax(table + h, mine);
          break;
        }
        h = (h + 1) & (nslots - 1);
      }
      slot_of[r] = (int)h;
    }
  }
#pragma unroll
  for (int off = 16; off; off >>= 1) {
    const float o = __shfl_xor_sync(0xffffffffu, local_max, off);
    local_max = o > local_max ? o : local_max;
  }
  if ((tid & 31) == 0 && local_max > 0.0f)
    atomicMax(reinterpret_cast<int *>(&U.state->max_priority), __float_as_int(local_max));
  __syncthreads();

  // winners (highest batch position of a leaf) write their leaf
#pragma unroll
  for (int r = 0; r < kFastItems; ++r) {
    if (leaf_of[r] >= 0) {
      const int64_t k = tid + (int64_t)r * nthr;
      if ((uint32_t)table[slot_of[r]] == (uint32_t)k) {
        __stcg(U.sum + U.cap2 + leaf_of[r], p_of[r]);
        __stcg(U.mn + U.cap2 + leaf_of[r], p_of[r]);
      }
    }
  }
  __syncthreads();

  // lower levels: h <= 3 levels per barrier
  for (int lvl = U.levels; lvl > split;) {
    const int h = (lvl - split) < 3 ? (lvl - split) : 3;
    const int gsz = 1 << h;
#pragma unroll
    for (int r = 0; r < kFastItems; ++r) {
      if (leaf_of[r] < 0) continue;
      const int64_t node = (U.cap2 + leaf_of[r]) >> (U.levels - lvl);
      const int64_t g = node & ~(int64_t)(gsz - 1);
      float vs[8], vm[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        vs[j] = 0.0f, vm[j] = FLT_MAX;
        if (j < gsz) {
          vs[j] = __ldcg(U.sum + g + j);
          vm[j] = __ldcg(U.mn + g + j);
        }
      }
      // rebuild h levels above the group; level lvl-1 nodes start at g>>1, lvl-2 at g>>2, ...
      int cnt = gsz;
      int64_t base = g;
#pragma unroll
      for (int step = 0; step < 3; ++step) {
        if (step < h) {
          cnt >>= 1;
          base >>= 1;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (j < cnt) {
              const float a = vs[2 * j], b = vs[2 * j + 1];
              const float ma = vm[2 * j], mb = vm[2 * j + 1];
              vs[j] = __fadd_rn(a, b);
              vm[j] = ma < mb ? ma : mb;
              __stcg(U.sum + base + j, vs[j]);
              __stcg(U.mn + base + j, vm[j]);
            }
          }
        }
      }
    }
    lvl -= h;
    __syncthreads();
  }

  // top `split` levels in shared memory
  const int width = 1 << split;
  for (int i = tid; i < width; i += nthr) {
    ssum[width + i] = __ldcg(U.sum + width + i);
    smin[width + i] = __ldcg(U.mn + width + i);
  }
  __syncthreads();
  for (int level = split - 1; level >= 0; --level) {
    const int w = 1 << level;
    for (int i = tid; i < w; i += nthr) {
      const int node = w + i;
      ssum[node] = __fadd_rn(ssum[2 * node], ssum[2 * node + 1]);
      const float ma = smin[2 * node], mb = smin[2 * node + 1];
      smin[node] = ma < mb ? ma : mb;
    }
    __syncthreads();
  }
  for (int i = tid + 1; i < width; i += nthr) {
    __stcg(U.sum + i, ssum[i]);
    __stcg(U.mn + i, smin[i]);
  }
}

// ---- push --------------------------------------------------------------------------------
__global__ void __launch_bounds__(32)
    prb_push_kernel(float *sum, float *mn, const PrbState *state, int64_t cap2, int32_t levels,
                    int64_t slot, const __grid_constant__ ColTable tab) {
  const int lane = threadIdx.x;
  constexpr unsigned FULL = 0xffffffffu;
  // row write (io -> ring)
  for (int c = 0; c < tab.ncols; ++c) {
    const int64_t rb = tab.row_bytes[c];
    const uint8_t *src = tab.io[c];
    uint8_t *dst = tab.col[c] + slot * rb;
    const int v = tab.vec[c];
    if (v == 16) {
      for (int64_t o = lane * 16; o < rb; o += 512)
        *reinterpret_cast<uint4 *>(dst + o) = *reinterpret_cast<const uint4 *>(src + o);
    } else if (v >= 4) {
      for (int64_t o = lane * 4; o < rb; o += 128)
        *reinterpret_cast<uint32_t *>(dst + o) = *reinterpret_cast<const uint32_t *>(src + o);
    } else {
      for (int64_t o = lane; o < rb; o += 32) dst[o] = src[o];
    }
  }
  const float p = state->max_priority;
  const int64_t leaf_node = cap2 + slot;
  // lane j holds the sibling of the path node at height j (levels <= 31)
  float sib_s = 0.0f, sib_m = FLT_MAX;
  if (lane < levels) {
    const int64_t sib = (leaf_node >> lane) ^ 1;
    sib_s = __ldcg(sum + sib);
    sib_m = __ldcg(mn + sib);
  }
  float vs = p, vm = p, my_s = 0.0f, my_m = 0.0f;
  for (int j = 0; j < levels; ++j) {
    const float s = __shfl_sync(FULL, sib_s, j);
    const float m = __shfl_sync(FULL, sib_m, j);
    vs = __fadd_rn(vs, s);
    vm = m < vm ? m : vm;
    if (lane == j) {
      my_s = vs;
      my_m = vm;
    }
  }
  if (lane < levels) {
    const int64_t parent = leaf_node >> (lane + 1);
    __stcg(sum + parent, my_s);
    __stcg(mn + parent, my_m);
  }
  if (lane == 0) {
    __stcg(sum + leaf_node, p);
    __stcg(mn + leaf_node, p);
  }
}

// leaves [lo, hi) (slots, no wrap) <- max_priority, then every ancestor of the range, one level at a
// time (single CTA; used by the batched push)
__global__ void __launch_bounds__(1024)
    prb_rebuild_kernel(float *sum, float *mn, const PrbState *state, int64_t cap2, int32_t levels,
                       int64_t lo, int64_t hi) {
  const float p = state->max_priority;
  for (int64_t s = lo + threadIdx.x; s < hi; s += blockDim.x) {
    __stcg(sum + cap2 + s, p);
    __stcg(mn + cap2 + s, p);
  }
  __syncthreads();
  int64_t a = cap2 + lo, b = cap2 + hi - 1;
  for (int l = 0; l < levels; ++l) {
    a >>= 1;
    b >>= 1;
    for (int64_t node = a + threadIdx.x; node <= b; node += blockDim.x) recompute_node(sum, mn, node);
    __syncthreads();
  }
}

// multi-CTA variant for large ranges: one launch per level
__global__ void __launch_bounds__(256)
    prb_set_leaves_kernel(float *sum, float *mn, const PrbState *state, int64_t cap2, int64_t lo,
                          int64_t hi) {
  const float p = state->max_priority;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t s = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < hi; s += stride) {
    sum[cap2 + s] = p;
    mn[cap2 + s] = p;
  }
}
__global__ void __launch_bounds__(256)
    prb_level_kernel(float *sum, float *mn, int64_t a, int64_t b) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t node = a + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; node <= b; node += stride) {
    const float x = sum[2 * node], y = sum[2 * node + 1];
    const float mx = mn[2 * node], my = mn[2 * node + 1];
    sum[node] = __fadd_rn(x, y);
    mn[node] = mx < my ? mx : my;
  }
}

__global__ void prb_stats_kernel(const float *sum, const float *mn, const PrbState *state,
                                 float size, float *out4) {
  out4[0] = __ldcg(sum + 1);
  out4[1] = __ldcg(mn + 1);
  out4[2] = size;
  out4[3] = state->max_priority;
}

static void fill_table(ColTable &tab, const PrbBuffer *b, void *const *io) {
  tab.ncols = b->ncols;
  for (int c = 0; c < b->ncols; ++c) {
    tab.col[c] = b->cols[c];
    tab.io[c] = io ? (uint8_t *)io[c] : nullptr;
    tab.row_bytes[c] = b->row_bytes[c];
    tab.vec[c] = align_unit((uintptr_t)b->cols[c], io ? (uintptr_t)io[c] : 0, (uint64_t)b->row_bytes[c]);
  }
}

static int rebuild_range(PrbBuffer *b, int64_t lo, int64_t hi, cudaStream_t st) {
  if (hi <= lo) return MRL_OK;
  if (hi - lo <= 4096) {
    prb_rebuild_kernel<<<1, 1024, 0, st>>>(b->sum, b->mn, b->state, b->cap2, b->levels, lo, hi);
    MRL_LAUNCHED();
    return MRL_OK;
  }
  int64_t blocks = (hi - lo + 255) / 256;
  if (blocks > kNumSM * 8) blocks = kNumSM * 8;
  prb_set_leaves_kernel<<<(unsigned)blocks, 256, 0, st>>>(b->sum, b->mn, b->state, b->cap2, lo, hi);
  MRL_LAUNCHED();
  int64_t a = b->cap2 + lo, e = b->cap2 + hi - 1;
  for (int l = 0; l < b->levels; ++l) {
    a >>= 1;
    e >>= 1;
    int64_t nb = (e - a + 1 + 255) / 256;
    if (nb > kNumSM * 8) nb = kNumSM * 8;
    prb_level_kernel<<<(unsigned)nb, 256, 0, st>>>(b->sum, b->mn, a, e);
    MRL_LAUNCHED();
  }
  return MRL_OK;
}

static int prb_push_batch_impl(PrbBuffer *b, int64_t nrows, void *const *src, cudaStream_t st) {
  const int64_t cap = b->capacity;
  if (nrows == 1) {
    const int64_t slot = (b->head + 1) % cap;
    ColTable tab;
    fill_table(tab, b, src);
    prb_push_kernel<<<1, 32, 0, st>>>(b->sum, b->mn, b->state, b->cap2, b->levels, slot, tab);
    MRL_LAUNCHED();
    b->head = slot;
    if (b->size < cap) b->size++;
    return MRL_OK;
  }
  const int64_t pos = (b->head + 1) % cap;
  const int64_t skip = nrows > cap ? nrows - cap : 0;
  const int64_t first = (pos + skip) % cap;
  const int64_t n = nrows - skip;
  ColTable tab;
  fill_table(tab, b, src);
  const int64_t run1 = n < cap - first ? n : cap - first;
  int rc = copy_segments_dir(tab, first, skip, run1, 1, st);
  if (rc) return rc;
  rc = rebuild_range(b, first, first + run1, st);
  if (rc) return rc;
  if (n > run1) {
    rc = copy_segments_dir(tab, 0, skip + run1, n - run1, 1, st);
    if (rc) return rc;
    rc = rebuild_range(b, 0, n - run1, st);
    if (rc) return rc;
  }
  b->head = (b->head + nrows) % cap;
  b->size = b->size + nrows > cap ? cap : b->size + nrows;
  return MRL_OK;
}

static int prb_sample_impl(PrbBuffer *b, int64_t n, float beta, const float *uniforms,
                           const float *gstats, int32_t world, int64_t *indices, float *weights,
                           void *const *out_cols, cudaStream_t st) {
  SampleArgs S;
  S.sum = b->sum, S.mn = b->mn, S.uniforms = uniforms, S.gstats = gstats;
  S.indices = indices, S.weights = weights;
  S.n = n, S.cap2 = b->cap2, S.capacity = b->capacity, S.size = b->size;
  S.levels = b->levels, S.world = world, S.beta = beta;
  S.seed = b->seed, S.ctr = b->ctr;
  // narrow columns are copied by the warp that found the leaf; wide ones go to the TMA gather
  uint32_t fused = 0, wide = 0;
  if (out_cols)
    for (int c = 0; c < b->ncols; ++c) {
      if (b->row_bytes[c] <= kFusedRowLimit) fused |= 1u << c;
      else wide |= 1u << c;
    }
  S.gather_mask = fused;
  ColTable tab;
  fill_table(tab, b, out_cols);
  const int64_t blocks = (n + kSampleWarps - 1) / kSampleWarps;
  prb_sample_kernel<<<(unsigned)blocks, kSampleWarps * 32, 0, st>>>(S, tab);
  MRL_LAUNCHED();
  if (!uniforms) b->ctr += (uint64_t)n;
  if (wide) return gather_rows(tab, indices, n, b->capacity, st, wide);
  return MRL_OK;
}

static int g_update_path = -1;  // MRL_UPDATE_PATH=tags forces the global-tag kernel

static int prb_update_impl(PrbBuffer *b, const int64_t *indices, const float *priorities,
                           int64_t n, cudaStream_t st) {
  if (g_update_path < 0) {
    const char *e = getenv("MRL_UPDATE_PATH");
    g_update_path = (e && e[0] == 't') ? 1 : 0;
  }
  UpdateArgs U;
  U.sum = b->sum, U.mn = b->mn, U.tag = b->tag, U.state = b->state;
  U.indices = indices, U.priorities = priorities;
  U.n = n, U.cap2 = b->cap2, U.levels = b->levels, U.alpha = b->alpha;
  U.epoch = ++b->epoch;
  if (n <= kFastUpdateMax && g_update_path == 0) {
    int bits = 6;
    while ((1 << bits) < 2 * n) ++bits;
    const size_t smem = (size_t)(1 << bits) * 8;
    static size_t configured = 0;
    if (smem > configured) {
      MRL_CUDA(cudaFuncSetAttribute(prb_update_fast_kernel,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(8192 * 8)));
      configured = 8192 * 8;
    }
    prb_update_fast_kernel<<<1, 1024, smem, st>>>(U, bits);
  } else {
    prb_update_kernel<<<1, 1024, 0, st>>>(U);
  }
  MRL_LAUNCHED();
  return MRL_OK;
}

}  // namespace mrl

using namespace mrl;

#define PRB_GET(b, h)                                                      \
  PrbBuffer *b = g_prb.get(h);                                             \
  if (!b) {                                                                \
    set_error("unknown PriorityReplayBuffer handle %lld", (long long)(h)); \
    return MRL_ERR_HANDLE;                                                 \
  }

static void prb_free(PrbBuffer *b) {
  for (int c = 0; c < b->ncols; ++c)
    if (b->cols[c]) cudaFree(b->cols[c]);
  if (b->sum) cudaFree(b->sum);
  if (b->mn) cudaFree(b->mn);
  if (b->tag) cudaFree(b->tag);
  if (b->state) cudaFree(b->state);
  b->stage.release();
  delete b;
}

extern "C" int mrl_prb_create(int64_t capacity, float alpha, int32_t ncols,
                              const int64_t *row_bytes, uint64_t seed0, uint64_t seed1,
                              int64_t *handle) {
  MRL_REQUIRE(handle, "mrl_prb_create: NULL handle");
  *handle = MRL_INVALID_HANDLE;
  MRL_REQUIRE(capacity > 0 && capacity <= (1LL << 30), "mrl_prb_create: capacity out of range");
  MRL_REQUIRE(ncols > 0 && ncols <= MRL_MAX_COLS && row_bytes, "mrl_prb_create: bad column table");
  PrbBuffer *b = new PrbBuffer();
  b->capacity = capacity;
  while (b->cap2 < capacity) {
    b->cap2 *= 2;
    b->levels++;
  }
  b->alpha = alpha;
  b->ncols = ncols;
  b->seed = mrl_mix_seed(seed0, seed1);
  int64_t total_row = 0;
  for (int c = 0; c < ncols; ++c) {
    b->row_bytes[c] = row_bytes[c];
    b->cols[c] = nullptr;
    total_row += row_bytes[c];
  }
  b->small_rows = total_row <= kFusedRowLimit;
  cudaError_t e = cudaSuccess;
  for (int c = 0; c < ncols && e == cudaSuccess; ++c) {
    if (row_bytes[c] <= 0) {
      prb_free(b);
      set_error("mrl_prb_create: row_bytes[%d] must be positive", c);
      return MRL_ERR_INVALID;
    }
    const size_t bytes = (size_t)capacity * (size_t)row_bytes[c];
    e = cudaMalloc((void **)&b->cols[c], bytes);
    if (e == cudaSuccess) e = cudaMemset(b->cols[c], 0, bytes);
  }
  if (e == cudaSuccess) e = cudaMalloc((void **)&b->sum, sizeof(float) * 2 * (size_t)b->cap2);
  if (e == cudaSuccess) e = cudaMalloc((void **)&b->mn, sizeof(float) * 2 * (size_t)b->cap2);
  if (e == cudaSuccess) e = cudaMalloc((void **)&b->tag, sizeof(unsigned long long) * (size_t)b->cap2);
  if (e == cudaSuccess) e = cudaMalloc((void **)&b->state, sizeof(PrbState));
  if (e != cudaSuccess) {
    prb_free(b);
    set_error("mrl_prb_create: allocation failed: %s", cudaGetErrorString(e));
    return MRL_ERR_CUDA;
  }
  prb_init_kernel<<<kNumSM * 4, 256>>>(b->sum, b->mn, b->tag, b->state, b->cap2);
  g_launches.fetch_add(1);
  e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    prb_free(b);
    set_error("mrl_prb_create: init failed: %s", cudaGetErrorString(e));
    return MRL_ERR_CUDA;
  }
  *handle = g_prb.put(b);
  return MRL_OK;
}

extern "C" int mrl_prb_destroy(int64_t handle) {
  PrbBuffer *b = g_prb.take(handle);
  if (!b) {
    set_error("unknown PriorityReplayBuffer handle %lld", (long long)handle);
    return MRL_ERR_HANDLE;
  }
  cudaDeviceSynchronize();
  prb_free(b);
  return MRL_OK;
}

extern "C" int mrl_prb_push(int64_t handle, const void *const *row_cols, void *stream) {
  PRB_GET(b, handle);
  MRL_REQUIRE(row_cols, "mrl_prb_push: NULL row table");
  return prb_push_batch_impl(b, 1, (void *const *)row_cols, as_stream(stream));
}

extern "C" int mrl_prb_push_batch(int64_t handle, int64_t nrows, const void *const *src_cols,
                                  void *stream) {
  PRB_GET(b, handle);
  MRL_REQUIRE(nrows >= 1 && src_cols, "mrl_prb_push_batch: need nrows >= 1 and a source table");
  return prb_push_batch_impl(b, nrows, (void *const *)src_cols, as_stream(stream));
}

extern "C" int mrl_prb_push_batch_host(int64_t handle, int64_t nrows, const void *const *src_cols,
                                       void *stream) {
  PRB_GET(b, handle);
  MRL_REQUIRE(nrows >= 1 && src_cols, "mrl_prb_push_batch_host: need nrows >= 1 and a source table");
  cudaStream_t st = as_stream(stream);
  size_t off[MRL_MAX_COLS + 1];
  off[0] = 0;
  for (int c = 0; c < b->ncols; ++c)
    off[c + 1] = off[c] + ((size_t)nrows * b->row_bytes[c] + 255) / 256 * 256;
  int rc = b->stage.wait();
  if (rc) return rc;
  rc = b->stage.ensure(off[b->ncols]);
  if (rc) return rc;
  void *src_dev[MRL_MAX_COLS];
  for (int c = 0; c < b->ncols; ++c) {
    memcpy(b->stage.host + off[c], src_cols[c], (size_t)nrows * b->row_bytes[c]);
    src_dev[c] = b->stage.dev + off[c];
  }
  MRL_CUDA(cudaMemcpyAsync(b->stage.dev, b->stage.host, off[b->ncols], cudaMemcpyHostToDevice, st));
  rc = prb_push_batch_impl(b, nrows, src_dev, st);
  if (rc) return rc;
  MRL_CUDA(cudaEventRecord(b->stage.done, st));
  return MRL_OK;
}

extern "C" int mrl_prb_push_host(int64_t handle, const void *const *row_cols, void *stream) {
  return mrl_prb_push_batch_host(handle, 1, row_cols, stream);
}

extern "C" int mrl_prb_sample(int64_t handle, int64_t n, float beta, const float *uniforms,
                              int64_t *indices, float *weights, void *const *out_cols,
                              void *stream) {
  PRB_GET(b, handle);
  MRL_REQUIRE(n >= 1 && indices && weights, "mrl_prb_sample: need n >= 1, indices and weights");
  return prb_sample_impl(b, n, beta, uniforms, nullptr, 0, indices, weights, out_cols,
                         as_stream(stream));
}

extern "C" int mrl_prb_sample_sharded(int64_t handle, int64_t n, float beta, const float *uniforms,
                                      const float *gstats, int32_t world, int64_t *indices,
                                      float *weights, void *const *out_cols, void *stream) {
  PRB_GET(b, handle);
  MRL_REQUIRE(n >= 1 && indices && weights && gstats && world >= 1,
              "mrl_prb_sample_sharded: bad arguments");
  return prb_sample_impl(b, n, beta, uniforms, gstats, world, indices, weights, out_cols,
                         as_stream(stream));
}

extern "C" int mrl_prb_sample_host(int64_t handle, int64_t n, float beta, const float *uniforms,
                                   int64_t *indices, float *weights, void *const *out_cols,
                                   void *stream) {
  PRB_GET(b, handle);
  MRL_REQUIRE(n >= 1 && indices && weights, "mrl_prb_sample_host: need n >= 1, indices and weights");
  cudaStream_t st = as_stream(stream);
  // device arena: [indices | weights | uniforms | col0 | col1 ...]
  size_t off[MRL_MAX_COLS + 4];
  const size_t o_idx = 0, o_w = (size_t)n * 8, o_u = o_w + (size_t)n * 4;
  off[0] = (o_u + (size_t)n * 4 + 255) / 256 * 256;
  for (int c = 0; c < b->ncols; ++c)
    off[c + 1] = off[c] + (out_cols ? ((size_t)n * b->row_bytes[c] + 255) / 256 * 256 : 0);
  const size_t total = off[b->ncols];
  int rc = b->stage.wait();
  if (rc) return rc;
  rc = b->stage.ensure(total);
  if (rc) return rc;
  uint8_t *d = b->stage.dev;
  if (uniforms) {
    memcpy(b->stage.host + o_u, uniforms, (size_t)n * 4);
    MRL_CUDA(cudaMemcpyAsync(d + o_u, b->stage.host + o_u, (size_t)n * 4, cudaMemcpyHostToDevice, st));
  }
  void *out_dev[MRL_MAX_COLS];
  for (int c = 0; c < b->ncols; ++c) out_dev[c] = d + off[c];
  rc = prb_sample_impl(b, n, beta, uniforms ? (const float *)(d + o_u) : nullptr, nullptr, 0,
                       (int64_t *)(d + o_idx), (float *)(d + o_w), out_cols ? out_dev : nullptr, st);
  if (rc) return rc;
  if (total <= (1u << 16)) {
    MRL_CUDA(cudaMemcpyAsync(b->stage.host, d, total, cudaMemcpyDeviceToHost, st));
    MRL_CUDA(cudaStreamSynchronize(st));
    memcpy(indices, b->stage.host + o_idx, (size_t)n * 8);
    memcpy(weights, b->stage.host + o_w, (size_t)n * 4);
    if (out_cols)
      for (int c = 0; c < b->ncols; ++c)
        memcpy(out_cols[c], b->stage.host + off[c], (size_t)n * b->row_bytes[c]);
  } else {
    MRL_CUDA(cudaMemcpyAsync(b->stage.host, d, o_u, cudaMemcpyDeviceToHost, st));
    if (out_cols)
      for (int c = 0; c < b->ncols; ++c)
        MRL_CUDA(cudaMemcpyAsync(out_cols[c], d + off[c], (size_t)n * b->row_bytes[c],
                                 cudaMemcpyDeviceToHost, st));
    MRL_CUDA(cudaStreamSynchronize(st));
    memcpy(indices, b->stage.host + o_idx, (size_t)n * 8);
    memcpy(weights, b->stage.host + o_w, (size_t)n * 4);
  }
  return MRL_OK;
}

extern "C" int mrl_prb_update(int64_t handle, const int64_t *indices, const float *priorities,
                              int64_t n, void *stream) {
  PRB_GET(b, handle);
  MRL_REQUIRE(n >= 0 && (n == 0 || (indices && priorities)), "mrl_prb_update: bad arguments");
  MRL_REQUIRE(n < (1LL << 32), "mrl_prb_update: batch too large");
  if (n == 0) return MRL_OK;
  return prb_update_impl(b, indices, priorities, n, as_stream(stream));
}

extern "C" int mrl_prb_update_host(int64_t handle, const int64_t *indices,
                                   const float *priorities, int64_t n, void *stream) {
  PRB_GET(b, handle);
  MRL_REQUIRE(n >= 0 && (n == 0 || (indices && priorities)), "mrl_prb_update_host: bad arguments");
  if (n == 0) return MRL_OK;
  cudaStream_t st = as_stream(stream);
  const size_t o_p = (size_t)n * 8, total = o_p + (size_t)n * 4;
  int rc = b->stage.wait();
  if (rc) return rc;
  rc = b->stage.ensure(total);
  if (rc) return rc;
  memcpy(b->stage.host, indices, (size_t)n * 8);
  memcpy(b->stage.host + o_p, priorities, (size_t)n * 4);
  MRL_CUDA(cudaMemcpyAsync(b->stage.dev, b->stage.host, total, cudaMemcpyHostToDevice, st));
  rc = prb_update_impl(b, (const int64_t *)b->stage.dev, (const float *)(b->stage.dev + o_p), n, st);
  if (rc) return rc;
  MRL_CUDA(cudaEventRecord(b->stage.done, st));
  return MRL_OK;
}

extern "C" int mrl_prb_info(int64_t handle, int64_t *size, int64_t *head, int64_t *leaf_capacity) {
  PRB_GET(b, handle);
  if (size) *size = b->size;
  if (head) *head = b->head;
  if (leaf_capacity) *leaf_capacity = b->cap2;
  return MRL_OK;
}

extern "C" int mrl_prb_max_priority(int64_t handle, float *max_priority, void *stream) {
  PRB_GET(b, handle);
  MRL_REQUIRE(max_priority, "mrl_prb_max_priority: NULL output");
  cudaStream_t st = as_stream(stream);
  MRL_CUDA(cudaMemcpyAsync(max_priority, &b->state->max_priority, sizeof(float),
                           cudaMemcpyDeviceToHost, st));
  MRL_CUDA(cudaStreamSynchronize(st));
  return MRL_OK;
}

extern "C" int mrl_prb_tree_dump(int64_t handle, float *sum_nodes, float *min_nodes, void *stream) {
  PRB_GET(b, handle);
  cudaStream_t st = as_stream(stream);
  const size_t bytes = sizeof(float) * 2 * (size_t)b->cap2;
  if (sum_nodes) MRL_CUDA(cudaMemcpyAsync(sum_nodes, b->sum, bytes, cudaMemcpyDeviceToHost, st));
  if (min_nodes) MRL_CUDA(cudaMemcpyAsync(min_nodes, b->mn, bytes, cudaMemcpyDeviceToHost, st));
  MRL_CUDA(cudaStreamSynchronize(st));
  return MRL_OK;
}

extern "C" int mrl_prb_columns(int64_t handle, void **col_ptrs_out) {
  PRB_GET(b, handle);
  for (int c = 0; c < b->ncols; ++c) col_ptrs_out[c] = b->cols[c];
  return MRL_OK;
}

extern "C" int mrl_prb_root_stats(int64_t handle, float *stats4, void *stream) {
  PRB_GET(b, handle);
  MRL_REQUIRE(stats4, "mrl_prb_root_stats: NULL output");
  prb_stats_kernel<<<1, 1, 0, as_stream(stream)>>>(b->sum, b->mn, b->state, (float)b->size, stats4);
  MRL_LAUNCHED();
  return MRL_OK;
}
