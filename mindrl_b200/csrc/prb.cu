// prb.cu -- PriorityReplayBuffer ops: PriorityReplayBuffer{Create,Push,Sample,Update,Destroy}
// (mindspore_rl/core/priority_replay_buffer.py:60-138).
//
// Layout in HBM
//   ring columns   [capacity, row_bytes[i]] bytes each, row-contiguous (hidden behind the handle,
//                  as in the reference)
//   sum[2*cap2], mn[2*cap2]   fp32 segment tree in heap order (root = 1, children 2i / 2i+1,
//                  leaves at cap2 + slot, cap2 = capacity rounded up to a power of two), stored as
//                  two arrays: the descent reads only `sum`, so the 128 descendants seven levels below
//                  a node are four consecutive 128-byte lines.
//   tag[cap2]      uint64 (call epoch << 32 | batch position): last-writer-wins arbiter of the one-CTA
//                  update kernel
//   state          {max_priority, done_ctas}
//
// Kernels (the defaults; DESIGN.md section 5 has the measurements)
//   prb_sample_kernel      warp per sample.  From the root the warp hops 7 levels per dependent load:
//                          float4 per lane (the 128 descendants), two levels rebuilt inside the lane and
//                          five with shuffles (bit-identical to the stored inner nodes because every
//                          parent is fl(left+right)), five shuffle-compare steps to the lane and two
//                          inside its quad: 3 dependent loads for 2^20 leaves.  Importance weights (two
//                          pows side by side on two lanes) and the narrow columns (one list of words,
//                          requested before the weights are computed) are fused behind the descent; in
//                          the sharded buffer one warp per CTA reads the peers' root statistics out of the
//                          local inbox (delivered a step earlier by the peers' courier kernels).
//   prb_sample_gather_kernel  the same descent and the TMA row gather of Atari-shaped rows in ONE kernel.
//   prb_update_sub_kernel  128..512 CTAs own the subtrees below level 11.  Normal case (a share of <= 32
//                          entries): one warp -- all path siblings in one round trip, who-meets-whom from
//                          one shuffle pass (highest differing bit of two leaf numbers), branch-free
//                          register walk, node stores afterwards.  Larger shares: thread per entry with a
//                          shared-memory list, or whole subtrees rebuilt in shared memory.  The last CTA
//                          (acq_rel ticket) rebuilds the top 11 levels.  Parents are always recomputed
//                          from both children (never delta atomics), so the tree equals the sequential
//                          oracle's bit for bit.
//   prb_update_kernel, prb_update_walk_kernel   one-CTA variants (small trees / MRL_UPDATE_PATH).
//   prb_push_kernel        one warp: row write, leaf <- max_priority, sibling loads in parallel, path
//                          recomputed through shuffles.
//   prb_rebuild_kernel     range rebuild after a batched push.
#include <float.h>
#include <stdlib.h>
#include <string.h>

#include <utility>

#include "bulk.cuh"

namespace mrl {

struct PrbState {
  float max_priority;
  unsigned int done_ctas;  // subtree update: CTAs that have finished their lower levels
  float pad[2];
};

// Peer-memory mailboxes of the sharded buffer.  Every rank owns an INBOX of [2][world][4] and an OUTBOX of [2][4]
// flag-in-payload words {value bits, tag} (8-byte accesses are single-copy atomic: a reader that sees the tag
// sees the value) in one allocation that every peer has mapped (CUDA IPC).  Two protocols share the words:
//   kPublishOnMutation (default): whoever CHANGES a rank's tree -- the last CTA of the update kernel, the
//       push kernel, the tail of a batched push / set_state -- stores that rank's {root_sum, root_min, size},
//       tag = the rank's mutation count, into the rank's OWN outbox: a local store.  A one-warp COURIER kernel on
//       a side stream of the handle (ordered behind the mutating kernel by an event) then copies the outbox words
//       into every peer's inbox over NVLink.  Neither the update nor the sample kernel ever touches peer memory:
//       measured (round 2, 2 GPUs), pushing the words from the update kernel itself made that kernel 2 us longer
//       (it cannot retire before the NVLink writes are acknowledged), and pulling them from the sample kernel
//       (remote loads issued before the descent) cost the step 6 us -- a kernel that has touched peer memory
//       retires late, whenever in its life it did so.  The courier's 2 us are off the critical path.
//       sample #j reads slot (j-1)&1 of its LOCAL inbox and expects tag == its own mutation count: ranks run the
//       same program (SPMD), so a peer's statistics "as of its own sample #j" are exactly its publication number
//       <own mutation count>, delivered a whole step before they are needed -- no wait in the sample step unless
//       a peer is more than a step behind.  Mutations issued after sample #j go to slot j&1 (outbox and inbox),
//       so a peer that runs ahead never overwrites what a slower rank has yet to read: it cannot pass ITS sample
//       #j+1 before the slower rank's courier has delivered the mutations that follow its sample #j, which are
//       ordered behind that sample.  For the same reason an outbox slot is never rewritten before its courier ran.
//   kPublishOnSample (round 1): the sample kernel itself pushes the statistics into all peers' inboxes, tag =
//       sample call number, and waits for all peers' words of the same call: every step ends when the slowest
//       rank has got there.
// Contract of both: all calls on a sharded handle go to ONE stream, every rank issues the same sequence of
// sample / mutating calls.  A poll that sees nothing for `shard_timeout_ms` gives up, raises the link's error
// word (mapped host memory) and the next call on the handle returns MRL_ERR_STATE.
constexpr int kMaxWorld = 32;
constexpr int kPublishOnSample = 0, kPublishOnMutation = 1;
struct ShardDiag {                    // device memory, read back by mrl_prb_shard_diag
  unsigned long long wait_cycles;     // SM cycles warp 0 of CTA 0 spent in the poll, summed over calls
  unsigned long long polled_calls;    // sample calls that polled
  unsigned long long waited_calls;    // ... whose first look did not find every peer's word
  unsigned long long timeouts;
};
struct ShardLink {
  int rank = -1, world = 0, mode = kPublishOnMutation;
  uint2 *mailbox = nullptr;     // this rank's mailbox (cudaMalloc, exported through CUDA IPC)
  uint2 *peer[kMaxWorld] = {};  // mapped mailboxes of all ranks (peer[rank] == mailbox)
  uint2 **peer_dev = nullptr;   // the same table in device memory
  uint32_t epoch = 0;           // sharded sample calls so far (identical on every rank)
  uint32_t version = 0;         // publications so far = mutating calls since connect + 1 (identical on every rank)
  uint32_t published_epoch = 0; // `epoch` at the last publication: a sample without a mutation since the last
                                // sample re-publishes into its own slot first
  ShardDiag *diag = nullptr;
  int *err_host = nullptr;      // pinned + mapped: raised by a kernel whose poll timed out
  int *err_dev = nullptr;
  cudaStream_t courier = nullptr;  // side stream of the courier kernels
  cudaEvent_t mutated = nullptr;   // recorded behind every mutating kernel, awaited by the courier
};
long long g_shard_timeout_ns = 10000000000ll;  // option "shard_timeout_ms" [10000]

// what a kernel needs to publish the root statistics (all zero = not sharded / not this protocol)
struct PublishArgs {
  uint2 *const *peers = nullptr;  // kPublishOnSample: push into every rank's inbox
  uint2 *outbox = nullptr;        // kPublishOnMutation: this rank's own outbox [2][4]
  int32_t rank = 0, world = 0;
  uint32_t slot = 0, tag = 0;
  float size = 0.0f;
};
// lanes of ONE warp.  Outbox: lane 0 stores this rank's three words locally.  Push: lane q stores them into rank
// q's inbox (NVLink peer stores).  Relaxed, system scope -- the words validate themselves.
__device__ __forceinline__ void publish_root(const PublishArgs &P, int lane, float root_sum, float root_min) {
  uint2 *dst = nullptr;
  if (P.outbox) {
    if (lane == 0) dst = P.outbox + (size_t)P.slot * 4;
  } else if (P.peers && lane < P.world) {
    dst = P.peers[lane] + ((size_t)P.slot * P.world + P.rank) * 4;
  }
  if (dst) {
    const float vals[3] = {root_sum, root_min, P.size};
#pragma unroll
    for (int q = 0; q < 3; ++q)
      asm volatile("st.relaxed.sys.global.v2.u32 [%0], {%1,%2};" ::"l"(dst + q), "r"(__float_as_uint(vals[q])),
                   "r"(P.tag)
                   : "memory");
  }
}
// the courier: lane q copies this rank's outbox words of `slot` into rank q's inbox (side stream)
__global__ void __launch_bounds__(32)
    prb_courier_kernel(const uint2 *outbox, uint2 *const *peers, int rank, int world, unsigned int slot) {
  const int q = threadIdx.x;
  if (q >= world || q == rank) return;
  const uint2 *src = outbox + (size_t)slot * 4;
  uint2 *dst = peers[q] + ((size_t)slot * world + rank) * 4;
#pragma unroll
  for (int w = 0; w < 3; ++w) {
    uint2 v;
    asm volatile("ld.relaxed.sys.global.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(src + w) : "memory");
    asm volatile("st.relaxed.sys.global.v2.u32 [%0], {%1,%2};" ::"l"(dst + w), "r"(v.x), "r"(v.y) : "memory");
  }
}

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
  ShardLink link;
};

static Registry<PrbBuffer> g_prb;

// debug: when set (mrl_set_option("debug_timers", device pointer to >= 32 int64)), thread 0 of the
// tree kernels stores clock64() at its phase boundaries
long long *g_debug_timers = nullptr;
__device__ __forceinline__ void dbg_mark(long long *dbg, int slot) {
  if (dbg && threadIdx.x == 0 && blockIdx.x == 0) {
    dbg[slot] = clock64();  // SM cycles (reading %globaltimer costs ~1 us and distorts the phases)
  }
}

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
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    state->max_priority = 1.0f;
    state->done_ctas = 0u;
  }
}

// global-timer variant (ns) for marks taken by different CTAs
__device__ __forceinline__ void dbg_mark_gt(long long *dbg, int slot) {
  if (dbg && threadIdx.x == 0) {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    dbg[slot] = t;
  }
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
  long long *dbg;
  // sharded, peer-memory exchange (mailbox == NULL: off)
  PublishArgs pub;       // kPublishOnSample: this kernel stores the rank's statistics itself (pub.peers != NULL)
  uint2 *mailbox;        // this rank's mailbox (inbox first)
  uint32_t slot, expect; // words to read: slot, and the tag they must carry
  ShardDiag *diag;
  int *err;              // raised when the poll times out
  long long timeout_ns;
  const float *beta_dev; // != NULL: beta is read from device memory (AOT shim: beta arrives as a tensor)
  int32_t exact_weights; // sharded: weights from the probability the sample was really drawn with (option "shard_weights")
  int32_t prefetch;  // option "tree_prefetch"
  int32_t word_units;  // > 0: every fused column is made of aligned 4-byte words; their total per row
};

constexpr int kHop = 7;  // tree levels per dependent load of the descent

// the 2^h descendants h levels below `node`, spread over the warp: 4 per lane for h = 7, 2 for h = 6, else 1
__device__ __forceinline__ float4 hop_load(const float *sum, int64_t node, int h, int lane) {
  float4 v = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
  const float *base = sum + (node << h);
  if (h == 7) {
    v = __ldcg(reinterpret_cast<const float4 *>(base) + lane);
  } else if (h == 6) {
    const float2 q = __ldcg(reinterpret_cast<const float2 *>(base) + lane);
    v.x = q.x, v.y = q.y;
  } else if (lane < (1 << h)) {
    v.x = __ldcg(base + lane);
  }
  return v;
}

// Warp descent from the root to the leaf that holds `mass`, h <= 7 levels per coalesced load of the 2^h
// descendants h levels below the node (512 bytes: lane l holds descendants 4l..4l+3).  The sums in between are
// rebuilt exactly as they are stored -- every parent is fl(left + right) -- two levels inside the lane, five with
// shuffles; the walk then takes five shuffle-compare steps to the lane and two inside its quad.  20 levels = 3
// dependent round trips.  (Staging the top 11 levels in shared memory first was measured slower: 3.3 us for the
// staging alone against ~0.5 us per hop; the top lines are shared by all warps.)  hop0 = the first hop's line,
// requested by the caller before the mass was known.  Returns the leaf (slot), its priority in leaf_p.
__device__ __forceinline__ int64_t prb_descend(const float *sum, int levels, int64_t cap2, int64_t size, float mass,
                                               const float4 hop0, float root_sum, int lane, float &leaf_p_out) {
  constexpr unsigned FULL = 0xffffffffu;
  int64_t node = 1;
  float leaf_p = root_sum;  // capacity 1: the root is the leaf
  int remaining = levels;
#pragma unroll 1
  while (remaining > 0) {
    const int h = remaining < kHop ? remaining : kHop;
    const int wl = h > 5 ? h - 5 : 0;  // levels inside a lane (0, 1 or 2)
    const int hx = h - wl;             // levels across lanes
    const int64_t base = node << h;
    const float4 v = node == 1 ? hop0 : hop_load(sum, node, h, lane);
    const float s01 = __fadd_rn(v.x, v.y), s23 = __fadd_rn(v.z, v.w);
    float p[6];
    p[0] = wl == 2 ? __fadd_rn(s01, s23) : (wl == 1 ? s01 : v.x);
#pragma unroll
    for (int k = 1; k <= 5; ++k) p[k] = __fadd_rn(p[k - 1], __shfl_xor_sync(FULL, p[k - 1], 1 << (k - 1)));
    int off = 0;
#pragma unroll
    for (int k = 4; k >= 0; --k) {
      if (k < hx) {  // left child at this step spans 2^k lanes starting at `off`
        const float left = __shfl_sync(FULL, p[k], off);
        if (!(mass <= left)) {
          mass = __fsub_rn(mass, left);
          off += 1 << k;
        }
      }
    }
    // inside lane `off`
    const float c0 = __shfl_sync(FULL, v.x, off), c1 = __shfl_sync(FULL, v.y, off);
    const float c2 = __shfl_sync(FULL, v.z, off), c3 = __shfl_sync(FULL, v.w, off);
    int sub = 0;
    if (wl == 2) {
      const float a = __fadd_rn(c0, c1);  // == s01 of lane `off`
      float left = c0;
      if (!(mass <= a)) {
        mass = __fsub_rn(mass, a);
        sub = 2, left = c2;
      }
      if (!(mass <= left)) {
        mass = __fsub_rn(mass, left);
        sub += 1;
      }
    } else if (wl == 1) {
      if (!(mass <= c0)) {
        mass = __fsub_rn(mass, c0);
        sub = 1;
      }
    }
    node = base + ((int64_t)off << wl) + sub;
    remaining -= h;
    if (remaining == 0) leaf_p = sub == 0 ? c0 : (sub == 1 ? c1 : (sub == 2 ? c2 : c3));
  }
  // A mass that rounds past the total walks into the empty part of the tree (leaves >= size, priority 0):
  // the newest valid leaf is returned instead, so that index, weight and row always belong together and
  // the index is safe to hand back to update_priorities (a decision of the parity checker too).
  int64_t leaf = node - cap2;
  if (leaf >= size) {
    leaf = size - 1;
    leaf_p = __ldcg(sum + cap2 + leaf);
  }
  leaf_p_out = leaf_p;
  return leaf;
}

__global__ void __launch_bounds__(kSampleWarps * 32)
    prb_sample_kernel(const __grid_constant__ SampleArgs S, const __grid_constant__ ColTable tab) {
  // Before anything is known about the masses: pull the lines the 2nd hop can land on (all of depth 14 =
  // 64 KB) into L2, spread over the grid.  L2 is the point of coherence, so this may run ahead of the
  // predecessor kernel (before griddepcontrol.wait); with a cold L2 one of the three dependent round
  // trips of the descent becomes an L2 hit.
  if (S.prefetch && S.levels > 2 * kHop) {
    const int64_t gthreads = (int64_t)gridDim.x * blockDim.x;
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t lines = (1LL << (2 * kHop)) / 32;
    for (int64_t q = gid; q < lines && q < gid + 4 * gthreads; q += gthreads)
      asm volatile("prefetch.global.L2 [%0];" ::"l"(S.sum + (1LL << (2 * kHop)) + 32 * q));
  }
  pdl_prologue();
  dbg_mark(S.dbg, 0);
  const int lane = threadIdx.x & 31;
  // (warps past the end of the batch shadow its last sample -- same loads, same values stored -- instead of
  // leaving: the sharded path meets at a block barrier)
  const int64_t i_raw = (int64_t)blockIdx.x * kSampleWarps + (threadIdx.x >> 5);
  const int64_t i = i_raw < S.n ? i_raw : S.n - 1;
  constexpr unsigned FULL = 0xffffffffu;
  __shared__ float s_tot[4];  // sharded: the global {sum, min, size}, formed by warp 0 of the CTA

  // everything that does not depend on the descent is requested first: the root, and the line of the
  // first hop (its address does not depend on the mass)
  const float root_sum = __ldcg(S.sum + 1);
  float w_min = __ldcg(S.mn + 1);
  float4 hop0 = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
  if (S.levels > 0) hop0 = hop_load(S.sum, 1, S.levels < kHop ? S.levels : kHop, lane);
  // kPublishOnSample: this rank's root statistics go straight into every rank's mailbox now, so that the
  // NVLink latency hides behind the descent
  if (blockIdx.x == 0 && threadIdx.x < 32) publish_root(S.pub, threadIdx.x, root_sum, w_min);
  // First look at the inbox NOW: with kPublishOnMutation the peers' words landed a step ago, so this look is
  // normally the only one, and its latency (the lines were written over NVLink: an L2 miss, 0.4 us; 2 us when the
  // 4096 warps of a C5-sized batch all ask for the same three words) runs behind the descent.
  // ONE warp per CTA looks (every warp asking -- 4096 warps x 7 peers x 3 words at C5's batch on 8 GPUs -- made
  // the sample kernel 4 us longer: they all hit the same few sectors); the CTA gets the totals through shared memory
  const bool polls = S.mailbox && threadIdx.x < 32 && lane < S.world && lane != S.pub.rank;
  const uint2 *inbox = nullptr;
  uint2 ia = make_uint2(0u, 0u), ib = ia, ic = ia;
  if (polls) {
    inbox = S.mailbox + ((size_t)S.slot * S.world + lane) * 4;
    asm volatile("ld.relaxed.sys.global.v2.u32 {%0,%1}, [%2];" : "=r"(ia.x), "=r"(ia.y) : "l"(inbox) : "memory");
    asm volatile("ld.relaxed.sys.global.v2.u32 {%0,%1}, [%2];" : "=r"(ib.x), "=r"(ib.y) : "l"(inbox + 1) : "memory");
    asm volatile("ld.relaxed.sys.global.v2.u32 {%0,%1}, [%2];" : "=r"(ic.x), "=r"(ic.y) : "l"(inbox + 2) : "memory");
  }
  const float u = S.uniforms ? __ldg(S.uniforms + i) : mrl_uniform01(S.seed, S.ctr + (uint64_t)i);
  // Denominators of the importance weights: (p / w_sum * w_size)^-beta for the sample, (w_min / m_sum * w_size)^-beta
  // for the largest weight.  One buffer: its own root.  Sharded, default: global sum, global minimum,
  // global size.  Sharded with "shard_weights" = 1: the probability the sample was REALLY drawn with -- every rank
  // draws its batch from its own tree, so P(i) = p_i / (world * S_rank) -- and the largest weight over all ranks,
  // (min_q / (world * S_q) * N)^-beta at the rank q where that ratio is smallest (first in rank order).
  float w_sum = root_sum, w_size = (float)S.size, m_sum = root_sum;
  if (S.gstats) {
    float t_sum = 0.0f, t_min = FLT_MAX, best = FLT_MAX, star_min = FLT_MAX, star_sum = 1.0f;
    w_size = 0.0f;
    for (int r = 0; r < S.world; ++r) {
      const float sr = __ldg(S.gstats + 4 * r), m = __ldg(S.gstats + 4 * r + 1);
      t_sum = __fadd_rn(t_sum, sr);
      t_min = m < t_min ? m : t_min;
      w_size = __fadd_rn(w_size, __ldg(S.gstats + 4 * r + 2));
      const float den = __fmul_rn((float)S.world, sr);
      const float ratio = __fdiv_rn(m, den);
      if (sr > 0.0f && ratio < best) best = ratio, star_min = m, star_sum = den;
    }
    if (S.exact_weights) w_sum = __fmul_rn((float)S.world, root_sum), w_min = star_min, m_sum = star_sum;
    else w_sum = t_sum, w_min = t_min, m_sum = t_sum;
  }
  const float seg = __fdiv_rn(root_sum, (float)S.n);
  float mass = __fmul_rn(__fadd_rn(u, (float)i), seg);
  dbg_mark(S.dbg, 1);

  float leaf_p;
  const int64_t leaf = prb_descend(S.sum, S.levels, S.cap2, S.size, mass, hop0, root_sum, lane, leaf_p);
  const int64_t slot = leaf;
  dbg_mark(S.dbg, 3);

  // The row: all 4-byte words of the narrow columns are ONE list, 4 per lane and pass, every load of a pass
  // in flight before the first store (column by column, each load waited for its predecessor's store -- the
  // pointers may alias as far as the compiler knows -- 1.5 us for the four columns of a 40-byte row).  The
  // first pass (512 bytes) is requested here and travels while the weights are computed.
  uint32_t val[4];
  uint32_t *dstp[4];
  auto row_pass_load = [&](int u0) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      int u = u0 + q * 32 + lane;
      dstp[q] = nullptr, val[q] = 0u;
      if (u < S.word_units) {
        int c = 0;
        for (;; ++c) {
          if (!((S.gather_mask >> c) & 1u)) continue;
          const int w = (int)(tab.row_bytes[c] >> 2);
          if (u < w) break;
          u -= w;
        }
        const int64_t rb = tab.row_bytes[c];
        val[q] = *reinterpret_cast<const uint32_t *>(tab.col[c] + slot * rb + 4 * u);
        dstp[q] = reinterpret_cast<uint32_t *>(tab.io[c] + i * rb + 4 * u);
      }
    }
  };
  auto row_pass_store = [&]() {
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (dstp[q]) *dstp[q] = val[q];
  };
  if (S.word_units > 0) {
    row_pass_load(0);
  } else if (S.gather_mask && lane < tab.ncols && ((S.gather_mask >> lane) & 1u)) {
    const uint8_t *src = tab.col[lane] + slot * tab.row_bytes[lane];
    asm volatile("prefetch.global.L1 [%0];" ::"l"(src));
  }
  if (S.mailbox) {
    // (sharded: everything that does not need the peers' statistics goes first -- the longer this rank
    // takes to get to its mailbox, the more skew between the ranks it absorbs)
    if (S.word_units > 0) {
      row_pass_store();
      for (int u0 = 128; u0 < S.word_units; u0 += 128) {
        row_pass_load(u0);
        row_pass_store();
      }
    }
    // Lane r polls rank r's three words in the LOCAL mailbox until they carry the expected tag (kPublishOn-
    // Mutation: they were stored by the peer's previous mutating kernel, a whole step ago; kPublishOnSample: by
    // the peer's matching sample kernel); this rank's own entry comes from its tree.  Then the totals are
    // formed in rank order (as the oracle does).  The poll is bounded: ADVICE r1 -- a rank that never makes
    // the matching call must not hang its peers' GPUs.
    if (threadIdx.x < 32) {
      float ps = root_sum, pm = w_min, pz = (float)S.size;
      const bool timing = S.diag && blockIdx.x == 0 && threadIdx.x == 0;
      long long c0 = 0;
      if (timing) c0 = clock64();
      bool first_miss = false;
      if (polls) {
        uint2 a = ia, b = ib, c = ic;
        long long t0 = 0;
        unsigned int spins = 0;
        while (!(a.y == S.expect && b.y == S.expect && c.y == S.expect)) {
          first_miss = true;
          if ((++spins & 1023u) == 0u) {  // the global timer costs ~1 us to read: only on the slow path
            long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (t0 == 0) t0 = now;
            if (now - t0 > S.timeout_ns) {
              if (S.err) *reinterpret_cast<volatile int *>(S.err) = 1 + lane;  // which peer was missing
              if (S.diag && blockIdx.x == 0 && threadIdx.x < 32) atomicAdd(&S.diag->timeouts, 1ull);
              a.x = b.x = c.x = 0x7fc00000u;  // NaN: the weights of this call are void
              break;
            }
          }
          asm volatile("ld.relaxed.sys.global.v2.u32 {%0,%1}, [%2];" : "=r"(a.x), "=r"(a.y) : "l"(inbox) : "memory");
          asm volatile("ld.relaxed.sys.global.v2.u32 {%0,%1}, [%2];" : "=r"(b.x), "=r"(b.y) : "l"(inbox + 1) : "memory");
          asm volatile("ld.relaxed.sys.global.v2.u32 {%0,%1}, [%2];" : "=r"(c.x), "=r"(c.y) : "l"(inbox + 2) : "memory");
        }
        ps = __uint_as_float(a.x), pm = __uint_as_float(b.x), pz = __uint_as_float(c.x);
      }
      if (S.diag && blockIdx.x == 0 && threadIdx.x < 32) {
        const bool any_miss = __any_sync(FULL, first_miss);
        if (timing) {
          atomicAdd(&S.diag->wait_cycles, (unsigned long long)(clock64() - c0));
          atomicAdd(&S.diag->polled_calls, 1ull);
          if (any_miss) atomicAdd(&S.diag->waited_calls, 1ull);
        }
      }
      float t_sum = 0.0f, t_min = FLT_MAX, t_size = 0.0f, best = FLT_MAX, star_min = FLT_MAX, star_sum = 1.0f;
      for (int r = 0; r < S.world; ++r) {  // rank order, as the oracle
        const float sr = __shfl_sync(FULL, ps, r), m = __shfl_sync(FULL, pm, r);
        t_sum = __fadd_rn(t_sum, sr);
        t_min = m < t_min ? m : t_min;
        t_size = __fadd_rn(t_size, __shfl_sync(FULL, pz, r));
        const float den = __fmul_rn((float)S.world, sr);
        const float ratio = __fdiv_rn(m, den);
        if (sr > 0.0f && ratio < best) best = ratio, star_min = m, star_sum = den;
      }
      if (lane == 0) {
        if (S.exact_weights) s_tot[0] = star_sum, s_tot[1] = star_min;
        else s_tot[0] = t_sum, s_tot[1] = t_min;
        s_tot[2] = t_size;
      }
    }
    __syncthreads();
    m_sum = s_tot[0], w_min = s_tot[1], w_size = s_tot[2];
    w_sum = S.exact_weights ? __fmul_rn((float)S.world, root_sum) : m_sum;
  }
  // lanes 0 and 1 evaluate the two powers side by side: (p/sum*size)^-beta and (min/sum*size)^-beta
  {
    const float x = lane == 0 ? leaf_p : w_min;
    float r = 0.0f;
    const float beta = S.beta_dev ? __ldg(S.beta_dev) : S.beta;
    if (lane < 2) r = mrl_is_weight(x, lane == 0 ? w_sum : m_sum, w_size, beta);
    const float max_w = __shfl_sync(FULL, r, 1);
    if (lane == 0) {
      S.indices[i] = leaf;
      S.weights[i] = __fdiv_rn(r, max_w);
    }
  }
  dbg_mark(S.dbg, 4);
  if (S.word_units > 0) {
    if (!S.mailbox) {
      row_pass_store();
      for (int u0 = 128; u0 < S.word_units; u0 += 128) {
        row_pass_load(u0);
        row_pass_store();
      }
    }
  } else if (S.gather_mask) {
    for (int c = 0; c < tab.ncols; ++c) {
      if (!((S.gather_mask >> c) & 1u)) continue;
      const int64_t rb = tab.row_bytes[c];
      const uint8_t *src = tab.col[c] + slot * rb;
      uint8_t *dst = tab.io[c] + i * rb;
      const int v = tab.vec[c];
      if (v >= 4) {
        for (int64_t o = lane * 4; o < rb; o += 128)
          *reinterpret_cast<uint32_t *>(dst + o) = *reinterpret_cast<const uint32_t *>(src + o);
      } else {
        for (int64_t o = lane; o < rb; o += 32) dst[o] = src[o];
      }
    }
  }
  dbg_mark(S.dbg, 5);
}
// ---- sample with the wide-row gather fused in (Atari-shaped buffers) -----------------------------------------
// The Atari-shaped step used to be three dependent launches (sample 9.9 us, TMA gather 14.3 us, update 12.3 us by
// events after a flush, 29.5 us with PDL overlap).  Here the gather's DMA ring and the descent share ONE kernel:
// the batch's copy items -- (sample, wide column, <= 8 KB chunk), ordered by sample -- are cut into gridDim.x
// contiguous ranges, a range touches at most two samples, and the CTA's two warps each run the descent of one of
// them (duplicated between the few CTAs that share a sample: two microseconds of L2-resident loads) and park the
// slot in shared memory; one elected thread then drives the same cp.async.bulk ring as gather_bulk_kernel
// (global -> shared with mbarrier complete_tx, shared -> global in bulk groups).  The copy starts as soon as the
// CTA's own two indices are known, and a launch with its ~4 us of latency is gone.  The CTA whose range holds a
// sample's first item also writes its index, weight and narrow columns.
struct FusedPlan {
  int32_t ncols;                         // wide columns
  int32_t col_id[MRL_MAX_COLS];
  int32_t chunk_bytes[MRL_MAX_COLS];
  int32_t item_off[MRL_MAX_COLS + 1];    // first item of wide column k within a sample's item list
  int32_t items_per_sample;
};
bool g_fused_gather = true;  // option "fused_gather"
bool g_exact_shard_weights = false;  // option "shard_weights"

template <int STAGES>
__global__ void __launch_bounds__(64)
    prb_sample_gather_kernel(const __grid_constant__ SampleArgs S, const __grid_constant__ ColTable tab,
                             const __grid_constant__ FusedPlan plan, int32_t stage_bytes) {
  extern __shared__ __align__(128) uint8_t stage_mem[];
  __shared__ __align__(8) uint64_t full_bar[STAGES];
  __shared__ int64_t s_slot[2];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr unsigned FULL = 0xffffffffu;
  if (S.prefetch && S.levels > 2 * kHop) {  // (as prb_sample_kernel: the lines the 2nd hop can land on)
    const int64_t gthreads = (int64_t)gridDim.x * blockDim.x;
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t lines = (1LL << (2 * kHop)) / 32;
    for (int64_t q = gid; q < lines && q < gid + 4 * gthreads; q += gthreads)
      asm volatile("prefetch.global.L2 [%0];" ::"l"(S.sum + (1LL << (2 * kHop)) + 32 * q));
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) mbar_init(&full_bar[s], 1);
    fence_async_smem();
  }
  pdl_prologue();
  const int ips = plan.items_per_sample;
  const int64_t total = S.n * ips;
  const int64_t lo = total * blockIdx.x / gridDim.x, hi = total * (blockIdx.x + 1) / gridDim.x;
  const int64_t i0 = lo / ips;
  const int64_t i = i0 + warp;  // this warp's sample
  const bool active = hi > lo && i <= (hi - 1) / ips;
  if (active) {
    const float root_sum = __ldcg(S.sum + 1);
    const float w_min = __ldcg(S.mn + 1);
    float4 hop0 = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    if (S.levels > 0) hop0 = hop_load(S.sum, 1, S.levels < kHop ? S.levels : kHop, lane);
    const float u = S.uniforms ? __ldg(S.uniforms + i) : mrl_uniform01(S.seed, S.ctr + (uint64_t)i);
    const float seg = __fdiv_rn(root_sum, (float)S.n);
    const float mass = __fmul_rn(__fadd_rn(u, (float)i), seg);
    float leaf_p;
    const int64_t leaf = prb_descend(S.sum, S.levels, S.cap2, S.size, mass, hop0, root_sum, lane, leaf_p);
    if (lane == 0) s_slot[warp] = leaf;
    if (i * ips >= lo) {  // the sample's first item is ours: index, weight, narrow columns
      const float x = lane == 0 ? leaf_p : w_min;
      float r = 0.0f;
      const float beta = S.beta_dev ? __ldg(S.beta_dev) : S.beta;
      if (lane < 2) r = mrl_is_weight(x, root_sum, (float)S.size, beta);
      const float max_w = __shfl_sync(FULL, r, 1);
      if (lane == 0) {
        S.indices[i] = leaf;
        S.weights[i] = __fdiv_rn(r, max_w);
      }
      for (int c = 0; c < tab.ncols; ++c) {
        if (!((S.gather_mask >> c) & 1u)) continue;
        const int64_t rb = tab.row_bytes[c];
        const uint8_t *src = tab.col[c] + leaf * rb;
        uint8_t *dst = tab.io[c] + i * rb;
        if (tab.vec[c] >= 4) {
          for (int64_t o = lane * 4; o < rb; o += 128)
            *reinterpret_cast<uint32_t *>(dst + o) = *reinterpret_cast<const uint32_t *>(src + o);
        } else {
          for (int64_t o = lane; o < rb; o += 32) dst[o] = src[o];
        }
      }
    }
  }
  __syncthreads();
  if (threadIdx.x != 0 || hi <= lo) return;

  // ---- the DMA ring over items [lo, hi) ---------------------------------------------------------------------
  const int64_t nmine = hi - lo;
  uint8_t *dst_of[STAGES];
  uint32_t bytes_of[STAGES];
#pragma unroll 1
  for (int64_t k = 0; k < nmine + STAGES - 1; ++k) {
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
    if (k < nmine) {  // fill: stage k % STAGES was last read by the store of item k - STAGES
      const int64_t item = lo + k;
      const int64_t si = item / ips;
      const int r = (int)(item - si * ips);
      int kc = 0;
      while (kc + 1 < plan.ncols && r >= plan.item_off[kc + 1]) ++kc;
      const int c = plan.col_id[kc];
      const int64_t rb = tab.row_bytes[c];
      const int64_t off = (int64_t)(r - plan.item_off[kc]) * plan.chunk_bytes[kc];
      const int64_t rem = rb - off;
      const uint32_t bytes = (uint32_t)(rem < plan.chunk_bytes[kc] ? rem : plan.chunk_bytes[kc]);
      const int64_t slot = s_slot[si - i0];
      const int s = (int)(k % STAGES);
      if (k >= STAGES) bulk_wait_read<1>();
#pragma unroll
      for (int q = 0; q < STAGES; ++q)
        if (q == s) {
          dst_of[q] = tab.io[c] + si * rb + off;
          bytes_of[q] = bytes;
        }
      mbar_expect_tx(&full_bar[s], bytes);
      bulk_g2s(stage_mem + (size_t)s * stage_bytes, tab.col[c] + slot * rb + off, bytes, &full_bar[s]);
    }
  }
  bulk_wait_all();
}

template <int STAGES>
static int launch_fused(const SampleArgs &S, const ColTable &tab, const FusedPlan &plan, int32_t stage_bytes,
                        int64_t blocks, cudaStream_t st) {
  const size_t smem = (size_t)stage_bytes * STAGES;
  if (first_use_on_device((const void *)prb_sample_gather_kernel<STAGES>)) {
    cudaFuncAttributes fa;
    MRL_CUDA(cudaFuncGetAttributes(&fa, prb_sample_gather_kernel<STAGES>));
    MRL_CUDA(cudaFuncSetAttribute(prb_sample_gather_kernel<STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  227 * 1024 - (int)fa.sharedSizeBytes));
  }
  MRL_CUDA(launch_pdl(prb_sample_gather_kernel<STAGES>, dim3((unsigned)blocks), dim3(64), smem, st, S, tab, plan,
                      stage_bytes));
  return MRL_OK;
}

// -> MRL_OK after launching the fused kernel, kNotFused when the shape does not fit it (the caller then runs
// prb_sample_kernel + gather_rows)
constexpr int kNotFused = -1001;
static int try_fused_sample(const SampleArgs &S, const ColTable &tab, uint32_t wide, cudaStream_t st) {
  if (!g_fused_gather || g_gather_path == 1 || S.levels <= 0) return kNotFused;
  FusedPlan plan;
  plan.ncols = 0;
  plan.item_off[0] = 0;
  int32_t stage_bytes = 0;
  for (int c = 0; c < tab.ncols; ++c) {
    if (!((wide >> c) & 1u)) continue;
    const int64_t rb = tab.row_bytes[c];
    if (tab.vec[c] != 16 || !(g_gather_path == 2 || rb >= g_bulk_min_row)) return kNotFused;  // (as gather_rows)
    int64_t nch = (rb + g_bulk_chunk - 1) / g_bulk_chunk;
    const int64_t chunk = ((rb + nch - 1) / nch + 15) / 16 * 16;
    nch = (rb + chunk - 1) / chunk;
    const int k = plan.ncols++;
    plan.col_id[k] = c;
    plan.chunk_bytes[k] = (int32_t)chunk;
    plan.item_off[k + 1] = plan.item_off[k] + (int32_t)nch;
    if (chunk > stage_bytes) stage_bytes = (int32_t)chunk;
  }
  if (plan.ncols == 0) return kNotFused;
  plan.items_per_sample = plan.item_off[plan.ncols];
  const int64_t total = S.n * plan.items_per_sample;
  int64_t blocks = (int64_t)sm_count() * g_bulk_ctas_per_sm;
  if (blocks > total) blocks = total;
  // a CTA's contiguous range must touch at most two samples (one descent per warp)
  if ((total + blocks - 1) / blocks > plan.items_per_sample) return kNotFused;
  stage_bytes = (stage_bytes + 127) / 128 * 128;
  int stages = (int)g_bulk_stages;
  while (stages > 2 && (size_t)stage_bytes * stages > 200 * 1024) stages = stages == 12 ? 8 : stages / 2;
  if (stages == 12) return launch_fused<12>(S, tab, plan, stage_bytes, blocks, st);
  if (stages == 8) return launch_fused<8>(S, tab, plan, stage_bytes, blocks, st);
  if (stages == 4) return launch_fused<4>(S, tab, plan, stage_bytes, blocks, st);
  return launch_fused<2>(S, tab, plan, stage_bytes, blocks, st);
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
  long long *dbg;
  int32_t prefetch;  // option "tree_prefetch"
  PublishArgs pub;   // sharded buffer, kPublishOnMutation: the kernel that finishes the root publishes it
};

__device__ __forceinline__ void recompute_node(float *sum, float *mn, int64_t node) {
  const float a = __ldcg(sum + 2 * node), b = __ldcg(sum + 2 * node + 1);
  const float ma = __ldcg(mn + 2 * node), mb = __ldcg(mn + 2 * node + 1);
  __stcg(sum + node, __fadd_rn(a, b));
  __stcg(mn + node, ma < mb ? ma : mb);
}

__device__ __noinline__ float prb_pow_fn(float priority, float alpha) {
  return mrl_priority_pow(priority, alpha);
}

__global__ void __launch_bounds__(1024) prb_update_kernel(const __grid_constant__ UpdateArgs U) {
  __shared__ float ssum[2 << kTopLevels];
  __shared__ float smin[2 << kTopLevels];
  const int tid = threadIdx.x, nthr = blockDim.x;
  dbg_mark(U.dbg, 0);

  // 0) arbitration: the highest batch position of a leaf wins; max over ALL priorities (also the
  //    overwritten ones) goes to max_priority, as in the sequential loop
  float local_max = 0.0f;
  for (int64_t k = tid; k < U.n; k += nthr) {
    const int64_t leaf = __ldg(U.indices + k);
    atomicMax(U.tag + leaf, (unsigned long long)((U.epoch << 32) | (uint64_t)k));
    const float p = prb_pow_fn(__ldg(U.priorities + k), U.alpha);
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
  dbg_mark(U.dbg, 1);

  // 1) winners write their leaf
  for (int64_t k = tid; k < U.n; k += nthr) {
    const int64_t leaf = __ldg(U.indices + k);
    if (__ldcg(U.tag + leaf) == (unsigned long long)((U.epoch << 32) | (uint64_t)k)) {
      const float p = prb_pow_fn(__ldg(U.priorities + k), U.alpha);
      __stcg(U.sum + U.cap2 + leaf, p);
      __stcg(U.mn + U.cap2 + leaf, p);
    }
  }
  __syncthreads();
  dbg_mark(U.dbg, 2);

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

  dbg_mark(U.dbg, 3);
  // 3) top `split` levels rebuilt in shared memory from the 2^split nodes of level `split`
  const int width = 1 << split;
  for (int i = tid; i < width; i += nthr) {
    ssum[width + i] = __ldcg(U.sum + width + i);
    smin[width + i] = __ldcg(U.mn + width + i);
  }
  __syncthreads();
  dbg_mark(U.dbg, 4);
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
  dbg_mark(U.dbg, 5);
  for (int i = tid + 1; i < width; i += nthr) {
    __stcg(U.sum + i, ssum[i]);
    __stcg(U.mn + i, smin[i]);
  }
  if (tid < 32) publish_root(U.pub, tid, ssum[1], smin[1]);
  dbg_mark(U.dbg, 6);
}

// ---- update, low-latency variant (n <= kWalkMax) -------------------------------------------
// Same result as prb_update_kernel, organised around DEPENDENT GLOBAL round trips.  Measured with
// %globaltimer: a level-synchronous step through global memory (store, barrier, load) costs ~1.3 us
// per tree level whether L2 is warm or cold, i.e. ~12 us for the 9 lower levels of a 2^20-leaf tree.
// Here every thread walks its own leaf-to-root path in registers and meets the other paths only in
// shared memory:
//   * all siblings along the path are loaded up front in one round trip (values of nodes no path of
//     this batch goes through cannot change);
//   * per level the thread enters its node into a shared-memory hash table (key = heap index), one
//     block barrier, then looks its sibling up: a hit means another path of the batch owns the sibling
//     and its fresh value is taken from shared memory, a miss means the prefetched value is current;
//   * two tables alternate between levels (entries carry the level as a tag, stale entries count as
//     free), so one barrier per level suffices;
//   * duplicates of a leaf are arbitrated at the leaf level (highest batch position wins);
//   * node values are stored to global memory as the walk passes them, nothing is read back;
//   * the top kTopLevels levels are rebuilt in shared memory from the 2^kTopLevels nodes below them.
// Parents are fl(left+right) / min(left,right) of their two children exactly as in the sequential
// oracle, so the tree is bit-identical.
constexpr int kWalkMax = 1 << 16;  // larger batches use the global-tag kernel
constexpr int kWalkThreads = 512;
constexpr int kWalkLevels = 12;  // lower levels handled by the walk (leaf capacity <= 2^(11+12))

struct WalkTab {
  unsigned long long *keys;
  float2 *vals;
  int mask, bits;
};
__device__ __forceinline__ uint32_t walk_hash(uint32_t node, int bits) {
  return (node * 2654435761u) >> (32 - bits);
}
// enters (tag, node) if absent; returns its slot.  Slots whose tag differs are free.
__device__ __forceinline__ int walk_insert(const WalkTab &t, uint32_t tag, uint32_t node) {
  const unsigned long long key = ((unsigned long long)tag << 32) | node;
  uint32_t h = walk_hash(node, t.bits);
  while (true) {
    unsigned long long cur = *reinterpret_cast<volatile unsigned long long *>(t.keys + h);
    while ((uint32_t)(cur >> 32) != tag) {
      const unsigned long long old = atomicCAS(t.keys + h, cur, key);
      if (old == cur) return (int)h;
      cur = old;
    }
    if (cur == key) return (int)h;
    h = (h + 1) & t.mask;
  }
}
__device__ __forceinline__ int walk_find(const WalkTab &t, uint32_t tag, uint32_t node) {
  const unsigned long long key = ((unsigned long long)tag << 32) | node;
  uint32_t h = walk_hash(node, t.bits);
  while (true) {
    const unsigned long long cur = t.keys[h];
    if ((uint32_t)(cur >> 32) != tag) return -1;
    if (cur == key) return (int)h;
    h = (h + 1) & t.mask;
  }
}

// Code size matters as much as round trips here: a one-CTA kernel fetches every instruction line cold
// (the 44 KB fully unrolled version of this kernel spent ~30 us on instruction fetch alone), so the
// level loop is a real loop over shared-memory arrays and p**alpha is a single out-of-line call.
__device__ __noinline__ float prb_pow_noinline(float priority, float alpha) {
  return mrl_priority_pow(priority, alpha);
}

__global__ void __launch_bounds__(kWalkThreads)
    prb_update_walk_kernel(const __grid_constant__ UpdateArgs U, int slot_bits) {
  __shared__ float ssum[2 << kTopLevels];
  __shared__ float smin[2 << kTopLevels];
  extern __shared__ __align__(16) unsigned char upd_dyn[];
  const int nslots = 1 << slot_bits;
  const int tid = threadIdx.x, nthr = blockDim.x;
  WalkTab tab[2];
  tab[0].keys = reinterpret_cast<unsigned long long *>(upd_dyn);
  tab[1].keys = tab[0].keys + nslots;
  tab[0].vals = reinterpret_cast<float2 *>(tab[1].keys + nslots);
  tab[1].vals = tab[0].vals + nslots;
  unsigned int *pos = reinterpret_cast<unsigned int *>(tab[1].vals + nslots);
  float2 *sib = reinterpret_cast<float2 *>(pos + nslots);  // [kWalkLevels][nthr] prefetched siblings
  tab[0].mask = tab[1].mask = nslots - 1;
  tab[0].bits = tab[1].bits = slot_bits;
  const int split = U.levels < kTopLevels ? U.levels : kTopLevels;
  const int nlev = U.levels - split;  // levels walked: nodes at level levels-d, d = 0 .. nlev-1
  const int width = 1 << split;
  dbg_mark(U.dbg, 8);

  for (int i = tid; i < 2 * nslots; i += nthr) tab[0].keys[i] = 0ull;
  for (int i = tid; i < nslots; i += nthr) pos[i] = 0u;
  // inputs of the top rebuild (patched by the walks below)
  for (int i = tid; i < width; i += nthr) {
    ssum[width + i] = __ldcg(U.sum + width + i);
    smin[width + i] = __ldcg(U.mn + width + i);
  }
  float local_max = 0.0f;

  // batches larger than the block are applied in order, blockDim indices at a time
  uint32_t pass = 0;
  for (int64_t base = 0; base < U.n; base += nthr, ++pass) {
    const int64_t k = base + tid;
    const bool on = k < U.n;
    // round trip 1: index and priority
    const int64_t leaf = on ? __ldg(U.indices + k) : 0;
    const float prio = on ? __ldg(U.priorities + k) : 1.0f;
    const uint32_t leaf_node = (uint32_t)(U.cap2 + leaf);
    // round trip 2: every sibling along the path, parked in shared memory
    if (on) {
#pragma unroll 4
      for (int d = 0; d < nlev; ++d) {
        const uint32_t sn = (leaf_node >> d) ^ 1u;
        sib[d * nthr + tid] = make_float2(__ldcg(U.sum + sn), __ldcg(U.mn + sn));
      }
    }
    const float p = prb_pow_noinline(prio, U.alpha);
    if (on) local_max = p > local_max ? p : local_max;
    __syncthreads();  // tables cleared (first pass) / previous pass done with them
    if (pass == 0) dbg_mark(U.dbg, 9);

    // leaf level: duplicates of a leaf -> highest batch position wins
    const uint32_t pass_tag = pass << 8;
    int slot = 0;
    if (on) {
      slot = walk_insert(tab[0], pass_tag | ((uint32_t)U.levels + 1u), leaf_node);
      atomicMax(pos + slot, (pass << 16) | ((uint32_t)tid + 1u));
    }
    __syncthreads();
    bool leaf_win = false;
    if (on) {
      leaf_win = pos[slot] == ((pass << 16) | ((uint32_t)tid + 1u));
      if (leaf_win) {
        tab[0].vals[slot] = make_float2(p, p);
        __stcg(U.sum + leaf_node, p);
        __stcg(U.mn + leaf_node, p);
      }
    }
    __syncthreads();
    if (pass == 0) dbg_mark(U.dbg, 10);
    float vs = 0.0f, vm = 0.0f;
    if (on) {
      const float2 me = tab[0].vals[slot];  // the winning duplicate's value
      vs = me.x, vm = me.y;
    }
    // walk up: one shared-memory exchange and one barrier per level
#pragma unroll 1
    for (int d = 0; d < nlev; ++d) {
      const WalkTab &t = tab[d & 1];
      const uint32_t tag = pass_tag | ((uint32_t)(U.levels - d) + 1u);
      const uint32_t node = leaf_node >> d;
      if (d > 0) {  // (the leaf level was entered above)
        if (on) {
          const int sl = walk_insert(t, tag, node);
          t.vals[sl] = make_float2(vs, vm);  // all paths through this node agree
          __stcg(U.sum + node, vs);
          __stcg(U.mn + node, vm);
        }
        __syncthreads();
      }
      if (on) {
        const int ss = walk_find(t, tag, node ^ 1u);
        const float2 o = ss >= 0 ? t.vals[ss] : sib[d * nthr + tid];
        vs = __fadd_rn(vs, o.x);
        vm = o.y < vm ? o.y : vm;
      }
    }
    if (pass == 0) dbg_mark(U.dbg, 11);
    // (vs, vm) now belong to the node at level `split`: patch the inputs of the top rebuild
    __syncthreads();
    if (on) {
      const uint32_t node = leaf_node >> nlev;  // in [width, 2*width)
      ssum[node] = vs;
      smin[node] = vm;
      if (nlev > 0) {
        __stcg(U.sum + node, vs);
        __stcg(U.mn + node, vm);
      }
    }
    __syncthreads();  // also orders this pass's global stores before the next pass's sibling loads
  }
#pragma unroll
  for (int off = 16; off; off >>= 1) {
    const float o = __shfl_xor_sync(0xffffffffu, local_max, off);
    local_max = o > local_max ? o : local_max;
  }
  if ((tid & 31) == 0 && local_max > 0.0f)
    atomicMax(reinterpret_cast<int *>(&U.state->max_priority), __float_as_int(local_max));
  dbg_mark(U.dbg, 12);
#pragma unroll 1
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
  dbg_mark(U.dbg, 13);
  for (int i = tid + 1; i < width; i += nthr) {
    __stcg(U.sum + i, ssum[i]);
    __stcg(U.mn + i, smin[i]);
  }
  if (tid < 32) publish_root(U.pub, tid, ssum[1], smin[1]);
  dbg_mark(U.dbg, 14);
}

// ---- update, subtree-partitioned over many SMs ------------------------------------------------
// Measured (scratch/lat_test2.cu): a level of the one-CTA kernel costs 0.5 us with 32 indices and
// 1.0 / 3.3 us with 256 / 1024 -- one SM's path to L2 serialises the scattered 32-byte requests.  So
// the batch is spread over kSubCtas CTAs by ownership of the subtrees below level kTopLevels: the
// lower levels of different subtrees never meet, every CTA scans the index list, keeps the entries
// that fall into its subtrees, loads all siblings of its paths in one round trip and walks the paths
// up in registers; paths that merge exchange values by shuffles (shares of <= 32 entries: one warp) or
// through a small shared-memory list.  The last CTA to finish (acq_rel ticket behind the block
// barrier) rebuilds the top levels.  Parents are fl(left+right)/min of their
// children exactly as in the sequential oracle; duplicates of a leaf resolve to the highest batch
// position; max_priority takes every processed priority.
constexpr int kSubCtas = 128;
constexpr int kSubThreads = 512;
constexpr int kSubMaxLocal = kSubThreads;  // indices scanned per pass = most entries a CTA can keep
constexpr int kSubDenseAt = 128;            // a share larger than this is rebuilt subtree by subtree

// INL: the batch travels in the kernel's own parameter block (InlineBatch, 2 KB) instead of device
// memory -- the *_host entry point then needs no host-to-device copy, no staging buffer and no event
// for a learner-sized batch; every CTA moves it from the constant bank to shared memory first.
constexpr int kInlineMax = 256;
struct InlineBatch {
  int32_t idx[kInlineMax];
  float pr[kInlineMax];
};

// PW: a warp per owned subtree (batches above 1024 entries; compiled apart so that the learner-sized
// batches keep the leaner kernel: 6.6 against 6.9 us at 256 entries)
template <bool INL, bool PW>
__global__ void __launch_bounds__(kSubThreads) prb_update_sub_kernel(const __grid_constant__ UpdateArgs U,
                                                                     const __grid_constant__ InlineBatch IB) {
  __shared__ float ssum[2 << kTopLevels];
  __shared__ float smin[2 << kTopLevels];
  __shared__ int s_raw[4 * kSubMaxLocal];  // sparse: entry lists; dense: per-leaf winner table
  __shared__ int s_cnt;
  __shared__ unsigned int s_dirty;  // bit j: owned subtree blockIdx.x + j*gridDim.x has an entry
  __shared__ unsigned int s_last;
  // one list per owned subtree = per warp (width / gridDim <= 16 subtrees): subtrees never meet below the
  // split, so every warp walks its own subtree's entries independently
  __shared__ int s_cw[PW ? kSubThreads / 32 : 1];
  __shared__ int s_kw[PW ? kSubThreads / 32 : 1][32];
  int *s_k = s_raw;                                                 // batch position of each entry
  uint32_t *s_node = reinterpret_cast<uint32_t *>(s_raw + kSubMaxLocal);
  float2 *s_val = reinterpret_cast<float2 *>(s_raw + 2 * kSubMaxLocal);
  __shared__ int32_t s_in_idx[INL ? kInlineMax : 1];
  __shared__ float s_in_pr[INL ? kInlineMax : 1];
  const int tid = threadIdx.x, nthr = blockDim.x;
  if (U.prefetch && tid == 0) {
    // the top rebuild at the end reads all of depth min(levels, 11) (2 x 8 KB): have it in L2 by then.
    // L2 is the point of coherence, so this may run ahead of the predecessor (before griddepcontrol.wait).
    const int sp = U.levels < kTopLevels ? U.levels : kTopLevels;
    const int lines = ((1 << sp) + 31) / 32;  // per array
    for (int q = blockIdx.x; q < 2 * lines; q += gridDim.x) {
      const float *line = (q < lines ? U.sum : U.mn) + (1 << sp) + 32 * (q < lines ? q : q - lines);
      asm volatile("prefetch.global.L2 [%0];" ::"l"(line));
    }
  }
  pdl_prologue();
  if (INL) {
    for (int i = tid; i < (int)U.n; i += nthr) s_in_idx[i] = IB.idx[i], s_in_pr[i] = IB.pr[i];
    __syncthreads();
  }
  auto ld_idx = [&](int64_t k) -> int64_t { return INL ? (int64_t)s_in_idx[k] : __ldg(U.indices + k); };
  auto ld_pr = [&](int64_t k) -> float { return INL ? s_in_pr[k] : __ldg(U.priorities + k); };
  const int G = (int)gridDim.x;  // a power of two (128..512): the ownership tests are masks
  // max_priority takes every processed priority (also those a later duplicate overwrites); every
  // entry is owned by exactly one CTA, which evaluates p**alpha for it anyway
  float local_max = 0.0f;
  const int split = U.levels < kTopLevels ? U.levels : kTopLevels;
  const int nlev = U.levels - split;
  const int width = 1 << split;
  if (blockIdx.x == 0) dbg_mark_gt(U.dbg, 8);

  // this CTA owns the subtrees (level-`split` nodes) congruent to blockIdx.x mod gridDim.x: a
  // contiguous run of leaves is spread over many CTAs
  if (tid == 0) s_cnt = 0, s_dirty = 0u;
  if (PW && tid < kSubThreads / 32) s_cw[tid] = 0;
  __syncthreads();
  const bool per_warp = PW && width % G == 0 && width / G <= nthr / 32 && nthr == kSubThreads;
  {
    // one scan: count this CTA's share, note its dirty subtrees, and optimistically build the entry lists
    unsigned int dirty = 0u;
    for (int64_t k = tid; k < U.n; k += nthr) {
      const int top = (int)(ld_idx(k) >> nlev);
      if ((top & (G - 1)) == (int)blockIdx.x) {
        dirty |= 1u << ((top / G) & 31);
        const int pos = atomicAdd(&s_cnt, 1);
        if (pos < kSubMaxLocal) s_k[pos] = (int)k;
        if (PW && per_warp) {
          const int w = (top / G) & (kSubThreads / 32 - 1);
          const int pw = atomicAdd(&s_cw[w], 1);
          if (pw < 32) s_kw[w][pw] = (int)k;
        }
      }
    }
    if (dirty) atomicOr(&s_dirty, dirty);
  }
  __syncthreads();
  // a subtree with more than 32 entries sends the whole CTA down the general paths (the counts are read
  // AFTER the barrier: a __syncthreads_or would evaluate its predicate before the other threads' atomics)
  bool crowded = false;
  if (PW && per_warp) {
#pragma unroll
    for (int w = 0; w < kSubThreads / 32; ++w) crowded = crowded || s_cw[w] > 32;
  }
  const int total_mine = s_cnt;
  const unsigned int dirty_mask = s_dirty;
  const bool warp_mode = per_warp && !crowded && total_mine > 0;
  const int wid = tid >> 5, lane = tid & 31;
  const int *my_list = PW && warp_mode ? s_kw[wid] : s_k;
  const int my_m = PW && warp_mode ? s_cw[wid] : (wid == 0 && total_mine <= 32 ? total_mine : 0);
  __syncthreads();
  if (blockIdx.x == 0) dbg_mark_gt(U.dbg, 9);

  const int S = 1 << nlev;  // leaves per subtree
  const bool dense = total_mine > kSubDenseAt && 2 * S <= (2 << kTopLevels) && S <= 4 * kSubMaxLocal &&
                     (width + G - 1) / G <= 32;
  if (__builtin_expect(warp_mode || (total_mine > 0 && total_mine <= 32), 1)) {
    // ---- the normal case: a warp per owned subtree (or one warp for the whole share when the CTA owns
    // more subtrees than it has warps), no barriers, no shared-memory exchange -------------------------
    // Lanes whose path nodes are siblings find each other by comparing against every entry of the list
    // (shuffles); duplicates of a leaf resolve to the highest batch position the same way.
    if (my_m > 0) {
      constexpr unsigned FULL = 0xffffffffu;
      const bool on = lane < my_m;
      const int k = on ? my_list[lane] : -1;
      const int64_t leaf = on ? ld_idx(k) : 0;
      const uint32_t leaf_node = (uint32_t)(U.cap2 + leaf);
      float p = 0.0f;
      float sib_s[kWalkLevels], sib_m[kWalkLevels];
      // The sibling loads must be IN FLIGHT while p**alpha is evaluated.  Behind a call (prb_pow_fn) the
      // compiler sinks them to their uses in the walk -- values that live across a call would have to be
      // spilled first, i.e. waited for -- and the walk becomes nine dependent L2 round trips (measured
      // 1.5 us for the walk).  So here the loads are volatile and the power is inlined.
#pragma unroll
      for (int d = 0; d < kWalkLevels; ++d) {
        sib_s[d] = 0.0f, sib_m[d] = FLT_MAX;
        if (on && d < nlev) {
          const uint32_t sn = (leaf_node >> d) ^ 1u;
          asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(sib_s[d]) : "l"(U.sum + sn) : "memory");
          asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(sib_m[d]) : "l"(U.mn + sn) : "memory");
        }
      }
      if (on) {
        p = mrl_priority_pow(ld_pr(k), U.alpha);
        local_max = p > local_max ? p : local_max;
      }
      // Who meets whom depends on the leaf numbers alone, so it is worked out from ONE pass over the list's
      // m <= 32 entries (a handful: n entries over 2048 subtrees), before any value is needed: the
      // paths of two leaves are siblings exactly at the level of the highest bit in which the leaf numbers
      // differ; equal leaves are duplicates, of which the highest batch position provides the value.
      // partner lane + 32 of level d sits in bits 6d..6d+5 of `meet`.  (match.any did the search in one
      // instruction per level, but costs ~0.3 us when all 32 keys differ; a shuffle loop per level still
      // cost 0.16 us per level.)
      const int m = my_m;
      int win_k = k, win_lane = lane;
      unsigned long long meet = 0ull;
      for (int j = 0; j < m; ++j) {
        const uint32_t lj = __shfl_sync(FULL, leaf_node, j);
        const int kj = __shfl_sync(FULL, k, j);
        const uint32_t x = lj ^ leaf_node;
        if (on) {
          if (x == 0u) {
            if (kj > win_k) win_k = kj, win_lane = j;
          } else {
            const int hb = 31 - __clz(x);
            if (hb < nlev) meet = (meet & ~(63ull << (6 * hb))) | ((unsigned long long)(32 | j) << (6 * hb));
          }
        }
      }
      float vs = __shfl_sync(FULL, p, win_lane), vm = vs;
      if (on && k == win_k) {
        __stcg(U.sum + leaf_node, vs);
        __stcg(U.mn + leaf_node, vm);
      }
      if (blockIdx.x == 0) dbg_mark_gt(U.dbg, 10);
      // the walk itself is branch-free register work (two shuffles, an add and a min per level); the
      // nodes are stored afterwards, all at once
      float lv_s[kWalkLevels + 1], lv_m[kWalkLevels + 1];
#pragma unroll
      for (int d = 0; d < kWalkLevels; ++d) {
        lv_s[d] = vs, lv_m[d] = vm;  // value of the path's node at level d (d = 0: the leaf)
        if (d < nlev) {
          // entries whose paths have merged carry identical values: any lane holding the sibling will do
          const int e = (int)(meet >> (6 * d)) & 63;
          const int src = e & 32 ? e & 31 : lane;
          const float o_s = __shfl_sync(FULL, vs, src), o_m = __shfl_sync(FULL, vm, src);
          const float os = e & 32 ? o_s : sib_s[d], om = e & 32 ? o_m : sib_m[d];
          vs = __fadd_rn(vs, os);
          vm = om < vm ? om : vm;
        }
      }
      lv_s[kWalkLevels] = vs, lv_m[kWalkLevels] = vm;
      if (on) {
#pragma unroll
        for (int d = 1; d <= kWalkLevels; ++d) {
          if (d <= nlev) {
            const uint32_t node = leaf_node >> d;
            __stcg(U.sum + node, lv_s[d]);
            __stcg(U.mn + node, lv_m[d]);
          }
        }
      }
    }
  } else if (dense) {
    // ---- dense share: rebuild every dirty subtree completely in shared memory -------------------
    // (coalesced loads of all its leaves instead of scattered sibling loads; cost independent of how
    // many of its leaves the batch touches)
    int *win = s_raw;  // per leaf of the subtree: highest batch position + 1
    for (int jsub = 0; jsub * G + (int)blockIdx.x < width; ++jsub) {
      if (!((dirty_mask >> (jsub & 31)) & 1u)) continue;
      const int sub = jsub * G + (int)blockIdx.x;
      const int64_t leaf0 = (int64_t)sub << nlev;
      for (int i = tid; i < S; i += nthr) win[i] = 0;
      __syncthreads();
      for (int64_t k = tid; k < U.n; k += nthr) {
        const int64_t leaf = ld_idx(k);
        if ((leaf >> nlev) == sub) {
          atomicMax(&win[leaf - leaf0], (int)k + 1);
          const float p = prb_pow_fn(ld_pr(k), U.alpha);
          local_max = p > local_max ? p : local_max;
        }
      }
      __syncthreads();
      for (int i = tid; i < S; i += nthr) {
        const int wk = win[i];
        float vs, vm;
        if (wk) {
          vs = vm = prb_pow_fn(ld_pr(wk - 1), U.alpha);
          __stcg(U.sum + U.cap2 + leaf0 + i, vs);
          __stcg(U.mn + U.cap2 + leaf0 + i, vm);
        } else {
          vs = __ldcg(U.sum + U.cap2 + leaf0 + i);
          vm = __ldcg(U.mn + U.cap2 + leaf0 + i);
        }
        ssum[S + i] = vs;
        smin[S + i] = vm;
      }
      __syncthreads();
#pragma unroll 1
      for (int level = nlev - 1; level >= 0; --level) {
        const int w = 1 << level;
        for (int i = tid; i < w; i += nthr) {
          const int node = w + i;
          ssum[node] = __fadd_rn(ssum[2 * node], ssum[2 * node + 1]);
          const float ma = smin[2 * node], mb = smin[2 * node + 1];
          smin[node] = ma < mb ? ma : mb;
        }
        __syncthreads();
      }
      for (int jn = tid + 1; jn < S; jn += nthr) {  // local heap node -> global heap node
        const int l = 31 - __clz(jn);
        const int64_t g = ((int64_t)(width + sub) << l) + (jn - (1 << l));
        __stcg(U.sum + g, ssum[jn]);
        __stcg(U.mn + g, smin[jn]);
      }
      __syncthreads();
    }
  } else if (total_mine > 0) {
    // ---- sparse share: one thread per entry walks its path in registers ---------------------------
    // single pass when the share fits the list, else passes of blockDim indices (batch order kept)
    const bool one_pass = total_mine <= kSubMaxLocal;  // the list built by the scan above is complete
    const int64_t span = one_pass ? U.n : kSubMaxLocal;
    for (int64_t base = 0; base < U.n; base += span) {
      const int64_t end = base + span < U.n ? base + span : U.n;
      if (!one_pass) {
        if (tid == 0) s_cnt = 0;
        __syncthreads();
        for (int64_t k = base + tid; k < end; k += nthr) {
          const int top = (int)(ld_idx(k) >> nlev);
          if ((top & (G - 1)) == (int)blockIdx.x) s_k[atomicAdd(&s_cnt, 1)] = (int)(k - base);
        }
        __syncthreads();
      }
      const int m = s_cnt;  // <= kSubMaxLocal == blockDim: one thread per local entry
      const int j = tid;
      const bool on = j < m;
      const int64_t k = on ? base + s_k[j] : 0;
      const int64_t leaf = on ? ld_idx(k) : 0;
      const uint32_t leaf_node = (uint32_t)(U.cap2 + leaf);
      float p = 0.0f;
      float sib_s[kWalkLevels], sib_m[kWalkLevels];
      if (on) {
#pragma unroll
        for (int d = 0; d < kWalkLevels; ++d) {
          sib_s[d] = 0.0f, sib_m[d] = FLT_MAX;
          if (d < nlev) {
            const uint32_t sn = (leaf_node >> d) ^ 1u;
            sib_s[d] = __ldcg(U.sum + sn);
            sib_m[d] = __ldcg(U.mn + sn);
          }
        }
        p = prb_pow_fn(ld_pr(k), U.alpha);
        local_max = p > local_max ? p : local_max;
      }
      // leaf level: among entries of the same leaf the highest batch position provides the value
      if (on) {
        s_node[j] = leaf_node;
        s_val[j] = make_float2(p, p);
      }
      __syncthreads();
      float vs = p, vm = p;
      if (on) {
        int best = s_k[j];
        int best_i = j;
        for (int i = 0; i < m; ++i)
          if (s_node[i] == leaf_node && s_k[i] > best) best = s_k[i], best_i = i;
        const float2 w = s_val[best_i];
        vs = w.x, vm = w.y;
        if (best_i == j) {
          __stcg(U.sum + leaf_node, p);
          __stcg(U.mn + leaf_node, p);
        }
      }
      __syncthreads();
      if (on) s_val[j] = make_float2(vs, vm);  // every entry of a leaf now carries the winner's value
      __syncthreads();
      if (blockIdx.x == 0 && base == 0) dbg_mark_gt(U.dbg, 10);
      // walk up; paths that merge exchange values through the shared list
#pragma unroll
      for (int d = 0; d < kWalkLevels; ++d) {
        if (d < nlev) {
          const uint32_t node = leaf_node >> d;
          if (d > 0) {
            if (on) {
              s_node[j] = node;
              s_val[j] = make_float2(vs, vm);
              __stcg(U.sum + node, vs);
              __stcg(U.mn + node, vm);
            }
            __syncthreads();
          }
          if (on) {
            float os = sib_s[d], om = sib_m[d];
            const uint32_t want = node ^ 1u;
            for (int i = 0; i < m; ++i)
              if (s_node[i] == want) {
                const float2 o = s_val[i];
                os = o.x, om = o.y;
                break;
              }
            vs = __fadd_rn(vs, os);
            vm = om < vm ? om : vm;
          }
          __syncthreads();
        }
      }
      if (on && nlev > 0) {
        const uint32_t node = leaf_node >> nlev;
        __stcg(U.sum + node, vs);
        __stcg(U.mn + node, vm);
      }
      __syncthreads();  // this pass's stores precede the next pass's sibling loads (same CTA)
    }
  }
#pragma unroll
  for (int off = 16; off; off >>= 1) {
    const float o = __shfl_xor_sync(0xffffffffu, local_max, off);
    local_max = o > local_max ? o : local_max;
  }
  if ((tid & 31) == 0 && local_max > 0.0f)
    atomicMax(reinterpret_cast<int *>(&U.state->max_priority), __float_as_int(local_max));

  if (blockIdx.x == 0) dbg_mark_gt(U.dbg, 11);
  // ---- the last CTA to get here rebuilds the top levels ------------------------------------------
  // every thread's tree stores happen-before the block barrier; thread 0's acq_rel ticket at gpu scope then
  // publishes them (release, cumulative over the barrier) and, in the last CTA, acquires the other CTAs'.
  // One fence in one thread instead of __threadfence() in all 512 (0.45 us).
  __syncthreads();
  if (tid == 0) {
    unsigned int ticket;
    asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(ticket) : "l"(&U.state->done_ctas) : "memory");
    s_last = ticket == gridDim.x - 1 ? 1u : 0u;
  }
  __syncthreads();
  if (blockIdx.x == 0) dbg_mark_gt(U.dbg, 12);
  if (!s_last) return;
  dbg_mark_gt(U.dbg, 13);
  // (thread 0 acquired with its ticket, the barrier hands that to the CTA, and the loads below bypass L1)
  if (__builtin_expect(width == 4 * kSubThreads && nthr == kSubThreads, 1)) {
    // 2048 level-11 nodes, 4 per thread: two levels in registers, five by warp shuffles, the last
    // four by one warp -- one block barrier instead of eleven, no shared-memory tree
    const float4 a = __ldcg(reinterpret_cast<const float4 *>(U.sum + width) + tid);
    const float4 m = __ldcg(reinterpret_cast<const float4 *>(U.mn + width) + tid);
    dbg_mark_gt(U.dbg, 14);
    const float s0 = __fadd_rn(a.x, a.y), s1 = __fadd_rn(a.z, a.w);
    const float m0 = m.x < m.y ? m.x : m.y, m1 = m.z < m.w ? m.z : m.w;
    __stcg(reinterpret_cast<float2 *>(U.sum + width / 2) + tid, make_float2(s0, s1));
    __stcg(reinterpret_cast<float2 *>(U.mn + width / 2) + tid, make_float2(m0, m1));
    float xs = __fadd_rn(s0, s1), xm = m0 < m1 ? m0 : m1;  // node width/4 + tid
    int node = width / 4 + tid;
    __stcg(U.sum + node, xs);
    __stcg(U.mn + node, xm);
    const int lane = tid & 31;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      const float os = __shfl_xor_sync(0xffffffffu, xs, 1 << k);
      const float om = __shfl_xor_sync(0xffffffffu, xm, 1 << k);
      // left child is the lane with the lower index: keep fl(left + right)
      xs = (lane >> k) & 1 ? __fadd_rn(os, xs) : __fadd_rn(xs, os);
      xm = om < xm ? om : xm;
      node >>= 1;
      if ((lane & ((2 << k) - 1)) == 0) {
        __stcg(U.sum + node, xs);
        __stcg(U.mn + node, xm);
      }
    }
    // warp roots = the 16 nodes of level 4
    float *r_s = ssum, *r_m = smin;
    if (lane == 0) r_s[tid >> 5] = xs, r_m[tid >> 5] = xm;
    __syncthreads();
    if (tid < 32) {
      xs = tid < 16 ? r_s[tid] : 0.0f;
      xm = tid < 16 ? r_m[tid] : FLT_MAX;
      node = 16 + tid;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float os = __shfl_xor_sync(0xffffffffu, xs, 1 << k);
        const float om = __shfl_xor_sync(0xffffffffu, xm, 1 << k);
        xs = (tid >> k) & 1 ? __fadd_rn(os, xs) : __fadd_rn(xs, os);
        xm = om < xm ? om : xm;
        node >>= 1;
        if (tid < 16 && (tid & ((2 << k) - 1)) == 0) {
          __stcg(U.sum + node, xs);
          __stcg(U.mn + node, xm);
        }
      }
      // lane 0 now holds the root: the sharded buffer's peers get it from here (kPublishOnMutation), a
      // whole step before their next sample needs it
      publish_root(U.pub, tid, __shfl_sync(0xffffffffu, xs, 0), __shfl_sync(0xffffffffu, xm, 0));
    }
    dbg_mark_gt(U.dbg, 16);
  } else {
    for (int i = tid; i < width; i += nthr) {
      ssum[width + i] = __ldcg(U.sum + width + i);
      smin[width + i] = __ldcg(U.mn + width + i);
    }
    __syncthreads();
#pragma unroll 1
    for (int level = split - 1; level >= 0; --level) {
      const int w = 1 << level;
      for (int i = tid; i < w; i += nthr) {
        const int nd = w + i;
        ssum[nd] = __fadd_rn(ssum[2 * nd], ssum[2 * nd + 1]);
        const float ma = smin[2 * nd], mb = smin[2 * nd + 1];
        smin[nd] = ma < mb ? ma : mb;
      }
      __syncthreads();
    }
    for (int i = tid + 1; i < width; i += nthr) {
      __stcg(U.sum + i, ssum[i]);
      __stcg(U.mn + i, smin[i]);
    }
    if (tid < 32) publish_root(U.pub, tid, ssum[1], smin[1]);
  }
  if (tid == 0) U.state->done_ctas = 0u;  // ready for the next launch (stream-ordered)
}

// ---- push --------------------------------------------------------------------------------
__global__ void __launch_bounds__(32)
    prb_push_kernel(float *sum, float *mn, const PrbState *state, int64_t cap2, int32_t levels,
                    int64_t slot, const __grid_constant__ ColTable tab, const PublishArgs pub) {
  pdl_prologue();  // (actor inserts between the learner's sample / update: the launch latencies overlap)
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
  // (vs, vm) is the root in every lane after the walk (levels == 0: the leaf is the root)
  publish_root(pub, lane, vs, vm);
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

// mrl_prb_shard_wait: one warp polls the local inbox until the words the NEXT sample will expect are there
// (bounded like the sample kernel's own poll)
__global__ void __launch_bounds__(32)
    prb_shard_wait_kernel(const uint2 *mailbox, int world, int rank, unsigned int slot, unsigned int expect,
                          long long timeout_ns, int *err) {
  const int lane = threadIdx.x;
  if (lane >= world || lane == rank) return;
  const uint2 *src = mailbox + ((size_t)slot * world + lane) * 4;
  long long t0 = 0;
  unsigned int spins = 0;
  while (true) {
    uint2 a, b, c;
    asm volatile("ld.relaxed.sys.global.v2.u32 {%0,%1}, [%2];" : "=r"(a.x), "=r"(a.y) : "l"(src) : "memory");
    asm volatile("ld.relaxed.sys.global.v2.u32 {%0,%1}, [%2];" : "=r"(b.x), "=r"(b.y) : "l"(src + 1) : "memory");
    asm volatile("ld.relaxed.sys.global.v2.u32 {%0,%1}, [%2];" : "=r"(c.x), "=r"(c.y) : "l"(src + 2) : "memory");
    if (a.y == expect && b.y == expect && c.y == expect) return;
    if ((++spins & 1023u) == 0u) {
      long long now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (t0 == 0) t0 = now;
      if (now - t0 > timeout_ns) {
        if (err) *reinterpret_cast<volatile int *>(err) = 1 + lane;
        return;
      }
    }
  }
}

// tail of the mutating calls that have no single kernel which ends with the root in hand (batched push,
// set_state, connect, and a sample that follows a sample): one warp reads the root and publishes it
__global__ void __launch_bounds__(32) prb_publish_kernel(const float *sum, const float *mn, const PublishArgs pub) {
  publish_root(pub, threadIdx.x, __ldcg(sum + 1), __ldcg(mn + 1));
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

// kPublishOnMutation: the publication a mutating call carries.  `bump` = this is a new state of the tree
// (false: the unchanged statistics are re-published into the slot of the coming sample).
static PublishArgs make_publication(PrbBuffer *b, int64_t size_after, bool bump = true) {
  PublishArgs P;
  ShardLink &k = b->link;
  if (k.peer_dev && k.mode == kPublishOnMutation) {
    P.outbox = k.mailbox + (size_t)2 * k.world * 4;
    P.rank = k.rank, P.world = k.world;
    P.slot = k.epoch & 1u;
    P.tag = bump ? ++k.version : k.version;
    P.size = (float)size_after;
    k.published_epoch = k.epoch;
  }
  return P;
}
// behind the kernel on `st` that has just written outbox slot P.slot: the courier on the handle's side stream
// carries the words to the peers (nothing to do when P is empty or the rank is alone)
static int dispatch_courier(PrbBuffer *b, const PublishArgs &P, cudaStream_t st) {
  ShardLink &k = b->link;
  if (!P.outbox || k.world < 2) return MRL_OK;
  MRL_CUDA(cudaEventRecord(k.mutated, st));
  MRL_CUDA(cudaStreamWaitEvent(k.courier, k.mutated, 0));
  prb_courier_kernel<<<1, 32, 0, k.courier>>>(P.outbox, k.peer_dev, k.rank, k.world, P.slot);
  MRL_LAUNCHED();
  return MRL_OK;
}
static int publish_tail(PrbBuffer *b, cudaStream_t st, bool bump = true) {
  const PublishArgs P = make_publication(b, b->size, bump);
  if (!P.outbox) return MRL_OK;
  prb_publish_kernel<<<1, 32, 0, st>>>(b->sum, b->mn, P);
  MRL_LAUNCHED();
  return dispatch_courier(b, P, st);
}

static int rebuild_range(PrbBuffer *b, int64_t lo, int64_t hi, cudaStream_t st) {
  if (hi <= lo) return MRL_OK;
  if (hi - lo <= 4096) {
    prb_rebuild_kernel<<<1, 1024, 0, st>>>(b->sum, b->mn, b->state, b->cap2, b->levels, lo, hi);
    MRL_LAUNCHED();
    return MRL_OK;
  }
  int64_t blocks = (hi - lo + 255) / 256;
  if (blocks > sm_count() * 8) blocks = sm_count() * 8;
  prb_set_leaves_kernel<<<(unsigned)blocks, 256, 0, st>>>(b->sum, b->mn, b->state, b->cap2, lo, hi);
  MRL_LAUNCHED();
  int64_t a = b->cap2 + lo, e = b->cap2 + hi - 1;
  for (int l = 0; l < b->levels; ++l) {
    a >>= 1;
    e >>= 1;
    int64_t nb = (e - a + 1 + 255) / 256;
    if (nb > sm_count() * 8) nb = sm_count() * 8;
    prb_level_kernel<<<(unsigned)nb, 256, 0, st>>>(b->sum, b->mn, a, e);
    MRL_LAUNCHED();
  }
  return MRL_OK;
}

static int prb_push_batch_impl(PrbBuffer *b, int64_t nrows, void *const *src, cudaStream_t st, int dir = 1) {
  const int64_t cap = b->capacity;
  if (nrows == 1 && dir == 1) {
    const int64_t slot = (b->head + 1) % cap;
    ColTable tab;
    fill_table(tab, b, src);
    const int64_t size_after = b->size < cap ? b->size + 1 : b->size;
    const PublishArgs P = make_publication(b, size_after);
    MRL_CUDA(launch_pdl(prb_push_kernel, dim3(1), dim3(32), 0, st, b->sum, b->mn, (const PrbState *)b->state, b->cap2,
                        b->levels, slot, tab, P));
    MRL_LAUNCHED();
    b->head = slot;
    b->size = size_after;
    return dispatch_courier(b, P, st);
  }
  const int64_t pos = (b->head + 1) % cap;
  const int64_t skip = nrows > cap ? nrows - cap : 0;
  const int64_t first = (pos + skip) % cap;
  const int64_t n = nrows - skip;
  ColTable tab;
  fill_table(tab, b, src);
  const int64_t run1 = n < cap - first ? n : cap - first;
  int rc = copy_segments_dir(tab, first, skip, run1, dir, st);
  if (rc) return rc;
  rc = rebuild_range(b, first, first + run1, st);
  if (rc) return rc;
  if (n > run1) {
    rc = copy_segments_dir(tab, 0, skip + run1, n - run1, dir, st);
    if (rc) return rc;
    rc = rebuild_range(b, 0, n - run1, st);
    if (rc) return rc;
  }
  b->head = (b->head + nrows) % cap;
  b->size = b->size + nrows > cap ? cap : b->size + nrows;
  return publish_tail(b, st);
}

static int prb_sample_impl(PrbBuffer *b, int64_t n, float beta, const float *uniforms,
                           const float *gstats, int32_t world, int64_t *indices, float *weights,
                           void *const *out_cols, cudaStream_t st, bool p2p = false,
                           const float *beta_dev = nullptr) {
  if (b->size <= 0) {
    set_error("PriorityReplayBuffer.sample: the buffer is empty");
    return MRL_ERR_STATE;
  }
  SampleArgs S;
  S.pub = PublishArgs(), S.mailbox = nullptr, S.slot = 0, S.expect = 0, S.diag = nullptr, S.err = nullptr;
  S.timeout_ns = g_shard_timeout_ns, S.beta_dev = beta_dev;
  if (p2p) {
    ShardLink &k = b->link;
    if (k.err_host && *reinterpret_cast<volatile int *>(k.err_host)) {
      set_error("sharded PriorityReplayBuffer: an earlier sample gave up waiting for the root statistics of rank %d "
                "(every rank must issue the same sequence of calls on one stream)", *k.err_host - 1);
      return MRL_ERR_STATE;
    }
    world = k.world;
    S.mailbox = k.mailbox, S.diag = k.diag, S.err = k.err_dev;
    S.pub.rank = k.rank;
    if (k.mode == kPublishOnSample) {
      const uint32_t e = ++k.epoch;
      S.pub.peers = k.peer_dev, S.pub.world = k.world, S.pub.slot = e & 1u, S.pub.tag = e;
      S.pub.size = (float)b->size;
      S.slot = e & 1u, S.expect = e;
    } else {
      // nothing has changed the tree since the previous sample: its publication sits in the other slot
      if (k.published_epoch != k.epoch) {
        const int rc = publish_tail(b, st, false);
        if (rc) return rc;
      }
      S.slot = k.epoch & 1u, S.expect = k.version;
      ++k.epoch;
    }
  }
  S.sum = b->sum, S.mn = b->mn, S.uniforms = uniforms, S.gstats = gstats;
  S.indices = indices, S.weights = weights;
  S.n = n, S.cap2 = b->cap2, S.capacity = b->capacity, S.size = b->size;
  S.levels = b->levels, S.world = world, S.beta = beta;
  S.exact_weights = g_exact_shard_weights ? 1 : 0;
  S.seed = b->seed, S.ctr = b->ctr;
  // narrow columns are copied by the warp that found the leaf; wide ones go to the TMA gather
  uint32_t fused = 0, wide = 0;
  if (out_cols)
    for (int c = 0; c < b->ncols; ++c) {
      if (b->row_bytes[c] <= kFusedRowLimit) fused |= 1u << c;
      else wide |= 1u << c;
    }
  S.gather_mask = fused;
  S.dbg = g_debug_timers;
  S.prefetch = g_tree_prefetch ? 1 : 0;
  ColTable tab;
  fill_table(tab, b, out_cols);
  S.word_units = 0;
  if (fused) {
    int64_t words = 0;
    bool ok = true;
    for (int c = 0; c < b->ncols; ++c)
      if ((fused >> c) & 1u) ok = ok && tab.vec[c] >= 4, words += b->row_bytes[c] / 4;
    if (ok) S.word_units = (int32_t)words;
  }
  if (wide && !p2p && !gstats) {  // Atari-shaped rows: descent and DMA ring in one kernel when the batch fits it
    const int rc = try_fused_sample(S, tab, wide, st);
    if (rc != kNotFused) {
      if (rc != MRL_OK) return rc;
      MRL_LAUNCHED();
      if (!uniforms) b->ctr += (uint64_t)n;
      return MRL_OK;
    }
  }
  const int64_t blocks = (n + kSampleWarps - 1) / kSampleWarps;
  MRL_CUDA(launch_pdl(prb_sample_kernel, dim3((unsigned)blocks), dim3(kSampleWarps * 32), 0, st, S, tab));
  MRL_LAUNCHED();
  if (!uniforms) b->ctr += (uint64_t)n;
  if (wide) return gather_rows(tab, indices, n, b->capacity, st, wide);
  return MRL_OK;
}

// 0 = subtree-partitioned kernel over 128 CTAs (default),
// 1 = one-CTA register walk with shared-memory hash exchange (MRL_UPDATE_PATH=walk; measured ~57 us),
// 2 = one-CTA global-tag level-synchronous kernel (MRL_UPDATE_PATH=tags; measured 18 us at n=256)
static int g_update_path = -1;

static bool update_takes_inline(const PrbBuffer *b, int64_t n) {
  return n <= kInlineMax && g_update_path <= 0 && b->levels - kTopLevels <= kWalkLevels && b->levels > kTopLevels &&
         b->cap2 < (1LL << 31);
}

bool g_no_inline_update = false;  // option "inline_update" = 0: always stage through device memory
bool g_host_zero_copy = true;     // option "host_zero_copy": small *_host sample results are written by the kernel itself

static int prb_update_impl(PrbBuffer *b, const int64_t *indices, const float *priorities,
                           int64_t n, cudaStream_t st, const InlineBatch *inl = nullptr) {
  if (g_update_path < 0) {
    const char *e = getenv("MRL_UPDATE_PATH");
    g_update_path = (e && e[0] == 'w') ? 1 : ((e && e[0] == 't') ? 2 : 0);
  }
  UpdateArgs U;
  U.sum = b->sum, U.mn = b->mn, U.tag = b->tag, U.state = b->state;
  U.indices = indices, U.priorities = priorities;
  U.n = n, U.cap2 = b->cap2, U.levels = b->levels, U.alpha = b->alpha;
  U.epoch = ++b->epoch;
  U.dbg = g_debug_timers;
  U.prefetch = g_tree_prefetch ? 1 : 0;
  U.pub = make_publication(b, b->size);
  if (n <= kWalkMax && b->levels - kTopLevels <= kWalkLevels && g_update_path == 1) {
    int threads = (int)((n + 31) / 32 * 32);
    threads = threads < 256 ? 256 : (threads > kWalkThreads ? kWalkThreads : threads);
    int bits = 6;
    while ((1 << bits) < 2 * (n < threads ? n : threads)) ++bits;
    const size_t smem = (size_t)(1 << bits) * (2 * 8 + 2 * 8 + 4) + (size_t)kWalkLevels * threads * 8;
    if (first_use_on_device((const void *)prb_update_walk_kernel))
      MRL_CUDA(cudaFuncSetAttribute(prb_update_walk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    1024 * 36 + kWalkLevels * kWalkThreads * 8));
    prb_update_walk_kernel<<<1, threads, smem, st>>>(U, bits);
  } else if (g_update_path == 0 && b->levels - kTopLevels <= kWalkLevels && b->levels > kTopLevels &&
             n < (1LL << 31)) {
    // 128 CTAs x 16 warps = one warp per level-11 subtree (n / 2048 entries each).  More CTAs (with their
    // launch ramp: 512 x 512 threads put up to four CTAs on an SM one after the other, and the last CTA of
    // the grid finished at 12.4 us where the first finished at 5.8) only for batches whose subtrees overflow
    // a warp anyway.
    int ctas = kSubCtas;
    if (n > 32768)
      while (ctas < 512 && n > 8 * (int64_t)ctas) ctas <<= 1;
    static const InlineBatch none = {};
    if (inl) MRL_CUDA(launch_pdl(prb_update_sub_kernel<true, false>, dim3(ctas), dim3(kSubThreads), 0, st, U, *inl));
    else if (n > 1024)
      MRL_CUDA(launch_pdl(prb_update_sub_kernel<false, true>, dim3(ctas), dim3(kSubThreads), 0, st, U, none));
    else MRL_CUDA(launch_pdl(prb_update_sub_kernel<false, false>, dim3(ctas), dim3(kSubThreads), 0, st, U, none));
  } else {
    prb_update_kernel<<<1, 1024, 0, st>>>(U);
  }
  MRL_LAUNCHED();
  return dispatch_courier(b, U.pub, st);
}

}  // namespace mrl

using namespace mrl;

#define PRB_GET(b, h)                                                      \
  PrbBuffer *b = g_prb.get(h);                                             \
  if (!b) {                                                                \
    set_error("unknown PriorityReplayBuffer handle %lld", (long long)(h)); \
    return MRL_ERR_HANDLE;                                                 \
  }

static void link_free(PrbBuffer *b) {
  ShardLink &k = b->link;
  for (int r = 0; r < k.world; ++r)
    if (k.peer[r] && r != k.rank) cudaIpcCloseMemHandle(k.peer[r]);
  if (k.peer_dev) cudaFree(k.peer_dev);
  if (k.mailbox) cudaFree(k.mailbox);
  if (k.diag) cudaFree(k.diag);
  if (k.err_host) cudaFreeHost(k.err_host);
  if (k.courier) {
    cudaStreamSynchronize(k.courier);
    cudaStreamDestroy(k.courier);
  }
  if (k.mutated) cudaEventDestroy(k.mutated);
  k = ShardLink();
}

static void prb_free(PrbBuffer *b) {
  link_free(b);
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
  prb_init_kernel<<<sm_count() * 4, 256>>>(b->sum, b->mn, b->tag, b->state, b->cap2);
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
  if (off[b->ncols] >= kDirectHostBytes) {
    // large batch: DMA from the caller's arrays straight into the ring rows, then the leaves and the tree;
    // the arrays are free again when this returns
    int rc = prb_push_batch_impl(b, nrows, (void *const *)src_cols, st, 2);
    if (rc) return rc;
    MRL_CUDA(cudaStreamSynchronize(st));
    return MRL_OK;
  }
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
  MRL_CUDA(b->stage.mark(st));
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

int mrl::prb_column_count(int64_t handle) {
  PrbBuffer *b = g_prb.get(handle);
  return b ? b->ncols : -1;
}

int mrl::prb_sample_beta_dev(int64_t handle, int64_t n, const float *beta_dev, int64_t *indices, float *weights,
                             void *const *out_cols, cudaStream_t st) {
  PRB_GET(b, handle);
  MRL_REQUIRE(n >= 1 && indices && weights && beta_dev, "PriorityReplayBufferSample: need n >= 1, beta, indices and weights");
  return prb_sample_impl(b, n, 0.0f, nullptr, nullptr, 0, indices, weights, out_cols, st, false, beta_dev);
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

// MRL_ERR_STATE when a kernel of this handle gave up polling its mailbox (the word is mapped host memory)
static int link_error(const PrbBuffer *b) {
  const ShardLink &k = b->link;
  if (k.err_host && *reinterpret_cast<volatile int *>(k.err_host)) {
    set_error("sharded PriorityReplayBuffer: the sample gave up waiting for the root statistics of rank %d "
              "(every rank must issue the same sequence of calls on one stream)", *k.err_host - 1);
    return MRL_ERR_STATE;
  }
  return MRL_OK;
}

static int prb_sample_host_impl(PrbBuffer *b, int64_t n, float beta, const float *uniforms,
                                int64_t *indices, float *weights, void *const *out_cols,
                                void *stream, bool p2p) {
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
  if (total <= (1u << 16) && g_host_zero_copy) {
    // small result: the kernel writes straight into the pinned staging buffer (it is mapped into the
    // device address space), so there is no device-to-host copy to launch -- only the stream wait.
    // (Measured and dropped: per-warp completion words in the same pinned memory, polled by the host
    // instead of the event -- the system-scope fences and 256 more PCIe writes cost 1.3 us more than
    // the event record + query they replace; and collecting a CTA's 8 results in shared memory to store
    // each output array's piece with full-width stores -- 1 us slower: the cost is the PCIe round trip
    // at the end of the kernel, not the number of transactions.)
    uint8_t *hz = b->stage.host;
    for (int c = 0; c < b->ncols; ++c) out_dev[c] = hz + off[c];
    rc = prb_sample_impl(b, n, beta, uniforms ? (const float *)(d + o_u) : nullptr, nullptr, 0,
                         (int64_t *)(hz + o_idx), (float *)(hz + o_w), out_cols ? out_dev : nullptr, st, p2p);
    if (rc) return rc;
    rc = stream_wait_spin(st);
    if (rc) return rc;
    if (p2p && (rc = link_error(b))) return rc;
    memcpy(indices, hz + o_idx, (size_t)n * 8);
    memcpy(weights, hz + o_w, (size_t)n * 4);
    if (out_cols)
      for (int c = 0; c < b->ncols; ++c) memcpy(out_cols[c], hz + off[c], (size_t)n * b->row_bytes[c]);
    return MRL_OK;
  }
  for (int c = 0; c < b->ncols; ++c) out_dev[c] = d + off[c];
  rc = prb_sample_impl(b, n, beta, uniforms ? (const float *)(d + o_u) : nullptr, nullptr, 0,
                       (int64_t *)(d + o_idx), (float *)(d + o_w), out_cols ? out_dev : nullptr, st, p2p);
  if (rc) return rc;
  if (total <= (1u << 16)) {  // (option "host_zero_copy" = 0) one DMA of the whole arena, then host copies
    MRL_CUDA(cudaMemcpyAsync(b->stage.host, d, total, cudaMemcpyDeviceToHost, st));
    rc = stream_wait_spin(st);
    if (rc) return rc;
    if (p2p && (rc = link_error(b))) return rc;
    const uint8_t *hz = b->stage.host;
    memcpy(indices, hz + o_idx, (size_t)n * 8);
    memcpy(weights, hz + o_w, (size_t)n * 4);
    if (out_cols)
      for (int c = 0; c < b->ncols; ++c) memcpy(out_cols[c], hz + off[c], (size_t)n * b->row_bytes[c]);
  } else {
    MRL_CUDA(cudaMemcpyAsync(b->stage.host, d, o_u, cudaMemcpyDeviceToHost, st));
    if (out_cols)
      for (int c = 0; c < b->ncols; ++c)
        MRL_CUDA(cudaMemcpyAsync(out_cols[c], d + off[c], (size_t)n * b->row_bytes[c],
                                 cudaMemcpyDeviceToHost, st));
    MRL_CUDA(cudaStreamSynchronize(st));
    if (p2p && (rc = link_error(b))) return rc;
    memcpy(indices, b->stage.host + o_idx, (size_t)n * 8);
    memcpy(weights, b->stage.host + o_w, (size_t)n * 4);
  }
  return MRL_OK;
}

extern "C" int mrl_prb_sample_host(int64_t handle, int64_t n, float beta, const float *uniforms,
                                   int64_t *indices, float *weights, void *const *out_cols,
                                   void *stream) {
  PRB_GET(b, handle);
  MRL_REQUIRE(n >= 1 && indices && weights, "mrl_prb_sample_host: need n >= 1, indices and weights");
  return prb_sample_host_impl(b, n, beta, uniforms, indices, weights, out_cols, stream, false);
}

extern "C" int mrl_prb_sample_sharded_p2p_host(int64_t handle, int64_t n, float beta,
                                               const float *uniforms, int64_t *indices,
                                               float *weights, void *const *out_cols, void *stream) {
  PRB_GET(b, handle);
  MRL_REQUIRE(n >= 1 && indices && weights, "mrl_prb_sample_sharded_p2p_host: need n >= 1, indices and weights");
  MRL_REQUIRE(b->link.peer_dev, "mrl_prb_sample_sharded_p2p_host: mailboxes are not connected");
  return prb_sample_host_impl(b, n, beta, uniforms, indices, weights, out_cols, stream, true);
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
  if (g_update_path < 0) {
    const char *e = getenv("MRL_UPDATE_PATH");
    g_update_path = (e && e[0] == 'w') ? 1 : ((e && e[0] == 't') ? 2 : 0);
  }
  // the indices are host memory here, so checking them costs nothing (the device entry point keeps the
  // reference's contract: "caller needs to ensure the validity of the indices", priority_replay_buffer.py:113)
  for (int64_t i = 0; i < n; ++i)
    MRL_REQUIRE(indices[i] >= 0 && indices[i] < b->capacity,
                "mrl_prb_update_host: index %lld at position %lld is outside [0, %lld)", (long long)indices[i],
                (long long)i, (long long)b->capacity);
  if (update_takes_inline(b, n) && !g_no_inline_update) {
    // learner-sized batch: it rides in the launch itself (no copy, no staging, the caller's arrays are free
    // again when this returns)
    InlineBatch ib;
    for (int64_t i = 0; i < n; ++i) {
      ib.idx[i] = (int32_t)indices[i];
      ib.pr[i] = priorities[i];
    }
    return prb_update_impl(b, nullptr, nullptr, n, st, &ib);
  }
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
  MRL_CUDA(b->stage.mark(st));
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

extern "C" int mrl_prb_shard_init(int64_t handle, int32_t rank, int32_t world, void *ipc_handle_out) {
  PRB_GET(b, handle);
  MRL_REQUIRE(world >= 1 && world <= kMaxWorld && rank >= 0 && rank < world && ipc_handle_out,
              "mrl_prb_shard_init: need 0 <= rank < world <= %d", kMaxWorld);
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle is 64 bytes in the ABI");
  link_free(b);
  ShardLink &k = b->link;
  k.rank = rank, k.world = world;
  const size_t bytes = sizeof(uint2) * (2 * (size_t)world * 4 + 2 * 4);  // inbox [2][world][4] + outbox [2][4]
  MRL_CUDA(cudaMalloc((void **)&k.mailbox, bytes));
  MRL_CUDA(cudaMemset(k.mailbox, 0, bytes));  // tag 0 never matches a call (tags start at 1)
  MRL_CUDA(cudaMalloc((void **)&k.diag, sizeof(ShardDiag)));
  MRL_CUDA(cudaMemset(k.diag, 0, sizeof(ShardDiag)));
  MRL_CUDA(cudaHostAlloc((void **)&k.err_host, sizeof(int), cudaHostAllocMapped));
  *k.err_host = 0;
  MRL_CUDA(cudaHostGetDevicePointer((void **)&k.err_dev, k.err_host, 0));
  MRL_CUDA(cudaStreamCreateWithFlags(&k.courier, cudaStreamNonBlocking));
  MRL_CUDA(cudaEventCreateWithFlags(&k.mutated, cudaEventDisableTiming));
  MRL_CUDA(cudaDeviceSynchronize());
  cudaIpcMemHandle_t h;
  MRL_CUDA(cudaIpcGetMemHandle(&h, k.mailbox));
  memcpy(ipc_handle_out, &h, sizeof(h));
  return MRL_OK;
}

extern "C" int mrl_prb_shard_set_mode(int64_t handle, int32_t mode) {
  PRB_GET(b, handle);
  MRL_REQUIRE(mode == kPublishOnSample || mode == kPublishOnMutation, "mrl_prb_shard_set_mode: mode is 0 or 1");
  MRL_REQUIRE(b->link.mailbox && !b->link.peer_dev,
              "mrl_prb_shard_set_mode: call it between mrl_prb_shard_init and mrl_prb_shard_connect");
  b->link.mode = mode;
  return MRL_OK;
}

extern "C" int mrl_prb_shard_connect(int64_t handle, const void *all_ipc_handles, void *stream) {
  PRB_GET(b, handle);
  ShardLink &k = b->link;
  MRL_REQUIRE(k.world >= 1 && k.mailbox && all_ipc_handles, "mrl_prb_shard_connect: call mrl_prb_shard_init first");
  MRL_REQUIRE(!k.peer_dev, "mrl_prb_shard_connect: already connected");
  for (int r = 0; r < k.world; ++r) {
    if (r == k.rank) {
      k.peer[r] = k.mailbox;
      continue;
    }
    cudaIpcMemHandle_t h;
    memcpy(&h, (const char *)all_ipc_handles + 64 * (size_t)r, sizeof(h));
    void *p = nullptr;
    const cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      // a rank on another node, or no peer access: the caller falls back to the all-gather path
      cudaGetLastError();
      for (int q = 0; q < r; ++q)
        if (q != k.rank && k.peer[q]) cudaIpcCloseMemHandle(k.peer[q]), k.peer[q] = nullptr;
      set_error("mrl_prb_shard_connect: cannot map the mailbox of rank %d: %s", r, cudaGetErrorString(e));
      return MRL_ERR_CUDA;
    }
    k.peer[r] = (uint2 *)p;
  }
  MRL_CUDA(cudaMalloc((void **)&k.peer_dev, sizeof(uint2 *) * kMaxWorld));
  MRL_CUDA(cudaMemcpy(k.peer_dev, k.peer, sizeof(uint2 *) * kMaxWorld, cudaMemcpyHostToDevice));
  k.epoch = 0, k.version = 0, k.published_epoch = 0;
  // kPublishOnMutation: the state the buffer has now is publication number 1 (every peer's mailbox exists
  // since its own mrl_prb_shard_init, which precedes the exchange of the handles)
  return publish_tail(b, as_stream(stream));
}

extern "C" int mrl_prb_shard_wait(int64_t handle, void *stream) {
  PRB_GET(b, handle);
  ShardLink &k = b->link;
  MRL_REQUIRE(k.peer_dev, "mrl_prb_shard_wait: mailboxes are not connected");
  if (k.mode != kPublishOnMutation || k.world < 2) return MRL_OK;  // nothing is in flight between samples
  cudaStream_t st = as_stream(stream);
  if (k.published_epoch != k.epoch) {  // the coming sample's slot holds nothing yet: what the sample call would do
    const int rc = publish_tail(b, st, false);
    if (rc) return rc;
  }
  prb_shard_wait_kernel<<<1, 32, 0, st>>>(k.mailbox, k.world, k.rank, k.epoch & 1u, k.version, g_shard_timeout_ns,
                                          k.err_dev);
  MRL_LAUNCHED();
  return MRL_OK;
}

extern "C" int mrl_prb_shard_diag(int64_t handle, int64_t *wait_cycles, int64_t *polled_calls,
                                  int64_t *waited_calls, int64_t *timeouts, int32_t reset, void *stream) {
  PRB_GET(b, handle);
  MRL_REQUIRE(b->link.diag, "mrl_prb_shard_diag: the buffer is not sharded");
  cudaStream_t st = as_stream(stream);
  ShardDiag d;
  MRL_CUDA(cudaMemcpyAsync(&d, b->link.diag, sizeof(d), cudaMemcpyDeviceToHost, st));
  if (reset) MRL_CUDA(cudaMemsetAsync(b->link.diag, 0, sizeof(ShardDiag), st));
  MRL_CUDA(cudaStreamSynchronize(st));
  if (wait_cycles) *wait_cycles = (int64_t)d.wait_cycles;
  if (polled_calls) *polled_calls = (int64_t)d.polled_calls;
  if (waited_calls) *waited_calls = (int64_t)d.waited_calls;
  if (timeouts) *timeouts = (int64_t)d.timeouts;
  if (b->link.err_host && *reinterpret_cast<volatile int *>(b->link.err_host)) {
    set_error("sharded PriorityReplayBuffer: a sample gave up waiting for the root statistics of rank %d",
              *b->link.err_host - 1);
    return MRL_ERR_STATE;
  }
  return MRL_OK;
}

extern "C" int mrl_prb_sample_sharded_p2p(int64_t handle, int64_t n, float beta, const float *uniforms,
                                          int64_t *indices, float *weights, void *const *out_cols,
                                          void *stream) {
  PRB_GET(b, handle);
  MRL_REQUIRE(n >= 1 && indices && weights, "mrl_prb_sample_sharded_p2p: need n >= 1, indices and weights");
  MRL_REQUIRE(b->link.peer_dev, "mrl_prb_sample_sharded_p2p: mailboxes are not connected");
  return prb_sample_impl(b, n, beta, uniforms, nullptr, 0, indices, weights, out_cols,
                         as_stream(stream), true);
}

extern "C" int mrl_prb_get_state(int64_t handle, int64_t *size, int64_t *head, uint64_t *rng_counter,
                                 float *max_priority, void *stream) {
  PRB_GET(b, handle);
  if (size) *size = b->size;
  if (head) *head = b->head;
  if (rng_counter) *rng_counter = b->ctr;
  if (max_priority) return mrl_prb_max_priority(handle, max_priority, stream);
  return MRL_OK;
}

extern "C" int mrl_prb_set_state(int64_t handle, int64_t size, int64_t head, uint64_t rng_counter,
                                 float max_priority, const float *sum_nodes, const float *min_nodes,
                                 void *stream) {
  PRB_GET(b, handle);
  MRL_REQUIRE(size >= 0 && size <= b->capacity && head >= -1 && head < b->capacity && sum_nodes && min_nodes,
              "mrl_prb_set_state: state does not fit this buffer");
  cudaStream_t st = as_stream(stream);
  const size_t bytes = sizeof(float) * 2 * (size_t)b->cap2;
  MRL_CUDA(cudaMemcpyAsync(b->sum, sum_nodes, bytes, cudaMemcpyHostToDevice, st));
  MRL_CUDA(cudaMemcpyAsync(b->mn, min_nodes, bytes, cudaMemcpyHostToDevice, st));
  MRL_CUDA(cudaMemcpyAsync(&b->state->max_priority, &max_priority, sizeof(float), cudaMemcpyHostToDevice, st));
  MRL_CUDA(cudaStreamSynchronize(st));
  b->size = size;
  b->head = head;
  b->ctr = rng_counter;
  return publish_tail(b, st);
}
