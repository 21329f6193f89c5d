/*
 * mindrl_b200.h -- C ABI of the B200-native experience data path.
 *
 * This is the drop-in boundary for the MindSpore primitives that MindRL's
 * experience data path calls (reference paths relative to /root/reference):
 *
 *   P.BufferAppend / BufferGetItem / BufferSample
 *       mindspore_rl/core/uniform_replay_buffer.py:82-89, 114, 128, 139
 *   _rl_inner_ops.PriorityReplayBuffer{Create,Push,Sample,Update,Destroy}
 *       mindspore_rl/core/priority_replay_buffer.py:62-76, 90, 106, 120, 138
 *   rl_ops.DiscountedReturn
 *       mindspore_rl/utils/discounted_return.py:71, 82 (CPU loop :84-92)
 *   in-learner reverse scans (GAE, discounted reward, V-trace, TD(lambda))
 *       mindspore_rl/algorithm/ppo/ppo.py:320-350
 *       mindspore_rl/algorithm/mappo/mappo.py:478-497
 *       mindspore_rl/algorithm/impala/vtrace.py:88-93
 *
 * Shape of the ABI follows the reference's only in-repo native operator ABI
 * (MindSpore AOT custom ops, mindspore_rl/utils/mcts/cpu/cpu_mcts_ops.cc:38-53,
 * mindspore_rl/utils/mcts/mcts_factory.h:39-82): extern "C", raw pointers,
 * an explicit cudaStream_t passed as void*, int return code (0 = ok),
 * state behind an int64 handle held in a process-global registry
 * (invalid handle = -1).
 *
 * Conventions
 *   - every pointer argument is a DEVICE pointer unless the function name ends
 *     in _host (then data pointers are host pointers; the library stages the
 *     host<->device copies on `stream` and the host buffers are reusable on
 *     return);
 *   - arrays of pointers / sizes (col tables) are always HOST arrays;
 *   - `stream` is a cudaStream_t (NULL = legacy default stream); every launch
 *     of a call goes to that stream, nothing synchronises unless stated;
 *   - calls on one handle must be issued from one thread at a time and are
 *     stream-ordered; different handles are independent (same contract as the
 *     reference registry, which has no locking);
 *   - the scans keep their workspaces (look-back words, moment partials, scratch)
 *     per (device, stream): calls on different streams or devices may run
 *     concurrently; every scan entry point can be captured in a CUDA graph once
 *     it has run on that stream for the same or a larger shape (workspace
 *     allocation is not a stream operation);
 *   - no exceptions, no exit(): errors are the return code, text through
 *     mrl_last_error().
 *   - there is NO CPU fallback: without a CUDA device every compute entry
 *     point returns MRL_ERR_CUDA.
 */
#ifndef MINDRL_B200_H_
#define MINDRL_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MRL_OK 0
#define MRL_ERR_INVALID 1 /* bad argument */
#define MRL_ERR_CUDA 2    /* CUDA runtime error (kErrorCode = 2 in cpu_mcts_ops.cc:25) */
#define MRL_ERR_HANDLE 3  /* unknown / destroyed handle */
#define MRL_ERR_STATE 4   /* e.g. sample size larger than buffer size */

#define MRL_INVALID_HANDLE (-1) /* mcts_factory.h:32 kInvalidHandle */
#define MRL_MAX_COLS 32

/* scan layouts */
#define MRL_TIME_MAJOR 0  /* [T, B, E...]  (DiscountedReturn, MAPPO, V-trace) */
#define MRL_BATCH_MAJOR 1 /* [B, T]        (PPO gae / discounted_reward) */
/* scan modes */
#define MRL_SCAN_AUTO 0
#define MRL_SCAN_SEQUENTIAL 1 /* one thread walks t = T-1..0: bit-identical to the reference loop order */
#define MRL_SCAN_PARALLEL 2   /* chunked / warp-shuffle affine scan: re-associates fp32, <=1e-5 rel */

/* ---- library / device ------------------------------------------------- */
int mrl_version(void);
const char *mrl_last_error(void);
int mrl_device_count(int *count);
int mrl_set_device(int device);
int mrl_get_device(int *device);
/* number of kernels this library has launched since load (per process) */
int64_t mrl_launch_count(void);

/* tunables (defaults in brackets; every one of them only selects among kernels that give the same results):
 *   "gather_path" [0] 0 auto, 1 register path only, 2 TMA ring whenever aligned; "bulk_min_row" [2048],
 *   "bulk_chunk" [8192], "bulk_ctas_per_sm" [4], "bulk_stages" [4] (2/4/8/12): the TMA-ring gather
 *   "fused_gather" [1]: PER sample of wide (TMA-gathered) rows runs the descent and the DMA ring in ONE kernel when
 *       the batch fits it (a CTA's share of the copy touches at most two samples: batch <= SMs * bulk_ctas_per_sm)
 *   "pdl" [1]: programmatic dependent launch of the PER sample / gather / update kernels
 *   "tree_prefetch" [1]: the PER sample / update kernels prefetch the tree lines of their later dependent
 *       round trips (depth 10 and 15 of the descent, depth 11 of the top rebuild) into L2 at kernel start
 *   "host_zero_copy" [1]: mrl_prb_sample_host results <= 64 KB are written by the kernel straight into mapped
 *       pinned memory (0: device arena + one DMA; measured 6 us slower per call)
 *   "inline_update" [1]: mrl_prb_update_host passes batches <= 256 in the kernel parameter block
 *   "tm_kernel" [0]: time-major scans: 0 auto, 2 two-pass, 3 128-column tiles, 5 warp tiles + look-back,
 *       6 TMA strips whenever aligned; "strip_cols" [0 = 32] (16, 28: measured slower), "strip_rpu" [0 = 32 steps per lane and tile, 16 for GAE with returns], "strip_groups" [0 = 1],
 *       "strip_stages" [0 = as many as fit], "strip_min_cols" [2048]
 *   "bm_kernel" [0]: batch-major scans: 0 auto, 1 warp-per-row register kernel, 2 TMA rows whenever aligned;
 *       "row_ctas_per_sm" [4], "row_stages" [0 = from row_ctas_per_sm], "row_spl" [8] steps per lane (16: four
 *       consumer warps per CTA; measured slower)
 *   "host_scan_chunks" [0 = ~16 MB of input per piece]: pieces of the *_host scans (1 = no pipelining)
 *   "shard_weights" [0]: importance weights of the SHARDED sample.  0: against the global sum,
 *       minimum and size, (p / sum_all * N)^-beta / (min_all / sum_all * N)^-beta -- exact when all shards carry the
 *       same priority mass.  1: against the probability the sample was really drawn with -- every rank draws its
 *       batch from its own tree, P(i) = p_i / (world * S_rank) -- normalised by the largest such weight over all
 *       ranks: unbiased for shards of unequal mass.  Set it identically on every rank.
 *   "shard_timeout_ms" [10000]: how long a sharded sample kernel polls for a peer's root statistics
 *   "debug_timers": device pointer to >= 32 int64 phase marks of the tree kernels (0 = off) */
int mrl_set_option(const char *key, int64_t value);

/* carrier-free memory helpers (PyTorch is optional; these are plain cudaMalloc etc.) */
int mrl_malloc(void **dptr, size_t bytes);
int mrl_free(void *dptr);
int mrl_host_alloc(void **hptr, size_t bytes); /* pinned */
int mrl_host_free(void *hptr);
int mrl_memcpy_h2d(void *dst, const void *src, size_t bytes, void *stream);
int mrl_memcpy_d2h(void *dst, const void *src, size_t bytes, void *stream);
int mrl_memcpy_d2d(void *dst, const void *src, size_t bytes, void *stream);
int mrl_memset(void *dst, int value, size_t bytes, void *stream);
int mrl_stream_sync(void *stream);
/* synthetic fill for benches: row r of a [nrows,row_bytes] column gets its slot id
 * (int64, little endian) in bytes 0..7 (when row_bytes>=8) and a hash of (seed,r,offset) after */
int mrl_fill_rows(void *col, int64_t nrows, int64_t row_bytes, uint64_t seed, void *stream);

/* ---- UniformReplayBuffer: BufferAppend / BufferGetItem / BufferSample --
 * (uniform_replay_buffer.py:75-172).  Storage is caller-visible, as in the
 * reference where `.buffer` Parameters are passed into every op: pass the
 * column base pointers in `col_ptrs` (each capacity*row_bytes[i] bytes), or
 * NULL to let the library allocate zeroed columns (read them back with
 * mrl_urb_columns).  `seed` seeds the BufferSample draw (:84-89). */
int mrl_urb_create(int64_t capacity, int32_t ncols, const int64_t *row_bytes,
                   void *const *col_ptrs, uint64_t seed, int64_t *handle);
int mrl_urb_destroy(int64_t handle);
int mrl_urb_columns(int64_t handle, void **col_ptrs_out);
/* BufferAppend (:114): append nrows>=1 rows; src_cols[i] holds nrows*row_bytes[i] bytes.
 * FIFO overwrite when full. */
int mrl_urb_append(int64_t handle, int64_t nrows, const void *const *src_cols, void *stream);
int mrl_urb_append_host(int64_t handle, int64_t nrows, const void *const *src_cols, void *stream);
/* Emplace of ReservoirReplayBufferPush (reservoir_replay_buffer.py:69-82): overwrite nrows rows at
 * PHYSICAL slot `slot`; count and head do not move. */
int mrl_urb_write_at(int64_t handle, int64_t slot, int64_t nrows, const void *const *src_cols,
                     void *stream);
int mrl_urb_write_at_host(int64_t handle, int64_t slot, int64_t nrows, const void *const *src_cols,
                          void *stream);
/* BufferGetItem (:128): row at LOGICAL position index (0 = oldest, negative = from the end) */
int mrl_urb_get_item(int64_t handle, int64_t index, void *const *out_cols, void *stream);
int mrl_urb_get_item_host(int64_t handle, int64_t index, void *const *out_cols, void *stream);
/* gather rows at the given PHYSICAL slots (the second half of BufferSample) */
int mrl_urb_gather(int64_t handle, const int64_t *slots, int64_t n, void *const *out_cols,
                   void *stream);
/* BufferSample (:139): draw n slots uniformly with replacement from the `count` valid rows and
 * gather them; slots_out may be NULL. MRL_ERR_STATE when n exceeds the valid rows. */
int mrl_urb_sample(int64_t handle, int64_t n, int64_t *slots_out, void *const *out_cols,
                   void *stream);
int mrl_urb_sample_host(int64_t handle, int64_t n, int64_t *slots_out, void *const *out_cols,
                        void *stream);
/* reset (:141-150): count <- 0, head and data untouched */
int mrl_urb_reset(int64_t handle, void *stream);
/* size()/full() (:152-172) read these; the host mirror is exact because the cursor rule is
 * deterministic in the number of rows appended */
int mrl_urb_state(int64_t handle, int32_t *count, int32_t *head);
int mrl_urb_set_state(int64_t handle, int32_t count, int32_t head);

/* MSRL.get_replay_buffer_elements(transpose=True, shape=(1, 0, 2)) (core/msrl.py:535-557): swap the two
 * leading axes of a column, out[b][a][:] = in[a][b][:], rows of row_bytes bytes.  (The scans take either
 * layout, so a learner can also skip this pass: see mrl_gae / mrl_discounted_return `layout`.) */
int mrl_transpose01(const void *in, void *out, int64_t A, int64_t B, int64_t row_bytes, void *stream);

/* ---- PriorityReplayBuffer (priority_replay_buffer.py:60-138) ---------- */
/* PriorityReplayBufferCreate(capacity, alpha, shapes, types, seed0, seed1) (:62-66) */
int mrl_prb_create(int64_t capacity, float alpha, int32_t ncols, const int64_t *row_bytes,
                   uint64_t seed0, uint64_t seed1, int64_t *handle);
/* PriorityReplayBufferDestroy (:138) */
int mrl_prb_destroy(int64_t handle);
/* PriorityReplayBufferPush (:90): one transition row; the new leaf gets max_priority */
int mrl_prb_push(int64_t handle, const void *const *row_cols, void *stream);
int mrl_prb_push_host(int64_t handle, const void *const *row_cols, void *stream);
/* nrows consecutive pushes in one call (same result as nrows calls of mrl_prb_push) */
int mrl_prb_push_batch(int64_t handle, int64_t nrows, const void *const *src_cols, void *stream);
int mrl_prb_push_batch_host(int64_t handle, int64_t nrows, const void *const *src_cols,
                            void *stream);
/* PriorityReplayBufferSample (:106): stratified proportional sampling of n transitions.
 * uniforms: n floats in [0,1) to use as the per-stratum draws (parity / reproducibility), or
 * NULL to draw them from the buffer's counter-based generator.
 * out: indices int64[n], weights float[n], out_cols[i] = n*row_bytes[i] bytes (out_cols may be
 * NULL to skip the row gather).  Every index is < size (a mass that rounds past the total returns
 * the newest row).  MRL_ERR_STATE when the buffer is empty. */
int mrl_prb_sample(int64_t handle, int64_t n, float beta, const float *uniforms,
                   int64_t *indices, float *weights, void *const *out_cols, void *stream);
int mrl_prb_sample_host(int64_t handle, int64_t n, float beta, const float *uniforms,
                        int64_t *indices, float *weights, void *const *out_cols, void *stream);
/* PriorityReplayBufferUpdate (:120): leaf[indices[k]] <- priorities[k]^alpha in order k=0..n-1
 * (duplicates: last occurrence wins), ancestors recomputed, max_priority tracked.  As in the
 * reference (:113) the caller guarantees 0 <= indices[k] < capacity for device arrays; the
 * _host variant checks them (MRL_ERR_INVALID). */
int mrl_prb_update(int64_t handle, const int64_t *indices, const float *priorities, int64_t n,
                   void *stream);
int mrl_prb_update_host(int64_t handle, const int64_t *indices, const float *priorities,
                        int64_t n, void *stream);
/* host-side facts: valid rows, ring cursor (slot of the newest row, -1 when empty), padded leaf
 * count (power of two) */
int mrl_prb_info(int64_t handle, int64_t *size, int64_t *head, int64_t *leaf_capacity);
/* reads max_priority from the device (synchronises `stream`) */
int mrl_prb_max_priority(int64_t handle, float *max_priority, void *stream);
/* parity helper: copies the 2*leaf_capacity node arrays (index 0 unused, root = 1) to HOST
 * buffers; synchronises `stream`. */
int mrl_prb_tree_dump(int64_t handle, float *sum_nodes, float *min_nodes, void *stream);
/* device pointers of the ring columns (capacity rows each) */
int mrl_prb_columns(int64_t handle, void **col_ptrs_out);
/* checkpoint support (the reference never saves its buffers): the scalar state
 * {size, head, rng counter, max_priority} and the 2*leaf_capacity node arrays.  Host pointers;
 * both calls synchronise `stream`.  Columns are copied by the caller through mrl_prb_columns. */
int mrl_prb_get_state(int64_t handle, int64_t *size, int64_t *head, uint64_t *rng_counter,
                      float *max_priority, void *stream);
int mrl_prb_set_state(int64_t handle, int64_t size, int64_t head, uint64_t rng_counter,
                      float max_priority, const float *sum_nodes, const float *min_nodes,
                      void *stream);
/* sharded buffer (one sub-buffer + tree per rank):
 *   mrl_prb_root_stats writes {root_sum, root_min, (float)size, max_priority} to stats4 (device);
 *   ranks all-gather those 16 bytes (NCCL) into gstats[world][4];
 *   mrl_prb_sample_sharded draws n local strata from the local tree and computes the
 *   importance weights against the GLOBAL sum / min / size. */
int mrl_prb_root_stats(int64_t handle, float *stats4, void *stream);
int mrl_prb_sample_sharded(int64_t handle, int64_t n, float beta, const float *uniforms,
                           const float *gstats, int32_t world, int64_t *indices, float *weights,
                           void *const *out_cols, void *stream);

/* Fused exchange for ranks on ONE node (one process per GPU): the root statistics travel as peer stores
 * over NVLink issued by the buffer's own kernels; no NCCL call, no extra launch in the learner loop.
 *   1. mrl_prb_shard_init     allocates this rank's mailbox, returns its 64-byte CUDA IPC handle
 *   2. (optional) mrl_prb_shard_set_mode
 *        1 = publish on mutation (default): the kernel that changes the tree (update_priorities' last CTA,
 *            push, the tail of a batched push / set_state) stores {root_sum, root_min, size}, tagged with the
 *            rank's mutation count, into the rank's own outbox, and a one-warp courier kernel on a side stream
 *            of the handle carries the words into every peer's inbox; sample #j uses, for every peer, the
 *            statistics that peer had at ITS sample #j -- which landed a whole step earlier, so the sample
 *            step does not wait for the slowest rank unless that rank is more than a step behind, and neither
 *            the sample nor the update kernel touches peer memory.
 *        0 = publish on sample (round 1): the sample kernel stores the statistics itself and waits for the
 *            matching words of all peers.
 *      Same results in both modes and as mrl_prb_root_stats + all-gather + mrl_prb_sample_sharded, PROVIDED
 *      every rank issues the same sequence of sample / mutating calls on the handle (SPMD) and all of them
 *      on one stream.  Programs whose ranks mutate their shards at different rates use the all-gather path.
 *   3. ranks all-gather the handles (any transport) into world*64 bytes, rank order
 *   4. mrl_prb_shard_connect  maps the peers' mailboxes (MRL_ERR_CUDA when a peer is not reachable through
 *      CUDA IPC, e.g. on another node: fall back to the all-gather path) and publishes the current state
 *   5. mrl_prb_sample_sharded_p2p.  A kernel that sees no word from a peer for "shard_timeout_ms" gives up
 *      (weights NaN) and raises the link's error word: the *_host variant and every later call on the handle
 *      return MRL_ERR_STATE instead of hanging the GPU.
 * Tear-down: the peers' kernels write into this rank's mailbox, so destroy a sharded buffer only after every rank
 * has drained its streams (synchronise, then barrier) -- CUDA IPC leaves freeing exported memory that a peer still
 * uses undefined.
 * mrl_prb_shard_diag reads (and optionally resets) the exchange counters: SM cycles warp 0 spent polling,
 * calls that polled, calls whose first look missed a word, timeouts.  Synchronises `stream`. */
int mrl_prb_shard_init(int64_t handle, int32_t rank, int32_t world, void *ipc_handle_out);
int mrl_prb_shard_set_mode(int64_t handle, int32_t mode);
int mrl_prb_shard_connect(int64_t handle, const void *all_ipc_handles, void *stream);
int mrl_prb_sample_sharded_p2p(int64_t handle, int64_t n, float beta, const float *uniforms,
                               int64_t *indices, float *weights, void *const *out_cols,
                               void *stream);
int mrl_prb_sample_sharded_p2p_host(int64_t handle, int64_t n, float beta, const float *uniforms,
                                    int64_t *indices, float *weights, void *const *out_cols,
                                    void *stream);
/* Puts the wait for the peers where the caller wants it: a one-warp kernel on `stream` that returns when the root
 * statistics the NEXT mrl_prb_sample_sharded_p2p will use have arrived from every peer (same bound and error word
 * as the sample's own poll).  A learner calls it before its own forward/backward work, a benchmark before its timed
 * region; the sample that follows finds its words in place.  No-op for mode 0. */
int mrl_prb_shard_wait(int64_t handle, void *stream);
int mrl_prb_shard_diag(int64_t handle, int64_t *wait_cycles, int64_t *polled_calls, int64_t *waited_calls,
                       int64_t *timeouts, int32_t reset, void *stream);

/* ---- reverse scans ----------------------------------------------------- */
/* DiscountedReturn (discounted_return.py:84-92):
 *   last <- reward[t] + ((1 - done[t]) * gamma) * last ; out[t] <- last,  t = T-1..0
 * time-major: reward/out [T,B,E], done [T,B] (uint8, broadcast over E), last_value [B,E]
 * batch-major: reward/out [B,T], done [B,T], last_value [B]  (E must be 1)
 * done may be NULL (PPO discounted_reward, ppo.py:320-332); last_value may be NULL (= 0). */
int mrl_discounted_return(const float *reward, const uint8_t *done, const float *last_value,
                          float *out, int64_t T, int64_t B, int64_t E, float gamma,
                          int32_t layout, int32_t mode, void *stream);
int mrl_discounted_return_host(const float *reward, const uint8_t *done, const float *last_value,
                               float *out, int64_t T, int64_t B, int64_t E, float gamma,
                               int32_t layout, int32_t mode, void *stream);
/* GAE (ppo.py:334-350 without done, mappo.py:478-497 with masks):
 *   m = 1 - done[t]
 *   delta = reward[t] + (gamma * V[t+1]) * m - V[t]
 *   adv[t] = delta + ((gamma*lambda) * m) * adv[t+1],  adv[T] = 0
 *   ret[t] = adv[t] + V[t]
 * V[t+1] comes from next_value[t] when next_value != NULL (PPO: critic(next_state)), otherwise
 * from value[t+1] with last_value[B] as V[T] (for a [T+1,B] value array pass value+T*B).
 * done, ret may be NULL. */
int mrl_gae(const float *reward, const float *value, const float *next_value,
            const float *last_value, const uint8_t *done, float *adv, float *ret, int64_t T,
            int64_t B, float gamma, float lambda, int32_t layout, int32_t mode, void *stream);
int mrl_gae_host(const float *reward, const float *value, const float *next_value,
                 const float *last_value, const uint8_t *done, float *adv, float *ret, int64_t T,
                 int64_t B, float gamma, float lambda, int32_t layout, int32_t mode, void *stream);
/* advantage normalisation that follows GAE in PPO / A2C (ppo.py:361-368):
 *   out = (x - mean(x)) / (sqrt(var(x)) + eps), mean and population variance over all n elements,
 * accumulated in float64 in a fixed order (deterministic).  out may alias x. */
int mrl_normalize(const float *x, float *out, int64_t n, float eps, void *stream);

/* The scan and the normalisation that follows it in the learners as ONE operation:
 *   PPO  (ppo.py:353-368): advantage = gae(...); (advantage - mean) / (sqrt(var) + eps)
 *   A2C  (a2c.py:189-192): returns = discount_return(...); (returns - mean) / (sqrt(var) + eps)
 * The moments are accumulated by the scan kernel while it stores its results (float64, fixed order: one
 * partial per CTA), so normalising costs one more pass instead of three.  adv / out receive the
 * NORMALISED values (ret of GAE stays un-normalised, as in PPO).  Shapes without a moment-producing
 * kernel (short, narrow or unaligned ones, MRL_SCAN_SEQUENTIAL) take the moments in a separate pass:
 * same results. */
int mrl_gae_normalized(const float *reward, const float *value, const float *next_value,
                       const float *last_value, const uint8_t *done, float *adv, float *ret, int64_t T,
                       int64_t B, float gamma, float lambda, float eps, int32_t layout, int32_t mode,
                       void *stream);
int mrl_discounted_return_normalized(const float *reward, const uint8_t *done, const float *last_value,
                                     float *out, int64_t T, int64_t B, int64_t E, float gamma, float eps,
                                     int32_t layout, int32_t mode, void *stream);

/* V-trace targets of IMPALA (algorithm/impala/vtrace.py:48-108, get_vs_and_advantages), time-major [T,B]
 * float arrays, bootstrap_value [B]; masks may be NULL (= 1); a clip threshold of +infinity disables that
 * clip.  Writes vs and pg_advantages.  MRL_SCAN_SEQUENTIAL: one kernel in the reference's operation order
 * (only exp() differs from a CPU libm by its last bit); MRL_SCAN_PARALLEL: prep, parallel scan, post;
 * MRL_SCAN_AUTO picks by shape. */
int mrl_vtrace(const float *log_rhos, const float *discounts, const float *rewards, const float *values,
               const float *bootstrap_value, const float *masks, float clip_rho_threshold,
               float clip_pg_threshold, float clip_cs_threshold, float *vs, float *pg_advantages, int64_t T,
               int64_t B, int32_t mode, void *stream);

/* TD(lambda) targets of COMA (algorithm/coma/coma.py:489-501, build_td_lambda_targets), batch-major:
 *   ret[:, t] = lambda_gamma * ret[:, t+1]
 *               + mask[:, t] * (rewards[:, t] + bootstrap_gamma * target_qs[:, t+1] * (1 - terminated[:, t]))
 * for t = max_t-1 .. 0, ret[:, t >= max_t] = 0.  target_qs / ret [B, Tq, A]; rewards, terminated, mask
 * [B, Tr, RA], RA = 1 (broadcast over agents) or A.  lambda_gamma = (float)(td_lambda * gamma) and
 * bootstrap_gamma = (float)((1 - td_lambda) * gamma), formed in double as the reference's Python scalars
 * are.  Same operation order as the reference loop: bit-identical. */
int mrl_td_lambda(const float *rewards, const float *terminated, const float *mask, const float *target_qs,
                  float *ret, int64_t B, int64_t Tq, int64_t Tr, int64_t A, int64_t RA, int64_t max_t,
                  float lambda_gamma, float bootstrap_gamma, void *stream);

/* generic first-order reverse recurrence x[t] = a[t] + b[t] * x[t+1], x[T] = x_last (NULL = 0):
 * V-trace (vtrace.py:88-93, b = discount*c*mask) and TD(lambda) (coma.py:489-501). */
int mrl_affine_scan(const float *a, const float *b, const float *x_last, float *out, int64_t T,
                    int64_t B, int32_t layout, int32_t mode, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* MINDRL_B200_H_ */
