/*
 * mrl_oracle.c -- CPU ORACLE for the experience data path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  Nothing under mindrl_b200/ imports, links or executes it.
 *
 * What it restates (paths relative to /root/reference):
 *   scans     -- mindspore_rl/utils/discounted_return.py:84-92 (in-repo CPU loop),
 *                mindspore_rl/algorithm/ppo/ppo.py:320-350, mindspore_rl/algorithm/mappo/mappo.py:478-497,
 *                mindspore_rl/algorithm/impala/vtrace.py:88-93.  PINNED by the two known-answer
 *                vectors of tests/st/test_discounted_return.py:39-53 and by golden vectors produced
 *                by running the reference's own Python loop (tests/golden/make_golden.py).
 *   uniform   -- the contract of mindspore_rl/core/uniform_replay_buffer.py:75-172 whose arithmetic
 *                lives in MindSpore 2.2 (P.BufferAppend/BufferGetItem/BufferSample; un-vendored,
 *                not installable here).  PARITY UNPINNED by the reference: its only test is never
 *                collected (tests/st/test_uniform_replay_buffer.py:58).  Cursor / logical-index
 *                rules follow MindSpore's published CPU kernels as recalled in SURVEY.md 8(a2-a4).
 *   priority  -- the contract of mindspore_rl/core/priority_replay_buffer.py:60-138 (MindSpore 2.2
 *                _rl_inner_ops.PriorityReplayBuffer*, un-vendored).  PARITY UNPINNED by the reference:
 *                its test is skipped and calls a method that does not exist
 *                (tests/st/test_priority_replay_buffer.py:26,46).  The algorithm restated is the
 *                published one: proportional prioritisation with a sum/min segment tree
 *                (Schaul et al. 2015, cited at priority_replay_buffer.py:33): root at node 1,
 *                children 2i/2i+1, leaves padded to a power of two with {sum 0, min FLT_MAX},
 *                max_priority starts at 1, new rows enter with max_priority, stratified draw
 *                mass = (u_i + i) * (sum / n), descent `mass <= left ? left : (mass -= left, right)`,
 *                weight = (p/sum*size)^-beta / (min/sum*size)^-beta.  The skipped test's
 *                behavioural expectations (indices < size, rows match indices, down-weighted rows are
 *                not resampled) are checked in tests/test_oracle.py.
 *
 * All arithmetic is sequential fp32 with one rounding per operation (build with
 * -ffp-contract=off); pow goes through include/mrl_arith.h (see there for why).
 */
#include <float.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "../include/mrl_arith.h"

#define ORC_MAX_COLS 32

/* ------------------------------------------------------------------------------------------
 * UniformReplayBuffer
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  int64_t capacity;
  int ncols;
  int64_t row_bytes[ORC_MAX_COLS];
  uint8_t *cols[ORC_MAX_COLS];
  int32_t count, head;
  uint64_t seed, ctr;
} orc_urb;

void *orc_urb_create(int64_t capacity, int32_t ncols, const int64_t *row_bytes, uint64_t seed) {
  if (capacity <= 0 || ncols <= 0 || ncols > ORC_MAX_COLS) return NULL;
  orc_urb *b = (orc_urb *)calloc(1, sizeof(orc_urb));
  b->capacity = capacity;
  b->ncols = ncols;
  b->seed = seed;
  for (int i = 0; i < ncols; ++i) {
    b->row_bytes[i] = row_bytes[i];
    b->cols[i] = (uint8_t *)calloc((size_t)capacity, (size_t)row_bytes[i]); /* zeros, :26-47 */
  }
  return b;
}

void orc_urb_destroy(void *h) {
  orc_urb *b = (orc_urb *)h;
  if (!b) return;
  for (int i = 0; i < b->ncols; ++i) free(b->cols[i]);
  free(b);
}

void *orc_urb_column(void *h, int32_t i) { return ((orc_urb *)h)->cols[i]; }

/* BufferAppend (uniform_replay_buffer.py:114; cursor rule SURVEY.md 8 a2) */
int orc_urb_append(void *h, int64_t nrows, const void *const *src_cols) {
  orc_urb *b = (orc_urb *)h;
  if (nrows <= 0) return 1;
  const int64_t cap = b->capacity;
  int64_t pos;
  if (b->head == 0 && b->count < cap) {
    pos = b->count;
    int64_t c = (int64_t)b->count + nrows;
    if (c > cap) {
      b->head = (int32_t)((c - cap) % cap);
      c = cap;
    }
    b->count = (int32_t)c;
  } else {
    pos = b->head;
    b->head = (int32_t)((b->head + nrows) % cap);
  }
  for (int i = 0; i < b->ncols; ++i) {
    const int64_t rb = b->row_bytes[i];
    const uint8_t *src = (const uint8_t *)src_cols[i];
    for (int64_t j = 0; j < nrows; ++j) { /* in order: when nrows > cap the last writer wins */
      memcpy(b->cols[i] + ((pos + j) % cap) * rb, src + j * rb, (size_t)rb);
    }
  }
  return 0;
}

/* BufferGetItem (:128): logical index -> physical slot */
int64_t orc_urb_logical_to_slot(void *h, int64_t index) {
  orc_urb *b = (orc_urb *)h;
  if (index < 0) index += b->count;
  if (index < 0 || index > b->capacity - 1) return -1;
  int64_t t = (int64_t)b->count - b->head;
  if (index < t) return index + b->head;
  return index - t;
}

int orc_urb_get_item(void *h, int64_t index, void *const *out_cols) {
  orc_urb *b = (orc_urb *)h;
  int64_t slot = orc_urb_logical_to_slot(h, index);
  if (slot < 0) return 1;
  for (int i = 0; i < b->ncols; ++i)
    memcpy(out_cols[i], b->cols[i] + slot * b->row_bytes[i], (size_t)b->row_bytes[i]);
  return 0;
}

int orc_urb_gather(void *h, const int64_t *slots, int64_t n, void *const *out_cols) {
  orc_urb *b = (orc_urb *)h;
  for (int64_t j = 0; j < n; ++j)
    for (int i = 0; i < b->ncols; ++i) {
      const int64_t rb = b->row_bytes[i];
      memcpy((uint8_t *)out_cols[i] + j * rb, b->cols[i] + slots[j] * rb, (size_t)rb);
    }
  return 0;
}

typedef struct {
  orc_urb *b;
  const int64_t *slots;
  int64_t lo, hi;
  void *const *out_cols;
} orc_gather_job;

static void *orc_gather_worker(void *p) {
  orc_gather_job *j = (orc_gather_job *)p;
  for (int64_t s = j->lo; s < j->hi; ++s)
    for (int i = 0; i < j->b->ncols; ++i) {
      const int64_t rb = j->b->row_bytes[i];
      memcpy((uint8_t *)j->out_cols[i] + s * rb, j->b->cols[i] + j->slots[s] * rb, (size_t)rb);
    }
  return NULL;
}

/* the upstream CPU kernel splits the batch loop over its thread pool */
int orc_urb_gather_mt(void *h, const int64_t *slots, int64_t n, void *const *out_cols,
                      int32_t nthreads) {
  if (nthreads <= 1 || n < 2 * nthreads) return orc_urb_gather(h, slots, n, out_cols);
  if (nthreads > 256) nthreads = 256;
  pthread_t th[256];
  orc_gather_job jobs[256];
  for (int t = 0; t < nthreads; ++t) {
    jobs[t].b = (orc_urb *)h;
    jobs[t].slots = slots;
    jobs[t].lo = n * t / nthreads;
    jobs[t].hi = n * (t + 1) / nthreads;
    jobs[t].out_cols = out_cols;
    pthread_create(&th[t], NULL, orc_gather_worker, &jobs[t]);
  }
  for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
  return 0;
}

/* BufferSample (:139): n slots uniformly with replacement from [0,count) (physical slots are
 * drawn directly, no head offset), then gather.  slots_out may be NULL. */
int orc_urb_sample(void *h, int64_t n, int64_t *slots_out, void *const *out_cols) {
  orc_urb *b = (orc_urb *)h;
  if (n <= 0) return 1;
  if ((b->head > 0 && n > b->capacity) || (b->head == 0 && n > b->count)) return 4;
  int64_t *slots = slots_out ? slots_out : (int64_t *)malloc(sizeof(int64_t) * (size_t)n);
  for (int64_t j = 0; j < n; ++j) slots[j] = mrl_uniform_below(b->seed, b->ctr + (uint64_t)j, b->count);
  b->ctr += (uint64_t)n;
  orc_urb_gather(h, slots, n, out_cols);
  if (!slots_out) free(slots);
  return 0;
}

void orc_urb_reset(void *h) { ((orc_urb *)h)->count = 0; } /* :141-150: head untouched */

void orc_urb_state(void *h, int32_t *count, int32_t *head) {
  *count = ((orc_urb *)h)->count;
  *head = ((orc_urb *)h)->head;
}

/* ------------------------------------------------------------------------------------------
 * PriorityReplayBuffer
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  int64_t capacity, cap2; /* cap2 = leaves, power of two >= capacity */
  int levels;             /* log2(cap2) */
  float alpha;
  int ncols;
  int64_t row_bytes[ORC_MAX_COLS];
  uint8_t *cols[ORC_MAX_COLS];
  float *sum, *mn; /* 2*cap2 nodes, root at 1 */
  float max_priority;
  int64_t size, head; /* head = slot of the newest row, -1 when empty */
  uint64_t seed, ctr;
} orc_prb;

void *orc_prb_create(int64_t capacity, float alpha, int32_t ncols, const int64_t *row_bytes,
                     uint64_t seed0, uint64_t seed1) {
  if (capacity <= 0 || ncols <= 0 || ncols > ORC_MAX_COLS) return NULL;
  orc_prb *b = (orc_prb *)calloc(1, sizeof(orc_prb));
  b->capacity = capacity;
  b->cap2 = 1;
  b->levels = 0;
  while (b->cap2 < capacity) {
    b->cap2 *= 2;
    b->levels++;
  }
  b->alpha = alpha;
  b->ncols = ncols;
  for (int i = 0; i < ncols; ++i) {
    b->row_bytes[i] = row_bytes[i];
    b->cols[i] = (uint8_t *)calloc((size_t)capacity, (size_t)row_bytes[i]);
  }
  b->sum = (float *)malloc(sizeof(float) * 2 * (size_t)b->cap2);
  b->mn = (float *)malloc(sizeof(float) * 2 * (size_t)b->cap2);
  for (int64_t i = 0; i < 2 * b->cap2; ++i) {
    b->sum[i] = 0.0f;
    b->mn[i] = FLT_MAX;
  }
  b->max_priority = 1.0f;
  b->size = 0;
  b->head = -1;
  b->seed = mrl_mix_seed(seed0, seed1);
  b->ctr = 0;
  return b;
}

void orc_prb_destroy(void *h) {
  orc_prb *b = (orc_prb *)h;
  if (!b) return;
  for (int i = 0; i < b->ncols; ++i) free(b->cols[i]);
  free(b->sum);
  free(b->mn);
  free(b);
}

/* segment-tree point update: set the leaf, recompute every ancestor from its two children */
static void orc_tree_set(orc_prb *b, int64_t leaf, float p) {
  int64_t idx = leaf + b->cap2;
  b->sum[idx] = p;
  b->mn[idx] = p;
  idx >>= 1;
  while (idx >= 1) {
    const float l = b->sum[2 * idx], r = b->sum[2 * idx + 1];
    b->sum[idx] = l + r;
    const float lm = b->mn[2 * idx], rm = b->mn[2 * idx + 1];
    b->mn[idx] = lm < rm ? lm : rm;
    idx >>= 1;
  }
}

/* PriorityReplayBufferPush (priority_replay_buffer.py:78-90) */
int orc_prb_push(void *h, const void *const *row_cols) {
  orc_prb *b = (orc_prb *)h;
  b->head = (b->head + 1) % b->capacity;
  if (b->size < b->capacity) b->size++;
  for (int i = 0; i < b->ncols; ++i)
    memcpy(b->cols[i] + b->head * b->row_bytes[i], row_cols[i], (size_t)b->row_bytes[i]);
  orc_tree_set(b, b->head, b->max_priority);
  return 0;
}

int orc_prb_push_batch(void *h, int64_t nrows, const void *const *src_cols) {
  orc_prb *b = (orc_prb *)h;
  const void *row[ORC_MAX_COLS];
  for (int64_t j = 0; j < nrows; ++j) {
    for (int i = 0; i < b->ncols; ++i) row[i] = (const uint8_t *)src_cols[i] + j * b->row_bytes[i];
    orc_prb_push(h, row);
  }
  return 0;
}

static int64_t orc_tree_descend(const orc_prb *b, float mass) {
  int64_t idx = 1;
  while (idx < b->cap2) {
    const float left = b->sum[2 * idx];
    if (mass <= left) {
      idx = 2 * idx;
    } else {
      mass = mass - left;
      idx = 2 * idx + 1;
    }
  }
  return idx - b->cap2;
}

/* PriorityReplayBufferSample (:92-106).  gstats == NULL: weights against the local tree.
 * gstats = world x {sum, min, size, max_priority}: weights against the global totals
 * (sharded buffer): Sum = sum over ranks in rank order, Min = min over ranks,
 * Size = sum of sizes. */
/* sharded buffers: 0 = weights against the global sum / minimum / size; 1 = against the
 * probability a sample is really drawn with, p_i / (world * S_rank), normalised by the largest such weight
 * over all ranks (the first form is biased when shards differ in priority mass) */
static int orc_exact_shard_weights = 0;
void orc_set_shard_weights(int exact) { orc_exact_shard_weights = exact; }

static int orc_prb_sample_impl(orc_prb *b, int64_t n, float beta, const float *uniforms,
                               const float *gstats, int32_t world, int64_t *indices,
                               float *weights, void *const *out_cols) {
  if (n <= 0 || b->size <= 0) return 1; /* nothing to sample from */
  const float root_sum = b->sum[1];
  float w_sum = root_sum, w_min = b->mn[1], w_size = (float)b->size, m_sum = root_sum;
  if (gstats) {
    float t_sum = 0.0f, t_min = FLT_MAX, best = FLT_MAX, star_min = FLT_MAX, star_sum = 1.0f;
    w_size = 0.0f;
    for (int r = 0; r < world; ++r) {
      const float sr = gstats[4 * r + 0], m = gstats[4 * r + 1];
      t_sum = t_sum + sr;
      t_min = m < t_min ? m : t_min;
      w_size = w_size + gstats[4 * r + 2];
      const float den = (float)world * sr;
      const float ratio = m / den;
      if (sr > 0.0f && ratio < best) {
        best = ratio;
        star_min = m;
        star_sum = den;
      }
    }
    if (orc_exact_shard_weights) {
      w_sum = (float)world * root_sum;
      w_min = star_min;
      m_sum = star_sum;
    } else {
      w_sum = t_sum;
      w_min = t_min;
      m_sum = t_sum;
    }
  }
  const float max_w = mrl_is_weight(w_min, m_sum, w_size, beta);
  const float seg = root_sum / (float)n;
  for (int64_t i = 0; i < n; ++i) {
    const float u = uniforms ? uniforms[i] : mrl_uniform01(b->seed, b->ctr + (uint64_t)i);
    const float mass = (u + (float)i) * seg;
    int64_t leaf = orc_tree_descend(b, mass);
    /* a mass that rounds past the total walks into the empty part of the tree (leaves >= size,
     * priority 0): the newest valid leaf is returned instead, so that index, weight and row
     * belong together (decision of this oracle, documented in DESIGN.md section 3) */
    if (leaf >= b->size) leaf = b->size - 1;
    indices[i] = leaf;
    const float p = b->sum[leaf + b->cap2];
    weights[i] = mrl_is_weight(p, w_sum, w_size, beta) / max_w;
    if (out_cols) {
      const int64_t slot = leaf;
      for (int c = 0; c < b->ncols; ++c) {
        const int64_t rb = b->row_bytes[c];
        memcpy((uint8_t *)out_cols[c] + i * rb, b->cols[c] + slot * rb, (size_t)rb);
      }
    }
  }
  if (!uniforms) b->ctr += (uint64_t)n;
  return 0;
}

int orc_prb_sample(void *h, int64_t n, float beta, const float *uniforms, int64_t *indices,
                   float *weights, void *const *out_cols) {
  return orc_prb_sample_impl((orc_prb *)h, n, beta, uniforms, NULL, 0, indices, weights, out_cols);
}

int orc_prb_sample_sharded(void *h, int64_t n, float beta, const float *uniforms,
                           const float *gstats, int32_t world, int64_t *indices, float *weights,
                           void *const *out_cols) {
  return orc_prb_sample_impl((orc_prb *)h, n, beta, uniforms, gstats, world, indices, weights,
                             out_cols);
}

/* PriorityReplayBufferUpdate (:108-120): sequential, so duplicates resolve last-writer-wins */
int orc_prb_update(void *h, const int64_t *indices, const float *priorities, int64_t n) {
  orc_prb *b = (orc_prb *)h;
  for (int64_t k = 0; k < n; ++k) {
    const float p = mrl_priority_pow(priorities[k], b->alpha);
    orc_tree_set(b, indices[k], p);
    if (p > b->max_priority) b->max_priority = p;
  }
  return 0;
}

void orc_prb_info(void *h, int64_t *size, int64_t *head, int64_t *leaf_capacity) {
  orc_prb *b = (orc_prb *)h;
  *size = b->size;
  *head = b->head;
  *leaf_capacity = b->cap2;
}

float orc_prb_max_priority(void *h) { return ((orc_prb *)h)->max_priority; }
float *orc_prb_sum_nodes(void *h) { return ((orc_prb *)h)->sum; }
float *orc_prb_min_nodes(void *h) { return ((orc_prb *)h)->mn; }
void *orc_prb_column(void *h, int32_t i) { return ((orc_prb *)h)->cols[i]; }

void orc_prb_root_stats(void *h, float *stats4) {
  orc_prb *b = (orc_prb *)h;
  stats4[0] = b->sum[1];
  stats4[1] = b->mn[1];
  stats4[2] = (float)b->size;
  stats4[3] = b->max_priority;
}

/* ------------------------------------------------------------------------------------------
 * reverse scans
 * layout 0: time-major [T,B,E] (done [T,B]); layout 1: batch-major [B,T] (E == 1)
 * ---------------------------------------------------------------------------------------- */
static inline int64_t orc_at(int32_t layout, int64_t t, int64_t col, int64_t T, int64_t N) {
  return layout == 0 ? t * N + col : col * T + t;
}

/* discounted_return.py:84-92:
 *   last_state_value = reward[step] + (1 - done[step]) * self.gamma * last_state_value */
int orc_discounted_return(const float *reward, const uint8_t *done, const float *last_value,
                          float *out, int64_t T, int64_t B, int64_t E, float gamma,
                          int32_t layout) {
  if (layout == 1 && E != 1) return 1;
  const int64_t N = B * E;
  for (int64_t col = 0; col < N; ++col) {
    const int64_t b = col / E;
    float last = last_value ? last_value[col] : 0.0f;
    for (int64_t t = T - 1; t >= 0; --t) {
      const float nd = done ? 1.0f - (float)(done[orc_at(layout, t, b, T, B)] != 0) : 1.0f;
      const float c = nd * gamma;
      const float prod = c * last;
      last = reward[orc_at(layout, t, col, T, N)] + prod;
      out[orc_at(layout, t, col, T, N)] = last;
    }
  }
  return 0;
}

/* ppo.py:334-350 (done == NULL, next_value given) and mappo.py:478-497 (mask = 1 - done):
 *   delta = reward + gamma * V' * m - V ; gae = delta + gamma * lambda * m * gae ; ret = gae + V */
int orc_gae(const float *reward, const float *value, const float *next_value,
            const float *last_value, const uint8_t *done, float *adv, float *ret, int64_t T,
            int64_t B, float gamma, float lambda, int32_t layout) {
  const float gl = gamma * lambda;
  for (int64_t col = 0; col < B; ++col) {
    float g = 0.0f;
    for (int64_t t = T - 1; t >= 0; --t) {
      const int64_t i = orc_at(layout, t, col, T, B);
      const float m = done ? 1.0f - (float)(done[i] != 0) : 1.0f;
      float nv;
      if (next_value) nv = next_value[i];
      else if (t == T - 1) nv = last_value ? last_value[col] : 0.0f;
      else nv = value[orc_at(layout, t + 1, col, T, B)];
      const float v = value[i];
      const float gv = gamma * nv;
      const float gvm = gv * m;
      const float rp = reward[i] + gvm;
      const float delta = rp - v;
      const float c = gl * m;
      const float cg = c * g;
      g = delta + cg;
      adv[i] = g;
      if (ret) ret[i] = g + v;
    }
  }
  return 0;
}

/* x[t] = a[t] + b[t] * x[t+1]  (vtrace.py:88-93 with b = discount*c*mask; coma.py:489-501) */
int orc_affine_scan(const float *a, const float *b, const float *x_last, float *out, int64_t T,
                    int64_t B, int32_t layout) {
  for (int64_t col = 0; col < B; ++col) {
    float x = x_last ? x_last[col] : 0.0f;
    for (int64_t t = T - 1; t >= 0; --t) {
      const int64_t i = orc_at(layout, t, col, T, B);
      const float prod = b[i] * x;
      x = a[i] + prod;
      out[i] = x;
    }
  }
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * helpers
 * ---------------------------------------------------------------------------------------- */
float orc_detpowf(float x, float y) { return mrl_detpowf(x, y); }
float orc_libm_powf_via_double(float x, float y) { return (float)pow((double)x, (double)y); }
void orc_detpowf_array(const float *x, const float *y, float *out, int64_t n) {
  for (int64_t i = 0; i < n; ++i) out[i] = mrl_detpowf(x[i], y[i]);
}
void orc_libm_pow_array(const float *x, const float *y, float *out, int64_t n) {
  for (int64_t i = 0; i < n; ++i) out[i] = (float)pow((double)x[i], (double)y[i]);
}
float orc_uniform01(uint64_t seed, uint64_t ctr) { return mrl_uniform01(seed, ctr); }
uint64_t orc_mix_seed(uint64_t s0, uint64_t s1) { return mrl_mix_seed(s0, s1); }

void orc_fill_rows(void *col, int64_t nrows, int64_t row_bytes, uint64_t seed) {
  uint8_t *p = (uint8_t *)col;
  for (int64_t r = 0; r < nrows; ++r) {
    for (int64_t o = 0; o < row_bytes; ++o) p[r * row_bytes + o] = mrl_fill_byte(seed, r, o);
    if (row_bytes >= 8) memcpy(p + r * row_bytes, &r, 8);
  }
}

/* the same rows as orc_fill_rows, but only those listed (full-size buffers are verified row by row) */
void orc_fill_rows_at(void *out, const int64_t *rows, int64_t n, int64_t row_bytes, uint64_t seed) {
  uint8_t *p = (uint8_t *)out;
  for (int64_t k = 0; k < n; ++k) {
    const int64_t r = rows[k];
    for (int64_t o = 0; o < row_bytes; ++o) p[k * row_bytes + o] = mrl_fill_byte(seed, r, o);
    if (row_bytes >= 8) memcpy(p + k * row_bytes, &r, 8);
  }
}

double orc_now(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* ------------------------------------------------------------------------------------------
 * timed loops for bench.py's cpu_baseline / --impl reference legs: the same oracle functions
 * as above, called from C so that the CPU number carries no Python-wrapper overhead
 * (C1's CPU leg was timed through the wrapper).  Return seconds (monotonic clock).
 * ---------------------------------------------------------------------------------------- */
/* `steps` iterations of { sample(batch, beta) -> update_priorities(indices, prio[step % pool]) } */
double orc_prb_bench_loop(void *h, int64_t steps, int64_t batch, float beta, const float *prio,
                          int64_t pool, int64_t *indices, float *weights, void *const *out_cols) {
  const double t0 = orc_now();
  for (int64_t s = 0; s < steps; ++s) {
    orc_prb_sample(h, batch, beta, NULL, indices, weights, out_cols);
    orc_prb_update(h, indices, prio + (s % pool) * batch, batch);
  }
  return orc_now() - t0;
}

/* `steps` iterations of { append one row (rows[step % pool]) -> sample(batch) }; the gather half of the
 * sample is split over `nthreads` threads when nthreads > 1 (as the upstream CPU kernel does) */
double orc_urb_bench_loop(void *h, int64_t steps, int64_t batch, const void *const *row_pool, int64_t pool,
                          int64_t *slots, void *const *out_cols, int32_t nthreads) {
  orc_urb *b = (orc_urb *)h;
  const void *row[ORC_MAX_COLS];
  const double t0 = orc_now();
  for (int64_t s = 0; s < steps; ++s) {
    for (int i = 0; i < b->ncols; ++i)
      row[i] = (const uint8_t *)row_pool[i] + (s % pool) * b->row_bytes[i];
    orc_urb_append(h, 1, row);
    if (nthreads <= 1) {
      orc_urb_sample(h, batch, slots, out_cols);
    } else {
      for (int64_t j = 0; j < batch; ++j) slots[j] = mrl_uniform_below(b->seed, b->ctr + (uint64_t)j, b->count);
      b->ctr += (uint64_t)batch;
      orc_urb_gather_mt(h, slots, batch, out_cols, nthreads);
    }
  }
  return orc_now() - t0;
}

/* `reps` gathers of n rows at the given slots over nthreads threads (wide rows: the all-cores CPU leg) */
double orc_urb_gather_bench(void *h, const int64_t *slots, int64_t n, void *const *out_cols, int32_t nthreads,
                            int64_t reps) {
  const double t0 = orc_now();
  for (int64_t r = 0; r < reps; ++r) orc_urb_gather_mt(h, slots, n, out_cols, nthreads);
  return orc_now() - t0;
}

/* ------------------------------------------------------------------------------------------
 * TD(lambda) targets of COMA, mindspore_rl/algorithm/coma/coma.py:489-501 (build_td_lambda_targets):
 *   ret[:, t] = td_lambda * gamma * ret[:, t + 1] + mask[:, t] * (
 *       rewards[:, t] + (1 - td_lambda) * gamma * target_qs[:, t + 1] * (1 - terminated[:, t]))
 * gamma and td_lambda are Python floats there (coma.py:395-396): the two products of scalars are
 * formed in double and meet the float32 tensors as float32 scalars (lg, bg).
 * PINNED by tests/golden/coma_td_lambda_ref.npz (the reference's own method, run by make_golden.py).
 * ---------------------------------------------------------------------------------------- */
int orc_td_lambda(const float *rewards, const float *terminated, const float *mask, const float *q,
                  float *ret, int64_t B, int64_t Tq, int64_t Tr, int64_t A, int64_t RA, int64_t max_t,
                  double gamma, double td_lambda) {
  if (max_t >= Tq || max_t > Tr || (RA != 1 && RA != A)) return 1;
  const float lg = (float)(td_lambda * gamma), bg = (float)((1.0 - td_lambda) * gamma);
  for (int64_t b = 0; b < B; ++b)
    for (int64_t a = 0; a < A; ++a) {
      for (int64_t t = max_t; t < Tq; ++t) ret[(b * Tq + t) * A + a] = 0.0f; /* F.zeros_like(target_qs) */
      float x = 0.0f;
      for (int64_t t = max_t - 1; t >= 0; --t) {
        const int64_t i = (b * Tr + t) * RA + (RA == 1 ? 0 : a);
        const float boot = (bg * q[(b * Tq + t + 1) * A + a]) * (1.0f - terminated[i]);
        const float inner = rewards[i] + boot;
        x = lg * x + mask[i] * inner;
        ret[(b * Tq + t) * A + a] = x;
      }
    }
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * V-trace, mindspore_rl/algorithm/impala/vtrace.py:79-105, time-major [T,B]; thresholds may be
 * INFINITY.  exp() is libm's expf here and CUDA's on the device: they may differ in the last bit,
 * so V-trace parity is held to the scans' tolerance, not bit for bit.
 * PINNED by tests/golden/vtrace_ref.npz (the reference's vtrace.py imported whole by make_golden.py).
 * ---------------------------------------------------------------------------------------- */
int orc_vtrace(const float *log_rhos, const float *discounts, const float *rewards, const float *values,
               const float *bootstrap, const float *masks, float clip_rho, float clip_pg, float clip_c,
               float *vs, float *pg, int64_t T, int64_t B) {
  for (int64_t col = 0; col < B; ++col) {
    float acc = 0.0f, v_next = bootstrap[col], vs_next = bootstrap[col];
    for (int64_t t = T - 1; t >= 0; --t) {
      const int64_t i = t * B + col;
      float rho = expf(log_rhos[i]);                    /* :79 */
      rho = rho < clip_rho ? rho : clip_rho;            /* :80 */
      const float c = rho < clip_c ? rho : clip_c;      /* :81 */
      const float pg_rho = rho < clip_pg ? rho : clip_pg; /* :82 */
      const float dv = discounts[i] * v_next;
      const float delta = rho * ((rewards[i] + dv) - values[i]); /* :86 */
      const float m = masks ? masks[i] : 1.0f;
      const float dc = discounts[i] * c;
      acc = delta + (dc * acc) * m;                     /* :91 */
      const float v = acc + values[i];                  /* :99 */
      vs[i] = v;
      const float dn = discounts[i] * vs_next;
      pg[i] = pg_rho * ((rewards[i] + dn) - values[i]); /* :105 */
      v_next = values[i];
      vs_next = v;
    }
  }
  return 0;
}
