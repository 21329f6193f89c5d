// urb.cu -- UniformReplayBuffer ops: BufferAppend / BufferGetItem / BufferSample
// (mindspore_rl/core/uniform_replay_buffer.py:75-172).
//
// Layout in HBM: one column per transition field, [capacity, row_bytes[i]] bytes, row-contiguous
// (exactly the `.buffer` ParameterTuple of the reference).  The cursor pair (count, head) is
// deterministic in the number of rows appended, so it is kept on the host and every launch gets
// plain integers: no device round trip on insert/get_item/sample.
#include <string.h>

#include "common.cuh"

namespace mrl {

struct UrbBuffer {
  int64_t capacity = 0;
  int32_t ncols = 0;
  int64_t row_bytes[MRL_MAX_COLS];
  uint8_t *cols[MRL_MAX_COLS];
  bool owned = false;
  int32_t count = 0, head = 0;
  uint64_t seed = 0, ctr = 0;
  int64_t *slots_dev = nullptr;  // scratch for sample()
  int64_t slots_cap = 0;
  Staging stage;
};

static Registry<UrbBuffer> g_urb;

int Staging::ensure(size_t need) {
  if (need <= bytes) return MRL_OK;
  release();
  size_t cap = 1 << 16;
  while (cap < need) cap <<= 1;
  MRL_CUDA(cudaMallocHost((void **)&host, cap));
  MRL_CUDA(cudaMalloc((void **)&dev, cap));
  MRL_CUDA(cudaEventCreateWithFlags(&done, cudaEventDisableTiming));
  bytes = cap;
  return MRL_OK;
}
int Staging::wait() {
  if (done && pending) MRL_CUDA(cudaEventSynchronize(done));  // (~2 us even when long complete: skipped when idle)
  pending = false;
  return MRL_OK;
}
void Staging::release() {
  if (host) cudaFreeHost(host);
  if (dev) cudaFree(dev);
  if (done) cudaEventDestroy(done);
  host = dev = nullptr;
  done = nullptr;
  pending = false;
  bytes = 0;
}

static void fill_table(ColTable &tab, int32_t ncols, uint8_t *const *cols, const int64_t *row_bytes,
                       void *const *io) {
  tab.ncols = ncols;
  for (int c = 0; c < ncols; ++c) {
    tab.col[c] = cols[c];
    tab.io[c] = (uint8_t *)io[c];
    tab.row_bytes[c] = row_bytes[c];
    tab.vec[c] = align_unit((uintptr_t)cols[c], (uintptr_t)io[c], (uint64_t)row_bytes[c]);
  }
}

// Philox-free counter generator: slot j of a draw = mrl_uniform_below(seed, ctr + j, count)
__global__ void __launch_bounds__(256)
    urb_draw_kernel(int64_t *slots, int64_t n, uint64_t seed, uint64_t ctr, int64_t count) {
  pdl_prologue();
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n) slots[j] = mrl_uniform_below(seed, ctr + (uint64_t)j, count);
}

// BufferSample for narrow rows in ONE launch: warp j draws its slot and copies the row of every column
__global__ void __launch_bounds__(256)
    urb_sample_fused_kernel(const __grid_constant__ ColTable tab, int64_t *slots_out, int64_t n,
                            uint64_t seed, uint64_t ctr, int64_t count, int word_units) {
  pdl_prologue();  // resident while the insert in front of it drains (DQN loop: insert 1 row, sample a batch)
  const int lane = threadIdx.x & 31;
  const int64_t j = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (j >= n) return;
  const int64_t slot = mrl_uniform_below(seed, ctr + (uint64_t)j, count);
  if (slots_out && lane == 0) slots_out[j] = slot;
  if (word_units > 0) {
    // every column is whole aligned words: the row is ONE list of words, 4 per lane and pass, all loads of
    // a pass before its stores (column by column each load waits for the previous column's store: the
    // pointers may alias as far as the compiler knows)
    for (int u0 = 0; u0 < word_units; u0 += 128) {
      uint32_t val[4];
      uint32_t *dstp[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        int u = u0 + q * 32 + lane;
        dstp[q] = nullptr, val[q] = 0u;
        if (u < word_units) {
          int c = 0;
          for (;; ++c) {
            const int w = (int)(tab.row_bytes[c] >> 2);
            if (u < w) break;
            u -= w;
          }
          const int64_t rb = tab.row_bytes[c];
          val[q] = *reinterpret_cast<const uint32_t *>(tab.col[c] + slot * rb + 4 * u);
          dstp[q] = reinterpret_cast<uint32_t *>(tab.io[c] + j * rb + 4 * u);
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (dstp[q]) *dstp[q] = val[q];
    }
    return;
  }
  for (int c = 0; c < tab.ncols; ++c) {
    const int64_t rb = tab.row_bytes[c];
    const uint8_t *src = tab.col[c] + slot * rb;
    uint8_t *dst = tab.io[c] + j * rb;
    if (tab.vec[c] >= 4) {
      for (int64_t o = lane * 4; o < rb; o += 128)
        *reinterpret_cast<uint32_t *>(dst + o) = *reinterpret_cast<const uint32_t *>(src + o);
    } else {
      for (int64_t o = lane; o < rb; o += 32) dst[o] = src[o];
    }
  }
}

static int urb_append_impl(UrbBuffer *b, int64_t nrows, void *const *src, cudaStream_t st, int dir = 1) {
  const int64_t cap = b->capacity;
  int64_t pos;
  // cursor rule of BufferAppend
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
  // only the last `cap` rows of an over-long batch survive (sequential last-writer-wins)
  int64_t skip = nrows > cap ? nrows - cap : 0;
  int64_t first = (pos + skip) % cap;
  int64_t n = nrows - skip;
  ColTable tab;
  fill_table(tab, b->ncols, b->cols, b->row_bytes, src);
  const int64_t run1 = n < cap - first ? n : cap - first;
  int rc = copy_segments_dir(tab, first, skip, run1, dir, st);
  if (rc) return rc;
  if (n > run1) rc = copy_segments_dir(tab, 0, skip + run1, n - run1, dir, st);
  return rc;
}

static int64_t urb_logical_to_slot(const UrbBuffer *b, int64_t index) {
  if (index < 0) index += b->count;
  if (index < 0 || index > b->capacity - 1) return -1;
  const int64_t t = (int64_t)b->count - b->head;
  return index < t ? index + b->head : index - t;
}

static int urb_ensure_slots(UrbBuffer *b, int64_t n) {
  if (n <= b->slots_cap) return MRL_OK;
  if (b->slots_dev) cudaFree(b->slots_dev);
  b->slots_dev = nullptr;
  b->slots_cap = 0;
  MRL_CUDA(cudaMalloc((void **)&b->slots_dev, sizeof(int64_t) * (size_t)n));
  b->slots_cap = n;
  return MRL_OK;
}

static int urb_sample_impl(UrbBuffer *b, int64_t n, int64_t *slots_out, void *const *out,
                           cudaStream_t st) {
  if ((b->head > 0 && n > b->capacity) || (b->head == 0 && n > b->count)) {
    set_error("The batch size %lld is larger than total buffer size %d", (long long)n, b->count);
    return MRL_ERR_STATE;
  }
  ColTable ftab;
  fill_table(ftab, b->ncols, b->cols, b->row_bytes, out);
  bool narrow = true;
  for (int c = 0; c < b->ncols; ++c) narrow = narrow && b->row_bytes[c] <= 512;
  if (narrow) {
    int64_t words = 0;
    bool whole = true;
    for (int c = 0; c < b->ncols; ++c) whole = whole && ftab.vec[c] >= 4, words += b->row_bytes[c] / 4;
    MRL_CUDA(launch_pdl(urb_sample_fused_kernel, dim3((unsigned)((n * 32 + 255) / 256)), dim3(256), 0, st, ftab, slots_out, n,
                        b->seed, b->ctr, (int64_t)b->count, whole ? (int)words : 0));
    MRL_LAUNCHED();
    b->ctr += (uint64_t)n;
    return MRL_OK;
  }
  int64_t *slots = slots_out;
  if (!slots) {
    int rc = urb_ensure_slots(b, n);
    if (rc) return rc;
    slots = b->slots_dev;
  }
  MRL_CUDA(launch_pdl(urb_draw_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, st, slots, n, b->seed, b->ctr,
                      (int64_t)b->count));
  MRL_LAUNCHED();
  b->ctr += (uint64_t)n;
  ColTable tab;
  fill_table(tab, b->ncols, b->cols, b->row_bytes, out);
  return gather_rows(tab, slots, n, b->capacity, st);
}

}  // namespace mrl

using namespace mrl;

#define URB_GET(b, h)                                        \
  UrbBuffer *b = g_urb.get(h);                               \
  if (!b) {                                                  \
    set_error("unknown UniformReplayBuffer handle %lld", (long long)(h)); \
    return MRL_ERR_HANDLE;                                   \
  }

extern "C" int mrl_urb_create(int64_t capacity, int32_t ncols, const int64_t *row_bytes,
                              void *const *col_ptrs, uint64_t seed, int64_t *handle) {
  MRL_REQUIRE(handle, "mrl_urb_create: NULL handle");
  *handle = MRL_INVALID_HANDLE;
  MRL_REQUIRE(capacity > 0 && capacity < (1LL << 31), "mrl_urb_create: capacity out of range");
  MRL_REQUIRE(ncols > 0 && ncols <= MRL_MAX_COLS && row_bytes, "mrl_urb_create: bad column table");
  UrbBuffer *b = new UrbBuffer();
  b->capacity = capacity;
  b->ncols = ncols;
  b->seed = seed;
  b->owned = col_ptrs == nullptr;
  for (int c = 0; c < ncols; ++c) {
    b->row_bytes[c] = row_bytes[c];
    b->cols[c] = nullptr;
  }
  for (int c = 0; c < ncols; ++c) {
    if (row_bytes[c] <= 0) {
      delete b;
      set_error("mrl_urb_create: row_bytes[%d] must be positive", c);
      return MRL_ERR_INVALID;
    }
    if (b->owned) {
      const size_t bytes = (size_t)capacity * (size_t)row_bytes[c];
      cudaError_t e = cudaMalloc((void **)&b->cols[c], bytes);
      if (e == cudaSuccess) e = cudaMemset(b->cols[c], 0, bytes);  // np.zeros, :26-47
      if (e != cudaSuccess) {
        for (int k = 0; k <= c; ++k)
          if (b->cols[k]) cudaFree(b->cols[k]);
        delete b;
        set_error("mrl_urb_create: allocating column %d failed: %s", c, cudaGetErrorString(e));
        return MRL_ERR_CUDA;
      }
    } else {
      b->cols[c] = (uint8_t *)col_ptrs[c];
    }
  }
  *handle = g_urb.put(b);
  return MRL_OK;
}

extern "C" int mrl_urb_destroy(int64_t handle) {
  UrbBuffer *b = g_urb.take(handle);
  if (!b) {
    set_error("unknown UniformReplayBuffer handle %lld", (long long)handle);
    return MRL_ERR_HANDLE;
  }
  if (b->owned)
    for (int c = 0; c < b->ncols; ++c) cudaFree(b->cols[c]);
  if (b->slots_dev) cudaFree(b->slots_dev);
  b->stage.release();
  delete b;
  return MRL_OK;
}

extern "C" int mrl_urb_columns(int64_t handle, void **col_ptrs_out) {
  URB_GET(b, handle);
  for (int c = 0; c < b->ncols; ++c) col_ptrs_out[c] = b->cols[c];
  return MRL_OK;
}

extern "C" int mrl_urb_append(int64_t handle, int64_t nrows, const void *const *src_cols,
                              void *stream) {
  URB_GET(b, handle);
  MRL_REQUIRE(nrows >= 1 && src_cols, "mrl_urb_append: need nrows >= 1 and a source table");
  return urb_append_impl(b, nrows, (void *const *)src_cols, as_stream(stream));
}

extern "C" int mrl_urb_append_host(int64_t handle, int64_t nrows, const void *const *src_cols,
                                   void *stream) {
  URB_GET(b, handle);
  MRL_REQUIRE(nrows >= 1 && src_cols, "mrl_urb_append_host: need nrows >= 1 and a source table");
  cudaStream_t st = as_stream(stream);
  size_t off[MRL_MAX_COLS + 1];
  off[0] = 0;
  for (int c = 0; c < b->ncols; ++c)
    off[c + 1] = off[c] + ((size_t)nrows * b->row_bytes[c] + 255) / 256 * 256;
  if (off[b->ncols] >= kDirectHostBytes) {
    // large batch: DMA from the caller's arrays straight into the ring rows; the arrays are free again
    // when this returns (staging 116 MB through a pinned copy first cost 12 ms instead of 2.3)
    int rc = urb_append_impl(b, nrows, (void *const *)src_cols, st, 2);
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
  rc = urb_append_impl(b, nrows, src_dev, st);
  if (rc) return rc;
  MRL_CUDA(b->stage.mark(st));
  return MRL_OK;
}

// Emplace of the reservoir buffer (reservoir_replay_buffer.py:69-82): overwrite the rows at PHYSICAL
// slots slot..slot+nrows-1; the cursors do not move.
extern "C" int mrl_urb_write_at(int64_t handle, int64_t slot, int64_t nrows, const void *const *src_cols,
                                void *stream) {
  URB_GET(b, handle);
  MRL_REQUIRE(src_cols && nrows >= 1 && slot >= 0 && slot + nrows <= b->capacity,
              "mrl_urb_write_at: rows [%lld, %lld) outside the buffer", (long long)slot, (long long)(slot + nrows));
  ColTable tab;
  fill_table(tab, b->ncols, b->cols, b->row_bytes, (void *const *)src_cols);
  return copy_segments_dir(tab, slot, 0, nrows, 1, as_stream(stream));
}

extern "C" int mrl_urb_write_at_host(int64_t handle, int64_t slot, int64_t nrows, const void *const *src_cols,
                                     void *stream) {
  URB_GET(b, handle);
  MRL_REQUIRE(src_cols && nrows >= 1 && slot >= 0 && slot + nrows <= b->capacity,
              "mrl_urb_write_at_host: rows [%lld, %lld) outside the buffer", (long long)slot,
              (long long)(slot + nrows));
  cudaStream_t st = as_stream(stream);
  size_t off[MRL_MAX_COLS + 1];
  off[0] = 0;
  for (int c = 0; c < b->ncols; ++c)
    off[c + 1] = off[c] + ((size_t)nrows * b->row_bytes[c] + 255) / 256 * 256;
  if (off[b->ncols] >= kDirectHostBytes) {
    ColTable htab;
    fill_table(htab, b->ncols, b->cols, b->row_bytes, (void *const *)src_cols);
    int rc = copy_segments_dir(htab, slot, 0, nrows, 2, st);
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
  ColTable tab;
  fill_table(tab, b->ncols, b->cols, b->row_bytes, src_dev);
  rc = copy_segments_dir(tab, slot, 0, nrows, 1, st);
  if (rc) return rc;
  MRL_CUDA(b->stage.mark(st));
  return MRL_OK;
}

extern "C" int mrl_urb_get_item(int64_t handle, int64_t index, void *const *out_cols,
                                void *stream) {
  URB_GET(b, handle);
  MRL_REQUIRE(out_cols, "mrl_urb_get_item: NULL output table");
  const int64_t slot = urb_logical_to_slot(b, index);
  MRL_REQUIRE(slot >= 0, "The index %lld is out of range", (long long)index);
  ColTable tab;
  fill_table(tab, b->ncols, b->cols, b->row_bytes, out_cols);
  return copy_segments_dir(tab, slot, 0, 1, 0, as_stream(stream));
}

extern "C" int mrl_urb_get_item_host(int64_t handle, int64_t index, void *const *out_cols,
                                     void *stream) {
  URB_GET(b, handle);
  MRL_REQUIRE(out_cols, "mrl_urb_get_item_host: NULL output table");
  const int64_t slot = urb_logical_to_slot(b, index);
  MRL_REQUIRE(slot >= 0, "The index %lld is out of range", (long long)index);
  cudaStream_t st = as_stream(stream);
  for (int c = 0; c < b->ncols; ++c)
    MRL_CUDA(cudaMemcpyAsync(out_cols[c], b->cols[c] + slot * b->row_bytes[c],
                             (size_t)b->row_bytes[c], cudaMemcpyDeviceToHost, st));
  MRL_CUDA(cudaStreamSynchronize(st));
  return MRL_OK;
}

extern "C" int mrl_urb_gather(int64_t handle, const int64_t *slots, int64_t n,
                              void *const *out_cols, void *stream) {
  URB_GET(b, handle);
  MRL_REQUIRE(n >= 0 && (n == 0 || (slots && out_cols)), "mrl_urb_gather: bad arguments");
  ColTable tab;
  fill_table(tab, b->ncols, b->cols, b->row_bytes, out_cols);
  return gather_rows(tab, slots, n, b->capacity, as_stream(stream));
}

extern "C" int mrl_urb_sample(int64_t handle, int64_t n, int64_t *slots_out,
                              void *const *out_cols, void *stream) {
  URB_GET(b, handle);
  MRL_REQUIRE(n >= 1 && out_cols, "mrl_urb_sample: need n >= 1 and an output table");
  return urb_sample_impl(b, n, slots_out, out_cols, as_stream(stream));
}

extern "C" int mrl_urb_sample_host(int64_t handle, int64_t n, int64_t *slots_out,
                                   void *const *out_cols, void *stream) {
  URB_GET(b, handle);
  MRL_REQUIRE(n >= 1 && out_cols, "mrl_urb_sample_host: need n >= 1 and an output table");
  cudaStream_t st = as_stream(stream);
  size_t off[MRL_MAX_COLS + 2];
  off[0] = 0;
  for (int c = 0; c < b->ncols; ++c)
    off[c + 1] = off[c] + ((size_t)n * b->row_bytes[c] + 255) / 256 * 256;
  off[b->ncols + 1] = off[b->ncols] + (size_t)n * 8;
  int rc = b->stage.wait();
  if (rc) return rc;
  rc = b->stage.ensure(off[b->ncols + 1]);
  if (rc) return rc;
  void *out_dev[MRL_MAX_COLS];
  const size_t total = off[b->ncols + 1];
  if (total <= (1u << 16)) {
    // small result: the kernels write straight into the pinned (device-mapped) staging buffer
    uint8_t *hz = b->stage.host;
    for (int c = 0; c < b->ncols; ++c) out_dev[c] = hz + off[c];
    rc = urb_sample_impl(b, n, (int64_t *)(hz + off[b->ncols]), out_dev, st);
    if (rc) return rc;
    rc = stream_wait_spin(st);
    if (rc) return rc;
    for (int c = 0; c < b->ncols; ++c)
      memcpy(out_cols[c], hz + off[c], (size_t)n * b->row_bytes[c]);
    if (slots_out) memcpy(slots_out, hz + off[b->ncols], (size_t)n * 8);
    return MRL_OK;
  }
  for (int c = 0; c < b->ncols; ++c) out_dev[c] = b->stage.dev + off[c];
  int64_t *slots_dev = (int64_t *)(b->stage.dev + off[b->ncols]);
  rc = urb_sample_impl(b, n, slots_dev, out_dev, st);
  if (rc) return rc;
  {
    for (int c = 0; c < b->ncols; ++c)
      MRL_CUDA(cudaMemcpyAsync(out_cols[c], b->stage.dev + off[c], (size_t)n * b->row_bytes[c],
                               cudaMemcpyDeviceToHost, st));
    if (slots_out)
      MRL_CUDA(cudaMemcpyAsync(slots_out, slots_dev, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
    MRL_CUDA(cudaStreamSynchronize(st));
  }
  return MRL_OK;
}

extern "C" int mrl_urb_reset(int64_t handle, void *stream) {
  (void)stream;
  URB_GET(b, handle);
  b->count = 0;  // head and data untouched (uniform_replay_buffer.py:141-150)
  return MRL_OK;
}

extern "C" int mrl_urb_state(int64_t handle, int32_t *count, int32_t *head) {
  URB_GET(b, handle);
  if (count) *count = b->count;
  if (head) *head = b->head;
  return MRL_OK;
}

extern "C" int mrl_urb_set_state(int64_t handle, int32_t count, int32_t head) {
  URB_GET(b, handle);
  MRL_REQUIRE(count >= 0 && count <= b->capacity && head >= 0 && head < b->capacity,
              "mrl_urb_set_state: out of range");
  b->count = count;
  b->head = head;
  return MRL_OK;
}
