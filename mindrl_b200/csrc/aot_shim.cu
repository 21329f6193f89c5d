// aot_shim.cu -- the experience data path behind MindSpore's "aot" custom-operator ABI.
//
// The reference mounts native code only one way:
//     ops.Custom("<lib>.so:<Func>", out_shapes, out_dtypes, "aot", reg_info=...)
// which binds
//     extern "C" int Func    (int nparam, void **params, int *ndims, int64_t **shapes, const char **dtypes,
//                             void *stream, void *extra);
//     extern "C" int FuncInit(int *ndims, int64_t **shapes, const char **dtypes, AotExtra *extra);
// (mindspore_rl/utils/mcts/cpu/cpu_mcts_ops.cc:38-53; loaded at utils/mcts/mcts.py:247-254).  params = inputs
// then outputs, device pointers on GPU; stream = cudaStream_t; 0 = ok, kErrorCode (2) = failure
// (cpu_mcts_ops.cc:25, 93); attributes through AotExtra::Attr<T>(name) (aot_extra.h).
//
// This file exports the operators of the path with exactly that signature, named after the MindSpore
// primitives they replace, so that a MindSpore build can mount libmindrl_b200.so in place of them:
//
//   BufferAppend   inputs  buffer[0..n) exp[0..n) count head          outputs count_out int32[1]
//                  (uniform_replay_buffer.py:114)
//   BufferGetItem  inputs  buffer[0..n) count head index               outputs row[0..n)         (:128)
//   BufferSample   inputs  buffer[0..n) count head                     outputs batch[0..n)       (:139)
//                  attrs   seed (int), unique (bool, only false)
//   PriorityReplayBufferCreate   attrs capacity alpha shapes(int 2-D) dtypes(str "float32,int32,...") seed0 seed1
//                                outputs handle int64[1]                (priority_replay_buffer.py:62-66)
//   PriorityReplayBufferPush     attrs handle; inputs transition[0..n); outputs handle int64[1]   (:90)
//   PriorityReplayBufferSample   attrs handle; inputs beta float[1];
//                                outputs indices int64[B] weights float[B] batch[0..n)            (:106)
//   PriorityReplayBufferUpdate   attrs handle; inputs indices int64[B] priorities float[B]; outputs handle (:120)
//   PriorityReplayBufferDestroy  attrs handle; outputs handle                                     (:138)
//   DiscountedReturn             attrs gamma; inputs reward done last_state_value; outputs returns
//                                (discounted_return.py:71, 82)
//   GaeScan                      attrs gamma lambda layout(0 [T,B] / 1 [B,T]) normalize(bool, optional)
//                                inputs reward value last_value done; outputs advantage returns
//                                (ppo.py:334-368, mappo.py:478-497)
//
// The uniform-buffer operators are STATELESS like the primitives they replace: the ring columns, count and
// head are the caller's tensors (the reference keeps them as Parameters and passes them into every call), so
// the cursors are read and advanced ON THE DEVICE -- no host mirror, no synchronisation, order given by the
// stream.  The priority-buffer operators go to the handle registry of prb.cu.
#include <string.h>

#include <map>
#include <sstream>

#include "aot_extra.h"
#include "common.cuh"

namespace mrl {
namespace {

constexpr int kAotError = 2;  // kErrorCode of the reference's operators (cpu_mcts_ops.cc:25)

int dtype_size(const char *name) {
  if (!name) return 0;
  std::string s(name);
  for (auto &ch : s) ch = (char)tolower(ch);
  const size_t dot = s.rfind('.');
  if (dot != std::string::npos) s = s.substr(dot + 1);  // "mindspore.float32"
  if (s == "bool" || s == "bool_" || s == "int8" || s == "uint8") return 1;
  if (s == "float16" || s == "half" || s == "bfloat16" || s == "int16" || s == "uint16") return 2;
  if (s == "float32" || s == "float" || s == "int32" || s == "int" || s == "uint32") return 4;
  if (s == "float64" || s == "double" || s == "int64" || s == "uint64" || s == "long") return 8;
  return 0;
}

int64_t numel_from(const int64_t *shape, int nd, int first) {
  int64_t n = 1;
  for (int i = first; i < nd; ++i) n *= shape[i];
  return n;
}

// ---- device-cursor kernels of the stateless uniform buffer ------------------------------------------
__device__ __forceinline__ void copy_units(uint8_t *dst, const uint8_t *src, int vec) {
  if (vec == 16) *reinterpret_cast<uint4 *>(dst) = *reinterpret_cast<const uint4 *>(src);
  else if (vec == 8) *reinterpret_cast<uint2 *>(dst) = *reinterpret_cast<const uint2 *>(src);
  else if (vec == 4) *reinterpret_cast<uint32_t *>(dst) = *reinterpret_cast<const uint32_t *>(src);
  else if (vec == 2) *reinterpret_cast<uint16_t *>(dst) = *reinterpret_cast<const uint16_t *>(src);
  else *dst = *src;
}

// BufferAppend: rows io[c][0..nrows) -> ring.  Cursor rule of SURVEY.md 8(a2), read from the device:
// while head == 0 and count < capacity rows go to `count`, afterwards to `head`; a run that crosses the end
// wraps; of an over-long batch only the last `capacity` rows survive.  blockIdx.y = column.
__global__ void __launch_bounds__(256)
    aot_append_kernel(const __grid_constant__ ColTable tab, int64_t nrows, int64_t cap, const int32_t *count,
                      const int32_t *head) {
  const int32_t cnt = *count, hd = *head;
  const int64_t pos = (hd == 0 && cnt < cap) ? cnt : hd;
  const int64_t skip = nrows > cap ? nrows - cap : 0;
  const int64_t first = (pos + skip) % cap, n = nrows - skip;
  const int c = blockIdx.y;
  const int64_t rb = tab.row_bytes[c];
  const int vec = tab.vec[c];
  const int64_t upr = rb / vec, total = n * upr;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; u < total; u += stride) {
    const int64_t r = u / upr, o = (u - r * upr) * vec;
    int64_t dr = first + r;
    dr = dr >= cap ? dr - cap : dr;
    copy_units(tab.col[c] + dr * rb + o, tab.io[c] + (skip + r) * rb + o, vec);
  }
}
// after the rows: advance the cursors (stream-ordered behind aot_append_kernel, which read the old ones)
__global__ void aot_cursor_kernel(int32_t *count, int32_t *head, int64_t nrows, int64_t cap, int32_t *count_out) {
  const int32_t cnt = *count, hd = *head;
  if (hd == 0 && cnt < cap) {
    int64_t c = (int64_t)cnt + nrows;
    if (c > cap) {
      *head = (int32_t)((c - cap) % cap);
      c = cap;
    }
    *count = (int32_t)c;
  } else {
    *head = (int32_t)((hd + nrows) % cap);
  }
  if (count_out) *count_out = *count;
}

// BufferGetItem: the row at LOGICAL position index (0 = oldest, negative = from the end); an index outside
// the valid rows is clamped into them (the operator cannot raise from the device).  One warp per column.
__global__ void __launch_bounds__(32)
    aot_get_item_kernel(const __grid_constant__ ColTable tab, int64_t cap, const int32_t *count, const int32_t *head,
                        const void *index, int index_bytes) {
  const int64_t cnt = *count, hd = *head;
  int64_t idx = index_bytes == 8 ? *reinterpret_cast<const int64_t *>(index)
                                 : (int64_t) * reinterpret_cast<const int32_t *>(index);
  if (idx < 0) idx += cnt;
  idx = idx < 0 ? 0 : (idx > cap - 1 ? cap - 1 : idx);
  const int64_t t = cnt - hd;
  const int64_t slot = idx < t ? idx + hd : idx - t;
  const int c = blockIdx.x;
  const int64_t rb = tab.row_bytes[c];
  const int vec = tab.vec[c];
  for (int64_t o = (int64_t)threadIdx.x * vec; o < rb; o += 32 * vec)
    copy_units(tab.io[c] + o, tab.col[c] + slot * rb + o, vec);
}

// BufferSample: warp j draws slot j uniformly (with replacement) from the `count` valid rows with the
// library's counter-based generator and copies the row of every column
__global__ void __launch_bounds__(256)
    aot_sample_kernel(const __grid_constant__ ColTable tab, int64_t n, uint64_t seed, uint64_t ctr, const int32_t *count) {
  const int lane = threadIdx.x & 31;
  const int64_t j = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (j >= n) return;
  const int64_t cnt = *count;
  const int64_t slot = cnt > 0 ? mrl_uniform_below(seed, ctr + (uint64_t)j, cnt) : 0;
  for (int c = 0; c < tab.ncols; ++c) {
    const int64_t rb = tab.row_bytes[c];
    const int vec = tab.vec[c];
    for (int64_t o = (int64_t)lane * vec; o < rb; o += 32 * vec)
      copy_units(tab.io[c] + j * rb + o, tab.col[c] + slot * rb + o, vec);
  }
}

__global__ void aot_write_i64_kernel(int64_t *dst, int64_t v) { *dst = v; }

// ---- per-operator data parked in AotExtra between Init and the calls ------------------------------------
struct UrbOpData : public AotKernelData {
  int ncols = 0;
  int64_t capacity = 0;
  int64_t row_bytes[MRL_MAX_COLS];
  uint64_t seed = 0, ctr = 0;
};
struct HandleOpData : public AotKernelData {
  int64_t handle = MRL_INVALID_HANDLE;
};
struct CreateOpData : public AotKernelData {
  int64_t capacity = 0;
  float alpha = 0.0f;
  int ncols = 0;
  int64_t row_bytes[MRL_MAX_COLS];
  uint64_t seed0 = 0, seed1 = 0;
};
struct ScanOpData : public AotKernelData {
  float gamma = 0.0f, lambda = 0.0f;
  int layout = MRL_TIME_MAJOR;
  bool normalize = false;
};

AotExtra *as_extra(void *p) { return static_cast<AotExtra *>(p); }

template <class T>
T *kernel_data(void *extra) {
  AotExtra *e = as_extra(extra);
  return e ? dynamic_cast<T *>(e->KernelData()) : nullptr;
}

// ring geometry of the n buffer columns that lead the parameter list
int urb_geometry(UrbOpData *d, int ncols, int *ndims, int64_t **shapes, const char **dtypes) {
  if (ncols < 1 || ncols > MRL_MAX_COLS) return kAotError;
  d->ncols = ncols;
  d->capacity = ndims[0] >= 1 ? shapes[0][0] : 0;
  if (d->capacity <= 0) return kAotError;
  for (int c = 0; c < ncols; ++c) {
    const int es = dtype_size(dtypes[c]);
    if (es == 0 || ndims[c] < 1 || shapes[c][0] != d->capacity) return kAotError;
    d->row_bytes[c] = numel_from(shapes[c], ndims[c], 1) * es;
  }
  return 0;
}

void fill(ColTable &tab, const UrbOpData *d, void **cols, void **io) {
  tab.ncols = d->ncols;
  for (int c = 0; c < d->ncols; ++c) {
    tab.col[c] = (uint8_t *)cols[c];
    tab.io[c] = (uint8_t *)io[c];
    tab.row_bytes[c] = d->row_bytes[c];
    tab.vec[c] = align_unit((uintptr_t)cols[c], (uintptr_t)io[c], (uint64_t)d->row_bytes[c]);
  }
}

int64_t attr_handle(AotExtra *e) {
  // the reference passes handles as float attributes (mcts.py:258-262 "tree_handle"); MindSpore's own
  // PriorityReplayBuffer ops use an int attribute: the int getter serves both registrations
  return e->Attr<int64_t>("handle");
}

}  // namespace
}  // namespace mrl

using namespace mrl;

#define AOT_CUDA_OK()                          \
  do {                                         \
    g_launches.fetch_add(1);                   \
    if (cudaPeekAtLastError() != cudaSuccess) { \
      set_error("%s", cudaGetErrorString(cudaGetLastError())); \
      return kAotError;                        \
    }                                          \
  } while (0)

// ---- BufferAppend -----------------------------------------------------------------------------------------
extern "C" int BufferAppendInit(int *ndims, int64_t **shapes, const char **dtypes, AotExtra *extra) {
  (void)ndims, (void)shapes, (void)dtypes;
  if (extra) extra->SetKernelData(new UrbOpData());
  return 0;
}
extern "C" int BufferAppend(int nparam, void **params, int *ndims, int64_t **shapes, const char **dtypes,
                            void *stream, void *extra) {
  if (nparam < 5 || (nparam - 3) % 2 != 0) return kAotError;
  const int n = (nparam - 3) / 2;
  UrbOpData local, *d = kernel_data<UrbOpData>(extra);
  if (!d) d = &local;
  if (urb_geometry(d, n, ndims, shapes, dtypes)) return kAotError;
  // "Support batch input" (README.md:428): exp[i] is one row shaped like buffer[i][0], or N rows of it
  int64_t nrows = 1;
  if (ndims[n] == ndims[0]) nrows = shapes[n][0];
  for (int c = 0; c < n; ++c) {
    const int es = dtype_size(dtypes[n + c]);
    if (es == 0 || numel_from(shapes[n + c], ndims[n + c], 0) * es != nrows * d->row_bytes[c]) return kAotError;
  }
  if (nrows < 1) return 0;
  cudaStream_t st = as_stream(stream);
  ColTable tab;
  fill(tab, d, params, params + n);
  int32_t *count = (int32_t *)params[2 * n], *head = (int32_t *)params[2 * n + 1];
  int64_t max_units = 1;
  for (int c = 0; c < n; ++c) {
    const int64_t u = (nrows < d->capacity ? nrows : d->capacity) * d->row_bytes[c] / tab.vec[c];
    max_units = u > max_units ? u : max_units;
  }
  int64_t blocks = (max_units + 255) / 256;
  blocks = blocks > (int64_t)sm_count() * 8 ? (int64_t)sm_count() * 8 : blocks;
  aot_append_kernel<<<dim3((unsigned)blocks, (unsigned)n), 256, 0, st>>>(tab, nrows, d->capacity, count, head);
  AOT_CUDA_OK();
  aot_cursor_kernel<<<1, 1, 0, st>>>(count, head, nrows, d->capacity, (int32_t *)params[2 * n + 2]);
  AOT_CUDA_OK();
  return 0;
}

// ---- BufferGetItem ----------------------------------------------------------------------------------------
extern "C" int BufferGetItemInit(int *ndims, int64_t **shapes, const char **dtypes, AotExtra *extra) {
  (void)ndims, (void)shapes, (void)dtypes;
  if (extra) extra->SetKernelData(new UrbOpData());
  return 0;
}
extern "C" int BufferGetItem(int nparam, void **params, int *ndims, int64_t **shapes, const char **dtypes,
                             void *stream, void *extra) {
  if (nparam < 5 || (nparam - 3) % 2 != 0) return kAotError;
  const int n = (nparam - 3) / 2;
  UrbOpData local, *d = kernel_data<UrbOpData>(extra);
  if (!d) d = &local;
  if (urb_geometry(d, n, ndims, shapes, dtypes)) return kAotError;
  const int ib = dtype_size(dtypes[n + 2]);
  if (ib != 4 && ib != 8) return kAotError;
  ColTable tab;
  fill(tab, d, params, params + n + 3);
  aot_get_item_kernel<<<n, 32, 0, as_stream(stream)>>>(tab, d->capacity, (const int32_t *)params[n],
                                                       (const int32_t *)params[n + 1], params[n + 2], ib);
  AOT_CUDA_OK();
  return 0;
}

// ---- BufferSample -----------------------------------------------------------------------------------------
extern "C" int BufferSampleInit(int *ndims, int64_t **shapes, const char **dtypes, AotExtra *extra) {
  (void)ndims, (void)shapes, (void)dtypes;
  if (!extra) return 0;
  UrbOpData *d = new UrbOpData();
  if (extra->HasAttr("seed")) d->seed = (uint64_t)extra->Attr<int64_t>("seed");
  const bool unique = extra->HasAttr("unique") && extra->Attr<bool>("unique");
  extra->SetKernelData(d);
  if (unique) {
    set_error("BufferSample: unique=True (sampling without replacement) is not provided; every caller in the "
              "reference keeps the default");
    return kAotError;
  }
  return 0;
}
extern "C" int BufferSample(int nparam, void **params, int *ndims, int64_t **shapes, const char **dtypes,
                            void *stream, void *extra) {
  if (nparam < 4 || (nparam - 2) % 2 != 0) return kAotError;
  const int n = (nparam - 2) / 2;
  UrbOpData local, *d = kernel_data<UrbOpData>(extra);
  if (!d) d = &local;
  if (urb_geometry(d, n, ndims, shapes, dtypes)) return kAotError;
  const int64_t batch = ndims[n + 2] >= 1 ? shapes[n + 2][0] : 0;
  if (batch < 1) return kAotError;
  ColTable tab;
  fill(tab, d, params, params + n + 2);
  aot_sample_kernel<<<(unsigned)((batch * 32 + 255) / 256), 256, 0, as_stream(stream)>>>(
      tab, batch, d->seed, d->ctr, (const int32_t *)params[n]);
  AOT_CUDA_OK();
  d->ctr += (uint64_t)batch;
  return 0;
}

// ---- PriorityReplayBufferCreate ---------------------------------------------------------------------------
extern "C" int PriorityReplayBufferCreateInit(int *ndims, int64_t **shapes, const char **dtypes, AotExtra *extra) {
  (void)ndims, (void)shapes, (void)dtypes;
  if (!extra) return kAotError;
  CreateOpData *d = new CreateOpData();
  extra->SetKernelData(d);
  d->capacity = extra->Attr<int64_t>("capacity");
  d->alpha = extra->Attr<float>("alpha");
  d->seed0 = extra->HasAttr("seed0") ? (uint64_t)extra->Attr<int64_t>("seed0") : 0;
  d->seed1 = extra->HasAttr("seed1") ? (uint64_t)extra->Attr<int64_t>("seed1") : 0;
  const std::vector<std::vector<int64_t>> shp = extra->Attr<std::vector<std::vector<int64_t>>>("shapes");
  std::vector<std::string> names;
  {
    std::stringstream ss(extra->Attr<std::string>("dtypes"));
    std::string tok;
    while (std::getline(ss, tok, ',')) {
      while (!tok.empty() && tok.front() == ' ') tok.erase(tok.begin());
      while (!tok.empty() && tok.back() == ' ') tok.pop_back();
      if (!tok.empty()) names.push_back(tok);
    }
  }
  if (shp.empty() || shp.size() != names.size() || shp.size() > MRL_MAX_COLS) {
    set_error("PriorityReplayBufferCreate: shapes and dtypes must name the same 1..%d columns", MRL_MAX_COLS);
    return kAotError;
  }
  d->ncols = (int)shp.size();
  for (int c = 0; c < d->ncols; ++c) {
    const int es = dtype_size(names[c].c_str());
    if (es == 0) return kAotError;
    int64_t ne = 1;
    for (int64_t v : shp[c]) ne *= v;
    d->row_bytes[c] = ne * es;
  }
  return 0;
}
extern "C" int PriorityReplayBufferCreate(int nparam, void **params, int *ndims, int64_t **shapes,
                                          const char **dtypes, void *stream, void *extra) {
  (void)ndims, (void)shapes, (void)dtypes;
  CreateOpData *d = kernel_data<CreateOpData>(extra);
  if (!d || nparam < 1) return kAotError;
  int64_t h = MRL_INVALID_HANDLE;
  if (mrl_prb_create(d->capacity, d->alpha, d->ncols, d->row_bytes, d->seed0, d->seed1, &h) != MRL_OK) return kAotError;
  aot_write_i64_kernel<<<1, 1, 0, as_stream(stream)>>>((int64_t *)params[nparam - 1], h);
  AOT_CUDA_OK();
  return 0;
}

// ---- the operators that carry the handle as an attribute ------------------------------------------------
static int handle_init(AotExtra *extra) {
  if (!extra) return kAotError;
  HandleOpData *d = new HandleOpData();
  d->handle = attr_handle(extra);
  extra->SetKernelData(d);
  return 0;
}
static int write_handle(void *out, int64_t h, void *stream) {
  aot_write_i64_kernel<<<1, 1, 0, as_stream(stream)>>>((int64_t *)out, h);
  AOT_CUDA_OK();
  return 0;
}

extern "C" int PriorityReplayBufferPushInit(int *ndims, int64_t **shapes, const char **dtypes, AotExtra *extra) {
  (void)ndims, (void)shapes, (void)dtypes;
  return handle_init(extra);
}
extern "C" int PriorityReplayBufferPush(int nparam, void **params, int *ndims, int64_t **shapes,
                                        const char **dtypes, void *stream, void *extra) {
  (void)ndims, (void)shapes, (void)dtypes;
  HandleOpData *d = kernel_data<HandleOpData>(extra);
  if (!d || nparam < 2) return kAotError;
  if (prb_column_count(d->handle) != nparam - 1) {
    set_error("PriorityReplayBufferPush: %d transition tensors for a buffer of %d columns", nparam - 1,
              prb_column_count(d->handle));
    return kAotError;
  }
  if (mrl_prb_push(d->handle, params, stream) != MRL_OK) return kAotError;
  return write_handle(params[nparam - 1], d->handle, stream);
}

extern "C" int PriorityReplayBufferSampleInit(int *ndims, int64_t **shapes, const char **dtypes, AotExtra *extra) {
  (void)ndims, (void)shapes, (void)dtypes;
  return handle_init(extra);
}
extern "C" int PriorityReplayBufferSample(int nparam, void **params, int *ndims, int64_t **shapes,
                                          const char **dtypes, void *stream, void *extra) {
  (void)dtypes;
  HandleOpData *d = kernel_data<HandleOpData>(extra);
  if (!d || nparam < 4) return kAotError;
  const int64_t batch = ndims[1] >= 1 ? shapes[1][0] : 0;  // indices [B]
  if (batch < 1) return kAotError;
  if (prb_column_count(d->handle) != nparam - 3) {
    set_error("PriorityReplayBufferSample: %d column outputs for a buffer of %d columns", nparam - 3,
              prb_column_count(d->handle));
    return kAotError;
  }
  if (prb_sample_beta_dev(d->handle, batch, (const float *)params[0], (int64_t *)params[1], (float *)params[2],
                          params + 3, as_stream(stream)) != MRL_OK)
    return kAotError;
  return 0;
}

extern "C" int PriorityReplayBufferUpdateInit(int *ndims, int64_t **shapes, const char **dtypes, AotExtra *extra) {
  (void)ndims, (void)shapes, (void)dtypes;
  return handle_init(extra);
}
extern "C" int PriorityReplayBufferUpdate(int nparam, void **params, int *ndims, int64_t **shapes,
                                          const char **dtypes, void *stream, void *extra) {
  HandleOpData *d = kernel_data<HandleOpData>(extra);
  if (!d || nparam != 3 || dtype_size(dtypes[0]) != 8 || dtype_size(dtypes[1]) != 4) return kAotError;
  const int64_t n = numel_from(shapes[0], ndims[0], 0);
  if (n != numel_from(shapes[1], ndims[1], 0)) return kAotError;
  if (mrl_prb_update(d->handle, (const int64_t *)params[0], (const float *)params[1], n, stream) != MRL_OK)
    return kAotError;
  return write_handle(params[2], d->handle, stream);
}

extern "C" int PriorityReplayBufferDestroyInit(int *ndims, int64_t **shapes, const char **dtypes, AotExtra *extra) {
  (void)ndims, (void)shapes, (void)dtypes;
  return handle_init(extra);
}
extern "C" int PriorityReplayBufferDestroy(int nparam, void **params, int *ndims, int64_t **shapes,
                                           const char **dtypes, void *stream, void *extra) {
  (void)ndims, (void)shapes, (void)dtypes;
  HandleOpData *d = kernel_data<HandleOpData>(extra);
  if (!d || nparam < 1) return kAotError;
  if (mrl_prb_destroy(d->handle) != MRL_OK) return kAotError;
  return write_handle(params[nparam - 1], d->handle, stream);
}

// ---- DiscountedReturn -------------------------------------------------------------------------------------
extern "C" int DiscountedReturnInit(int *ndims, int64_t **shapes, const char **dtypes, AotExtra *extra) {
  (void)ndims, (void)shapes, (void)dtypes;
  if (!extra) return kAotError;
  ScanOpData *d = new ScanOpData();
  d->gamma = extra->Attr<float>("gamma");
  extra->SetKernelData(d);
  return 0;
}
extern "C" int DiscountedReturn(int nparam, void **params, int *ndims, int64_t **shapes, const char **dtypes,
                                void *stream, void *extra) {
  ScanOpData *d = kernel_data<ScanOpData>(extra);
  if (!d || nparam != 4 || dtype_size(dtypes[0]) != 4 || dtype_size(dtypes[1]) != 1 || ndims[0] < 1) return kAotError;
  // reward [T, B, ...], done [T, B] (trailing 1s of done broadcast), last_state_value [B, ...]
  const int64_t T = shapes[0][0];
  int nd_done = ndims[1];
  while (nd_done > 1 && shapes[1][nd_done - 1] == 1) --nd_done;
  const int64_t B = numel_from(shapes[1], nd_done, 1);
  const int64_t N = numel_from(shapes[0], ndims[0], 1);
  if (B <= 0 || N % B != 0 || shapes[1][0] != T) return kAotError;
  if (mrl_discounted_return((const float *)params[0], (const uint8_t *)params[1], (const float *)params[2],
                            (float *)params[3], T, B, N / B, d->gamma, MRL_TIME_MAJOR, MRL_SCAN_AUTO, stream) != MRL_OK)
    return kAotError;
  return 0;
}

// ---- GaeScan ------------------------------------------------------------------------------------------------
extern "C" int GaeScanInit(int *ndims, int64_t **shapes, const char **dtypes, AotExtra *extra) {
  (void)ndims, (void)shapes, (void)dtypes;
  if (!extra) return kAotError;
  ScanOpData *d = new ScanOpData();
  d->gamma = extra->Attr<float>("gamma");
  d->lambda = extra->Attr<float>("lambda");
  d->layout = extra->HasAttr("layout") ? (int)extra->Attr<int64_t>("layout") : MRL_TIME_MAJOR;
  d->normalize = extra->HasAttr("normalize") && extra->Attr<bool>("normalize");
  extra->SetKernelData(d);
  return d->layout == MRL_TIME_MAJOR || d->layout == MRL_BATCH_MAJOR ? 0 : kAotError;
}
extern "C" int GaeScan(int nparam, void **params, int *ndims, int64_t **shapes, const char **dtypes, void *stream,
                       void *extra) {
  ScanOpData *d = kernel_data<ScanOpData>(extra);
  if (!d || nparam != 6 || ndims[0] < 2 || dtype_size(dtypes[0]) != 4 || dtype_size(dtypes[3]) != 1) return kAotError;
  const int64_t d0 = shapes[0][0], d1 = numel_from(shapes[0], ndims[0], 1);
  const int64_t T = d->layout == MRL_TIME_MAJOR ? d0 : d1, B = d->layout == MRL_TIME_MAJOR ? d1 : d0;
  int rc;
  if (d->normalize)
    rc = mrl_gae_normalized((const float *)params[0], (const float *)params[1], nullptr, (const float *)params[2],
                            (const uint8_t *)params[3], (float *)params[4], (float *)params[5], T, B, d->gamma,
                            d->lambda, 1e-8f, d->layout, MRL_SCAN_AUTO, stream);
  else
    rc = mrl_gae((const float *)params[0], (const float *)params[1], nullptr, (const float *)params[2],
                 (const uint8_t *)params[3], (float *)params[4], (float *)params[5], T, B, d->gamma, d->lambda,
                 d->layout, MRL_SCAN_AUTO, stream);
  return rc == MRL_OK ? 0 : kAotError;
}

// =================================================================================================================
// A concrete AotExtra for hosts that are not MindSpore (tests, other frameworks): attributes in a map.
// =================================================================================================================
namespace {
class MapAotExtra final : public AotExtra {
 public:
  std::map<std::string, bool> b;
  std::map<std::string, int64_t> i;
  std::map<std::string, float> f;
  std::map<std::string, std::string> s;
  std::map<std::string, std::vector<int64_t>> iv;
  std::map<std::string, std::vector<float>> fv;
  std::map<std::string, std::vector<std::vector<int64_t>>> i2;
  ~MapAotExtra() override { DestructKernelData(); }
  bool HasAttr(std::string n) override {
    return b.count(n) || i.count(n) || f.count(n) || s.count(n) || iv.count(n) || fv.count(n) || i2.count(n);
  }

 private:
  bool GetAttrBool(std::string n) override { return b.count(n) ? b[n] : (i.count(n) ? i[n] != 0 : false); }
  int64_t GetAttrInt(std::string n) override { return i.count(n) ? i[n] : (f.count(n) ? (int64_t)f[n] : 0); }
  float GetAttrFloat(std::string n) override { return f.count(n) ? f[n] : (i.count(n) ? (float)i[n] : 0.0f); }
  std::string GetAttrStr(std::string n) override { return s.count(n) ? s[n] : std::string(); }
  std::vector<int64_t> GetAttrIntVec(std::string n) override { return iv.count(n) ? iv[n] : std::vector<int64_t>(); }
  std::vector<float> GetAttrFloatVec(std::string n) override { return fv.count(n) ? fv[n] : std::vector<float>(); }
  std::vector<std::vector<int64_t>> GetAttrInt2DVec(std::string n) override {
    return i2.count(n) ? i2[n] : std::vector<std::vector<int64_t>>();
  }
  std::vector<std::vector<float>> GetAttrFloat2DVec(std::string) override { return {}; }
};
MapAotExtra *as_map(void *p) { return static_cast<MapAotExtra *>(static_cast<AotExtra *>(p)); }
}  // namespace

extern "C" void *mrl_aot_extra_create(void) { return static_cast<AotExtra *>(new MapAotExtra()); }
extern "C" void mrl_aot_extra_destroy(void *extra) {
  if (extra) delete as_map(extra);
}
extern "C" int mrl_aot_extra_set_bool(void *extra, const char *name, int32_t v) {
  MRL_REQUIRE(extra && name, "mrl_aot_extra_set_bool: NULL argument");
  as_map(extra)->b[name] = v != 0;
  return MRL_OK;
}
extern "C" int mrl_aot_extra_set_int(void *extra, const char *name, int64_t v) {
  MRL_REQUIRE(extra && name, "mrl_aot_extra_set_int: NULL argument");
  as_map(extra)->i[name] = v;
  return MRL_OK;
}
extern "C" int mrl_aot_extra_set_float(void *extra, const char *name, float v) {
  MRL_REQUIRE(extra && name, "mrl_aot_extra_set_float: NULL argument");
  as_map(extra)->f[name] = v;
  return MRL_OK;
}
extern "C" int mrl_aot_extra_set_str(void *extra, const char *name, const char *v) {
  MRL_REQUIRE(extra && name && v, "mrl_aot_extra_set_str: NULL argument");
  as_map(extra)->s[name] = v;
  return MRL_OK;
}
// rows of a 2-D integer attribute, flattened: row r has lens[r] entries
extern "C" int mrl_aot_extra_set_int2d(void *extra, const char *name, const int64_t *flat, const int32_t *lens,
                                       int32_t nrows) {
  MRL_REQUIRE(extra && name && nrows >= 0 && (nrows == 0 || (flat && lens)), "mrl_aot_extra_set_int2d: bad arguments");
  std::vector<std::vector<int64_t>> rows;
  for (int r = 0; r < nrows; ++r) {
    rows.emplace_back(flat, flat + lens[r]);
    flat += lens[r];
  }
  as_map(extra)->i2[name] = rows;
  return MRL_OK;
}
