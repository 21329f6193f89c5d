// common.cuh -- shared host/device helpers of libmindrl_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>
#include <mutex>
#include <string>
#include <unordered_map>
#include <utility>

#include "../../include/mindrl_b200.h"
#include "../../include/mrl_arith.h"

namespace mrl {

// SMs of the current device (cudaDevAttrMultiProcessorCount, cached per device; 148 on a B200).  Grids are
// sized from this, never from a constant.
int sm_count();
// true exactly once per (kernel, device): guards cudaFuncSetAttribute, which is a per-device setting
bool first_use_on_device(const void *kernel);

void set_error(const char *fmt, ...);
extern std::atomic<int64_t> g_launches;

inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }

// (a reported error is also CLEARED -- cudaGetLastError -- so that one failed call does not fail every later
// launch check of the process)
#define MRL_CUDA(expr)                                                                   \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) {                                                             \
      ::mrl::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                       __LINE__);                                                        \
      cudaGetLastError();                                                                \
      return MRL_ERR_CUDA;                                                               \
    }                                                                                    \
  } while (0)

#define MRL_REQUIRE(cond, ...)      \
  do {                              \
    if (!(cond)) {                  \
      ::mrl::set_error(__VA_ARGS__); \
      return MRL_ERR_INVALID;       \
    }                               \
  } while (0)

// call right after every <<<>>> launch: counts it and surfaces launch-config errors
#define MRL_LAUNCHED()                          \
  do {                                          \
    ::mrl::g_launches.fetch_add(1);             \
    MRL_CUDA(cudaGetLastError());               \
  } while (0)

#ifdef __CUDACC__
// Programmatic dependent launch: the tree kernels of a learner loop are 10-16 us long and follow each
// other (sample -> update -> sample ...), so the ~3 us between one kernel draining and the next one being
// resident is a fifth of the step.  Each kernel lets its successor become resident as soon as its own CTAs
// have all started (launch_dependents) and waits for its predecessor to complete, memory flushed, before it
// touches global memory (wait; a no-op when the kernel was launched the ordinary way).
__device__ __forceinline__ void pdl_prologue() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}
extern bool g_pdl;  // option "pdl" (copy.cu)
extern bool g_tree_prefetch;  // option "tree_prefetch" (copy.cu)

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              Args &&...args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid, cfg.blockDim = block, cfg.dynamicSmemBytes = smem, cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr, cfg.numAttrs = g_pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

#endif

// process-global handle registry (mcts_factory.h:39-82 is the reference's model)
template <class T>
class Registry {
 public:
  int64_t put(T *obj) {
    std::lock_guard<std::mutex> g(mu_);
    int64_t h = next_++;
    map_[h] = obj;
    return h;
  }
  T *get(int64_t h) {
    std::lock_guard<std::mutex> g(mu_);
    auto it = map_.find(h);
    return it == map_.end() ? nullptr : it->second;
  }
  T *take(int64_t h) {
    std::lock_guard<std::mutex> g(mu_);
    auto it = map_.find(h);
    if (it == map_.end()) return nullptr;
    T *o = it->second;
    map_.erase(it);
    return o;
  }

 private:
  std::mutex mu_;
  std::unordered_map<int64_t, T *> map_;
  int64_t next_ = 1;
};

// ------------------------------------------------------------------------------------------
// column table passed by value to the copy kernels (<= MRL_MAX_COLS columns)
// ------------------------------------------------------------------------------------------
struct ColTable {
  uint8_t *col[MRL_MAX_COLS];       // ring column base (capacity rows)
  uint8_t *io[MRL_MAX_COLS];        // per-call source (append) or destination (gather)
  int64_t row_bytes[MRL_MAX_COLS];
  int32_t vec[MRL_MAX_COLS];        // widest unit (16/8/4/2/1 bytes) every row address is aligned to
  int32_t ncols;
};

inline int32_t align_unit(uintptr_t a, uintptr_t b, uint64_t row_bytes) {
  uint64_t x = a | b | row_bytes | 16u;
  return (int32_t)(x & (~x + 1));  // lowest set bit, capped at 16
}

// pinned + device staging used by the *_host entry points
struct Staging {
  uint8_t *host = nullptr;
  uint8_t *dev = nullptr;
  size_t bytes = 0;
  cudaEvent_t done = nullptr;  // last async use of `host`
  bool pending = false;        // `done` was recorded since the last wait()
  cudaError_t mark(cudaStream_t st) {
    pending = true;
    return cudaEventRecord(done, st);
  }
  int ensure(size_t need);
  int wait();
  void release();
};

// gathers the columns selected by col_mask (bit c = column c)
// waits for `st` by polling an event: the blocking cudaStreamSynchronize adds several microseconds of
// wake-up latency, which is most of a small *_host call
int stream_wait_spin(cudaStream_t st);
int set_scan_option(const char *key, int64_t value);  // scan.cu

// gathers the columns selected by col_mask (bit c = column c)
int gather_rows(const ColTable &tab, const int64_t *slots, int64_t n, int64_t capacity,
                cudaStream_t st, uint32_t col_mask = 0xffffffffu);
int copy_segments(const ColTable &tab, int64_t dst_row, int64_t src_row, int64_t nrows,
                  cudaStream_t st);
// to_ring = 1: io rows -> ring rows; 0: ring rows -> io rows; 2: io rows are HOST memory -> ring rows (DMA)
// batches of at least this many bytes go from host arrays straight into the ring (the *_host entry points)
constexpr size_t kDirectHostBytes = 256 << 10;
int copy_segments_dir(const ColTable &tab, int64_t ring_row, int64_t io_row, int64_t nrows,
                      int to_ring, cudaStream_t st);
// tunables of the TMA-ring copies (copy.cu; options "gather_path", "bulk_*")
extern int64_t g_gather_path, g_bulk_min_row, g_bulk_chunk, g_bulk_ctas_per_sm, g_bulk_stages;
// number of columns of a PriorityReplayBuffer, -1 for an unknown handle (the AOT shims check their parameter lists)
int prb_column_count(int64_t handle);
// mrl_prb_sample with beta read from DEVICE memory by the kernel (the AOT shim gets beta as a tensor)
int prb_sample_beta_dev(int64_t handle, int64_t n, const float *beta_dev, int64_t *indices, float *weights,
                        void *const *out_cols, cudaStream_t st);

}  // namespace mrl
