// raw inbound rate of 2-D TMA box loads per SM: one producer thread keeps a ring full, one consumer thread frees
// the stages at once.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_rate tma_rate.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)
__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t *b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t ph) {
  asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(s32(b)), "r"(ph) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *m, int x, int y, uint64_t *bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(s32(dst)), "l"(m), "r"(x), "r"(y), "r"(s32(bar)) : "memory");
}
constexpr int MAXS = 16;
__global__ void __launch_bounds__(64) k(const __grid_constant__ CUtensorMap tm, int box_cols, int box_rows, int n_tiles, int n_stages, int stage_bytes, int cols_per_cta) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t full[MAXS], empty[MAXS];
  if (threadIdx.x == 0) {
    for (int s = 0; s < n_stages; ++s) mbar_init(&full[s], 1), mbar_init(&empty[s], 1);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  const int col0 = blockIdx.x * cols_per_cta;
  if (threadIdx.x == 0) {
    int s = 0; uint32_t ph = 1;
    for (int t = 0; t < n_tiles; ++t) {
      if (t >= n_stages) mbar_wait(&empty[s], ph);
      mbar_expect_tx(&full[s], (uint32_t)(box_cols * box_rows * 4));
      tma_load_2d(smem + (size_t)s * stage_bytes, &tm, col0, t * box_rows, &full[s]);
      if (++s == n_stages) s = 0, ph ^= 1u;
    }
  } else if (threadIdx.x == 32) {
    int s = 0; uint32_t ph = 0;
    for (int t = 0; t < n_tiles; ++t) {
      mbar_wait(&full[s], ph);
      mbar_arrive(&empty[s]);
      if (++s == n_stages) s = 0, ph ^= 1u;
    }
  }
}
typedef CUresult (*Enc)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main() {
  void *fp = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
  Enc enc = (Enc)fp;
  const int64_t T = 2048, N = 8192;
  float *d; CK(cudaMalloc(&d, T * N * 4)); CK(cudaMemset(d, 0, T * N * 4));
  char *fl; CK(cudaMalloc(&fl, 512 << 20));
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  struct Cfg { int bc, br, ctas, stages; } cfgs[] = {
      {32, 128, 128, 8}, {32, 128, 8, 8}, {64, 128, 64, 4}, {64, 128, 128, 4}, {32, 128, 256, 4}};
  const CUtensorMapL2promotion promos[3] = {CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B};
  const char *pname[3] = {"none", "128B", "256B"};
  for (int pi = 0; pi < 3; ++pi)
  for (auto c : cfgs) {
    CUtensorMap tm;
    cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)T}; cuuint64_t str[1] = {(cuuint64_t)N * 4};
    cuuint32_t box[2] = {(cuuint32_t)c.bc, (cuuint32_t)c.br}; cuuint32_t es[2] = {1, 1};
    if (enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, promos[pi], CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode failed\n"); continue; }
    const int stage = c.bc * c.br * 4, tiles = (int)(T / c.br);
    if (c.ctas * c.bc > N || c.stages * stage > 200 * 1024) { printf("skip\n"); continue; }
    float best = 1e9;
    for (int rep = 0; rep < 5; ++rep) {
      CK(cudaMemset(fl, rep, 512 << 20));
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      cudaEventRecord(e0);
      k<<<c.ctas, 64, c.stages * stage>>>(tm, c.bc, c.br, tiles, c.stages, stage, c.bc);
      cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
      float ms; cudaEventElapsedTime(&ms, e0, e1); best = ms < best ? ms : best;
    }
    const double bytes = (double)c.ctas * c.bc * T * 4;
    printf("L2 promotion %s: box %2d cols x %3d rows, %3d CTAs, %2d stages (%3d KB in flight/CTA): %7.2f us  %6.0f GB/s total  %5.1f GB/s per CTA\n", pname[pi], c.bc, c.br, c.ctas, c.stages, c.stages * stage / 1024, best * 1e3, bytes / best / 1e6, bytes / best / 1e6 / c.ctas);
  }
  return 0;
}
