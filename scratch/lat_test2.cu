// model of the level-synchronous tree update: where do 1.25 us per level go?
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
__device__ __forceinline__ void recompute(float* sum, float* mn, long node) {
  float a = __ldcg(sum + 2 * node), b = __ldcg(sum + 2 * node + 1);
  float ma = __ldcg(mn + 2 * node), mb = __ldcg(mn + 2 * node + 1);
  __stcg(sum + node, a + b); __stcg(mn + node, ma < mb ? ma : mb);
}
template <int MODE>
__global__ void k(float* sum, float* mn, const long* idx, int n, long cap2, int levels, long long* out) {
  int tid = threadIdx.x;
  long long c[24];
  c[0] = clock64();
  for (int level = levels - 1, j = 1; level >= 11; --level, ++j) {
    int shift = levels - level;
    for (int q = tid; q < n; q += blockDim.x) {
      long node = (cap2 + (MODE == 2 ? idx[q] : __ldg(idx + q))) >> shift;
      if (MODE == 1) {  // loads only, no stores
        float a = __ldcg(sum + 2 * node), b = __ldcg(sum + 2 * node + 1);
        if (a + b == 12345.f) sum[0] = a;
      } else recompute(sum, mn, node);
    }
    __syncthreads();
    c[j] = clock64();
  }
  if (tid == 0) for (int j = 0; j < 10; ++j) out[j] = c[j] - c[0];
}
int main() {
  long cap2 = 1 << 20; int levels = 20, n = 256;
  float *sum, *mn; long* idx; long long* out;
  cudaMalloc(&sum, 8 * cap2); cudaMalloc(&mn, 8 * cap2); cudaMalloc(&idx, 8 * 4096); cudaMalloc(&out, 256);
  cudaMemset(sum, 0, 8 * cap2); cudaMemset(mn, 0, 8 * cap2);
  long h[4096]; srand(1); for (int i = 0; i < 4096; ++i) h[i] = (((long)rand() << 10) ^ rand()) % cap2;
  cudaMemcpy(idx, h, sizeof(h), cudaMemcpyHostToDevice);
  char* flush; cudaMalloc(&flush, 256 << 20);
  long long o[10];
  auto report = [&](const char* name) { cudaMemcpy(o, out, 80, cudaMemcpyDeviceToHost); printf("%-34s", name); for (int j = 1; j < 10; ++j) printf(" %5.2f", (o[j] - o[j - 1]) / 1965.0); printf("  total %.2f us\n", o[9] / 1965.0); };
  for (int cold = 0; cold < 2; ++cold) {
    for (int nn : {32, 256, 1024}) {
      for (int threads : {256, 1024}) {
        for (int rep = 0; rep < 3; ++rep) { if (cold) cudaMemset(flush, rep, 256 << 20); k<0><<<1, threads>>>(sum, mn, idx + 1024 * (rep % 3), nn, cap2, levels, out); cudaDeviceSynchronize(); }
        char name[64]; snprintf(name, 64, "%s n=%d thr=%d ld+st", cold ? "cold" : "warm", nn, threads); report(name);
      }
    }
    for (int rep = 0; rep < 3; ++rep) { if (cold) cudaMemset(flush, rep, 256 << 20); k<1><<<1, 1024>>>(sum, mn, idx + 1024 * (rep % 3), 256, cap2, levels, out); cudaDeviceSynchronize(); }
    report(cold ? "cold n=256 thr=1024 loads only" : "warm n=256 thr=1024 loads only");
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
