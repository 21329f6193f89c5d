// single-CTA latency calibration: what do barriers, shared atomics and global round trips cost?
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ long long gt() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__global__ void k(float* g, unsigned long long* out, int iters) {
  __shared__ unsigned long long tab[1024];
  __shared__ float sh[1024];
  int tid = threadIdx.x;
  for (int i = tid; i < 1024; i += blockDim.x) { tab[i] = 0; sh[i] = 0; }
  __syncthreads();
  long long t0 = gt(); long long c0 = clock64();
  for (int i = 0; i < iters; ++i) __syncthreads();
  long long t1 = gt(); long long c1 = clock64();
  // store + barrier + load of a neighbour's value (global, L2)
  float v = tid;
  for (int i = 0; i < iters; ++i) { __stcg(g + tid, v); __syncthreads(); v += __ldcg(g + ((tid + 1) % blockDim.x)); }
  long long t2 = gt(); long long c2 = clock64();
  // same through shared memory
  for (int i = 0; i < iters; ++i) { sh[tid] = v; __syncthreads(); v += sh[(tid + 1) % blockDim.x]; __syncthreads(); }
  long long t3 = gt(); long long c3 = clock64();
  // shared 64-bit CAS, no contention
  for (int i = 0; i < iters; ++i) { atomicCAS(&tab[(tid * 7 + i) & 1023], (unsigned long long)i, (unsigned long long)(i + 1)); }
  long long t4 = gt(); long long c4 = clock64();
  // dependent global loads (pointer chase in L2)
  int idx = tid;
  for (int i = 0; i < iters; ++i) idx = ((int)__ldcg(g + 4096 + (idx & 1023))) & 1023;
  long long t5 = gt(); long long c5 = clock64();
  if (tid == 0) {
    out[0] = t1 - t0; out[1] = t2 - t1; out[2] = t3 - t2; out[3] = t4 - t3; out[4] = t5 - t4;
    out[5] = c1 - c0; out[6] = c2 - c1; out[7] = c3 - c2; out[8] = c4 - c3; out[9] = c5 - c4;
  }
  g[8192 + tid] = v + idx;
}
int main() {
  float* g; unsigned long long* out; cudaMalloc(&g, 1 << 20); cudaMemset(g, 0, 1 << 20); cudaMalloc(&out, 128);
  unsigned long long h[10];
  for (int threads : {256, 1024}) {
    for (int rep = 0; rep < 3; ++rep) {
      k<<<1, threads>>>(g, out, 200);
      cudaMemcpy(h, out, 80, cudaMemcpyDeviceToHost);
    }
    printf("threads %d (ns per iter / cycles per iter): barrier %.0f/%.0f  gst+bar+gld %.0f/%.0f  sst+bar+sld+bar %.0f/%.0f  smemCAS64 %.0f/%.0f  gld-chase %.0f/%.0f\n",
           threads, h[0] / 200.0, h[5] / 200.0, h[1] / 200.0, h[6] / 200.0, h[2] / 200.0, h[7] / 200.0, h[3] / 200.0, h[8] / 200.0, h[4] / 200.0, h[9] / 200.0);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
