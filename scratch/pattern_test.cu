// access-pattern microbenchmark: copy a [T, N] fp32 matrix with (a) tile4-like pieces: a warp reads
// S rows x 512 B (stride N*4) ; (b) contiguous 4 KB per warp.  Prints GB/s (read+write).
#include <cstdio>
#include <cuda_runtime.h>
template <int S>
__global__ void __launch_bounds__(128) tile_copy(const float4* __restrict__ in, float4* __restrict__ out, int T, int N4, int n_groups, int W) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int tile = blockIdx.x;
  int chunk = tile / n_groups, group = tile % n_groups;
  long col4 = (long)group * 32 + lane;
  long t_base = (long)chunk * (W * S) + (long)w * S;
  float4 v[S];
#pragma unroll
  for (int s = 0; s < S; ++s) v[s] = __ldcs(in + (t_base + s) * N4 + col4);
#pragma unroll
  for (int s = 0; s < S; ++s) { v[s].x += 1.f; __stcs(out + (t_base + s) * N4 + col4, v[s]); }
}
// row-wide variant: CTA of W warps side by side covers W*128 columns, S rows
template <int S>
__global__ void __launch_bounds__(256) row_copy(const float4* __restrict__ in, float4* __restrict__ out, int T, int N4, int n_groups) {
  int tile = blockIdx.x;
  int chunk = tile / n_groups, group = tile % n_groups;
  long col4 = (long)group * blockDim.x + threadIdx.x;
  long t_base = (long)chunk * S;
  float4 v[S];
#pragma unroll
  for (int s = 0; s < S; ++s) v[s] = __ldcs(in + (t_base + s) * N4 + col4);
#pragma unroll
  for (int s = 0; s < S; ++s) { v[s].x += 1.f; __stcs(out + (t_base + s) * N4 + col4, v[s]); }
}
// scalar variant: a warp owns 32 columns (one float per lane) x S rows, like scan_tm_wtile_kernel
template <int S>
__global__ void __launch_bounds__(128) wtile_copy(const float* __restrict__ in, float* __restrict__ out, int T, int N, int n_wgroups) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int tile = blockIdx.x;                       // CTA = 4 adjacent column groups of one chunk
  int chunk = tile / (n_wgroups / 4), cg = tile % (n_wgroups / 4);
  long col = ((long)cg * 4 + w) * 32 + lane;
  long t_base = (long)chunk * S;
  float v[S];
#pragma unroll
  for (int s = 0; s < S; ++s) v[s] = __ldcs(in + (t_base + s) * N + col);
#pragma unroll
  for (int s = 0; s < S; ++s) __stcs(out + (t_base + s) * N + col, v[s] + 1.f);
}
__global__ void __launch_bounds__(256) flat_copy(const float4* __restrict__ in, float4* __restrict__ out, long n4) {
  long i = ((long)blockIdx.x * blockDim.x + threadIdx.x) * 8;
  float4 v[8];
  // each warp: 8 x 512 B contiguous = 4 KB
  long base = (i / 256) * 256 + (threadIdx.x & 31);
#pragma unroll
  for (int k = 0; k < 8; ++k) v[k] = __ldcs(in + base + k * 32);
#pragma unroll
  for (int k = 0; k < 8; ++k) { v[k].x += 1.f; __stcs(out + base + k * 32, v[k]); }
}
int main() {
  const int T = 2048, N = 4096, N4 = N / 4;
  size_t bytes = (size_t)T * N * 4;
  float4 *in[4], *out[4];
  for (int i = 0; i < 4; ++i) { cudaMalloc(&in[i], bytes); cudaMalloc(&out[i], bytes); cudaMemset(in[i], 0, bytes); }
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  auto run = [&](const char* name, auto launch) {
    for (int i = 0; i < 3; ++i) launch(i % 4);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    const int reps = 40;
    for (int i = 0; i < reps; ++i) launch(i % 4);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    printf("%-40s %8.2f us  %8.1f GB/s  (%s)\n", name, ms / reps * 1e3, 2.0 * bytes / (ms / reps * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
  };
  run("tile 128col x (4 warps x 8 rows)", [&](int k) { tile_copy<8><<<(T / 32) * (N / 128), 128>>>(in[k], out[k], T, N4, N / 128, 4); });
  run("tile 128col x (4 warps x 16 rows)", [&](int k) { tile_copy<16><<<(T / 64) * (N / 128), 128>>>(in[k], out[k], T, N4, N / 128, 4); });
  run("row 1024col x 8 rows (256 thr)", [&](int k) { row_copy<8><<<(T / 8) * (N / 1024), 256>>>(in[k], out[k], T, N4, N / 1024); });
  run("row 1024col x 16 rows (256 thr)", [&](int k) { row_copy<16><<<(T / 16) * (N / 1024), 256>>>(in[k], out[k], T, N4, N / 1024); });
  run("row 512col x 8 rows (128 thr)", [&](int k) { row_copy<8><<<(T / 8) * (N / 512), 128>>>(in[k], out[k], T, N4, N / 512); });
  run("wtile 32col x 32 rows scalar (128 thr)", [&](int k) { wtile_copy<32><<<(T / 32) * (N / 128), 128>>>((const float*)in[k], (float*)out[k], T, N, N / 32); });
  run("wtile 32col x 16 rows scalar (128 thr)", [&](int k) { wtile_copy<16><<<(T / 16) * (N / 128), 128>>>((const float*)in[k], (float*)out[k], T, N, N / 32); });
  run("flat 4KB per warp", [&](int k) { flat_copy<<<(unsigned)((size_t)T * N / 4 / 8 / 256), 256>>>(in[k], out[k], (long)T * N4); });
  return 0;
}
