// microbench.cu -- design probes for the aggregation kernels (not product code):
//   gather : random 2 KB row gathers (LDG.128), table resident in L2 vs HBM
//   red    : random 2 KB row red.global.add.v4.f32, table resident in L2 vs HBM
//   wread  : random 16 KB contiguous chunk reads from a 32 MB table (the per-run weight fetch)
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__global__ void k_gather(const float4* __restrict__ T, const int* __restrict__ idx, int n, int row4, float4* out) {
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  int nw = (gridDim.x * blockDim.x) >> 5;
  float4 acc = make_float4(0, 0, 0, 0);
  for (int i = warp * 4; i + 3 < n; i += nw * 4) {
    float4 v[4][4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float4* r = T + (size_t)idx[i + u] * row4;
#pragma unroll
      for (int k = 0; k < 4; ++k) v[u][k] = (lane + 32 * k < row4) ? __ldg(r + lane + 32 * k) : make_float4(0, 0, 0, 0);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int k = 0; k < 4; ++k) { acc.x += v[u][k].x; acc.y += v[u][k].y; acc.z += v[u][k].z; acc.w += v[u][k].w; }
  }
  if (acc.x == 12345.f) out[0] = acc;
}
__global__ void k_red(float* T, const int* __restrict__ idx, int n, int row4) {
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  int nw = (gridDim.x * blockDim.x) >> 5;
  for (int i = warp; i < n; i += nw) {
    float* r = T + (size_t)idx[i] * row4 * 4;
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (lane + 32 * k < row4)
        asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(r + 4 * (lane + 32 * k)), "f"(1.f), "f"(1.f), "f"(1.f), "f"(1.f) : "memory");
  }
}
__global__ void k_red_scalar(float* T, const int* __restrict__ idx, int n, int row4) {
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  int nw = (gridDim.x * blockDim.x) >> 5;
  for (int i = warp; i < n; i += nw) {
    float* r = T + (size_t)idx[i] * row4 * 4;
    for (int k = lane; k < row4 * 4; k += 32) atomicAdd(r + k, 1.f);
  }
}
__global__ void k_wread(const float4* __restrict__ T, const int* __restrict__ idx, int n, int chunk4, float4* out) {
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  int nw = (gridDim.x * blockDim.x) >> 5;
  float4 acc = make_float4(0, 0, 0, 0);
  for (int i = warp; i < n; i += nw) {
    const float4* r = T + (size_t)idx[i] * chunk4;
    for (int k = lane; k < chunk4; k += 128) {
      float4 a = __ldg(r + k), b = (k + 32 < chunk4) ? __ldg(r + k + 32) : a, c = (k + 64 < chunk4) ? __ldg(r + k + 64) : a, d = (k + 96 < chunk4) ? __ldg(r + k + 96) : a;
      acc.x += a.x + b.x + c.x + d.x;
    }
  }
  if (acc.x == 12345.f) out[0] = acc;
}
template <typename F> float timeit(F f, int reps = 5) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) { cudaEventRecord(a); f(); cudaEventRecord(b); CK(cudaEventSynchronize(b)); float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms; }
  return best;
}
int main() {
  const int row4 = 128;  // 2 KB rows
  const int n = 4 << 20; // 4M row accesses = 8.6 GB
  std::vector<int> h(n);
  float4* out; CK(cudaMalloc(&out, 64));
  int* idx; CK(cudaMalloc(&idx, n * sizeof(int)));
  size_t sizes[] = {14541, 200000, 2000000};
  for (size_t rows : sizes) {
    for (int i = 0; i < n; ++i) h[i] = (int)(((unsigned long long)rand() * 2654435761ULL + i * 40503ULL) % rows);
    CK(cudaMemcpy(idx, h.data(), n * sizeof(int), cudaMemcpyHostToDevice));
    float* T; CK(cudaMalloc(&T, rows * row4 * 16)); CK(cudaMemset(T, 0, rows * row4 * 16));
    for (int bps : {4, 8}) {
      int blocks = 148 * bps;
      float ms = timeit([&] { k_gather<<<blocks, 256>>>((const float4*)T, idx, n, row4, out); });
      printf("gather  rows=%8zu (%7.1f MB) blocks/SM=%d : %7.3f ms  %8.1f GB/s\n", rows, rows * 2048 / 1e6, bps, ms, n * 2048.0 / ms / 1e6);
    }
    int nr = n / 4;
    float ms = timeit([&] { k_red<<<148 * 8, 256>>>(T, idx, nr, row4); });
    printf("red.v4  rows=%8zu (%7.1f MB)            : %7.3f ms  %8.1f GB/s payload, %6.1f M vec-red/ms\n", rows, rows * 2048 / 1e6, ms, nr * 2048.0 / ms / 1e6, nr * 128.0 / ms / 1e6);
    ms = timeit([&] { k_red_scalar<<<148 * 8, 256>>>(T, idx, nr / 4, row4); });
    printf("red.f32 rows=%8zu (%7.1f MB)            : %7.3f ms  %8.1f GB/s payload\n", rows, rows * 2048 / 1e6, ms, nr / 4 * 2048.0 / ms / 1e6);
    CK(cudaFree(T));
  }
  // weight chunk reads: 2000 chunks x 16 KB = 32 MB and 474 x 10 KB = 4.7 MB
  for (int cfg = 0; cfg < 2; ++cfg) {
    int chunks = cfg ? 474 : 2000, chunk4 = cfg ? 625 : 1024;
    int nn = 1 << 20;
    for (int i = 0; i < nn; ++i) h[i] = rand() % chunks;
    CK(cudaMemcpy(idx, h.data(), nn * sizeof(int), cudaMemcpyHostToDevice));
    float* T; CK(cudaMalloc(&T, (size_t)chunks * chunk4 * 16)); CK(cudaMemset(T, 0, (size_t)chunks * chunk4 * 16));
    float ms = timeit([&] { k_wread<<<148 * 8, 256>>>((const float4*)T, idx, nn, chunk4, out); });
    printf("wread   chunks=%d x %d B (%5.1f MB)     : %7.3f ms  %8.1f GB/s\n", chunks, chunk4 * 16, chunks * chunk4 * 16 / 1e6, ms, (double)nn * chunk4 * 16 / ms / 1e6);
    CK(cudaFree(T));
  }
  return 0;
}
