// microbench3.cu -- random row-gather rate against TABLE SIZE (design probe, not product code).
// The aggregation kernels gather 2 KB rows at random from H.  At the full benchmark size H is 20 GB and the
// kernels run at ~1.2 G rows/s although the same code does 2.2 G rows/s on a 0.4-2 GB table.  This probe separates the
// memory system from the kernels: a plain LDG.128 gather (8 blocks/SM x 8 warps, 4 rows in flight per warp) over
// tables of growing size, and with rows of 2 KB / 4 KB / 8 KB at the largest size (a per-access limit -- address
// translation -- shows up as bandwidth proportional to the row size).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

template <int Q>   // Q float4 per lane per row: row bytes = Q * 512
__global__ void __launch_bounds__(256) k_gather(const float4* __restrict__ T, const int* __restrict__ src, int n, int row4, float4* sink) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int nw = (gridDim.x * blockDim.x) >> 5;
  float4 acc = make_float4(0, 0, 0, 0);
  for (int i = warp * 4; i + 4 <= n; i += nw * 4) {
    float4 v[4][Q];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float4* r = T + (size_t)__ldg(src + i + u) * row4;
#pragma unroll
      for (int k = 0; k < Q; ++k) v[u][k] = __ldg(r + lane + 32 * k);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int k = 0; k < Q; ++k) { acc.x += v[u][k].x; acc.y += v[u][k].y; acc.z += v[u][k].z; acc.w += v[u][k].w; }
  }
  if (acc.x == 12345.f) sink[0] = acc;
}

template <typename F> float timeit(F f, int reps = 3) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) { cudaEventRecord(a); f(); cudaEventRecord(b); CK(cudaEventSynchronize(b)); float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms; }
  return best;
}

int main() {
  const int n = 8 << 20;
  std::vector<int> h(n);
  int* src; float4* sink; CK(cudaMalloc(&src, (size_t)n * 4)); CK(cudaMalloc(&sink, 64));
  float* T; const size_t cap = (size_t)40 << 30;
  CK(cudaMalloc(&T, cap)); CK(cudaMemset(T, 0, cap));
  uint64_t x = 88172645463325252ull;
  auto rnd = [&]() { x ^= x << 13; x ^= x >> 7; x ^= x << 17; return x; };
  const double gbs[] = {0.4, 1, 2, 4, 8, 12, 16, 20, 30, 40};
  for (double g : gbs) {
    const size_t rows = (size_t)(g * (1ull << 30)) / 2048;
    for (int i = 0; i < n; ++i) h[i] = (int)(rnd() % rows);
    CK(cudaMemcpy(src, h.data(), (size_t)n * 4, cudaMemcpyHostToDevice));
    float ms = timeit([&] { k_gather<4><<<148 * 8, 256>>>((const float4*)T, src, n, 128, sink); });
    printf("table %5.1f GB  rows of 2 KB : %8.3f ms  %6.2f G rows/s  %7.1f GB/s\n", g, ms, n / ms / 1e6, n * 2048.0 / ms / 1e6);
    fflush(stdout);
  }
  for (int q : {8, 16}) {   // 4 KB and 8 KB rows over 40 GB
    const size_t rows = cap / (q * 512);
    const int nn = n / (q / 4);
    for (int i = 0; i < nn; ++i) h[i] = (int)(rnd() % rows);
    CK(cudaMemcpy(src, h.data(), (size_t)nn * 4, cudaMemcpyHostToDevice));
    float ms = q == 8 ? timeit([&] { k_gather<8><<<148 * 8, 256>>>((const float4*)T, src, nn, 256, sink); })
                      : timeit([&] { k_gather<16><<<148 * 4, 256>>>((const float4*)T, src, nn, 512, sink); });
    printf("table  40.0 GB  rows of %d KB : %8.3f ms  %6.2f G rows/s  %7.1f GB/s\n", q / 2, ms, nn / ms / 1e6, (double)nn * q * 512 / ms / 1e6);
  }
  // sorted-by-page access order over 40 GB (same rows, ascending): the DRAM-side ceiling without translation misses
  {
    const size_t rows = cap / 2048;
    for (int i = 0; i < n; ++i) h[i] = (int)((size_t)i * rows / n);
    CK(cudaMemcpy(src, h.data(), (size_t)n * 4, cudaMemcpyHostToDevice));
    float ms = timeit([&] { k_gather<4><<<148 * 8, 256>>>((const float4*)T, src, n, 128, sink); });
    printf("table  40.0 GB  2 KB rows, ASCENDING order: %8.3f ms  %6.2f G rows/s  %7.1f GB/s\n", ms, n / ms / 1e6, n * 2048.0 / ms / 1e6);
  }
  return 0;
}
