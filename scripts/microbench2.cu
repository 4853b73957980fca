// microbench2.cu -- design probes for the round-2 aggregation kernels (not product code).
// Question: per message the weight-id-major kernels gather one 2 KB row from HBM and reduce one 2 KB row into an
// L2-resident window.  Do the two streams overlap, and is the TMA path (cp.async.bulk gather into shared memory,
// cp.reduce.async.bulk add.f32 out of it) faster than LDG.128 + red.global.add.v4.f32?
//   A  ldg+red      : LDG.128 gather (U rows in flight per warp) -> red.v4 into the window        (today's data path)
//   B  bulk+red     : cp.async.bulk gather into a smem ring (1 producer lane) -> LDS -> red.v4   (7 consumer warps)
//   D  bulk+bulkred : cp.async.bulk gather -> cp.reduce.async.bulk.add.f32 from the same slot     (no registers)
//   E  bulk gather only, F bulk reduce only, G ldg gather only, H red.v4 only
// Sizes: table T = 2M rows x 2 KB (4 GB, HBM), accumulation window 8192 rows (16 MB) sliding over a 4 GB output.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>
#include <vector>
#include <algorithm>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tWAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra WAIT_DONE;\n\tbra WAIT_LOOP;\n\tWAIT_DONE:\n\t}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_red_s2g(void* dst, uint32_t src, uint32_t bytes) {
  asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void red4(float* p, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// ---- A / G / H: register path -------------------------------------------------------------------------------
template <int U, bool GATHER, bool RED>
__global__ void __launch_bounds__(256) k_ldg_red(const float4* __restrict__ T, float* __restrict__ O, const int* __restrict__ src,
                                                  const int* __restrict__ dst, int n, float4* sink) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int nw = (gridDim.x * blockDim.x) >> 5;
  // items of 64 consecutive messages dealt round-robin to the warps (like the product's work items): concurrently
  // running warps sit in the same 16 MB window of the output
  float4 acc = make_float4(0, 0, 0, 0);
  for (int item = warp; item * 64 < n; item += nw)
  for (int i = item * 64; i + U <= min(n, item * 64 + 64); i += U) {
    float4 v[U][4];
    int dr[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      dr[u] = __ldg(dst + i + u);
      if (GATHER) {
        const float4* r = T + (size_t)__ldg(src + i + u) * 128;
#pragma unroll
        for (int k = 0; k < 4; ++k) v[u][k] = __ldg(r + lane + 32 * k);
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) v[u][k] = make_float4(1.f, 2.f, 3.f, 4.f);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (RED) {
        float* o = O + (size_t)dr[u] * 512;
#pragma unroll
        for (int k = 0; k < 4; ++k) red4(o + 4 * (lane + 32 * k), v[u][k]);
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) { acc.x += v[u][k].x; acc.y += v[u][k].y; acc.z += v[u][k].z; acc.w += v[u][k].w; }
      }
    }
  }
  if (acc.x == 12345.f) sink[0] = acc;
}

// ---- B: bulk gather ring + consumer warps that red.v4 ---------------------------------------------------------
constexpr int B_SLOTS = 48;      // 96 KB ring
constexpr int B_CONS = 7;        // consumer warps
__global__ void __launch_bounds__(256, 1) k_bulk_red(const float4* __restrict__ T, float* __restrict__ O, const int* __restrict__ src,
                                                      const int* __restrict__ dst, int n, int mode /*0: red.v4, 1: none*/, float4* sink) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t full[B_SLOTS], empty[B_SLOTS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // CTA b owns chunks b, b + grid, ... of CH consecutive messages (window locality like the product)
  constexpr int CH = 224;  // multiple of 32 and of B_CONS
  const int n_chunks = n / CH;
  int my_chunks = 0;
  for (int c = blockIdx.x; c < n_chunks; c += gridDim.x) ++my_chunks;
  const int total = my_chunks * CH;  // messages this CTA handles; message k lives at chunk (k / CH), offset k % CH
  auto gidx = [&](int k) { return (blockIdx.x + (k / CH) * gridDim.x) * CH + (k % CH); };
  if (threadIdx.x == 0) {
    for (int s = 0; s < B_SLOTS; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (warp == B_CONS) {
    for (int k0 = 0; k0 < total; k0 += 32) {   // CH % 32 == 0: a batch never straddles chunks
      const int mine = __ldg(src + gidx(k0 + lane));
      for (int u = 0; u < 32; ++u) {
        const int k = k0 + u, s = k % B_SLOTS;
        const int row = __shfl_sync(0xffffffffu, mine, u);
        if (lane == 0) {
          if (k >= B_SLOTS) mbar_wait(&empty[s], ((k / B_SLOTS) - 1) & 1);
          mbar_expect_tx(&full[s], 2048);
          bulk_g2s(smem_u32(smem + s * 2048), T + (size_t)row * 128, 2048, &full[s]);
        }
        __syncwarp();
      }
    }
  } else {
    float4 acc = make_float4(0, 0, 0, 0);
    for (int k = warp; k < total; k += B_CONS) {
      const int s = k % B_SLOTS;
      const int dr = __ldg(dst + gidx(k));
      mbar_wait(&full[s], (k / B_SLOTS) & 1);
      const float4* row = reinterpret_cast<const float4*>(smem + s * 2048);
      float4 v[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) v[q] = row[lane + 32 * q];
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);
      if (mode == 0) {
        float* o = O + (size_t)dr * 512;
#pragma unroll
        for (int q = 0; q < 4; ++q) red4(o + 4 * (lane + 32 * q), v[q]);
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) { acc.x += v[q].x; acc.y += v[q].y; acc.z += v[q].z; acc.w += v[q].w; }
      }
    }
    if (acc.x == 12345.f) sink[0] = acc;
  }
}

// ---- D / E / F: one lane drives a private ring: bulk gather -> bulk reduce -------------------------------------
constexpr int D_SLOTS = 16, D_AHEAD = 8, D_RINGS = 6;  // 6 rings x 32 KB = 192 KB
template <bool GATHER, bool RED>
__global__ void __launch_bounds__(32 * D_RINGS, 1) k_bulk_bulk(const float4* __restrict__ T, float* __restrict__ O, const int* __restrict__ src,
                                                                const int* __restrict__ dst, int n) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t full[D_RINGS][D_SLOTS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nr = gridDim.x * D_RINGS, ring = blockIdx.x * D_RINGS + warp;
  constexpr int CH = 64;
  const int n_chunks = n / CH;
  int my_chunks = 0;
  for (int c = ring; c < n_chunks; c += nr) ++my_chunks;
  const int cnt = my_chunks * CH;
  auto gidx = [&](int k) { return (ring + (k / CH) * nr) * CH + (k % CH); };
  uint8_t* my = smem + warp * D_SLOTS * 2048;
  if (lane == 0) {
    for (int s = 0; s < D_SLOTS; ++s) mbar_init(&full[warp][s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  if (!GATHER) {  // F: fill the ring once
    for (int k = lane; k < D_SLOTS * 512; k += 32) reinterpret_cast<float*>(my)[k] = 1.f;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
  }
  // the whole warp walks the stream (indices are loaded 32 at a time and broadcast); lane 0 issues the copies
  int src_cur = 0, dst_cur = 0, dst_prev = 0;
  for (int k0 = 0; k0 < cnt + 32; k0 += 32) {
    dst_prev = dst_cur;
    if (k0 < cnt) { src_cur = __ldg(src + gidx(k0 + lane)); dst_cur = __ldg(dst + gidx(k0 + lane)); }
    for (int u = 0; u < 32; ++u) {
      const int k = k0 + u;
      const int srow = __shfl_sync(0xffffffffu, src_cur, u);
      const int j = k - D_AHEAD;  // D_AHEAD < 32: j lies in this batch or the previous one
      const int drow = (u >= D_AHEAD) ? __shfl_sync(0xffffffffu, dst_cur, (u - D_AHEAD) & 31)
                                      : __shfl_sync(0xffffffffu, dst_prev, (u + 32 - D_AHEAD) & 31);
      if (lane == 0) {
        if (k < cnt) {
          const int s = k % D_SLOTS;
          if (RED && k >= D_SLOTS) bulk_wait_read<D_SLOTS - 1 - D_AHEAD>();  // the reduce that last used slot s has read it
          if (GATHER) {
            mbar_expect_tx(&full[warp][s], 2048);
            bulk_g2s(smem_u32(my + s * 2048), T + (size_t)srow * 128, 2048, &full[warp][s]);
          }
        }
        if (j >= 0 && j < cnt) {
          const int s = j % D_SLOTS;
          if (GATHER) mbar_wait(&full[warp][s], (j / D_SLOTS) & 1);
          if (RED) {
            bulk_red_s2g(O + (size_t)drow * 512, smem_u32(my + s * 2048), 2048);
            bulk_commit();
          }
        }
      }
      __syncwarp();
    }
  }
  if (RED && lane == 0) bulk_wait_read<0>();
  __syncwarp();
}

template <typename F> float timeit(F f, int reps = 5) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) { cudaEventRecord(a); f(); cudaEventRecord(b); CK(cudaEventSynchronize(b)); float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms; }
  return best;
}

int main(int argc, char** argv) {
  // usage: microbench2 [n_messages] [rows] [tests: any of A B D]   (small sizes for compute-sanitizer)
  const int n = argc > 1 ? atoi(argv[1]) : (4 << 20);   // 4M messages = 8.6 GB gathered + 8.6 GB reduced
  const size_t rows = argc > 2 ? (size_t)atol(argv[2]) : 2000000;  // 4 GB table, 4 GB output
  const char* tests = argc > 3 ? argv[3] : "ABD";
  auto want = [&](char c) { for (const char* p = tests; *p; ++p) if (*p == c) return true; return false; };
  std::vector<int> hs(n), hd(n);
  uint64_t x = 88172645463325252ull;
  auto rnd = [&]() { x ^= x << 13; x ^= x >> 7; x ^= x << 17; return x; };
  for (int i = 0; i < n; ++i) {
    hs[i] = (int)(rnd() % rows);
    const size_t win = ((size_t)i * rows / n) / 8192 * 8192;  // sliding 16 MB window (supertile order)
    hd[i] = (int)std::min(rows - 1, win + rnd() % 8192);
  }
  int *src, *dst; float4* sink; float *T, *O;
  CK(cudaMalloc(&src, n * 4)); CK(cudaMalloc(&dst, n * 4)); CK(cudaMalloc(&sink, 64));
  CK(cudaMemcpy(src, hs.data(), n * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dst, hd.data(), n * 4, cudaMemcpyHostToDevice));
  CK(cudaMalloc(&T, rows * 2048)); CK(cudaMemset(T, 0, rows * 2048));
  CK(cudaMalloc(&O, rows * 2048)); CK(cudaMemset(O, 0, rows * 2048));
  const double gb = n * 2048.0 / 1e9;
  auto rep = [&](const char* name, float ms, double bytes_gb) { printf("%-34s %8.3f ms  %8.1f GB/s (per stream %.1f GB)\n", name, ms, bytes_gb / ms * 1e3, gb); };
  if (want('A')) for (int bps : {2, 4, 8}) {
    char nm[96];
    float ms = timeit([&] { k_ldg_red<4, true, false><<<148 * bps, 256>>>((const float4*)T, O, src, dst, n, sink); });
    snprintf(nm, 96, "G ldg gather only   U=4 bps=%d", bps); rep(nm, ms, gb);
    ms = timeit([&] { k_ldg_red<4, false, true><<<148 * bps, 256>>>((const float4*)T, O, src, dst, n, sink); });
    snprintf(nm, 96, "H red.v4 only       U=4 bps=%d", bps); rep(nm, ms, gb);
    ms = timeit([&] { k_ldg_red<4, true, true><<<148 * bps, 256>>>((const float4*)T, O, src, dst, n, sink); });
    snprintf(nm, 96, "A ldg + red.v4      U=4 bps=%d", bps); rep(nm, ms, 2 * gb);
    ms = timeit([&] { k_ldg_red<2, true, true><<<148 * bps, 256>>>((const float4*)T, O, src, dst, n, sink); });
    snprintf(nm, 96, "A ldg + red.v4      U=2 bps=%d", bps); rep(nm, ms, 2 * gb);
    ms = timeit([&] { k_ldg_red<8, true, true><<<148 * bps, 256>>>((const float4*)T, O, src, dst, n, sink); });
    snprintf(nm, 96, "A ldg + red.v4      U=8 bps=%d", bps); rep(nm, ms, 2 * gb);
  }
  if (want('B')) {
    const int smemB = B_SLOTS * 2048;
    CK(cudaFuncSetAttribute(k_bulk_red, cudaFuncAttributeMaxDynamicSharedMemorySize, smemB));
    float ms = timeit([&] { k_bulk_red<<<148, 256, smemB>>>((const float4*)T, O, src, dst, n, 1, sink); });
    rep("E' bulk gather -> LDS (no red)", ms, gb);
    ms = timeit([&] { k_bulk_red<<<148, 256, smemB>>>((const float4*)T, O, src, dst, n, 0, sink); });
    rep("B bulk gather -> LDS -> red.v4", ms, 2 * gb);
    ms = timeit([&] { k_bulk_red<<<296, 256, smemB>>>((const float4*)T, O, src, dst, n, 0, sink); });
    rep("B (2 CTAs/SM)", ms, 2 * gb);
  }
  if (want('D')) {
    const int smemD = D_RINGS * D_SLOTS * 2048;
    CK(cudaFuncSetAttribute(k_bulk_bulk<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smemD));
    CK(cudaFuncSetAttribute(k_bulk_bulk<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smemD));
    CK(cudaFuncSetAttribute(k_bulk_bulk<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smemD));
    float ms = timeit([&] { k_bulk_bulk<true, false><<<148, 32 * D_RINGS, smemD>>>((const float4*)T, O, src, dst, n); });
    rep("E bulk gather only", ms, gb);
    ms = timeit([&] { k_bulk_bulk<false, true><<<148, 32 * D_RINGS, smemD>>>((const float4*)T, O, src, dst, n); });
    rep("F bulk reduce only", ms, gb);
    ms = timeit([&] { k_bulk_bulk<true, true><<<148, 32 * D_RINGS, smemD>>>((const float4*)T, O, src, dst, n); });
    rep("D bulk gather -> bulk reduce", ms, 2 * gb);
  }
  // sanity: O must hold n*512 adds of... (only check that the bulk reduce really added: sum of one window)
  std::vector<float> h(512);
  CK(cudaMemcpy(h.data(), O + (size_t)hd[0] * 512, 2048, cudaMemcpyDeviceToHost));
  printf("O[dst0][0..3] = %g %g %g %g\n", h[0], h[1], h[2], h[3]);
  return 0;
}
