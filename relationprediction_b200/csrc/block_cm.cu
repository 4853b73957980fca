// block_cm.cu -- EXPERIMENTAL component-major path for the block-diagonal layer (block_algo = 2, opt-in;
// written without GPU access at the end of round 1: not yet validated or timed on a B200).
//
// Why: with 5x5 blocks a float4 of a row-major feature row straddles two blocks, so the weight-id-major kernel
// (rgcn_kernels.cu, k_block_relg) stages every pre-summed row in shared memory and reads it back with scalar
// loads; ncu shows it L1-wavefront / issue bound (l1tex 76 %, ~400 warp instructions per message).
// In the COMPONENT-MAJOR layout a row is stored as [component i][block b] (position i*B + b), so the lane that
// owns blocks 4q..4q+3 finds, for every component j, the four inputs it needs in ONE aligned float4
// (x[j][4q..4q+3]) and computes its 4 x S outputs entirely in registers:
//     y[i][b] = sum_j W[b][i][j] * x[j][b]           (S*S elementwise float4 FMAs, weights resident per item)
// No shared memory, no barrier, no cross-lane traffic; loads and reductions stay 128-bit and coalesced
// (B/4 consecutive lanes cover B*4 contiguous bytes per component).  Per (row, weight id) run that is
// S gathers + S*S*4 FMAs + S RED.128 per lane for 4 blocks, against 4 warps x (STS.128 + bar + 4S LDS + 4S FMA
// + RED.128) in the row-major kernel.
//
// Pieces: k_to_cm / k_cm_add (row-major <-> component-major, fused with the epilogue), k_relayout_cm /
// k_unlayout_cm (weight tables), k_block_cm<S, DW> (transform, or weight-gradient accumulation).
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.cuh"

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int CM_WARPS = 4;

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 zero4() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ void red4(float* p, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}
__device__ __forceinline__ void fma4s(float4& a, float s, const float4& x) {  // a += s * x
  a.x = fmaf(s, x.x, a.x);
  a.y = fmaf(s, x.y, a.y);
  a.z = fmaf(s, x.z, a.z);
  a.w = fmaf(s, x.w, a.w);
}
__device__ __forceinline__ void fma4v(float4& a, const float4& w, const float4& x) {  // a += w (.) x
  a.x = fmaf(w.x, x.x, a.x);
  a.y = fmaf(w.y, x.y, a.y);
  a.z = fmaf(w.z, x.z, a.z);
  a.w = fmaf(w.w, x.w, a.w);
}

// Xc[row][i*B + b] = X[row][b*S + i]; one thread per component-major quad (row, i, 4 consecutive blocks)
__global__ void __launch_bounds__(256)
    k_to_cm(const float* __restrict__ X, int64_t rows, int B, int S, float* __restrict__ Xc) {
  const int d = B * S, qpr = d / 4;  // quads per row
  const int64_t n = rows * qpr;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = t / qpr;
    const int p = (int)(t - row * qpr) * 4;  // component-major position of the quad's first element
    const int i = p / B, b = p - i * B;      // B % 4 == 0: the quad stays inside component i
    const float* xr = X + row * d;
    float4 v;
    v.x = __ldg(xr + (b + 0) * S + i);
    v.y = __ldg(xr + (b + 1) * S + i);
    v.z = __ldg(xr + (b + 2) * S + i);
    v.w = __ldg(xr + (b + 3) * S + i);
    *reinterpret_cast<float4*>(Xc + row * d + p) = v;
  }
}

// out[row][b*S + i] = act(out[row][b*S + i] + Mc[row][i*B + b]); every row-major element is touched by one thread
__global__ void __launch_bounds__(256)
    k_cm_add(const float* __restrict__ Mc, int64_t rows, int B, int S, int relu, float* __restrict__ out) {
  const int d = B * S, qpr = d / 4;
  const int64_t n = rows * qpr;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = t / qpr;
    const int p = (int)(t - row * qpr) * 4;
    const int i = p / B, b = p - i * B;
    const float4 v = *reinterpret_cast<const float4*>(Mc + row * d + p);
    float* o = out + row * d;
    const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float r = o[(b + c) * S + i] + vv[c];
      if (relu) r = fmaxf(r, 0.f);
      o[(b + c) * S + i] = r;
    }
  }
}

// Wc[((w*S + a)*S + c)*B + b] = W_dir[r][b][i][j]   with (i, j) = transpose ? (c, a) : (a, c)
__global__ void __launch_bounds__(256)
    k_relayout_cm(const float* __restrict__ Wf, const float* __restrict__ Wb, int R, int B, int S, int transpose,
                  float* __restrict__ Wc) {
  const int64_t per = (int64_t)S * S * B, n = 2 * (int64_t)R * per;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
    const int w = (int)(t / per);
    int rem = (int)(t - w * per);
    const int a = rem / (S * B);
    rem -= a * S * B;
    const int c = rem / B, b = rem - c * B;
    const int i = transpose ? c : a, j = transpose ? a : c;
    const float* src = (w < R) ? Wf + (int64_t)w * per : Wb + (int64_t)(w - R) * per;
    Wc[t] = __ldg(src + ((int64_t)b * S + i) * S + j);
  }
}

// dW_dir[r][b][i][j] = dWc[((w*S + i)*S + j)*B + b]
__global__ void __launch_bounds__(256)
    k_unlayout_cm(const float* __restrict__ dWc, int R, int B, int S, float* __restrict__ dWf,
                  float* __restrict__ dWb) {
  const int64_t per = (int64_t)S * S * B, n = 2 * (int64_t)R * per;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
    const int w = (int)(t / per);
    int rem = (int)(t - w * per);  // row-major destination index inside the table: (b*S + i)*S + j
    const int b = rem / (S * S);
    rem -= b * S * S;
    const int i = rem / S, j = rem - i * S;
    float* dst = (w < R) ? dWf + (int64_t)w * per : dWb + (int64_t)(w - R) * per;
    dst[((int64_t)b * S + i) * S + j] = dWc[(((int64_t)w * S + i) * S + j) * B + b];
  }
}

// One warp per work item (<= 128 messages of ONE weight id, sorted by row); lane = 4 consecutive blocks.
//   DW = false:  outc[row] += Wc[w] . (sum over the row's run of norm * Xc[nbr])           (vector reductions)
//   DW = true :  dWc[w]    += (sum over the run of norm * Xc[nbr]) (x) Hc[row]               (item-local, one flush)
template <int S, bool DW>
__global__ void __launch_bounds__(CM_WARPS * 32, 2)
    k_block_cm(const WorkItem* __restrict__ items, int n_items, const int32_t* __restrict__ r_row,
               const int32_t* __restrict__ r_nbr, const float* __restrict__ r_norm, const float* __restrict__ Xc,
               int ldx, int B, const float* __restrict__ Wc, float* __restrict__ outc, int ldo,
               const float* __restrict__ Hc, int ldh, float* __restrict__ dWc) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int item = blockIdx.x * CM_WARPS + warp;
  if (item >= n_items) return;
  const int q = blockIdx.y * 32 + lane;  // block quad owned by this lane
  const bool act = 4 * q < B;
  const int col = act ? 4 * q : 0;       // lanes past the last quad shadow quad 0 and never store
  const int4 itv = __ldg(reinterpret_cast<const int4*>(items) + item);
  const int beg = itv.x, end = itv.y, w = itv.z;

  float4 wreg[S][S];  // DW = false: the item's weights, resident for all its messages
  float4 acc[S][S];   // DW = true : the item's weight-gradient accumulators   (the unused array is never touched)
  float4 xs[S], hcur[S];
#pragma unroll
  for (int i = 0; i < S; ++i) {
    xs[i] = zero4();
    if constexpr (DW) hcur[i] = zero4();
#pragma unroll
    for (int j = 0; j < S; ++j) {
      if constexpr (DW)
        acc[i][j] = zero4();
      else
        wreg[i][j] = ldg4(Wc + (((size_t)w * S + i) * S + j) * B + col);
    }
  }
  int cur = -1;

  auto flush = [&](int row) {
    if constexpr (!DW) {
      float* po = outc + (size_t)row * ldo + col;
#pragma unroll
      for (int i = 0; i < S; ++i) {
        float4 y = zero4();
#pragma unroll
        for (int j = 0; j < S; ++j) fma4v(y, wreg[i][j], xs[j]);
        if (act) red4(po + (size_t)i * B, y);
      }
    } else {
#pragma unroll
      for (int i = 0; i < S; ++i)
#pragma unroll
        for (int j = 0; j < S; ++j) fma4v(acc[i][j], xs[i], hcur[j]);
    }
  };

  constexpr int U = DW ? 1 : 2;  // rows in flight per lane (S 128-bit loads each); DW already holds 140 floats of state
  for (int base = beg; base < end; base += 32) {
    const int n = min(32, end - base);
    int my_row = 0, my_nbr = 0;
    float my_nm = 0.f;
    if (lane < n) {
      my_row = __ldg(r_row + base + lane);
      my_nbr = __ldg(r_nbr + base + lane);
      my_nm = __ldg(r_norm + base + lane);
    }
    for (int t = 0; t < n; t += U) {
      float4 x[U][S], hx[U][S];
      int rv[U];
      float nm[U];
      bool starts[U];
      int prev = cur;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int tt = min(t + u, n - 1);
        const int src = __shfl_sync(FULL, my_nbr, tt);
        rv[u] = __shfl_sync(FULL, my_row, tt);
        nm[u] = __shfl_sync(FULL, my_nm, tt);
        starts[u] = (t + u < n) && (rv[u] != prev);  // warp-uniform
        prev = rv[u];
        const float* xr = Xc + (size_t)src * ldx + col;
#pragma unroll
        for (int j = 0; j < S; ++j) x[u][j] = ldg4(xr + (size_t)j * B);
        if constexpr (DW) {  // the run's own layer-input row travels with the run's first gathered row
          const float* hr = Hc + (size_t)rv[u] * ldh + col;
#pragma unroll
          for (int j = 0; j < S; ++j) hx[u][j] = starts[u] ? ldg4(hr + (size_t)j * B) : zero4();
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (t + u < n) {
          if (starts[u]) {
            if (cur >= 0) flush(cur);
            cur = rv[u];
#pragma unroll
            for (int j = 0; j < S; ++j) xs[j] = zero4();
            if constexpr (DW) {
#pragma unroll
              for (int j = 0; j < S; ++j) hcur[j] = hx[u][j];
            }
          }
#pragma unroll
          for (int j = 0; j < S; ++j) fma4s(xs[j], nm[u], x[u][j]);
        }
      }
    }
  }
  if (cur >= 0) flush(cur);
  if constexpr (DW) {
    if (act) {
#pragma unroll
      for (int i = 0; i < S; ++i)
#pragma unroll
        for (int j = 0; j < S; ++j) red4(dWc + (((size_t)w * S + i) * S + j) * B + col, acc[i][j]);
    }
  }
}

int grid_for(int64_t n) {
  int64_t b = (n + 255) / 256;
  if (b > 148 * 16) b = 148 * 16;
  return (int)(b < 1 ? 1 : b);
}

int cm_check(const char* what) {
  ++g_rgcn_launches;
  return rgcn_check_cuda(cudaGetLastError(), what);
}

}  // namespace

bool block_cm_supported(int d, int s) {
  if (s != 5 || d % s != 0) return false;
  const int B = d / s;
  return B % 4 == 0 && B >= 4;
}

int launch_to_cm(const float* X, int64_t rows, int B, int s, float* Xc, cudaStream_t st) {
  if (rows == 0) return RGCN_OK;
  k_to_cm<<<grid_for(rows * (int64_t)(B * s / 4)), 256, 0, st>>>(X, rows, B, s, Xc);
  return cm_check("k_to_cm");
}

int launch_cm_add(const float* Mc, int64_t rows, int B, int s, int relu, float* out, cudaStream_t st) {
  if (rows == 0) return RGCN_OK;
  k_cm_add<<<grid_for(rows * (int64_t)(B * s / 4)), 256, 0, st>>>(Mc, rows, B, s, relu, out);
  return cm_check("k_cm_add");
}

int launch_relayout_cm(const float* Wf, const float* Wb, int R, int B, int s, int transpose, float* Wc,
                       cudaStream_t st) {
  if (R == 0) return RGCN_OK;
  k_relayout_cm<<<grid_for(2 * (int64_t)R * s * s * B), 256, 0, st>>>(Wf, Wb, R, B, s, transpose, Wc);
  return cm_check("k_relayout_cm");
}

int launch_unlayout_cm(const float* dWc, int R, int B, int s, float* dWf, float* dWb, cudaStream_t st) {
  if (R == 0) return RGCN_OK;
  k_unlayout_cm<<<grid_for(2 * (int64_t)R * s * s * B), 256, 0, st>>>(dWc, R, B, s, dWf, dWb);
  return cm_check("k_unlayout_cm");
}

int launch_block_cm(const WorkItem* items, int n_items, const int32_t* r_row, const int32_t* r_nbr,
                    const float* r_norm, const float* Xc, int ldx, int B, int s, const float* Wc, float* outc,
                    int ldo, const float* Hc, int ldh, float* dWc, cudaStream_t st) {
  if (n_items == 0) return RGCN_OK;
  if (s != 5) {
    rgcn_set_error("component-major block kernel: only 5x5 blocks are instantiated");
    return RGCN_ERR_INVALID;
  }
  dim3 grid((n_items + CM_WARPS - 1) / CM_WARPS, (B / 4 + 31) / 32);
  if (dWc)
    k_block_cm<5, true><<<grid, CM_WARPS * 32, 0, st>>>(items, n_items, r_row, r_nbr, r_norm, Xc, ldx, B, Wc, outc,
                                                         ldo, Hc, ldh, dWc);
  else
    k_block_cm<5, false><<<grid, CM_WARPS * 32, 0, st>>>(items, n_items, r_row, r_nbr, r_norm, Xc, ldx, B, Wc,
                                                          outc, ldo, Hc, ldh, dWc);
  return cm_check("k_block_cm");
}
