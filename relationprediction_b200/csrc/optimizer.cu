// optimizer.cu -- next row N1: global-norm gradient clipping + Adam with TensorFlow-1.x
// semantics, as the reference's optimizer stack applies them
// (optimization/tensorflow_backend/algorithms.py:65-68 tf.clip_by_global_norm, :36-42 tf.train.AdamOptimizer):
//   scale = clip_norm * min(1 / global_norm, 1 / clip_norm)            (global_norm = sqrt(sum over ALL tensors g^2))
//   lr_t  = lr * sqrt(1 - beta2^t) / (1 - beta1^t)
//   m <- beta1 m + (1 - beta1) g ;  v <- beta2 v + (1 - beta2) g^2 ;  p <- p - lr_t * m / (sqrt(v) + eps)
// Bandwidth-bound elementwise kernels; the clip scale is read from a device scalar so no host sync is needed.
#include <cuda_runtime.h>

#include "kernels.cuh"

namespace {

__global__ void __launch_bounds__(256)
    k_sumsq(const float* __restrict__ g, int64_t n, float* __restrict__ acc) {
  double local = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float x = g[i];
    local += (double)x * (double)x;
  }
  for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
  __shared__ double sh[8];
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += sh[w];
    atomicAdd(acc, (float)t);
  }
}

__global__ void __launch_bounds__(256)
    k_adam(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
           int64_t n, float lr_t, float beta1, float beta2, float eps, const float* __restrict__ sumsq,
           float max_norm) {
  float scale = 1.f;
  if (sumsq && max_norm > 0.f) {
    const float gn = sqrtf(__ldg(sumsq));
    // tf.clip_by_global_norm: clip_norm * min(1/global_norm, 1/clip_norm); global_norm = 0 -> 1
    scale = max_norm * fminf(1.f / gn, 1.f / max_norm);
  }
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float gi = g[i] * scale;
    const float mi = beta1 * m[i] + (1.f - beta1) * gi;
    const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] -= lr_t * mi / (sqrtf(vi) + eps);
  }
}

int blocks_for(int64_t n) {
  int64_t b = (n + 255) / 256;
  if (b > 148 * 8) b = 148 * 8;
  return (int)(b < 1 ? 1 : b);
}

}  // namespace

extern "C" int rgcn_sumsq_accumulate(const float* g, int64_t n, float* acc_dev, void* stream) {
  if (!g || !acc_dev || n < 0) {
    rgcn_set_error("rgcn_sumsq_accumulate: bad arguments");
    return RGCN_ERR_INVALID;
  }
  if (n == 0) return RGCN_OK;
  k_sumsq<<<blocks_for(n), 256, 0, (cudaStream_t)stream>>>(g, n, acc_dev);
  ++g_rgcn_launches;
  return rgcn_check_cuda(cudaGetLastError(), "k_sumsq");
}

extern "C" int rgcn_adam_update(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1,
                                float beta2, float eps, int64_t step, const float* sumsq_dev, float max_norm,
                                void* stream) {
  if (!p || !g || !m || !v || n < 0 || step < 1) {
    rgcn_set_error("rgcn_adam_update: bad arguments (step is 1-based)");
    return RGCN_ERR_INVALID;
  }
  if (n == 0) return RGCN_OK;
  const double lr_t = (double)lr * sqrt(1.0 - pow((double)beta2, (double)step)) / (1.0 - pow((double)beta1, (double)step));
  k_adam<<<blocks_for(n), 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, (float)lr_t, beta1, beta2, eps, sumsq_dev,
                                                          max_norm);
  ++g_rgcn_launches;
  return rgcn_check_cuda(cudaGetLastError(), "k_adam");
}
