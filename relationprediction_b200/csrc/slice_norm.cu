// slice_norm.cu -- squared norms of the UN-AGGREGATED gradient slices of gathered variables (sm_100a).
//
// In the reference every variable that is read through tf.nn.embedding_lookup (the block tables W_forward /
// W_backward, gcn_basis_concat.py:38-39; the relation table of the decoder, bilinear_diag.py:18) receives an
// IndexedSlices gradient from tf.gradients: one slice per edge / per triple, duplicates NOT summed.
// tf.clip_by_global_norm (optimization/tensorflow_backend/algorithms.py:65-68) takes the global norm over those
// slice VALUES, so the clipping scale is  max_norm / sqrt( sum_dense |g|^2 + sum_sparse sum_slices |slice|^2 ),
// which is not the norm of the summed dense gradients.  These kernels produce the sparse terms:
//   block tables : slice of message m = norm_m * G[dst_m]_b (outer) H[src_m]_b per block b, so
//                  |slice_m|^2 = norm_m^2 * sum_b |G[dst_m]_b|^2 |H[src_m]_b|^2
//                  -> per-node per-block squared norms (one pass over H and G) + a B-wide dot per message.
// The relation-table term is produced inside k_distmult_bwd (distmult.cu).
#include <cuda_runtime.h>

#include "kernels.cuh"

#define FULL 0xffffffffu

namespace {

// XB[v][b] = sum_i X[v][b*s+i]^2 ; one warp per row
__global__ void __launch_bounds__(256)
    k_block_sqnorm(const float* __restrict__ X, int64_t rows, int ld, int B, int s, float* __restrict__ XB) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int64_t v = (int64_t)blockIdx.x * 8 + warp; v < rows; v += (int64_t)gridDim.x * 8) {
    const float* x = X + (size_t)v * ld;
    for (int b = lane; b < B; b += 32) {
      float a = 0.f;
      for (int i = 0; i < s; ++i) {
        const float t = __ldg(x + b * s + i);
        a = fmaf(t, t, a);
      }
      XB[(size_t)v * B + b] = a;
    }
  }
}

// one warp per weight-id-major work item (all its messages share the weight id, i.e. the table)
__global__ void __launch_bounds__(256)
    k_block_slice_sumsq(const WorkItem* __restrict__ items, int n_items, const int32_t* __restrict__ r_row,
                        const int32_t* __restrict__ r_nbr, const float* __restrict__ r_norm,
                        const float* __restrict__ GB, const float* __restrict__ HB, int B, int half,
                        float* __restrict__ out2) {
  __shared__ float sh[2][8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float acc_f = 0.f, acc_b = 0.f;
  for (int it = blockIdx.x * 8 + warp; it < n_items; it += gridDim.x * 8) {
    const int4 iv = __ldg(reinterpret_cast<const int4*>(items) + it);
    float a = 0.f;
    for (int m = iv.x; m < iv.y; ++m) {
      const float* g = GB + (size_t)__ldg(r_row + m) * B;
      const float* h = HB + (size_t)__ldg(r_nbr + m) * B;
      const float nm = __ldg(r_norm + m);
      float p = 0.f;
      for (int b = lane; b < B; b += 32) p = fmaf(__ldg(g + b), __ldg(h + b), p);
      a = fmaf(nm * nm, p, a);
    }
    if (iv.z >= half)
      acc_b += a;
    else
      acc_f += a;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    acc_f += __shfl_xor_sync(FULL, acc_f, o);
    acc_b += __shfl_xor_sync(FULL, acc_b, o);
  }
  if (lane == 0) {
    sh[0][warp] = acc_f;
    sh[1][warp] = acc_b;
  }
  __syncthreads();
  if (threadIdx.x < 2) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += sh[threadIdx.x][w];
    if (t != 0.f) atomicAdd(out2 + threadIdx.x, t);
  }
}

int grid_rows(int64_t n) {
  int64_t b = (n + 7) / 8;
  if (b > 148 * 8) b = 148 * 8;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace

int launch_block_sqnorm(const float* X, int64_t rows, int ld, int B, int s, float* XB, cudaStream_t st) {
  if (rows == 0) return RGCN_OK;
  k_block_sqnorm<<<grid_rows(rows), 256, 0, st>>>(X, rows, ld, B, s, XB);
  ++g_rgcn_launches;
  return rgcn_check_cuda(cudaGetLastError(), "k_block_sqnorm");
}

int launch_block_slice_sumsq(const WorkItem* items, int n_items, const int32_t* r_row, const int32_t* r_nbr,
                             const float* r_norm, const float* GB, const float* HB, int B, int half, float* out2,
                             cudaStream_t st) {
  if (n_items == 0) return RGCN_OK;
  k_block_slice_sumsq<<<grid_rows(n_items), 256, 0, st>>>(items, n_items, r_row, r_nbr, r_norm, GB, HB, B, half, out2);
  ++g_rgcn_launches;
  return rgcn_check_cuda(cudaGetLastError(), "k_block_slice_sumsq");
}
