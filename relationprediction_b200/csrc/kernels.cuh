// kernels.cuh -- device helpers + launcher declarations shared by rgcn_kernels.cu / api.cu
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "graph.h"

#define RGCN_WARPS_PER_BLOCK 8
#define RGCN_THREADS (RGCN_WARPS_PER_BLOCK * 32)

extern int64_t g_rgcn_launches;  // counted on the host at every kernel launch of this library

struct AggLaunch {
  const WorkItem* items;
  int n_items;
  const int32_t* nbr;   // gather row per message
  const int32_t* relw;  // weight id per message
  const float* norm;    // per message
  const float* X;       // gathered feature matrix, row-major, leading dimension ldx
  int ldx;
  int d;                // feature width
  const int32_t* split_nitems;
  float* scratch;  // [n_split, d]  zeroed
  int* counters;   // [n_split * n_slabs] zeroed
};

// Block-diagonal aggregation (forward, and backward-w.r.t.-H with the transposed table):
//   out[row,:] = act( out[row,:] (*mask/keep) + sum_m norm_m * Wt[relw_m] (.) X[nbr_m,:] )
// Wt layout: [n_relw][s][d] with Wt[w][j][b*s+i] = coefficient multiplying x[b*s+j] in y[b*s+i].
int launch_block_agg(const AggLaunch& a, int s, const float* Wt, float* out, const uint8_t* mask,
                     float inv_keep, int relu, cudaStream_t st);

// Weight-id major variant of the above (weights in registers, vector reductions into `out`, which
// must already hold the self-loop term).  Supported block sizes: block_rel_supported().
bool block_rel_supported(int d, int s);
bool block_rel_fuse_dw_supported(int d, int s);
// dWt != nullptr (backward pass, X = G, rows = sources, Hrow = layer input): additionally accumulates
// the block weight gradient in the j-major layout (dWt zeroed by the caller) in the same walk.
int launch_block_rel(const WorkItem* items, int n_items, const int32_t* r_row, const int32_t* r_nbr,
                     const float* r_norm, const float* X, int ldx, int d, int s, const float* Wt,
                     float* out, const float* Hrow, int ldh, float* dWt, cudaStream_t st);

// Same contract as launch_block_rel with the gathered rows staged through shared memory by TMA bulk copies
// (block_staged.cu); block sizes 4, 8, 16.
bool block_stg_supported(int d, int s);
int launch_block_stg(const WorkItem* items, int n_items, const int32_t* r_row, const int32_t* r_nbr,
                     const float* r_norm, const float* X, int ldx, int d, int s, const float* Wt, float* out,
                     const float* Hrow, int ldh, float* dWt, cudaStream_t st);

// Block-diagonal weight gradient, weight-id major:
//   dWt[w][j][b*s+i] += sum_{m: relw_m = w} norm_m * G[dst_m, b*s+i] * H[src_m, b*s+j]
int launch_block_dw(const WorkItem* items, int n_items, const int32_t* r_dst, const int32_t* r_src,
                    const float* r_norm, const float* H, int ldh, const float* G, int ldg, int d,
                    int s, float* dWt, cudaStream_t st);

// Re-layout of the reference weight tables [R,B,s,s] (W.x orientation) into the kernel tables.
//   transpose = 0:  Wt[w][j][b*s+i] = W[w][b][i][j]      (forward)
//   transpose = 1:  Wt[w][i][b*s+j] = W[w][b][i][j]      (backward w.r.t. H: y = W^T g)
int launch_block_relayout(const float* Wf, const float* Wb, int R, int B, int s, int transpose,
                          float* Wt, cudaStream_t st);
// inverse of transpose=0 layout: dW[w][b][i][j] = dWt[w][j][b*s+i]
int launch_block_unlayout(const float* dWt, int R, int B, int s, float* dWf, float* dWb,
                          int accumulate, int table_t, cudaStream_t st);

// Basis aggregation: Agg[row][dir][...] = sum_m norm_m * C[relw_m, b] * X[nbr_m, k]
//   layout 0 (interleaved): index k*B + b      (matches V.reshape(d_in*B, d_out) rows)
//   layout 1 (planar):      index b*d + k      (matches V.reshape(d_in, B*d_out) columns)
// dir = relw >= n_relw/2.  Row stride of Agg is 2*d*B, direction stride d*B.
int launch_basis_agg(const AggLaunch& a, const float* C, int B, int n_relw, int layout, float* Agg,
                     cudaStream_t st);

// Basis coefficient gradient (destination major):
//   dC[w][b] += sum_{m into row, relw_m = w} norm_m * < H[src_m,:], dAgg[row][dir][:, b] >
int launch_basis_dc(const AggLaunch& a, const float* dAgg, int B, int n_relw, float* dC,
                    cudaStream_t st);

// slice_norm.cu: squared norms of the un-aggregated IndexedSlices gradients (tf.clip_by_global_norm semantics)
int launch_block_sqnorm(const float* X, int64_t rows, int ld, int B, int s, float* XB, cudaStream_t st);
// out2[0] += sum over forward-table messages, out2[1] += backward-table messages of norm^2 * <GB[dst], HB[src]>
int launch_block_slice_sumsq(const WorkItem* items, int n_items, const int32_t* r_row, const int32_t* r_nbr,
                             const float* r_norm, const float* GB, const float* HB, int B, int half, float* out2,
                             cudaStream_t st);

// Elementwise helpers
// G = dOut * (out > 0 if relu);  dS = G * mask * inv_keep (only if mask != null, else dS untouched)
int launch_grad_prologue(const float* dOut, const float* out, const uint8_t* mask, float inv_keep,
                         int relu, int64_t n, float* G, float* dS, cudaStream_t st);
// x = x * mask * inv_keep (if mask) ; x = relu(x) (if relu)
int launch_mask_relu(float* x, const uint8_t* mask, float inv_keep, int relu, int64_t n,
                     cudaStream_t st);
// dst[rows[i], :] += src[i, :], rows unique within the call
int launch_rows_add(float* dst, const int64_t* rows, const float* src, int64_t n, int d, cudaStream_t st);
int launch_rows_gather(float* dst, const float* src, const int64_t* rows, int64_t n, int d, int max_ctas,
                       cudaStream_t st);
// zero rows listed in `rows` of a [*, width] matrix
int launch_zero_rows(float* A, int64_t width, const int32_t* rows, int n_rows, cudaStream_t st);

// fp32-accurate tensor-core GEMM (gemm_tf32x3.cu): C[M,N] (+)= A[M,K] * Bt[N,K]^T with Bt pre-split
int launch_gemm_split_b(const float* B, int64_t ldb, int N, int K, int transposed, float* hi, float* lo,
                        cudaStream_t st);
int launch_gemm_tf32x3(const float* A, int64_t lda, const float* Bt_hi, const float* Bt_lo, int64_t ldb,
                       float* C, int64_t ldc, int M, int N, int K, int accumulate, cudaStream_t st);

int launch_gemm_rank_tf32x3(const float* Q, int64_t ldq, const float* Bt_hi, const float* Bt_lo, int64_t ldb, int M,
                            int N, int K, const float* gold_sig, const int32_t* gold_col, const uint32_t* known,
                            int words, int32_t* raw_cnt, int32_t* known_cnt, cudaStream_t st);

int launch_gemm_tn_tf32x3(const float* A, int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc,
                          int M, int N, int K, int accumulate, cudaStream_t st);

// DistMult
// queries + gold scores of the fused scorer/ranker: side 0 (subjects corrupted): Q[t] = rel[r] * codes[o], gold = s;
// side 1 (objects corrupted): Q[t] = codes[s] * rel[r], gold = o.  gold_sig[t] = sigmoid(<Q[t], codes[gold]>)
int launch_distmult_rank_prepare(const float* codes, const float* rel, int d, const int32_t* X, int64_t n, int side,
                                 float* Q, float* gold_sig, int32_t* gold_col, cudaStream_t st);
int launch_distmult_rank_finalize(const int32_t* raw_cnt, const int32_t* known_cnt, int64_t n, int32_t* raw_rank,
                                  int32_t* filtered_rank, cudaStream_t st);
int launch_distmult_forward(const float* codes, const float* rel, int d, const int32_t* X, int64_t N,
                            const float* Y, float* energies, float* loss_out, cudaStream_t st);
int launch_distmult_backward(const float* codes, const float* rel, int d, const int32_t* X,
                             int64_t N, const float* Y, const float* energies, float g_loss,
                             float g_reg, const float* g_scale_dev, const float* g_energy,
                             float* dcodes, float* drel, float* rel_slice_sumsq, cudaStream_t st);
