/*
 * rgcn_b200.h -- C-ABI of librgcn_b200.so: the B200-native (sm_100a) R-GCN relational
 * message-passing hot path (block-diagonal + basis decomposition, forward and backward) and the
 * DistMult triple scorer.
 *
 * The reference (MichSchli/RelationPrediction) is pure Python/TensorFlow-1 and has no FFI; its
 * "operator interface" for this path is the set of plugin hooks that build a TF graph.  Each entry
 * point below cites the reference hook(s) (file:line under /root/reference/code) whose *executed*
 * TF ops it replaces.  INTEGRATION.md shows the ctypes binding a reference maintainer would add.
 *
 * Conventions
 *   - every function returns int: 0 = ok, negative = error (RGCN_ERR_*); rgcn_last_error() gives text.
 *   - no exceptions cross the boundary; no torch types; plain pointers + explicit sizes.
 *   - the caller owns every tensor (device pointers unless the name ends in _host); the library
 *     owns only the opaque graph handle.  `stream` is a cudaStream_t passed as void*.
 *   - all feature / weight tensors are dense row-major fp32; all indices are int32.
 *   - all calls are asynchronous with respect to the host (work is enqueued on `stream`).
 *   - a "message" is one (source row -> destination row, relation-weight id) item.  A triple
 *     (s, r, o) yields two messages: forward  s->o with weight id r       (W_forward[r])
 *                                    backward o->s with weight id r + R   (W_backward[r])
 *     (reference: extras/graph_representations.py:21-27, message_gcn.py:28-42).
 */
#ifndef RGCN_B200_H
#define RGCN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RGCN_OK 0
#define RGCN_ERR_INVALID (-1)   /* bad argument (shape, null pointer, index out of range)      */
#define RGCN_ERR_CUDA (-2)      /* CUDA runtime / cuBLAS failure; text in rgcn_last_error()    */
#define RGCN_ERR_NOMEM (-3)     /* host or device allocation failed                            */
#define RGCN_ERR_WORKSPACE (-4) /* caller workspace too small (see *_workspace_bytes)          */
#define RGCN_ERR_NODEVICE (-5)  /* device entry point called on a host-only graph / no GPU     */

/* normalisation modes for rgcn_graph_create (extras/graph_representations.py:84-93,124-133) */
#define RGCN_NORM_CANONICAL 0 /* 1 / (#messages of that direction into the destination)        */
#define RGCN_NORM_EXPLICIT 1  /* caller supplies norm_f[E], norm_b[E] (e.g. tf_unsorted_compat) */
#define RGCN_NORM_NONE 2      /* all ones ('none' branch, :70-82)                               */

typedef struct rgcn_graph rgcn_graph_t; /* opaque */

int rgcn_version(void);
const char* rgcn_last_error(void);
/* number of CUDA kernels this library has launched in this process (bench.py "gpu_launches") */
int64_t rgcn_launch_count(void);

/* Library options.  "block_algo": 0 = destination-major aggregation (deterministic summation order,
 * epilogue fused), 1 = weight-id-major aggregation with the gathered rows in registers (block weights in
 * registers, vector reductions in L2; fp32 summation order not reproducible run to run), 3 = the same walk with the
 * gathered rows staged through shared memory by TMA bulk copies / cp.async (block sizes 4, 8, 16; fastest),
 * -1 = auto (default): 3 where the block size allows, else 1, else 0.
 * The environment variable RGCN_BLOCK_ALGO overrides the option.
 * "graph_views": which sorted views GPU-prepared graphs created AFTER the call get: 1 = the two CSR views
 * (deterministic block mode, basis layers), 2 = the two weight-id-major views (default block kernels),
 * 3 = all four (default).  A 200 M-message graph saves ~5 GB and half its preparation time with 2; entry
 * points return RGCN_ERR_INVALID when the view they walk is absent. */
int rgcn_set_option(const char* name, int64_t value);

/* Dense fp32-accurate GEMM on the tcgen05 tensor cores (3xTF32 split, TMEM accumulators):
 *   C[M,N] = (accumulate ? C : 0) + A[M,K] * op(B),   op(B) = B[K,N] (b_is_nk = 0) or B[N,K]^T (b_is_nk = 1)
 * all row-major fp32 device pointers; K, N and the leading dimensions must be multiples of 4.
 * workspace: 2*N*K floats (the hi/lo split of B).  This is the kernel the layer entry points use for
 * the self-loop terms (gcn_basis.py:70-71 / gcn_basis_concat.py:65-66). */
int rgcn_gemm_tf32x3(const float* A, int64_t lda, const float* B, int64_t ldb, int b_is_nk, float* C,
                     int64_t ldc, int32_t M, int32_t N, int32_t K, int accumulate, void* workspace,
                     int64_t workspace_bytes, void* stream);

/* C[M,N] = (accumulate ? C : 0) + A^T B with A [K,M], B [K,N] row-major (the V-long reductions of
 * the backward pass: dW_self = H^T dS).  Same tensor-core path, both operands MN-major, split-K with
 * vector reductions into C (fp32 summation order across splits not reproducible run to run).
 * M, N and the leading dimensions must be multiples of 4. */
int rgcn_gemm_tn_tf32x3(const float* A, int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc,
                        int32_t M, int32_t N, int32_t K, int accumulate, void* stream);

/* Host-side neighbourhood-expansion edge sampler (next row N2; train.py:161-198): writes sample_size
 * DISTINCT edge ids.  Same stochastic process as the reference (vertex ~ unpicked-degree x seen, then a
 * uniform unpicked incident edge), O(log V) per draw instead of O(V); its own random stream (seed). */
int rgcn_sample_edge_neighborhood(const int32_t* triples_host, int64_t E, int32_t V, int64_t sample_size,
                                  uint64_t seed, int32_t* out_edges_host);

/* The same sampler with the per-dataset incidence structure built once: create a handle for the training
 * triples, then draw any number of samples from it.  Draws are thread-compatible (the handle is read-only, every
 * draw owns its scratch copies), so host threads can prepare samples concurrently.  A draw with the same seed
 * returns exactly what rgcn_sample_edge_neighborhood returns. */
typedef struct rgcn_sampler rgcn_sampler_t;
int rgcn_sampler_create(const int32_t* triples_host, int64_t E, int32_t V, rgcn_sampler_t** out);
int rgcn_sampler_draw(const rgcn_sampler_t* sampler, int64_t sample_size, uint64_t seed, int32_t* out_edges_host);
/* One whole training sample of train.py:140-198 in one call (so that the host threads preparing samples hold no
 * interpreter lock): batch = rgcn_sampler_draw(batch) edges; graph_split_host [split,3] = `split` of them uniformly
 * without replacement (np.random.choice(ids, split, replace=False)); X_host [(neg_rate+1)*batch, 3] / Y_host = the batch
 * followed by neg_rate corrupted copies with labels 1 / 0 (common/auxilliaries.py NegativeSampler.transform: fair coin
 * object-or-subject, uniform replacement entity).  Same stochastic process, own random stream. */
int rgcn_sampler_draw_batch(const rgcn_sampler_t* sampler, int32_t batch, int32_t split, int32_t neg_rate, uint64_t seed,
                            int32_t* graph_split_host, int32_t* X_host, float* Y_host);
void rgcn_sampler_destroy(rgcn_sampler_t* sampler);

/* Next row N1: global-norm clipping + Adam with TensorFlow-1.x semantics
 * (optimization/tensorflow_backend/algorithms.py:65-68 and :36-42).  Call rgcn_sumsq_accumulate on every
 * gradient tensor into one zeroed device float, then rgcn_adam_update on every (param, grad, m, v) with that
 * scalar: scale = max_norm * min(1/sqrt(sumsq), 1/max_norm) (skipped when sumsq_dev is NULL or max_norm <= 0);
 * lr_t = lr*sqrt(1-beta2^step)/(1-beta1^step); p -= lr_t * m / (sqrt(v) + eps).  step is 1-based. */
int rgcn_sumsq_accumulate(const float* g, int64_t n, float* acc_dev, void* stream);
int rgcn_adam_update(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1,
                     float beta2, float eps, int64_t step, const float* sumsq_dev, float max_norm,
                     void* stream);

/* Optional per-kernel timing (bench.py roofline): when enabled, every layer entry point records a
 * CUDA event on its stream after each internal stage.  rgcn_profile_read() synchronises, writes
 * the stage durations (ms) and their '\n'-separated names, clears the log and returns the count. */
int rgcn_profile_enable(int enable);
int rgcn_profile_read(float* ms_out, int max_entries, char* names_out, int names_cap);

/* ------------------------------------------------------------------------------------------------
 * Graph preparation.  Replaces Representation/MessageGraph
 * (extras/graph_representations.py:21-27 index vectors, :84-93 / :124-133 normalised incidence).
 *
 * triples_host : int32 [E,3], columns (subject, relation, object), host memory.
 * V            : number of entities (rows of H); R: number of relations.
 * device       : CUDA device ordinal, or -1 to build the host-side structure only (CPU tests).
 * Builds, deterministically (stable counting sorts; message id of triple k is k forward, E+k
 * backward): destination-CSR sorted by (dst, weight-id), source-CSR sorted by (src, weight-id),
 * two weight-id-major lists sorted by (supertile(dst), weight-id, dst) and (supertile(src),
 * weight-id, src), per-message norm, and warp work lists.
 * ---------------------------------------------------------------------------------------------- */
int rgcn_graph_create(const int32_t* triples_host, int64_t E, int32_t V, int32_t R, int norm_mode,
                      const float* norm_f_host, const float* norm_b_host, int device, void* stream,
                      rgcn_graph_t** out);

/* Generic message-list constructor used by the 1-D node-sharded path: rank-local
 * destinations [0,V_dst), sources index an extended row space [0,V_src) (local rows then halo
 * rows).  All arrays host, length M.  relw in [0, n_relw). */
int rgcn_graph_create_messages(const int32_t* dst_host, const int32_t* src_host,
                               const int32_t* relw_host, const float* norm_host, int64_t M,
                               int32_t V_dst, int32_t V_src, int32_t n_relw, int device,
                               void* stream, rgcn_graph_t** out);

/* The same two constructors for index arrays that ALREADY live on `device` (int32 / float device pointers,
 * same shapes and meaning as the _host arguments above): the node-sharded path partitions the edge list on
 * the GPU and bench.py generates its synthetic graphs there, so a 100 M-edge list never visits the host.
 * GPU preparation only (device must be >= 0); the arrays are read on `stream` and may be freed by the caller
 * as soon as the call returns.  Replaces the same reference code as rgcn_graph_create
 * (extras/graph_representations.py:21-27, :84-93, :124-133). */
int rgcn_graph_create_device(const int32_t* triples_dev, int64_t E, int32_t V, int32_t R, int norm_mode,
                             const float* norm_f_dev, const float* norm_b_dev, int device, void* stream,
                             rgcn_graph_t** out);
int rgcn_graph_create_messages_device(const int32_t* dst_dev, const int32_t* src_dev,
                                      const int32_t* relw_dev, const float* norm_dev, int64_t M,
                                      int32_t V_dst, int32_t V_src, int32_t n_relw, int device,
                                      void* stream, rgcn_graph_t** out);

/* Opt-in stream-ordered destroy: GPU-prepared graphs return their arrays with cudaFreeAsync on `stream` (no
 * device synchronisation); every kernel that used the graph must be ordered before `stream`'s tail.  Host-prepared
 * graphs take the synchronous path of rgcn_graph_destroy. */
int rgcn_graph_destroy_async(rgcn_graph_t* graph, void* stream);
int rgcn_graph_destroy(rgcn_graph_t* g);

/* info[0]=M messages, [1]=V_dst, [2]=V_src, [3]=n_relw, [4]=#dst work items, [5]=#src work items,
 * [6]=#relw work items, [7]=#split dst rows, [8]=#split src rows, [9]=#(dst,relw) groups,
 * [10]=device, [11]=bytes resident on device, [12]=item_max, [13]=supertile rows, [14]=#supertiles,
 * [15]=#relw work items of the source-keyed view. */
int rgcn_graph_info(const rgcn_graph_t* g, int64_t info[16]);

/* Export of the prepared structure to host memory, for bit-exact index tests. */
enum {
  RGCN_X_DST_ROWPTR = 0, /* int32 [V_dst+1] */
  RGCN_X_DST_SRC = 1,    /* int32 [M]  source row of each message, destination-major order */
  RGCN_X_DST_RELW = 2,   /* int32 [M]  */
  RGCN_X_DST_NORM = 3,   /* float [M]  */
  RGCN_X_DST_MID = 4,    /* int32 [M]  original message id (k or E+k)                      */
  RGCN_X_SRC_ROWPTR = 5, /* int32 [V_src+1] */
  RGCN_X_SRC_DST = 6,
  RGCN_X_SRC_RELW = 7,
  RGCN_X_SRC_NORM = 8,
  RGCN_X_SRC_MID = 9,
  RGCN_X_REL_PTR = 10, /* int32 [n_super*n_relw+1]: weight-id major view keyed (supertile(dst), relw, dst) */
  RGCN_X_REL_DST = 11,
  RGCN_X_REL_SRC = 12,
  RGCN_X_REL_NORM = 13,
  RGCN_X_REL_MID = 14,
  RGCN_X_MSG_NORM = 15, /* float [M] norm in original message order */
  RGCN_X_REL2_PTR = 16, /* second weight-id major view keyed (supertile(src), relw, src) */
  RGCN_X_REL2_SRC = 17,
  RGCN_X_REL2_DST = 18,
  RGCN_X_REL2_NORM = 19,
  RGCN_X_REL2_MID = 20
};
int64_t rgcn_graph_export_bytes(const rgcn_graph_t* g, int which);
int rgcn_graph_export(const rgcn_graph_t* g, int which, void* dst_host, int64_t nbytes);

/* ------------------------------------------------------------------------------------------------
 * Block-diagonal R-GCN layer ("ConcatGcn", encoders/message_gcns/gcn_basis_concat.py:35-83 +
 * message_gcn.py:49-79).
 *
 *   out[v,:] = act( sum_{messages m into v} norm_m * blockdiag(W[relw_m]) . H[src_m,:]
 *                   + dropout(H[v,:] @ W_self) )
 *
 * H      : [V_src, d]   (rows [0,V_dst) are the local nodes: the self-loop uses those)
 * Wf, Wb : [R, B, s, s], s = d / B, "W . x" orientation (gcn_basis_concat.py:46-47):
 *          y[b*s+i] = sum_j W[r,b,i,j] * x[b*s+j].      n_relw of the graph must equal 2R.
 * Wself  : [d, d]
 * drop_mask : uint8 [V_dst, d] keep-mask (1 = keep) or NULL; survivors are scaled by 1/keep
 *          (tf.nn.dropout semantics, message_gcn.py:64; self-loop only).
 * relu   : 1 for hidden layers, 0 for the last layer (common/model_builder.py:275).
 * out    : [V_dst, d].
 * ---------------------------------------------------------------------------------------------- */
int64_t rgcn_block_workspace_bytes(const rgcn_graph_t* g, int32_t d, int32_t B, int backward);

int rgcn_block_forward(const rgcn_graph_t* g, int32_t d, int32_t B, const float* H,
                       const float* Wf, const float* Wb, const float* Wself,
                       const uint8_t* drop_mask, float keep, int relu, float* out, void* workspace,
                       int64_t workspace_bytes, void* stream);

/* Backward of the above (what tf.gradients, optimization/abstract.py:117-118, derives):
 *   G = dOut * (out > 0 if relu);  dS = G * mask / keep
 *   dH      [V_src, d]  (overwritten)   dWf, dWb [R,B,s,s] (overwritten)   dWself [d,d] (overwritten)
 * `out` is the forward result (needed for the ReLU mask only when relu != 0). */
int rgcn_block_backward(const rgcn_graph_t* g, int32_t d, int32_t B, const float* H,
                        const float* Wf, const float* Wb, const float* Wself,
                        const uint8_t* drop_mask, float keep, int relu, const float* out,
                        const float* dOut, float* dH, float* dWf, float* dWb, float* dWself,
                        void* workspace, int64_t workspace_bytes, void* stream);

/* Messages-only parts of the block layer, for graphs whose sources live in a separate row space
 * (the halo rows of the node-sharded path: V_src may be smaller than V_dst):
 *   aggregate          : out[dst,:] += sum_m norm_m * blockdiag(W[relw_m]) . X[src_m,:]      (no self loop)
 *   aggregate_backward : dX [V_src,d] (overwritten) = sum_m norm_m W^T G[dst_m];
 *                        dWf, dWb (overwritten, or += when accumulate_dW) = block outer products.
 * Same per-message arithmetic as gcn_basis_concat.py:35-52 + the SpMMs of :69-75. */
int64_t rgcn_block_aggregate_workspace_bytes(const rgcn_graph_t* g, int32_t d, int32_t B, int backward);
int rgcn_block_aggregate(const rgcn_graph_t* g, int32_t d, int32_t B, const float* X, const float* Wf,
                         const float* Wb, float* out, void* workspace, int64_t workspace_bytes,
                         void* stream);
int rgcn_block_aggregate_backward(const rgcn_graph_t* g, int32_t d, int32_t B, const float* X,
                                  const float* Wf, const float* Wb, const float* G, float* dX,
                                  float* dWf, float* dWb, int accumulate_dW, void* workspace,
                                  int64_t workspace_bytes, void* stream);

/* dst[rows[i], :] += src[i, :] for i < n, rows (int64, device) UNIQUE within the call -- no atomics.  The
 * node-sharded path returns halo gradients grouped by peer; a local row gets at most one contribution per peer,
 * so each peer segment is one call (replaces an atomic index_add over 16 GB per rank at 8 GPUs). */
int rgcn_rows_add(float* dst, const int64_t* rows, const float* src, int64_t n, int32_t d, void* stream);

/* G[i] = out[i] > 0 ? dOut[i] : 0 for i < n (n % 4 == 0): the ReLU gradient of message_gcn.py:64-66 as its own pass.
 * rgcn_block_backward applies it internally; the node-sharded layers need G before their first backward kernel
 * (the halo-source messages run first so that their gradients can travel while the local work runs). */
int rgcn_relu_backward(const float* dOut, const float* out, float* G, int64_t n, void* stream);

/* dst[i, :] = src[rows[i], :] for i < n (rows int64, device).  The halo PUSH of the node-sharded path: `dst` may be
 * (and in that path is) a PEER GPU's buffer mapped into this process (CUDA symmetric / IPC memory), so the rows a
 * peer needs go from H straight over NVLink into the buffer its aggregation kernel reads -- no packed send buffer,
 * no all-to-all (the reference has no multi-device path at all: model.py builds one tf.Session graph).
 * max_ctas > 0 bounds the grid so the push shares the GPU with the layer's local work; 0 = library default. */
int rgcn_rows_gather(float* dst, const float* src, const int64_t* rows, int64_t n, int32_t d, int32_t max_ctas,
                     void* stream);

/* ------------------------------------------------------------------------------------------------
 * Basis-decomposition R-GCN layer ("BasisGcn", encoders/message_gcns/gcn_basis.py:39-88).
 *
 *   m_f[k] = sum_b Cf[r_k,b] * (H[s_k,:] @ Vf[:,b,:])   (and likewise backward with Vb, Cb)
 *   out    = act( A_f m_f + A_b m_b + dropout(H @ W_self) )
 *
 * computed re-associated (aggregate-then-transform): Agg_dir[v,k,b] = sum_m norm_m C[relw_m,b] H[src_m,k]
 * followed by dense GEMMs with Vf/Vb viewed as [d_in*B, d_out].
 * Vf, Vb : [d_in, B, d_out] (gcn_basis.py:18);  Cf, Cb : [R, B] (gcn_basis.py:17).  d_in == d_out == d.
 * `saved` (float [V_dst, 2*d*B]) receives Agg_f | Agg_b and must be handed unchanged to backward.
 * ---------------------------------------------------------------------------------------------- */
int64_t rgcn_basis_workspace_bytes(const rgcn_graph_t* g, int32_t d, int32_t B, int backward);

int rgcn_basis_forward(const rgcn_graph_t* g, int32_t d, int32_t B, const float* H,
                       const float* Vf, const float* Vb, const float* Cf, const float* Cb,
                       const float* Wself, const uint8_t* drop_mask, float keep, int relu,
                       float* out, float* saved, void* workspace, int64_t workspace_bytes,
                       void* stream);

int rgcn_basis_backward(const rgcn_graph_t* g, int32_t d, int32_t B, const float* H,
                        const float* Vf, const float* Vb, const float* Cf, const float* Cb,
                        const float* Wself, const uint8_t* drop_mask, float keep, int relu,
                        const float* out, const float* saved, const float* dOut, float* dH,
                        float* dVf, float* dVb, float* dCf, float* dCb, float* dWself,
                        void* workspace, int64_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * DistMult triple scorer ("BilinearDiag", decoders/bilinear_diag.py:14-34, :63-69).
 *
 *   energy[n] = sum_k codes[X[n,0],k] * rel[X[n,1],k] * codes[X[n,2],k]
 *   loss_out[0] = mean_n( (1-y)x + log1p(exp(-|x|)) + max(-x,0) )            (only if Y != NULL)
 *   loss_out[1] = mean(e1^2) + mean(r^2) + mean(e2^2) over the gathered rows  (un-scaled; the
 *                 caller multiplies by RegularizationParameter, bilinear_diag.py:69)
 * codes : [V, d]; rel : [Vrel, d] (the reference sizes it [EntityCount, d], model_builder.py:134);
 * X : int32 [N,3] device; Y : float [N] device or NULL; energies : float [N]; loss_out : float [2].
 * ---------------------------------------------------------------------------------------------- */
int distmult_forward(const float* codes, const float* rel, int32_t V, int32_t Vrel, int32_t d,
                     const int32_t* X, int64_t N, const float* Y, float* energies, float* loss_out,
                     void* stream);

/* Backward: given upstream scalars g_loss (d total / d loss_out[0]) and g_reg (d total / d loss_out[1]),
 * optionally multiplied by DEVICE scalars g_scale_dev[2] (NULL = 1,1; lets an autograd engine pass
 * its upstream gradients without a host sync), and optionally a per-triple upstream gradient
 * g_energy[N] (NULL = none), ACCUMULATES (+=) into dcodes [V,d] and drel [Vrel,d] (the caller
 * zeroes them when it wants plain gradients). */
int distmult_backward(const float* codes, const float* rel, int32_t V, int32_t Vrel, int32_t d,
                      const int32_t* X, int64_t N, const float* Y, const float* energies,
                      float g_loss, float g_reg, const float* g_scale_dev, const float* g_energy,
                      float* dcodes, float* drel, void* stream);

/* Same as distmult_backward, additionally accumulating (+=) into the device float rel_slice_sumsq (may be NULL) the
 * sum over triples of |gradient slice of the gathered relation row|^2 -- the contribution of the relation table to
 * tf.clip_by_global_norm, which sees that gradient as un-aggregated IndexedSlices (bilinear_diag.py:18,
 * optimization/tensorflow_backend/algorithms.py:65-68). */
int distmult_backward_slices(const float* codes, const float* rel, int32_t V, int32_t Vrel, int32_t d,
                             const int32_t* X, int64_t N, const float* Y, const float* energies, float g_loss,
                             float g_reg, const float* g_scale_dev, const float* g_energy, float* dcodes,
                             float* drel, float* rel_slice_sumsq, void* stream);

/* IndexedSlices norm of the block tables' gradients: sumsq2[0] (W_forward) and sumsq2[1] (W_backward), overwritten,
 * receive  sum_messages norm_m^2 * sum_b |G[dst_m]_b|^2 |H[src_m]_b|^2  -- the squared norm of the per-edge gradient
 * slices tf.gradients hands to tf.clip_by_global_norm for variables read through tf.nn.embedding_lookup
 * (gcn_basis_concat.py:38-39).  H [V_src,d] layer input, G [V_dst,d] = dOut * relu'(out). */
int64_t rgcn_block_slice_sumsq_workspace_bytes(const rgcn_graph_t* g, int32_t d, int32_t B);
int rgcn_block_slice_sumsq(const rgcn_graph_t* g, int32_t d, int32_t B, const float* H, const float* G,
                           float* sumsq2, void* workspace, int64_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * DistMult all-entity scoring + ranking, fused (next row N3): decoders/bilinear_diag.py:51-61
 * (predict_all_subject_scores / predict_all_object_scores) feeding the rank counts of
 * common/evaluation.py:148-159 and :355-367.  For every triple t of X and the corrupted side,
 *   score(t, v)      = sigmoid( sum_k q[t,k] * codes[v,k] ),  q = rel[r]*codes[o] (side 0: subjects) or
 *                      codes[s]*rel[r] (side 1: objects)
 *   raw_rank[t]      = #{ v : score(t, v) >= score(t, gold_t) }               (the gold entity counts itself)
 *   filtered_rank[t] = raw_rank[t] - #{ v in known(t) : score(t, v) >= score(t, gold_t) } + 1
 * The [n, V] score matrix the reference materialises per 1000-triple chunk is never written: the energies come
 * out of the tcgen05 3xTF32 GEMM tile by tile and are compared in its epilogue.
 * known_mask : uint32 [n, ceil(V/32)] device, bit v of row t = v is a known true answer of t (the reference's
 *              known sets contain the evaluated triple itself, train.py:103-105), or NULL (then filtered_rank
 *              must be NULL).  raw_rank / filtered_rank : int32 [n] device.
 * workspace  : distmult_rank_workspace_bytes(V, d, n); its head holds the hi/lo split of `codes`:
 *              reuse_split != 0 skips re-splitting when the same workspace is passed again with unchanged codes.
 * ---------------------------------------------------------------------------------------------- */
int64_t distmult_rank_workspace_bytes(int32_t V, int32_t d, int64_t n);
int distmult_rank(const float* codes, const float* rel, int32_t V, int32_t Vrel, int32_t d, const int32_t* X,
                  int64_t n, int side, const uint32_t* known_mask, int reuse_split, int32_t* raw_rank,
                  int32_t* filtered_rank, void* workspace, int64_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RGCN_B200_H */
