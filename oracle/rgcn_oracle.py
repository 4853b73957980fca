"""CPU ORACLE for the R-GCN hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The product path (relationprediction_b200/) never does.

PARITY STATUS: **pinned through the reference's own code executed over a TF-op stand-in; NOT pinned by
TensorFlow itself.**  The reference's implementation of this path needs TensorFlow 1.4 (README.md:9;
`import tensorflow` in every hot-path file), which is not installable here (no network, no py3.12 wheel), and
the reference ships no tests, golden vectors or fixtures.  What pins this oracle:
  * tests/golden/reference_model_golden.npz -- outputs of the reference's unmodified model classes
    (model_builder, Representation, AffineTransform, ConcatGcn, BasisGcn, RelationEmbedding, BilinearDiag)
    run over tests/golden/tf1_shim.py (eager float64 restatement of the ~30 TF ops they call); loss,
    regularisation, every weight gradient and the test-mode scores agree with this oracle to 1e-10
    (tests/test_reference_golden.py).  Residual assumption: the shim's per-op TF semantics;
  * the integer goldens derived by hand from data/Toy and listed in SURVEY.md 8(c)
    (tests/golden/toy_golden.json, generated with the reference's own numpy-only loaders
    common/io.py and common/settings_reader.py, which DO import here);
  * algebraic known-answer tests (identity blocks => neighbourhood means, B=1 basis == plain GCN,
    block with B=1 == dense W_r, rows of the incidence sum to 1);
  * torch.autograd.gradcheck in float64 of every restated op.

Everything below restates the reference op for op (the TF op sequence of SURVEY.md section 2.3) in
torch-CPU; backward is torch.autograd over the restated forward, standing in for tf.gradients
(optimization/abstract.py:117-118).  Each function cites the reference file:line it follows
(paths relative to /root/reference/code).
"""
import numpy as np
import torch


# --------------------------------------------------------------------------------------------
# F1: index vectors                                   extras/graph_representations.py:21-27
# --------------------------------------------------------------------------------------------
def process_triples(triples):
    """triplets = transpose(X); sender = row 0, receiver = row 2, type = row 1 (:22-25)."""
    t = np.asarray(triples, dtype=np.int32).reshape(-1, 3)
    return t[:, 0].copy(), t[:, 1].copy(), t[:, 2].copy()


# --------------------------------------------------------------------------------------------
# F2: incidence normalisation           extras/graph_representations.py:84-93 and :124-133
# --------------------------------------------------------------------------------------------
def incidence_norm(rows, n_vertices, mode="canonical", norm_dtype=np.float32):
    """Values of tf.sparse_softmax(SparseTensor([rows, arange(E)], ones, [V,E])).

    canonical          : softmax over a row of all-ones = 1 / (#entries in that row)  -- what the
                         op is documented to compute and what the paper's 1/c_i means.
    tf_unsorted_compat : quirk Q1: the TF 1.x kernel computes on a copy sorted into
                         canonical (row, col) order and the Python wrapper re-attaches the values
                         to the ORIGINAL (unsorted) indices, so entry k receives
                         1 / count(row of the k-th entry in sorted order).
    none               : all ones (:70-82).
    """
    rows = np.asarray(rows, dtype=np.int64)
    counts = np.bincount(rows, minlength=n_vertices).astype(norm_dtype)   # float32 as TF; float64 for goldens
    if mode == "canonical":
        return (norm_dtype(1.0) / counts[rows]).astype(norm_dtype)
    if mode == "tf_unsorted_compat":
        order = np.argsort(rows, kind="stable")  # canonical order: by row, then by column (=k)
        return (norm_dtype(1.0) / counts[rows[order]]).astype(norm_dtype)
    if mode == "none":
        return np.ones(rows.shape[0], dtype=norm_dtype)
    raise ValueError(mode)


def graph_norms(triples, n_vertices, mode="canonical", norm_dtype=np.float32):
    """(norm_f[E], norm_b[E]): forward matrix rows = receivers (:85-87), backward rows = senders
    (:125-127); normalised per direction, not per relation ('global' branch)."""
    s, _, o = process_triples(triples)
    return incidence_norm(o, n_vertices, mode, norm_dtype), incidence_norm(s, n_vertices, mode, norm_dtype)


def messages_from_triples(triples, n_relations, n_vertices, mode="canonical"):
    """The 2E messages the layer consumes: forward message k: src=s_k, dst=o_k, weight id r_k;
    backward message E+k: src=o_k, dst=s_k, weight id r_k + R (separate W_backward table,
    gcn_basis_concat.py:38-39)."""
    s, r, o = process_triples(triples)
    nf, nb = graph_norms(triples, n_vertices, mode)
    dst = np.concatenate([o, s]).astype(np.int32)
    src = np.concatenate([s, o]).astype(np.int32)
    relw = np.concatenate([r, r + n_relations]).astype(np.int32)
    norm = np.concatenate([nf, nb]).astype(np.float32)
    return dst, src, relw, norm


def view_supertile_rows(base_rows, n_rows, M, n_relw, fixed=False):
    """Rows per supertile of one weight-id-major view (the library's view_supertile_rows, csrc/graph.h): the base
    size doubles, up to 32768 rows, while the view averages fewer than 48 messages per (supertile, weight id) item;
    `fixed` = the size was forced by $RGCN_SUPERTILE_ROWS."""
    rows = int(base_rows)
    if fixed or rows <= 0:
        return rows
    while rows < 32768:
        n_super = max(1, -(-int(n_rows) // rows))
        if M >= 48 * n_super * max(int(n_relw), 1):
            break
        rows *= 2
    return rows


def sorted_views(dst, src, relw, norm, V_dst, V_src, n_relw, supertile_rows=32768, adaptive=False):
    """Reference (numpy lexsort, stable) for the three sorted message views the library builds.
    Not in the TF reference (it uses COO matrices); this pins the bit-exact index work of the
    graph-prep step against an independent implementation.  adaptive: apply view_supertile_rows per view
    (what the library does unless the size is forced through the environment)."""
    base_rows = supertile_rows
    M = dst.shape[0]
    mid = np.arange(M, dtype=np.int32)
    out = {}
    p = np.lexsort((mid, relw, dst))
    out["dst_rowptr"] = np.concatenate([[0], np.cumsum(np.bincount(dst, minlength=V_dst))]).astype(np.int32)
    out["dst_src"], out["dst_relw"], out["dst_norm"], out["dst_mid"] = src[p], relw[p], norm[p], mid[p]
    p = np.lexsort((mid, relw, src))
    out["src_rowptr"] = np.concatenate([[0], np.cumsum(np.bincount(src, minlength=V_src))]).astype(np.int32)
    out["src_dst"], out["src_relw"], out["src_norm"], out["src_mid"] = dst[p], relw[p], norm[p], mid[p]
    # weight-id major views keyed (supertile(row), relw, row)
    for name, row, nbr, n_rows in (("rel", dst, src, V_dst), ("rel2", src, dst, V_src)):
        supertile_rows = view_supertile_rows(base_rows, n_rows, M, n_relw, fixed=not adaptive)
        n_super = max(1, -(-n_rows // supertile_rows))
        key = (row // supertile_rows).astype(np.int64) * n_relw + relw
        p = np.lexsort((mid, row, key))
        out[name + "_ptr"] = np.concatenate([[0], np.cumsum(np.bincount(key, minlength=n_super * n_relw))]).astype(np.int32)
        out[name + "_row"], out[name + "_nbr"], out[name + "_norm"], out[name + "_mid"] = row[p], nbr[p], norm[p], mid[p]
    return out


# --------------------------------------------------------------------------------------------
# Q2: initialisers                                              common/shared_functions.py:12-29
# --------------------------------------------------------------------------------------------
def glorot_variance(shape):
    return 3 / np.sqrt(shape[0] + shape[1])  # :12-13 -- used as a std-dev by make_tf_variable


def make_variable(rng, mean, std, shape):
    return rng.normal(mean, std, size=shape).astype(np.float32)  # :16-18 ("normal" init)


def init_block_layer(rng, n_relations, d, n_blocks):
    """ConcatGcn.local_initialize_train (gcn_basis_concat.py:17-28)."""
    s = int(d / n_blocks)
    std = glorot_variance([n_relations, s])  # :22  (vertex_matrix_shape[0], [2])
    return {
        "W_forward": make_variable(rng, 0, std, (n_relations, n_blocks, s, s)),
        "W_backward": make_variable(rng, 0, std, (n_relations, n_blocks, s, s)),
        "W_self": make_variable(rng, 0, std, (d, d)),
        "b": np.zeros(d, dtype=np.float32),  # created, listed, never added (quirk, 8a F6)
    }


def init_basis_layer(rng, n_relations, d, n_bases):
    """BasisGcn.local_initialize_train (gcn_basis.py:15-30)."""
    std = glorot_variance([d, d])  # :21
    return {
        "W_forward": make_variable(rng, 0, std, (d, n_bases, d)),
        "W_backward": make_variable(rng, 0, std, (d, n_bases, d)),
        "W_self": make_variable(rng, 0, std, (d, d)),
        "C_forward": make_variable(rng, 0, 1, (n_relations, n_bases)),   # :26-28
        "C_backward": make_variable(rng, 0, 1, (n_relations, n_bases)),
        "b": np.zeros(d, dtype=np.float32),
    }


# --------------------------------------------------------------------------------------------
# TF op restatements
# --------------------------------------------------------------------------------------------
def _t(x, dtype):
    if isinstance(x, torch.Tensor):
        return x.to(dtype)
    return torch.as_tensor(np.asarray(x)).to(dtype)


def sparse_dense_matmul(rows, values, dense, n_rows):
    """tf.sparse_tensor_dense_matmul(SparseTensor([rows, arange(E)], values, [V,E]), dense[E,d])
    = COO scatter-add of values[k]*dense[k] into row rows[k] (gcn_basis.py:78-79)."""
    out = torch.zeros(n_rows, dense.shape[1], dtype=dense.dtype)
    return out.index_add(0, rows, dense * values[:, None])


def dropout_with_mask(x, mask, keep):
    """tf.nn.dropout(x, keep): zero w.p. 1-keep, survivors scaled by 1/keep (message_gcn.py:64).
    The Bernoulli draw is an explicit input (TF's Philox stream cannot be reproduced)."""
    if mask is None:
        return x
    return x * mask.to(x.dtype) / keep


# --------------------------------------------------------------------------------------------
# F0: input embedding                                   encoders/affine_transform.py:63-82
# --------------------------------------------------------------------------------------------
def affine_onehot(W, b, use_bias=True, use_nonlinearity=True):
    h = W
    if use_bias:
        h = h + b
    if use_nonlinearity:
        h = torch.relu(h)
    return h


# --------------------------------------------------------------------------------------------
# F3-F6 block-diagonal layer         message_gcn.py:49-79 + gcn_basis_concat.py:35-83
# --------------------------------------------------------------------------------------------
def concat_gcn_forward(H, triples, W_forward, W_backward, W_self, norm_f, norm_b, drop_mask=None,
                       keep=1.0, use_nonlinearity=True, dtype=torch.float32):
    H = _t(H, dtype)
    Wf, Wb, Ws = _t(W_forward, dtype), _t(W_backward, dtype), _t(W_self, dtype)
    s_idx, r_idx, o_idx = (torch.as_tensor(a.astype(np.int64)) for a in process_triples(triples))
    V, d = H.shape
    n_coeff, sub = Wf.shape[1], Wf.shape[2]
    # message_gcn.py:39-40   embedding_lookup of sender / receiver rows
    sender_features = H[s_idx]
    receiver_features = H[o_idx]
    # gcn_basis_concat.py:38-39   per-edge weight gather [E,B,s,s]
    forward_transforms = Wf[r_idx]
    backward_transforms = Wb[r_idx]
    # :42-43 reshape, :46-47 batched matmul W . x, :50-51 reshape back
    rs = sender_features.reshape(-1, n_coeff, sub)
    rr = receiver_features.reshape(-1, n_coeff, sub)
    forward_messages = torch.matmul(forward_transforms, rs.unsqueeze(-1)).squeeze(-1).reshape(-1, d)
    backward_messages = torch.matmul(backward_transforms, rr.unsqueeze(-1)).squeeze(-1).reshape(-1, d)
    # :65-66 self loop; message_gcn.py:60-64 dropout on the self loop only, train mode only
    self_loop = dropout_with_mask(H @ Ws, None if drop_mask is None else _t(drop_mask, dtype), keep)
    # :70-74 two SpMMs: forward matrix rows = receivers, backward matrix rows = senders
    cf = sparse_dense_matmul(o_idx, _t(norm_f, dtype), forward_messages, V)
    cb = sparse_dense_matmul(s_idx, _t(norm_b, dtype), backward_messages, V)
    upd = cf + cb
    # :78-81  (the bias self.b is never added)
    return torch.relu(upd + self_loop) if use_nonlinearity else upd + self_loop


# --------------------------------------------------------------------------------------------
# F3-F6 basis layer                        message_gcn.py:49-79 + gcn_basis.py:39-88
# --------------------------------------------------------------------------------------------
def basis_gcn_forward(H, triples, W_forward, W_backward, C_forward, C_backward, W_self, norm_f,
                      norm_b, drop_mask=None, keep=1.0, use_nonlinearity=True, dtype=torch.float32):
    H = _t(H, dtype)
    Vf, Vb, Ws = _t(W_forward, dtype), _t(W_backward, dtype), _t(W_self, dtype)
    Cf, Cb = _t(C_forward, dtype), _t(C_backward, dtype)
    s_idx, r_idx, o_idx = (torch.as_tensor(a.astype(np.int64)) for a in process_triples(triples))
    V, d = H.shape
    B = Vf.shape[1]
    sender_features = H[s_idx]
    receiver_features = H[o_idx]
    # gcn_basis.py:48-52 coefficients
    forward_type_scaling = Cf[r_idx]
    backward_type_scaling = Cb[r_idx]
    # :60-68 dot_or_tensor_mul: [E,d] @ [d, B*d] -> [E,B,d]
    sender_terms = (sender_features @ Vf.reshape(Vf.shape[0], -1)).reshape(-1, B, Vf.shape[2])
    receiver_terms = (receiver_features @ Vb.reshape(Vb.shape[0], -1)).reshape(-1, B, Vb.shape[2])
    # :43-44
    forward_messages = (sender_terms * forward_type_scaling.unsqueeze(-1)).sum(1)
    backward_messages = (receiver_terms * backward_type_scaling.unsqueeze(-1)).sum(1)
    self_loop = dropout_with_mask(H @ Ws, None if drop_mask is None else _t(drop_mask, dtype), keep)
    cf = sparse_dense_matmul(o_idx, _t(norm_f, dtype), forward_messages, V)
    cb = sparse_dense_matmul(s_idx, _t(norm_b, dtype), backward_messages, V)
    upd = cf + cb
    return torch.relu(upd + self_loop) if use_nonlinearity else upd + self_loop


# --------------------------------------------------------------------------------------------
# D1/D2 DistMult                                          decoders/bilinear_diag.py:14-34, :63-69
# --------------------------------------------------------------------------------------------
def distmult_energies(codes, rel, X, dtype=torch.float32):
    codes, rel = _t(codes, dtype), _t(rel, dtype)
    X = torch.as_tensor(np.asarray(X).astype(np.int64)) if not isinstance(X, torch.Tensor) else X.long()
    e1s, rs, e2s = codes[X[:, 0]], rel[X[:, 1]], codes[X[:, 2]]  # :19-21
    return (e1s * rs * e2s).sum(1), (e1s, rs, e2s)                # :30


def weighted_cross_entropy_with_logits(targets, logits, pos_weight=1):
    """TF formula, stable form: (1-z)x + l*(log1p(exp(-|x|)) + max(-x,0)), l = 1+(q-1)z."""
    l = 1 + (pos_weight - 1) * targets
    return (1 - targets) * logits + l * (torch.log1p(torch.exp(-logits.abs())) + torch.relu(-logits))


def distmult_loss(codes, rel, X, Y, dtype=torch.float32):
    """Returns (loss, reg_unscaled, energies): loss = reduce_mean(weighted CE, pos_weight forced to
    1) (:32-34); reg = mean(e1^2)+mean(r^2)+mean(e2^2) over the gathered rows (:65-67)."""
    energies, (e1s, rs, e2s) = distmult_energies(codes, rel, X, dtype)
    Yt = _t(Y, dtype)
    loss = weighted_cross_entropy_with_logits(Yt, energies, 1).mean()
    reg = (e1s ** 2).mean() + (rs ** 2).mean() + (e2s ** 2).mean()
    return loss, reg, energies


def distmult_predict_all_objects(codes, rel, X, dtype=torch.float32):
    """predict_all_object_scores (:57-61): sigmoid((e1*r) @ codes^T)."""
    _, (e1s, rs, _) = distmult_energies(codes, rel, X, dtype)
    return torch.sigmoid((e1s * rs) @ _t(codes, dtype).T)


def distmult_predict_all_subjects(codes, rel, X, dtype=torch.float32):
    """predict_all_subject_scores (:51-55): sigmoid((codes @ (r*e2)^T)^T)."""
    _, (_, rs, e2s) = distmult_energies(codes, rel, X, dtype)
    return torch.sigmoid((_t(codes, dtype) @ (rs * e2s).T).T)


# --------------------------------------------------------------------------------------------
# End-to-end encoder/decoder as the reference wires it
# common/model_builder.py:121-184 (gcn_basis branch), :273-309 (layer stacking, last layer linear),
# train.py:262 (loss = CE + regularisation)
# --------------------------------------------------------------------------------------------
def encoder_forward(params, triples, n_vertices, n_relations, variant, mode="train", drop_masks=None,
                    keep=1.0, norm_mode="canonical", dtype=torch.float32, norm_dtype=np.float32):
    """params: {'W_in','b_in','layers':[{...}], 'W_relation'}; returns final codes [V,d]."""
    nf, nb = graph_norms(triples, n_vertices, norm_mode, norm_dtype)
    H = affine_onehot(_t(params["W_in"], dtype), _t(params["b_in"], dtype))  # model_builder.py:140-146
    n_layers = len(params["layers"])
    for li, lp in enumerate(params["layers"]):
        relu = li < n_layers - 1  # model_builder.py:275
        mask = None
        if mode == "train" and drop_masks is not None:
            mask = drop_masks[li]
        k = keep if mode == "train" else 1.0
        if variant == "block":
            H = concat_gcn_forward(H, triples, lp["W_forward"], lp["W_backward"], lp["W_self"], nf, nb,
                                   mask, k, relu, dtype)
        else:
            H = basis_gcn_forward(H, triples, lp["W_forward"], lp["W_backward"], lp["C_forward"],
                                  lp["C_backward"], lp["W_self"], nf, nb, mask, k, relu, dtype)
    return H


# --------------------------------------------------------------------------------------------
# Convenience: forward+backward of one layer with autograd (stands in for tf.gradients)
# --------------------------------------------------------------------------------------------
def layer_fwd_bwd(variant, H, triples, weights, norm_f, norm_b, dOut, drop_mask=None, keep=1.0,
                  use_nonlinearity=True, dtype=torch.float32):
    """Returns (out, grads dict) with grads for H and every weight in `weights`."""
    Ht = _t(H, dtype).clone().requires_grad_(True)
    wt = {k: _t(v, dtype).clone().requires_grad_(True) for k, v in weights.items() if k != "b"}
    if variant == "block":
        out = concat_gcn_forward(Ht, triples, wt["W_forward"], wt["W_backward"], wt["W_self"], norm_f,
                                 norm_b, drop_mask, keep, use_nonlinearity, dtype)
    else:
        out = basis_gcn_forward(Ht, triples, wt["W_forward"], wt["W_backward"], wt["C_forward"],
                                wt["C_backward"], wt["W_self"], norm_f, norm_b, drop_mask, keep,
                                use_nonlinearity, dtype)
    names = ["H"] + list(wt.keys())
    gs = torch.autograd.grad(out, [Ht] + list(wt.values()), grad_outputs=_t(dOut, dtype),
                             allow_unused=True)
    grads = {n: (g if g is not None else torch.zeros_like(t))
             for n, g, t in zip(names, gs, [Ht] + list(wt.values()))}
    return out.detach(), {k: v.detach() for k, v in grads.items()}


# --------------------------------------------------------------------------------------------
# N1: optimizer step           optimization/tensorflow_backend/algorithms.py:36-42 and :65-68
# (TensorFlow 1.x formulas restated: tf.clip_by_global_norm, tf.train.AdamOptimizer dense update)
# --------------------------------------------------------------------------------------------
def tf_clip_by_global_norm(grads, clip_norm):
    gn = np.sqrt(sum(float((np.asarray(g, dtype=np.float64) ** 2).sum()) for g in grads))
    scale = clip_norm * min(1.0 / gn if gn > 0 else np.inf, 1.0 / clip_norm)
    return [np.asarray(g, dtype=np.float64) * scale for g in grads], gn


def tf_adam_step(params, grads, ms, vs, t, lr=0.01, beta1=0.9, beta2=0.999, eps=1e-8):
    """In place on float64 numpy arrays; t is the 1-based step count."""
    lr_t = lr * np.sqrt(1 - beta2 ** t) / (1 - beta1 ** t)
    for p, g, m, v in zip(params, grads, ms, vs):
        m[...] = beta1 * m + (1 - beta1) * g
        v[...] = beta2 * v + (1 - beta2) * g * g
        p[...] = p - lr_t * m / (np.sqrt(v) + eps)
