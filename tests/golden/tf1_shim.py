"""A minimal EAGER stand-in for the ~30 TensorFlow-1.x ops the reference's hot path calls, so that the
reference's OWN Python (model_builder, Representation/MessageGraph, AffineTransform, ConcatGcn, BasisGcn,
RelationEmbedding, BilinearDiag -- imported unmodified from /root/reference/code) can be executed here to
produce golden vectors (tests/golden/make_reference_golden.py).  TensorFlow 1.4 itself is not installable in
this environment; what this file restates is therefore the documented semantics of each TF op, NOT the
reference's algorithm -- the composition of the ops (which rows are gathered, which transposes, the order of
messages, the class-level caches, the loss/regularisation formulae) is the reference's code running as is.

Values are torch float64 CPU tensors (int64 for integer tensors); tf.gradients is torch.autograd.  Graph
construction is eager, so placeholders must be fed (Placeholder.feed) before the first op that reads them.

Semantics restated per op (TF 1.4 API docs):
  transpose (reverse dims), stack, range, shape, reshape (-1 allowed), squeeze (all size-1 dims), expand_dims,
  to_float / to_int64 / ones_like, square, reduce_sum(axis) / reduce_mean (all elements), matmul (batched over
  leading dims), nn.embedding_lookup (params[ids]), nn.relu, nn.sigmoid,
  nn.dropout(x, keep): x / keep * floor(keep + U[0,1))      -- the uniform draws come from `dropout_rng`, and
      every mask is appended to `dropout_masks` so the same mask can be replayed on the CUDA path,
  nn.weighted_cross_entropy_with_logits(targets, logits, pos_weight):
      (1 - z) * x + (1 + (q - 1) * z) * (log1p(exp(-|x|)) + max(-x, 0)),
  SparseTensor / sparse_tensor_dense_matmul (out[i] += v * dense[j]; any index order),
  sparse_softmax: softmax over the last dimension within each group of entries sharing the leading indices.
      Two readings, selected by SPARSE_SOFTMAX_GROUPING (SURVEY.md quirk Q1 -- the reference feeds indices
      that are NOT in canonical row-major order, which the op's documentation requires):
        "canonical" : entry k gets softmax within the entries that share its leading indices -- what the op
                      is documented to compute (and computes on sorted input); the paper's 1/c_i.
        "tf_kernel" : our reading of the TF 1.x kernel (sparse_softmax_op.cc): it deep-copies the input,
                      Reorder()s the COPY into canonical order, writes the per-group softmax values in that
                      sorted order, and the Python wrapper re-attaches them to the ORIGINAL indices -- so
                      entry k receives the value computed for the k-th entry of the sorted copy.  Not
                      verifiable here (no TensorFlow); kept so that both behaviours have goldens.
"""
import builtins
import sys
import types

import numpy as np
import torch

SPARSE_SOFTMAX_GROUPING = "canonical"
dropout_rng = np.random.RandomState(0)
dropout_masks = []

int32, int64, float32 = "int32", "int64", "float32"


def _raw(x):
    if isinstance(x, T):
        if x.t is None:
            raise RuntimeError("placeholder read before it was fed")
        return x.t
    if isinstance(x, torch.Tensor):
        return x
    if isinstance(x, np.ndarray):
        return torch.from_numpy(x.astype(np.float64) if x.dtype.kind == "f" else x.astype(np.int64))
    if isinstance(x, (list, tuple)):
        return torch.stack([_raw(v) for v in x])
    if isinstance(x, (int, np.integer)):
        return torch.tensor(int(x), dtype=torch.int64)
    return torch.tensor(float(x), dtype=torch.float64)


class T(object):
    """Eager tensor handle with the operator overloads the reference uses."""

    def __init__(self, t=None):
        self.t = t

    def feed(self, value):
        self.t = _raw(np.asarray(value))

    def __getitem__(self, idx):
        return T(_raw(self)[idx])

    def __add__(self, o):
        return T(_raw(self) + _raw(o))

    __radd__ = __add__

    def __iadd__(self, o):
        return T(_raw(self) + _raw(o))

    def __sub__(self, o):
        return T(_raw(self) - _raw(o))

    def __mul__(self, o):
        return T(_raw(self) * _raw(o))

    __rmul__ = __mul__

    def __neg__(self):
        return T(-_raw(self))

    def numpy(self):
        return _raw(self).detach().numpy()


class SparseTensor(object):
    def __init__(self, indices, values, dense_shape):
        self.indices, self.values, self.dense_shape = _raw(indices), _raw(values), _raw(dense_shape)


def placeholder(dtype, shape=None, name=None):
    return T(None)


def Variable(initial_value, *a, **k):
    v = _raw(np.asarray(initial_value)).clone()
    if v.dtype == torch.float64:
        v.requires_grad_(True)
    return T(v)


def transpose(x):
    r = _raw(x)
    return T(r.permute(*reversed(builtins.range(r.dim()))))


def shape(x):
    return [int(s) for s in _raw(x).shape]


def stack(values):
    return T(_raw(list(values)))


def range(n):  # noqa: A001 (mirrors tf.range)
    return T(torch.arange(int(_raw(n)) if not isinstance(n, int) else n, dtype=torch.int64))


def to_float(x):
    return T(_raw(x).to(torch.float64))


def to_int64(x):
    return T(_raw(x).to(torch.int64))


def ones_like(x):
    return T(torch.ones_like(_raw(x)))


def reshape(x, shp):
    return T(_raw(x).reshape([int(s) for s in shp]))


def squeeze(x):
    return T(_raw(x).squeeze())


def expand_dims(x, axis):
    return T(_raw(x).unsqueeze(axis))


def square(x):
    return T(_raw(x) ** 2)


def reduce_sum(x, axis=None):
    return T(_raw(x).sum() if axis is None else _raw(x).sum(dim=axis))


def reduce_mean(x, axis=None):
    return T(_raw(x).mean() if axis is None else _raw(x).mean(dim=axis))


def matmul(a, b):
    return T(torch.matmul(_raw(a), _raw(b)))


def sparse_softmax(sp):
    idx = sp.indices.numpy()
    vals = sp.values.numpy().astype(np.float64)
    lead = idx[:, :-1]
    out = np.empty_like(vals)
    groups = {}
    for i, key in enumerate(map(tuple, lead)):
        groups.setdefault(key, []).append(i)
    for members in groups.values():
        e = np.exp(vals[members] - vals[members].max())
        out[members] = e / e.sum()
    if SPARSE_SOFTMAX_GROUPING == "tf_kernel":
        order = np.lexsort(idx.T[::-1])          # canonical (row-major) order of the copy, stable
        out = out[order]                          # values stay in sorted order, indices stay original
    elif SPARSE_SOFTMAX_GROUPING != "canonical":
        raise ValueError(SPARSE_SOFTMAX_GROUPING)
    return SparseTensor(sp.indices, torch.from_numpy(out), sp.dense_shape)


def sparse_tensor_dense_matmul(sp, dense):
    d = _raw(dense)
    rows, cols = sp.indices[:, 0], sp.indices[:, 1]
    out = torch.zeros(int(sp.dense_shape[0]), d.shape[1], dtype=torch.float64)
    return T(out.index_add(0, rows, sp.values[:, None] * d[cols]))


def sparse_reduce_sum_sparse(*a, **k):
    raise NotImplementedError("'local' normalisation is not on the accelerated path")


def _embedding_lookup(params, ids):
    return T(_raw(params)[_raw(ids).long()])


def _relu(x):
    return T(torch.relu(_raw(x)))


def _sigmoid(x):
    return T(torch.sigmoid(_raw(x)))


def _dropout(x, keep_prob):
    r = _raw(x)
    u = dropout_rng.uniform(size=tuple(r.shape))
    mask = np.floor(keep_prob + u)
    dropout_masks.append(mask.astype(np.uint8))
    return T(r / keep_prob * torch.from_numpy(mask))


def _weighted_cross_entropy_with_logits(targets, logits, pos_weight):
    z, x, q = _raw(targets).to(torch.float64), _raw(logits), float(pos_weight)
    l = 1 + (q - 1) * z
    return T((1 - z) * x + l * (torch.log1p(torch.exp(-x.abs())) + torch.clamp(-x, min=0)))


def gradients(ys, xs):
    g = torch.autograd.grad(_raw(ys), [_raw(x) for x in xs], allow_unused=True)
    return [None if gi is None else T(gi) for gi in g]


def install():
    """Register this module as `tensorflow` (only inside the golden-generation process)."""
    me = sys.modules[__name__]
    nn = types.ModuleType("tensorflow.nn")
    nn.embedding_lookup = _embedding_lookup
    nn.relu = _relu
    nn.sigmoid = _sigmoid
    nn.dropout = _dropout
    nn.weighted_cross_entropy_with_logits = _weighted_cross_entropy_with_logits
    me.nn = nn
    sys.modules["tensorflow"] = me
    sys.modules["tensorflow.nn"] = nn
    return me
