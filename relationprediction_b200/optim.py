"""Optimizer step of the reference's stack (GradientClipping + Adam, optimization/optimize.py:152-203 wiring,
tensorflow_backend/algorithms.py:36-42, :65-68) on the library's kernels (csrc/optimizer.cu): TensorFlow-1.x
clip_by_global_norm and AdamOptimizer formulas, no host synchronisation.

The clipping norm follows what tf.clip_by_global_norm actually sees in the reference: variables read through
tf.nn.embedding_lookup (block tables W_forward / W_backward, the decoder's relation table) carry IndexedSlices
gradients, whose norm is taken over the un-aggregated per-edge / per-triple slice VALUES.  When the backward passes
were run with ops.set_slice_norms(True) they leave that sum of squares on the parameter (`_slice_sumsq`) and it is
used here; parameters without it (dense gradients) contribute the sum of squares of their gradient.  The Adam update
itself is dense in both worlds (TF 1.4 sums duplicate indices before _apply_sparse and decays every row).
Not covered: the basis coefficients C_forward / C_backward (also embedding_lookup variables) still use the dense norm."""
import torch

from . import _lib
from .ops import _ptr, _stream


class ClippedAdam(object):
    def __init__(self, params, lr=0.01, beta1=0.9, beta2=0.999, eps=1e-8, max_norm=None):
        self.params = [p for p in params]
        self.lr, self.beta1, self.beta2, self.eps = float(lr), float(beta1), float(beta2), float(eps)
        self.max_norm = None if max_norm is None else float(max_norm)
        self.step_count = 0
        self.m = [torch.zeros_like(p) for p in self.params]
        self.v = [torch.zeros_like(p) for p in self.params]

    def zero_grad(self):
        for p in self.params:
            p.grad = None

    @torch.no_grad()
    def step(self):
        lib = _lib.load()
        live = [(p, m, v) for p, m, v in zip(self.params, self.m, self.v) if p.grad is not None]
        if not live:
            return
        self.step_count += 1
        dev = live[0][0].device
        st = _stream(dev)
        sumsq = None
        if self.max_norm is not None:
            sumsq = torch.zeros(1, dtype=torch.float32, device=dev)
            for p, _, _ in live:
                ss = getattr(p, "_slice_sumsq", None)
                if ss is not None:   # IndexedSlices in the reference: norm over the un-aggregated slices
                    sumsq += ss.to(dev)
                    p._slice_sumsq = None
                    continue
                g = p.grad.contiguous()
                _lib.check(lib.rgcn_sumsq_accumulate(_ptr(g), g.numel(), _ptr(sumsq), st), "rgcn_sumsq_accumulate")
        for p, m, v in live:
            if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous()):
                raise _lib.RgcnError("ClippedAdam: parameters must be contiguous CUDA float32 tensors")
            g = p.grad.contiguous()
            rc = lib.rgcn_adam_update(_ptr(p), _ptr(g), _ptr(m), _ptr(v), p.numel(), self.lr, self.beta1, self.beta2,
                                      self.eps, self.step_count, _ptr(sumsq),
                                      self.max_norm if self.max_norm is not None else 0.0, st)
            _lib.check(rc, "rgcn_adam_update")
