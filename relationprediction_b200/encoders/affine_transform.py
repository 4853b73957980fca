"""AffineTransform (reference: encoders/affine_transform.py:24-82).  With onehot_input=True it is the
free entity embedding relu(W + b) feeding the first R-GCN layer (model_builder.py:140-146)."""
import torch

from ..model import Model
from ..common.shared_functions import glorot_variance, make_variable, make_bias


class AffineTransform(Model):
    def __init__(self, shape, settings, next_component=None, use_nonlinearity=False, onehot_input=False,
                 use_bias=True):
        Model.__init__(self, next_component, settings)
        self.shape = shape
        self.use_nonlinearity = use_nonlinearity
        self.use_bias = use_bias
        self.onehot_input = onehot_input

    def local_initialize_train(self):
        dev = self.get_device()
        self.W = make_variable(0, glorot_variance(self.shape), self.shape, dev)
        self.b = make_bias(self.shape[1], dev)

    def local_get_weights(self):
        return [self.W, self.b]

    def _finish(self, hidden):
        if self.use_bias:
            hidden = hidden + self.b
        if self.use_nonlinearity:
            hidden = torch.relu(hidden)
        return hidden

    def _one_side(self, getter, mode):
        if self.onehot_input:      # one-hot input times W is W itself
            return self._finish(self.W)
        return self._finish(getattr(self.next_component, getter)(mode=mode) @ self.W)

    def get_all_subject_codes(self, mode='train'):
        return self._one_side('get_all_subject_codes', mode)

    def get_all_object_codes(self, mode='train'):
        return self._one_side('get_all_object_codes', mode)

    def get_all_codes(self, mode='train'):
        if self.onehot_input:
            h = self._finish(self.W)
            return h, None, h
        codes = self.next_component.get_all_codes(mode=mode)
        if codes[0] is codes[2]:   # one entity-code matrix (every R-GCN encoder): project it once, and keep it ONE
            h = self._finish(codes[0] @ self.W)   # tensor so the DistMult kernel can gather both ends from it
            return h, codes[1], h
        return self._finish(codes[0] @ self.W), codes[1], self._finish(codes[2] @ self.W)
