"""Numeric helpers with the reference's initialisation semantics (common/shared_functions.py:5-29)."""
import numpy as np
import torch


def glorot_variance(shape):
    """3 / sqrt(fan_in + fan_out); the reference passes it to np.random.normal as the STD-DEV (:12-18)."""
    return 3 / np.sqrt(shape[0] + shape[1])


def make_variable(mean, std, shape, device, init="normal"):
    """make_tf_variable (:16-22): numpy-drawn initial value (np.random global stream, like the
    reference) turned into a trainable device tensor."""
    if init == "normal":
        value = np.random.normal(mean, std, size=shape).astype(np.float32)
    elif init == "uniform":
        value = np.random.uniform(mean, std, size=shape).astype(np.float32)
    else:
        raise ValueError(init)
    return torch.tensor(value, device=device, requires_grad=True)


def make_bias(shape, device, init=0):
    """make_tf_bias (:25-29)."""
    value = np.zeros(shape, dtype=np.float32) if init == 0 else np.ones(shape, dtype=np.float32)
    return torch.tensor(value, device=device, requires_grad=True)


def dot_or_lookup(features, weights, onehot_input=False):
    """:5-9 -- row lookup for one-hot inputs, dense matmul otherwise."""
    if onehot_input:
        return weights[features.long()]
    return features @ weights
