"""CPU: the next-row host pieces against goldens produced by the reference's own importable numpy
modules (tests/golden/make_golden.py): evaluation ranking (common/evaluation.py) and the negative
sampler (common/auxilliaries.py)."""
import os

import numpy as np

from relationprediction_b200.common import auxilliaries, evaluation
from conftest import GOLDEN


def test_ranking_matches_reference_evaluation():
    g = np.load(os.path.join(GOLDEN, "eval_golden.npz"))
    T, train = g["T"], g["train"]

    class FakeModel:
        def score_all_subjects(self, tr):
            return np.stack([T[r, :, o] for s, r, o in tr])

        def score_all_objects(self, tr):
            return np.stack([T[r, s, :] for s, r, o in tr])
    sc = evaluation.Scorer()
    sc.register_data(train)
    sc.register_model(FakeModel())
    res = sc.compute_scores(train[:40]).get_summary().results
    np.testing.assert_allclose([res['Raw'][m] for m in ('MRR', 'H@1', 'H@3', 'H@10')], g["raw"], rtol=1e-12)
    np.testing.assert_allclose([res['Filtered'][m] for m in ('MRR', 'H@1', 'H@3', 'H@10')], g["filtered"], rtol=1e-12)


def test_negative_sampler_matches_reference_stream():
    g = np.load(os.path.join(GOLDEN, "negsample_golden.npz"))
    np.random.seed(123)   # same numpy global stream, same draw order (binomial, then randint)
    idx, lab = auxilliaries.NegativeSampler(3, 50).transform(g["batch"])
    np.testing.assert_array_equal(idx, g["idx"])
    np.testing.assert_array_equal(lab, g["labels"])
    assert idx.dtype == np.int32 and lab.dtype == np.float32
