"""GPU: the fused DistMult scorer/ranker (distmult_rank: scoring GEMM with a rank-counting epilogue) against the
counting rules of the reference's Scorer (common/evaluation.py:148-159, :355-367) restated in numpy."""
import numpy as np
import pytest
import torch

from relationprediction_b200 import ops
from relationprediction_b200.common import evaluation
from relationprediction_b200.decoders.bilinear_diag import BilinearDiag

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def reference_ranks(codes, rel, X, side, known_lists, sigmoid=True):
    """raw = #{score >= gold}, filtered = raw - #{known with score >= gold} + 1, float64 energies."""
    c, r = codes.astype(np.float64), rel.astype(np.float64)
    s, p, o = X[:, 0], X[:, 1], X[:, 2]
    q = r[p] * c[o] if side == 0 else c[s] * r[p]
    gold = s if side == 0 else o
    e = q @ c.T
    if sigmoid:
        e = (1.0 / (1.0 + np.exp(-e.astype(np.float32)))).astype(np.float32)
    g = e[np.arange(len(X)), gold]
    raw = (e >= g[:, None]).sum(1)
    kn = np.array([int((e[i, np.asarray(k, dtype=np.int64)] >= g[i]).sum()) if len(k) else 0 for i, k in enumerate(known_lists)])
    return raw, raw - kn + 1


def make_known(rng, X, V, side, include_gold=True):
    lists = []
    for t in X:
        k = set(rng.randint(0, V, rng.randint(0, 12)).tolist())
        if include_gold:
            k.add(int(t[0] if side == 0 else t[2]))
        lists.append(sorted(k))
    return lists


@pytest.mark.parametrize("V,d,n", [(1000, 64, 300), (4133, 500, 777), (129, 200, 5)])
def test_integer_codes_give_exact_ranks_with_ties(V, d, n):
    """Small integer codes: every energy is exact in fp32 (and in the 3xTF32 GEMM), so ties are exact ties and the
    ranks must equal the numpy restatement for every triple, raw and filtered, both sides."""
    rng = np.random.RandomState(0)
    codes = rng.randint(-1, 2, (V, d)).astype(np.float32) * (rng.uniform(size=(V, d)) < 0.05)
    rel = rng.randint(-1, 2, (V, d)).astype(np.float32)
    X = np.stack([rng.randint(0, V, n), rng.randint(0, V, n), rng.randint(0, V, n)], 1).astype(np.int32)
    ranker = ops.DistMultRanker(torch.as_tensor(codes, device=DEV), torch.as_tensor(rel, device=DEV))
    for side in (0, 1):
        known = make_known(rng, X, V, side)
        mask = torch.as_tensor(BilinearDiag.known_bit_mask(known, V), device=DEV)
        raw, filt = ranker.rank(torch.as_tensor(X, device=DEV), side, mask)
        ref_raw, ref_filt = reference_ranks(codes, rel, X, side, known, sigmoid=False)   # integer energies: order == sigmoid order
        np.testing.assert_array_equal(raw.cpu().numpy(), ref_raw)
        np.testing.assert_array_equal(filt.cpu().numpy(), ref_filt)
        raw2, none = ranker.rank(torch.as_tensor(X, device=DEV), side, None)            # raw only
        assert none is None
        np.testing.assert_array_equal(raw2.cpu().numpy(), ref_raw)


def test_float_codes_ranks_match_float64_up_to_near_ties():
    rng = np.random.RandomState(1)
    V, d, n = 14541, 500, 1000            # FB15k-237 sizes, one reference chunk
    codes = rng.normal(0, 0.3, (V, d)).astype(np.float32)
    rel = rng.normal(0, 1, (V, d)).astype(np.float32)
    X = np.stack([rng.randint(0, V, n), rng.randint(0, 237, n), rng.randint(0, V, n)], 1).astype(np.int32)
    ranker = ops.DistMultRanker(torch.as_tensor(codes, device=DEV), torch.as_tensor(rel, device=DEV))
    for side in (0, 1):
        known = make_known(rng, X, V, side)
        mask = torch.as_tensor(BilinearDiag.known_bit_mask(known, V), device=DEV)
        raw, filt = ranker.rank(torch.as_tensor(X, device=DEV), side, mask)
        ref_raw, ref_filt = reference_ranks(codes, rel, X, side, known)
        dr = np.abs(raw.cpu().numpy() - ref_raw)
        df = np.abs(filt.cpu().numpy() - ref_filt)
        # fp32 rounding can swap entities whose scores agree to ~1e-7 or that saturate together; nothing else may move
        assert (dr == 0).mean() > 0.97 and dr.max() <= max(3, 0.002 * V), (dr.mean(), dr.max())
        assert (df == 0).mean() > 0.97 and df.max() <= max(3, 0.002 * V)
        mrr = lambda r: float(np.mean(1.0 / r))
        assert abs(mrr(filt.cpu().numpy()) - mrr(ref_filt)) < 1e-4


def test_scorer_uses_the_fused_path_and_agrees_with_the_matrix_path(toy, tmp_path):
    """The whole chain: Scorer -> Model.rank_all_entities -> BilinearDiag.rank_all -> distmult_rank, against the
    score_all_subjects / score_all_objects path on the same trained-from-init Toy model."""
    from relationprediction_b200 import train as driver
    from test_gpu_train import TOY_EXP, write_toy
    write_toy(toy, tmp_path)
    exp = tmp_path / "toy.exp"
    exp.write_text(TOY_EXP.format(layers=2, concat="Yes"))
    np.random.seed(0)
    torch.manual_seed(0)
    model, scorer = driver.main(["--settings", str(exp), "--dataset", str(tmp_path), "--max-iterations", "30",
                                 "--no-periodic-eval", "--no-save"])
    test = np.array(toy["train"])
    assert model.supports_fused_ranking()
    fused = scorer.compute_scores(test).get_summary().results
    model.supports_fused_ranking = lambda: False
    matrix = scorer.compute_scores(test).get_summary().results
    for kind in ("Raw", "Filtered"):
        for k in ("MRR", "H@1", "H@3", "H@10"):
            assert abs(fused[kind][k] - matrix[kind][k]) < 2e-2, (kind, k, fused[kind][k], matrix[kind][k])
