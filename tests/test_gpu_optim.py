"""GPU: clip-by-global-norm + Adam kernels (csrc/optimizer.cu) against the float64 restatement of the
TensorFlow-1.x formulas the reference's optimizer stack executes."""
import numpy as np
import pytest
import torch

from oracle import rgcn_oracle as oracle
from relationprediction_b200.optim import ClippedAdam

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("max_norm", [1.0, None, 1e6])
def test_clipped_adam_matches_tf_formulas(max_norm):
    rng = np.random.RandomState(0)
    shapes = [(237, 100, 5, 5), (500, 500), (500,), (14541, 40)]
    ps = [rng.normal(size=s) for s in shapes]
    ms = [np.zeros(s) for s in shapes]
    vs = [np.zeros(s) for s in shapes]
    tp = [torch.tensor(p, dtype=torch.float32, device="cuda") for p in ps]
    opt = ClippedAdam(tp, lr=0.01, max_norm=max_norm)
    for t in range(1, 7):
        gs = [rng.normal(size=s) * (10.0 if t % 2 else 0.01) for s in shapes]
        for p, g in zip(tp, gs):
            p.grad = torch.tensor(g, dtype=torch.float32, device="cuda")
        if t == 4:
            tp[2].grad = None   # a weight without gradient this step (e.g. the unused bias) is skipped
            gs_used, idx = [g for i, g in enumerate(gs) if i != 2], [0, 1, 3]
        else:
            gs_used, idx = gs, [0, 1, 2, 3]
        cl = gs_used if max_norm is None else oracle.tf_clip_by_global_norm(gs_used, max_norm)[0]
        opt.step()
        oracle.tf_adam_step([ps[i] for i in idx], cl, [ms[i] for i in idx], [vs[i] for i in idx], opt.step_count)
        for p, ref in zip(tp, ps):
            err = float(np.abs(p.cpu().numpy() - ref).max() / np.abs(ref).max())
            assert err < 2e-6, (t, err)


def test_indexed_slices_norms_match_the_per_slice_definition():
    """tf.clip_by_global_norm over IndexedSlices: sum over edges / triples of the squared norm of the UN-aggregated
    gradient slice (block tables: norm_m * G[dst]_b (x) H[src]_b per block; relation table: per-triple row gradient).
    The backward passes park those sums on the parameters when ops.set_slice_norms(True); ClippedAdam then clips with
    them.  Compared with a float64 restatement that materialises every slice."""
    from relationprediction_b200 import ops
    from conftest import synthetic_kg
    rng = np.random.RandomState(0)
    V, R, E, d, B = 300, 7, 2500, 40, 8
    s = d // B
    tr = synthetic_kg(V, R, E, seed=4, skewed=True)
    g = ops.Graph(tr, V, R, device=0)
    H = torch.tensor(rng.normal(0, 1, (V, d)), dtype=torch.float32, device="cuda:0", requires_grad=True)
    Wf = torch.tensor(rng.normal(0, .3, (R, B, s, s)), dtype=torch.float32, device="cuda:0", requires_grad=True)
    Wb = torch.tensor(rng.normal(0, .3, (R, B, s, s)), dtype=torch.float32, device="cuda:0", requires_grad=True)
    Ws = torch.tensor(rng.normal(0, .1, (d, d)), dtype=torch.float32, device="cuda:0", requires_grad=True)
    dOut = torch.tensor(rng.normal(0, 1, (V, d)), dtype=torch.float32, device="cuda:0")
    ops.set_slice_norms(True)
    try:
        out = ops.block_layer(H, Wf, Wb, Ws, g, B, None, 1.0, True)
        out.backward(dOut)
        got_f, got_b = float(Wf._slice_sumsq), float(Wb._slice_sumsq)
        # float64 restatement: every message's slice, squared
        G = (dOut * (out > 0)).double().cpu().numpy().reshape(V, B, s)
        Hn = H.detach().double().cpu().numpy().reshape(V, B, s)
        so, oo = tr[:, 0], tr[:, 2]
        cf = np.bincount(oo, minlength=V).astype(np.float64)
        cb = np.bincount(so, minlength=V).astype(np.float64)
        sl_f = (1.0 / cf[oo])[:, None, None, None] * np.einsum("mbi,mbj->mbij", G[oo], Hn[so])
        sl_b = (1.0 / cb[so])[:, None, None, None] * np.einsum("mbi,mbj->mbij", G[so], Hn[oo])
        assert abs(got_f - (sl_f ** 2).sum()) <= 1e-4 * (sl_f ** 2).sum()
        assert abs(got_b - (sl_b ** 2).sum()) <= 1e-4 * (sl_b ** 2).sum()
        # the slices sum to the dense gradient the layer returns
        dWf = np.zeros((R, B, s, s))
        np.add.at(dWf, tr[:, 1], sl_f)
        assert np.abs(dWf - Wf.grad.double().cpu().numpy()).max() <= 1e-4 * np.abs(dWf).max()
        # DistMult relation table
        N = 900
        X = torch.tensor(np.stack([rng.randint(0, V, N), rng.randint(0, R, N), rng.randint(0, V, N)], 1), dtype=torch.int32,
                         device="cuda:0")
        Y = torch.tensor((rng.uniform(size=N) < 0.3).astype(np.float32), device="cuda:0")
        codes = torch.tensor(rng.normal(0, .5, (V, d)), dtype=torch.float32, device="cuda:0", requires_grad=True)
        rel = torch.tensor(rng.normal(0, 1, (V, d)), dtype=torch.float32, device="cuda:0", requires_grad=True)
        en, loss, reg = ops.distmult(codes, rel, X, Y)
        (loss + 0.01 * reg).backward()
        c, r = codes.detach().double().cpu().numpy(), rel.detach().double().cpu().numpy()
        xs = X.cpu().numpy()
        e = (c[xs[:, 0]] * r[xs[:, 1]] * c[xs[:, 2]]).sum(1)
        gx = (1 / (1 + np.exp(-e)) - Y.cpu().numpy()) / N
        sl = gx[:, None] * c[xs[:, 0]] * c[xs[:, 2]] + 0.01 * 2.0 / (N * d) * r[xs[:, 1]]
        ref = (sl ** 2).sum()
        assert abs(float(rel._slice_sumsq) - ref) <= 1e-4 * ref
        # and the optimizer consumes them
        opt = ClippedAdam([Wf, Wb, Ws, rel], lr=0.01, max_norm=1.0)
        before = Wf.detach().clone()
        opt.step()
        assert Wf._slice_sumsq is None and rel._slice_sumsq is None and not torch.equal(before, Wf.detach())
    finally:
        ops.set_slice_norms(False)
