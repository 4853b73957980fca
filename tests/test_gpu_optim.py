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
