"""GPU: the plugin chain built by the Name= factory from the reference's .exp settings, end to end
(embedding -> R-GCN layers -> DistMult loss + regularisation, and its gradients through every weight)
against the oracle's restatement of the same chain with the same weights."""
import numpy as np
import pytest
import torch

from oracle import rgcn_oracle as oracle
from relationprediction_b200.common import model_builder
from conftest import synthetic_kg
from test_plugin_host import merged_settings

pytestmark = pytest.mark.gpu


def relerr(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


def build(toy, name, V, R, E, triples, overrides):
    enc, dec = merged_settings(toy, name, V, R, E)
    for k, v in overrides.items():
        enc.put(k, v)
        dec.put(k, v)
    model = model_builder.build_decoder(model_builder.build_encoder(enc, triples), dec)
    model.set_device("cuda:0")
    np.random.seed(0)
    model.initialize_train()
    return model


def oracle_params(model, variant, n_layers):
    ws = [w.detach().cpu().numpy() for w in model.get_weights()]
    per = 4 if variant == "block" else 6
    params = {"W_in": ws[0], "b_in": ws[1], "layers": [], "W_relation": ws[2 + per * n_layers]}
    for l in range(n_layers):
        c = ws[2 + per * l: 2 + per * (l + 1)]
        if variant == "block":
            params["layers"].append({"W_forward": c[0], "W_backward": c[1], "W_self": c[2]})
        else:
            params["layers"].append({"W_forward": c[0], "W_backward": c[1], "C_forward": c[2],
                                     "C_backward": c[3], "W_self": c[4]})
    return params


def oracle_loss_and_grads(params, variant, triples, V, R, X, Y, lam):
    leaves = {}

    def leaf(name, v):
        t = torch.tensor(v, dtype=torch.float64, requires_grad=True)
        leaves[name] = t
        return t
    p = {"W_in": leaf("W_in", params["W_in"]), "b_in": leaf("b_in", params["b_in"]), "layers": []}
    for i, lp in enumerate(params["layers"]):
        p["layers"].append({k: leaf("L%d.%s" % (i, k), v) for k, v in lp.items()})
    Wr = leaf("W_relation", params["W_relation"])
    codes = oracle.encoder_forward(p, triples, V, R, variant, mode="test", dtype=torch.float64)
    loss, reg, _ = oracle.distmult_loss(codes, Wr, X, Y, torch.float64)
    total = loss + lam * reg
    total.backward()
    return total.item(), {k: v.grad.numpy() for k, v in leaves.items()}, codes.detach().numpy()


@pytest.mark.parametrize("name,variant,over", [
    ("gcn_block.exp", "block", {"InternalEncoderDimension": "40", "CodeDimension": "40",
                                "NumberOfBasisFunctions": "8", "DropoutKeepProbability": "1.0"}),
    ("gcn_basis.exp", "basis", {"InternalEncoderDimension": "24", "CodeDimension": "24",
                                "NumberOfBasisFunctions": "2", "DropoutKeepProbability": "1.0",
                                "NumberOfLayers": "1"}),   # BASELINE configs[0]: Toy 1-layer basis + DistMult
])
def test_toy_model_loss_and_all_gradients(toy, name, variant, over):
    V, R = toy["V"], toy["R"]
    tr = np.array(toy["train"], dtype=np.int32)
    model = build(toy, name, V, R, len(tr), tr, over)
    n_layers = int(over.get("NumberOfLayers", 2))
    rng = np.random.RandomState(1)
    N = 3 * len(tr)
    X = np.tile(tr, (3, 1)).astype(np.int32)
    X[len(tr):, 2] = rng.randint(0, V, N - len(tr))
    Y = np.zeros(N, np.float32)
    Y[:len(tr)] = 1
    loss = model.train_loss(tr, X, Y)
    loss.backward()
    ws = model.get_weights()
    ref_loss, ref_g, ref_codes = oracle_loss_and_grads(oracle_params(model, variant, n_layers), variant, tr,
                                                       V, R, X, Y, 0.01)
    assert abs(loss.item() - ref_loss) / abs(ref_loss) < 1e-4
    per = 4 if variant == "block" else 6
    order = ["W_in", "b_in"]
    for l in range(n_layers):
        order += ["L%d.%s" % (l, k) for k in (["W_forward", "W_backward", "W_self", None] if variant == "block"
                                               else ["W_forward", "W_backward", "C_forward", "C_backward", "W_self", None])]
    order += ["W_relation"]
    assert len(order) == len(ws)
    for nm, w in zip(order, ws):
        if nm is None or nm.endswith(".None"):   # the unused bias: no gradient flows (reference: created, never added)
            assert w.grad is None or float(w.grad.abs().max()) == 0.0
            continue
        assert w.grad is not None, "no gradient reached " + nm
        assert relerr(w.grad.cpu().numpy(), ref_g[nm]) < 1e-4, nm
    # scoring API: score / score_all_subjects / score_all_objects (model.py:46-81)
    model.preprocess(tr)
    model.register_for_test(tr)
    test = np.array(toy["test"], dtype=np.int32)
    sc = model.score(test)
    ref_e, _ = oracle.distmult_energies(ref_codes, model.get_weights()[-1].detach().cpu().numpy(), test, torch.float64)
    assert relerr(sc, torch.sigmoid(ref_e).numpy()) < 1e-4
    so = model.score_all_objects(test)
    ss = model.score_all_subjects(test)
    Wr = model.get_weights()[-1].detach().cpu().numpy()
    assert so.shape == ss.shape == (len(test), V)
    assert relerr(so, oracle.distmult_predict_all_objects(ref_codes, Wr, test, torch.float64).numpy()) < 1e-4
    assert relerr(ss, oracle.distmult_predict_all_subjects(ref_codes, Wr, test, torch.float64).numpy()) < 1e-4
    # the gold triple's own score appears in the all-entity rows
    np.testing.assert_allclose(so[np.arange(len(test)), test[:, 2]], sc, rtol=1e-4)


def test_gcn_block_exp_unchanged_on_fb15k237_shaped_graph(toy):
    """settings/gcn_block.exp exactly as shipped (d=500, B=100, 2 layers, dropout 0.8) on a synthetic KG;
    test mode (dropout off) codes against the oracle, train mode statistics of the dropout."""
    V, R, E = 2000, 237, 15000
    tr = synthetic_kg(V, R, E, seed=2, skewed=True)
    model = build(toy, "gcn_block.exp", V, R, E, tr, {})
    model.preprocess(tr)
    model.register_for_test(tr)
    X = tr[:64]
    sc = model.score(X)
    params = oracle_params(model, "block", 2)
    codes = oracle.encoder_forward(params, tr, V, R, "block", mode="test", dtype=torch.float64)
    e, _ = oracle.distmult_energies(codes, params["W_relation"], X, torch.float64)
    assert relerr(sc, torch.sigmoid(e).numpy()) < 1e-4
    # train mode: self-loop dropout keeps 80 % and rescales by 1/0.8 -> E[loss] finite, grads flow
    Y = np.ones(64, np.float32)
    l1 = model.train_loss(tr, X, Y)
    l2 = model.train_loss(tr, X, Y)
    assert torch.isfinite(l1) and torch.isfinite(l2) and l1.item() != l2.item()   # fresh mask each step
    l2.backward()
    assert all(torch.isfinite(w.grad).all() for w in model.get_weights() if w.grad is not None)
