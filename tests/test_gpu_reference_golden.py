"""GPU: the product path (plugin classes over the C-ABI, CUDA kernels) against golden vectors produced by
running the REFERENCE'S OWN model code (tests/golden/make_reference_golden.py; see tests/golden/tf1_shim.py
for how the reference was executed).  Same weights, same fed graph / batch, the reference's dropout masks
replayed; compared: train loss + regularisation, the gradient of every weight, test-mode scores.
Tolerance 1e-4 relative (fp32 kernels vs the float64 golden) -- north_star's floating-point tolerance."""
import numpy as np
import pytest
import torch

from relationprediction_b200.common import model_builder
from relationprediction_b200.encoders.message_gcns.message_gcn import MessageGcn
from test_plugin_host import merged_settings
from test_reference_golden import ALL, CASE_SETTINGS, load_case, split_weights

pytestmark = pytest.mark.gpu



def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-300))


def layers_of(model):
    out, c = [], model
    while c is not None:
        if isinstance(c, MessageGcn):
            out.append(c)
        c = c.next_component
    return out[::-1]          # input side first


@pytest.mark.parametrize("name,variant,grouping,norm_mode", ALL)
def test_product_matches_reference_code_outputs(toy, name, variant, grouping, norm_mode):
    c = load_case(name + "_" + grouping)
    settings_file, overrides = CASE_SETTINGS[name]
    V, R = int(c["V"]), int(c["R"])
    graph_split, X, Y = c["graph_split"], c["X"], c["Y"]
    enc, dec = merged_settings(toy, settings_file, V, R, len(c["test_graph"]))
    for s in (enc, dec):
        for k, v in overrides.items():
            s.put(k, v)
        s.put("NormalizationMode", norm_mode)
    model = model_builder.build_decoder(model_builder.build_encoder(enc, c["test_graph"]), dec)
    model.set_device("cuda:0")
    model.initialize_train()
    names, n_layers = split_weights(c, variant)
    ws = model.get_weights()
    assert len(ws) == len(names)
    with torch.no_grad():
        for i, w in enumerate(ws):
            g = torch.tensor(c["w%d" % i], dtype=torch.float32, device=w.device)
            assert tuple(w.shape) == tuple(g.shape), (names[i], tuple(w.shape), tuple(g.shape))
            w.copy_(g)
    # replay the reference's dropout masks (layer 1 was drawn first, then layer 2)
    masks = [torch.tensor(c["mask%d" % i], dtype=torch.uint8, device="cuda:0") for i in range(int(c["n_masks"]))]
    for layer, m in zip(layers_of(model), masks):
        assert abs(layer.dropout_keep_probability - 0.8) < 1e-12
        layer.make_drop_mask = (lambda rows, mode, m=m, k=layer.dropout_keep_probability:
                                (m, k) if mode == 'train' else (None, 1.0))
    total = model.train_loss(*((graph_split, X, Y) if model.needs_graph() else (X, Y)))
    total.backward()
    ref_total = float(c["loss"]) + float(c["reg"])
    assert abs(total.item() - ref_total) <= 1e-4 * abs(ref_total)
    for i, (nm, w) in enumerate(zip(names, ws)):
        if bool(c["g%d_unused" % i]):
            assert w.grad is None or float(w.grad.abs().max()) == 0.0, nm
            continue
        assert w.grad is not None, nm
        assert rel(w.grad.cpu().numpy(), c["g%d" % i]) < 1e-4, nm
    # test mode through the reference's scoring API (model.py:46-81), full training graph fed
    model.preprocess(c["test_graph"])
    model.register_for_test(c["test_graph"])
    tX = c["test_X"]
    # the reference returns sigmoid(energy): compare the PRE-sigmoid energies (logit of both sides) at the 1e-4
    # relative bar wherever the sigmoid is not saturated (1e-3 < score < 1 - 1e-3, where float32 scores still
    # resolve the logit), and every entry -- saturated ones included -- by a tight absolute error
    for got, ref in ((model.score(tX), c["predict"]), (model.score_all_objects(tX), c["all_objects"]),
                     (model.score_all_subjects(tX), c["all_subjects"])):
        got = np.asarray(got, np.float64)
        ref = np.asarray(ref, np.float64)
        assert got.shape == ref.shape
        assert np.abs(got - ref).max() < 2e-4
        live = (ref > 1e-3) & (ref < 1 - 1e-3) & (got > 0) & (got < 1)
        if live.any():   # tiny cases can be saturated throughout (5 triples of the Toy model)
            lg, lr = np.log(got[live] / (1 - got[live])), np.log(ref[live] / (1 - ref[live]))
            assert np.abs(lg - lr).max() / max(1.0, np.abs(lr).max()) < 1e-4
