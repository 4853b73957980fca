"""CPU: the WHOLE host side of the plugin chain (factory, Model protocol, placeholders, per-instance caches,
Representation, AffineTransform, ConcatGcn / BasisGcn, RelationEmbedding, BilinearDiag, scoring API) against
the goldens produced by the reference's own code (tests/golden/reference_model_golden.npz) -- with the three
library calls (`ops.Graph`, `ops.block_layer` / `ops.basis_layer`, `ops.distmult`) replaced by the CPU oracle.
This is a test of the HOST LOGIC only (the CUDA kernels are checked against the same goldens in
tests/test_gpu_reference_golden.py); the substitution lives in this test file, the product has no CPU path."""
import numpy as np
import pytest
import torch

from oracle import rgcn_oracle as oracle
from relationprediction_b200 import ops
from relationprediction_b200.common import evaluation, model_builder
from relationprediction_b200.encoders.message_gcns.message_gcn import MessageGcn
from test_plugin_host import merged_settings
from test_reference_golden import ALL, CASE_SETTINGS, load_case, split_weights

DT = torch.float64


class OracleGraph(object):
    """Stands in for ops.Graph: keeps the triples and the per-direction norms."""

    def __init__(self, triples, n_entities, n_relations, norm_mode="canonical", norm_f=None, norm_b=None,
                 device=None):
        self.triples = np.asarray(triples, dtype=np.int32).reshape(-1, 3)
        self.V_dst = self.V_src = int(n_entities)
        self.M = 2 * len(self.triples)
        if norm_mode == "explicit":
            self.nf, self.nb = np.asarray(norm_f), np.asarray(norm_b)
        else:
            self.nf, self.nb = oracle.graph_norms(self.triples, n_entities, norm_mode, np.float64)


def oracle_block_layer(H, Wf, Wb, Ws, graph, n_blocks, drop_mask=None, keep=1.0, use_nonlinearity=True):
    assert Wf.shape[1] == n_blocks
    return oracle.concat_gcn_forward(H, graph.triples, Wf, Wb, Ws, graph.nf, graph.nb, drop_mask, keep,
                                     use_nonlinearity, DT)


def oracle_basis_layer(H, Vf, Vb, Cf, Cb, Ws, graph, drop_mask=None, keep=1.0, use_nonlinearity=True):
    return oracle.basis_gcn_forward(H, graph.triples, Vf, Vb, Cf, Cb, Ws, graph.nf, graph.nb, drop_mask, keep,
                                    use_nonlinearity, DT)


def oracle_distmult(codes, rel, X, Y=None):
    if Y is None:
        e, (e1s, rs, e2s) = oracle.distmult_energies(codes, rel, X, DT)
        return e, torch.zeros((), dtype=DT), (e1s ** 2).mean() + (rs ** 2).mean() + (e2s ** 2).mean()
    loss, reg, e = oracle.distmult_loss(codes, rel, X, Y, DT)
    return e, loss, reg


@pytest.fixture
def oracle_backed_ops(monkeypatch):
    monkeypatch.setattr(ops, "Graph", OracleGraph)
    monkeypatch.setattr(ops, "block_layer", oracle_block_layer)
    monkeypatch.setattr(ops, "basis_layer", oracle_basis_layer)
    monkeypatch.setattr(ops, "distmult", oracle_distmult)
    monkeypatch.setattr(torch.cuda, "current_device", lambda: 0)


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-300))


@pytest.mark.parametrize("name,variant,grouping,norm_mode", ALL)
def test_host_chain_reproduces_reference_code_outputs(toy, oracle_backed_ops, name, variant, grouping, norm_mode):
    c = load_case(name + "_" + grouping)
    settings_file, overrides = CASE_SETTINGS[name]
    V, R = int(c["V"]), int(c["R"])
    enc, dec = merged_settings(toy, settings_file, V, R, len(c["test_graph"]))
    for s in (enc, dec):
        for k, v in overrides.items():
            s.put(k, v)
        s.put("NormalizationMode", norm_mode)
    model = model_builder.build_decoder(model_builder.build_encoder(enc, c["test_graph"]), dec)
    model.set_device("cpu")
    model.initialize_train()
    names, n_layers = split_weights(c, variant)
    ws = model.get_weights()
    assert len(ws) == len(names)
    for i, w in enumerate(ws):
        assert tuple(w.shape) == c["w%d" % i].shape, names[i]
        w.data = torch.tensor(c["w%d" % i], dtype=DT)
    layers, comp = [], model
    while comp is not None:
        if isinstance(comp, MessageGcn):
            layers.append(comp)
        comp = comp.next_component
    for layer, i in zip(layers[::-1], range(int(c["n_masks"]))):
        m = torch.tensor(c["mask%d" % i])
        layer.make_drop_mask = (lambda rows, mode, m=m, k=layer.dropout_keep_probability:
                                (m, k) if mode == 'train' else (None, 1.0))
    feed = (c["graph_split"], c["X"], c["Y"]) if model.needs_graph() else (c["X"], c["Y"])
    assert model.needs_graph() == (variant != "embedding")
    total = model.train_loss(*feed)
    total.backward()
    ref_total = float(c["loss"]) + float(c["reg"])
    # the tf_unsorted_compat norms travel through the product as float32 values (explicit norm arrays)
    tol = 1e-10 if norm_mode == "canonical" else 1e-6
    assert abs(total.item() - ref_total) <= tol * abs(ref_total)
    for i, (nm, w) in enumerate(zip(names, ws)):
        if bool(c["g%d_unused" % i]):
            assert w.grad is None or float(w.grad.abs().max()) == 0.0, nm
        else:
            assert rel(w.grad.numpy(), c["g%d" % i]) < tol, nm
    model.preprocess(c["test_graph"])
    model.register_for_test(c["test_graph"])
    for got, ref in ((model.score(c["test_X"]), c["predict"]),
                     (model.score_all_objects(c["test_X"]), c["all_objects"]),
                     (model.score_all_subjects(c["test_X"]), c["all_subjects"])):
        assert got.shape == ref.shape and np.abs(np.asarray(got, np.float64) - ref).max() < 100 * tol
    # ranking: this repository's Scorer over this repository's model == the reference's Scorer over the
    # reference's model (same weights): raw / filtered MRR and Hits@1/3/10 on test + 40 training triples
    sc = evaluation.Scorer({'Metric': 'MRR'})
    sc.register_data(c["test_graph"])
    sc.register_data(c["ranked"])
    sc.register_model(model)
    res = sc.compute_scores(c["ranked"]).get_summary().results
    got = np.array([[float(res[f][k]) for k in ('MRR', 'H@1', 'H@3', 'H@10')] for f in ('Raw', 'Filtered')])
    if norm_mode == "canonical":       # float64 scores identical to the golden's => identical ranks
        assert np.abs(got - c["ranking"]).max() < 1e-12
    else:                              # float32 norm values can reorder near-ties
        assert np.abs(got - c["ranking"]).max() < 5e-3
