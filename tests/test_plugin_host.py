"""CPU: host-side mirror of the reference plugin surface -- settings parsing equals the reference
reader's output (golden), dataset loaders round-trip the reference file formats, the Name= factory
builds the same component chain and weight order as common/model_builder.py."""
import os

import numpy as np
import pytest
import torch

from relationprediction_b200.common import io, model_builder, settings_reader
from relationprediction_b200.decoders.bilinear_diag import BilinearDiag
from relationprediction_b200.encoders.affine_transform import AffineTransform
from relationprediction_b200.encoders.message_gcns.gcn_basis import BasisGcn
from relationprediction_b200.encoders.message_gcns.gcn_basis_concat import ConcatGcn
from relationprediction_b200.encoders.relation_embedding import RelationEmbedding
from relationprediction_b200.extras.graph_representations import Representation


def test_settings_reader_matches_reference_parse(toy):
    for name, text in toy["settings_text"].items():
        got = settings_reader.read_string(text)
        assert got.to_dict() == toy["settings_parsed"][name], name
    s = settings_reader.read_string(toy["settings_text"]["gcn_block.exp"])
    assert s["Encoder"]["NumberOfBasisFunctions"] == "100"      # every value stays a string
    assert s["Optimizer"]["EarlyStopping"]["CheckEvery"] == "2000"
    assert "Concatenation" in s["Encoder"] and "Nope" not in s["Encoder"]


def test_dataset_loaders_roundtrip_reference_formats(toy, tmp_path):
    ent = {int(k): v for k, v in toy["entities"].items()}
    rel = {int(k): v for k, v in toy["relations"].items()}
    (tmp_path / "entities.dict").write_text("".join("%d\t%s\n" % kv for kv in sorted(ent.items())))
    (tmp_path / "relations.dict").write_text("".join("%d\t%s\n" % kv for kv in sorted(rel.items())))
    (tmp_path / "train.txt").write_text("".join("%s\t%s\t%s\n" % (ent[s], rel[r], ent[o]) for s, r, o in toy["train"]))
    e, r = str(tmp_path / "entities.dict"), str(tmp_path / "relations.dict")
    assert io.read_dictionary(e) == ent
    assert io.read_dictionary(r, id_lookup=False) == {v: k for k, v in rel.items()}
    assert io.read_triplets_as_list(str(tmp_path / "train.txt"), e, r) == toy["train"]
    arr = io.read_triplets_as_array(str(tmp_path / "train.txt"), e, r)
    assert arr.dtype == np.int32 and arr.shape == (43, 3)


def merged_settings(toy, name, V, R, E):
    """train.py:69-86: counts into General, Shared+General merged into Encoder/Decoder."""
    s = settings_reader.read_string(toy["settings_text"][name])
    general = s["General"]
    general.put("EntityCount", V)
    general.put("RelationCount", R)
    general.put("EdgeCount", E)
    enc, dec = s["Encoder"], s["Decoder"]
    for t in (enc, dec):
        t.merge(s["Shared"])
        t.merge(general)
    return enc, dec


def test_factory_builds_reference_chain_for_gcn_block(toy):
    V, R, E = 16, 9, 43
    enc, dec = merged_settings(toy, "gcn_block.exp", V, R, E)
    encoder = model_builder.build_encoder(enc, toy["train"])
    model = model_builder.build_decoder(encoder, dec)
    chain, c = [], model
    while c is not None:
        chain.append(type(c))
        c = c.next_component
    assert chain == [BilinearDiag, RelationEmbedding, ConcatGcn, ConcatGcn, AffineTransform, Representation]
    assert model.needs_graph()
    top, bottom = model.next_component.next_component, model.next_component.next_component.next_component
    assert top.use_nonlinearity is False and bottom.use_nonlinearity is True   # last layer linear
    assert top.n_coefficients == 100 and top.submatrix_d == 5 and top.dropout_keep_probability == 0.8
    assert model.regularization_parameter == 0.01
    model.set_device("cpu")
    np.random.seed(0)
    model.initialize_train()
    shapes = [tuple(w.shape) for w in model.get_weights()]
    d = 500
    layer = [(R, 100, 5, 5), (R, 100, 5, 5), (d, d), (d,)]
    assert shapes == [(V, d), (d,)] + layer + layer + [(V, d)]   # deepest component first; relation table [V,d]
    assert [p.name for p in model.get_train_input_variables()] == ["graph_edges", "X", "Y"]
    assert [p.name for p in model.get_test_input_variables()] == ["graph_edges", "X"]
    # initialisers: glorot_variance([R, s]) used as std (gcn_basis_concat.py:22)
    Wf = model.get_weights()[2]
    assert abs(float(Wf.detach().std()) - 3 / np.sqrt(R + 5)) < 0.05


def test_factory_gcn_basis_embedding_and_unknown(toy):
    enc, dec = merged_settings(toy, "gcn_basis.exp", 16, 9, 43)
    model = model_builder.build_decoder(model_builder.build_encoder(enc, toy["train"]), dec)
    layer = model.next_component.next_component
    assert isinstance(layer, BasisGcn) and layer.n_coefficients == 5
    model.set_device("cpu")
    model.initialize_train()
    names = [tuple(w.shape) for w in layer.local_get_weights()]
    assert names == [(500, 5, 500), (500, 5, 500), (9, 5), (9, 5), (500, 500), (500,)]
    enc, dec = merged_settings(toy, "distmult.exp", 16, 9, 43)
    emb = model_builder.build_encoder(enc, toy["train"])
    assert isinstance(emb, RelationEmbedding) and isinstance(emb.next_component, AffineTransform)
    assert not emb.needs_graph()
    enc.put("Name", "no-such-encoder")
    assert model_builder.build_encoder(enc, toy["train"]) is None      # reference behaviour (:270)
    dec.put("Name", "no-such-decoder")
    assert model_builder.build_decoder(emb, dec) is None               # (:320)
    enc, dec = merged_settings(toy, "gcn_block.exp", 16, 9, 43)
    enc.put("AddDiagonal", "Yes")
    with pytest.raises(NotImplementedError):
        model_builder.build_encoder(enc, toy["train"])


def test_no_cpu_fallback_in_plugin_layers(toy):
    enc, dec = merged_settings(toy, "gcn_block.exp", 16, 9, 43)
    model = model_builder.build_decoder(model_builder.build_encoder(enc, toy["train"]), dec)
    model.set_device("cpu")
    model.initialize_train()
    layer = model.next_component.next_component
    with pytest.raises(NotImplementedError):
        layer.compute_messages(None, None)
    if not torch.cuda.is_available():
        with pytest.raises(Exception):   # graph handle / kernels need a CUDA device: loud failure
            model.train_loss(np.array(toy["train"]), np.array(toy["train"]), np.ones(43, np.float32))


def test_chain_protocol_disciplines_with_dummy_components():
    """The three call disciplines of the plugin chain (code/model.py:148-182) on a GPU-free dummy chain."""
    from relationprediction_b200.model import Model
    S = {'EntityCount': 3, 'RelationCount': 2, 'EdgeCount': 1}
    log = []

    class Bottom(Model):
        def local_initialize_train(self):
            log.append('bottom-init')

        def local_get_weights(self):
            return ['w_bottom']

        def local_get_regularization(self):
            return 2.0

        def get_all_codes(self, mode='train'):
            return ('codes', mode)

        def get_loss(self, mode='train'):
            return 'loss-' + mode

        def needs_graph(self):
            return True

    class Middle(Model):          # defines nothing: everything passes through
        pass

    class Top(Model):
        def local_initialize_train(self):
            log.append('top-init')

        def local_get_weights(self):
            return ['w_top']

        def local_set_variable(self, name, value):
            log.append((name, value))

        def local_get_regularization(self):
            return 0.5
    chain = Top(Middle(Bottom(None, S), S), S)
    chain.initialize_train()
    assert log == ['top-init', 'bottom-init']                      # local first, then down the chain
    assert chain.get_weights() == ['w_bottom', 'w_top']            # deepest component first
    assert chain.get_regularization() == 2.5                       # base 0, summed along the chain
    assert chain.get_train_input_variables() == [] and chain.get_additional_ops() == []
    assert chain.get_all_codes() == ('codes', 'train')             # defaults of the defining component apply
    assert chain.get_all_codes(mode='test') == ('codes', 'test') and chain.get_all_codes('test') == ('codes', 'test')
    assert chain.get_loss(mode='test') == 'loss-test'
    assert chain.predict() is None and chain.get_graph() is None   # nobody defines them: None at the chain end
    assert chain.needs_graph() is True and Middle(None, S).needs_graph() is False
    chain.set_variable('x', 1)
    assert log[-1] == ('x', 1)
    assert chain.get_weights.__name__ == 'get_weights'


def test_real_chain_weight_and_feed_order_on_cpu(toy):
    """gcn_block.exp chain with its weights created on the CPU device: get_weights() order and the feed lists
    are what the reference produces (AffineTransform, layer 1, layer 2, RelationEmbedding; [graph, X, Y])."""
    V, R, E = 16, 9, 43
    enc, dec = merged_settings(toy, "gcn_block.exp", V, R, E)
    for s in (enc, dec):
        s.put("InternalEncoderDimension", "20")
        s.put("CodeDimension", "20")
        s.put("NumberOfBasisFunctions", "4")
    model = model_builder.build_decoder(model_builder.build_encoder(enc, toy["train"]), dec)
    model.set_device("cpu")
    model.initialize_train()
    shapes = [tuple(w.shape) for w in model.get_weights()]
    assert shapes == [(16, 20), (20,), (9, 4, 5, 5), (9, 4, 5, 5), (20, 20), (20,),
                      (9, 4, 5, 5), (9, 4, 5, 5), (20, 20), (20,), (16, 20)]
    assert [p.name for p in model.get_train_input_variables()] == ['graph_edges', 'X', 'Y']
    assert [p.name for p in model.get_test_input_variables()] == ['graph_edges', 'X']
    assert model.needs_graph()


def test_known_bit_mask_layout():
    """uint32 [n, ceil(V/32)] masks for the fused ranker: bit v of row t set iff v is in lists[t]."""
    from relationprediction_b200.decoders.bilinear_diag import BilinearDiag
    lists = [[0, 31, 32, 99], [], [64], [5, 5, 7]]
    m = BilinearDiag.known_bit_mask(lists, 100).view(np.uint32)
    assert m.shape == (4, 4) and m.dtype == np.uint32
    dense = np.zeros((4, 128), bool)
    for t, l in enumerate(lists):
        dense[t, l] = True
    got = ((m[:, :, None] >> np.arange(32, dtype=np.uint32)) & 1).astype(bool).reshape(4, 128)
    np.testing.assert_array_equal(got, dense)
