"""CPU: pins the oracle against the committed goldens and algebraic known-answer tests
(SURVEY.md 8c).  The reference ships no tests; these are the vectors that stand in."""
import numpy as np
import torch

from oracle import rgcn_oracle as oracle


def test_toy_integer_goldens(toy):
    tr = np.array(toy["train"])
    assert tr.shape == (43, 3) and toy["V"] == 16 and toy["R"] == 9
    assert tr[:3].tolist() == [[10, 0, 3], [10, 7, 12], [10, 8, 12]]
    assert tr[-2:].tolist() == [[11, 0, 14], [11, 7, 6]]
    assert toy["valid"] == toy["test"] == [[12, 8, 9], [4, 3, 3], [9, 7, 6], [13, 2, 9], [6, 0, 14]]
    s, r, o = oracle.process_triples(tr)
    assert np.bincount(o, minlength=16).tolist() == [1, 1, 4, 7, 1, 1, 7, 2, 4, 4, 0, 0, 3, 3, 1, 4]
    assert np.bincount(s, minlength=16).tolist() == [6, 2, 0, 0, 5, 4, 6, 0, 0, 8, 3, 2, 6, 0, 0, 1]


def test_toy_norm_goldens(toy):
    tr = np.array(toy["train"])
    nf, nb = oracle.graph_norms(tr, 16, "canonical")
    np.testing.assert_array_equal(nf[:8], np.float32([1 / 7, 1 / 3, 1 / 3, 1 / 7, 1, 1 / 3, 1 / 4, 1 / 7]))
    np.testing.assert_array_equal(nf, np.float32(toy["norm_f_canonical"]))
    np.testing.assert_array_equal(nb, np.float32(toy["norm_b_canonical"]))
    cf, cb = oracle.graph_norms(tr, 16, "tf_unsorted_compat")
    np.testing.assert_array_equal(cf[:8], np.float32([1, 1, 1 / 4, 1 / 4, 1 / 4, 1 / 4, 1 / 7, 1 / 7]))
    np.testing.assert_array_equal(cf, np.float32(toy["norm_f_tf_unsorted_compat"]))
    np.testing.assert_array_equal(cb, np.float32(toy["norm_b_tf_unsorted_compat"]))
    assert int((nf != cf).sum()) == 35  # 35 of 43 forward norms differ between readings


def test_incidence_rows_sum_to_one(toy):
    tr = np.array(toy["train"])
    s, r, o = oracle.process_triples(tr)
    nf, nb = oracle.graph_norms(tr, 16)
    rows_f = np.bincount(o, weights=nf, minlength=16)
    rows_b = np.bincount(s, weights=nb, minlength=16)
    zero_f, zero_b = [10, 11], [2, 3, 7, 8, 13, 14]
    for v in range(16):
        assert abs(rows_f[v] - (0 if v in zero_f else 1)) < 1e-6
        assert abs(rows_b[v] - (0 if v in zero_b else 1)) < 1e-6


def test_identity_blocks_give_neighbourhood_means(toy):
    tr = np.array(toy["train"])
    V, R, d, B = 16, 9, 8, 2
    rng = np.random.RandomState(1)
    H = rng.normal(size=(V, d))
    eye = np.tile(np.eye(d // B)[None, None], (R, B, 1, 1))
    nf, nb = oracle.graph_norms(tr, V)
    out = oracle.concat_gcn_forward(H, tr, eye, eye, np.zeros((d, d)), nf, nb, None, 1.0, False,
                                    torch.float64).numpy()
    exp = np.zeros_like(H)
    for v in range(V):
        fin = [s for s, r, o in tr if o == v]
        bin_ = [o for s, r, o in tr if s == v]
        if fin:
            exp[v] += H[fin].mean(0)
        if bin_:
            exp[v] += H[bin_].mean(0)
    np.testing.assert_allclose(out, exp, rtol=1e-6, atol=1e-6)


def test_block_B1_equals_dense_relation_matrices_and_basis_B1_equals_plain_gcn(toy):
    tr = np.array(toy["train"])
    V, R, d = 16, 9, 8
    rng = np.random.RandomState(2)
    H = rng.normal(size=(V, d))
    Wf, Wb = rng.normal(size=(R, 1, d, d)), rng.normal(size=(R, 1, d, d))
    Ws = rng.normal(size=(d, d))
    nf, nb = oracle.graph_norms(tr, V)
    out = oracle.concat_gcn_forward(H, tr, Wf, Wb, Ws, nf, nb, None, 1.0, True, torch.float64).numpy()
    exp = H @ Ws
    for k, (s, r, o) in enumerate(tr):
        exp[o] += nf[k] * (Wf[r, 0] @ H[s])
        exp[s] += nb[k] * (Wb[r, 0] @ H[o])
    np.testing.assert_allclose(out, np.maximum(exp, 0), rtol=1e-6, atol=1e-6)
    # basis, B = 1, C == 1: a plain (relation-blind) GCN with one matrix per direction
    Vf, Vb = rng.normal(size=(d, 1, d)), rng.normal(size=(d, 1, d))
    ones = np.ones((R, 1))
    out = oracle.basis_gcn_forward(H, tr, Vf, Vb, ones, ones, Ws, nf, nb, None, 1.0, False,
                                   torch.float64).numpy()
    exp = H @ Ws
    for k, (s, r, o) in enumerate(tr):
        exp[o] += nf[k] * (H[s] @ Vf[:, 0])
        exp[s] += nb[k] * (H[o] @ Vb[:, 0])
    np.testing.assert_allclose(out, exp, rtol=1e-6, atol=1e-6)


def test_oracle_reproduces_committed_layer_goldens(toy, layer_golden):
    g = layer_golden
    tr = g["triples"]
    nf, nb = oracle.graph_norms(tr, 16)
    for variant in ("block", "basis"):
        w = {k[len(variant) + 1:]: v for k, v in g.items()
             if k.startswith(variant + "_") and k.split("_")[1] in ("W", "C", "b")}
        for tag, m, keep, relu in (("plain", None, 1.0, True), ("drop", g["mask"], 0.8, False)):
            o, gr = oracle.layer_fwd_bwd(variant, g["H"], tr, w, nf, nb, g["dOut"], m, keep, relu,
                                         torch.float64)
            np.testing.assert_allclose(o.numpy(), g["%s_%s_out" % (variant, tag)], rtol=1e-12, atol=1e-12)
            for k, v in gr.items():
                np.testing.assert_allclose(v.numpy(), g["%s_%s_d%s" % (variant, tag, k)], rtol=1e-12,
                                           atol=1e-12)
            # float32 oracle (the cpu baseline path) agrees with float64 to fp32 accuracy
            o32, _ = oracle.layer_fwd_bwd(variant, g["H"], tr, w, nf, nb, g["dOut"], m, keep, relu,
                                          torch.float32)
            np.testing.assert_allclose(o32.numpy(), g["%s_%s_out" % (variant, tag)], rtol=1e-4, atol=1e-5)


def test_oracle_gradcheck_float64(toy):
    tr = np.array(toy["train"])[:12]
    V, R, d, B = 16, 9, 4, 2
    rng = np.random.RandomState(3)
    nf, nb = oracle.graph_norms(tr, V)
    H = torch.tensor(rng.normal(size=(V, d)), requires_grad=True)
    Wf = torch.tensor(rng.normal(size=(R, B, d // B, d // B)), requires_grad=True)
    Wb = torch.tensor(rng.normal(size=(R, B, d // B, d // B)), requires_grad=True)
    Ws = torch.tensor(rng.normal(size=(d, d)), requires_grad=True)
    f = lambda h, a, b, c: oracle.concat_gcn_forward(h, tr, a, b, c, nf, nb, None, 1.0, False, torch.float64)
    assert torch.autograd.gradcheck(f, (H, Wf, Wb, Ws), eps=1e-6, atol=1e-5)
    Vf = torch.tensor(rng.normal(size=(d, B, d)), requires_grad=True)
    Vb = torch.tensor(rng.normal(size=(d, B, d)), requires_grad=True)
    Cf = torch.tensor(rng.normal(size=(R, B)), requires_grad=True)
    Cb = torch.tensor(rng.normal(size=(R, B)), requires_grad=True)
    f = lambda h, a, b, c, e, s: oracle.basis_gcn_forward(h, tr, a, b, c, e, s, nf, nb, None, 1.0, False,
                                                          torch.float64)
    assert torch.autograd.gradcheck(f, (H, Vf, Vb, Cf, Cb, Ws), eps=1e-6, atol=1e-5)


def test_distmult_golden_and_formula(layer_golden):
    g = layer_golden
    loss, reg, en = oracle.distmult_loss(g["dm_codes"], g["dm_rel"], g["dm_X"], g["dm_Y"], torch.float64)
    np.testing.assert_allclose(en.numpy(), g["dm_energies"], rtol=1e-12)
    np.testing.assert_allclose(loss.item(), g["dm_loss"], rtol=1e-12)
    np.testing.assert_allclose(reg.item(), g["dm_reg"], rtol=1e-12)
    # weighted CE with pos_weight 1 == plain sigmoid cross entropy
    x = torch.tensor(g["dm_energies"])
    y = torch.tensor(g["dm_Y"], dtype=torch.float64)
    ref = torch.nn.functional.binary_cross_entropy_with_logits(x, y)
    np.testing.assert_allclose(loss.item(), ref.item(), rtol=1e-10)
