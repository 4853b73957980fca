"""GPU parity tests: the sm_100a path (through the C-ABI) against the CPU oracle and the committed
goldens.  Tolerance: fp32 features within 1e-4 relative (BASELINE.json north_star), measured as
max|a-b| / max|b| per tensor; integer/index work is covered bit-exactly in test_cabi_host.py."""
import os

import numpy as np
import pytest
import torch

from oracle import rgcn_oracle as oracle
from relationprediction_b200 import _lib
from relationprediction_b200 import ops
from conftest import synthetic_kg

pytestmark = pytest.mark.gpu

TOL = 1e-4
DEV = "cuda:0"


@pytest.fixture(params=[0, 1, 3], ids=["dst-major", "rel-major", "staged"], autouse=True)
def block_algo(request):
    """Every test runs under the three aggregation algorithms (rgcn_set_option "block_algo"): deterministic
    destination-major, weight-id-major with rows in registers, weight-id-major with TMA-staged rows (falls back to the
    register path for block sizes the staged kernels do not cover)."""
    _lib.set_option("block_algo", request.param)
    yield request.param
    _lib.set_option("block_algo", -1)


def relerr(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


def assert_close(name, got, ref, tol=TOL):
    e = relerr(got, ref)
    assert np.isfinite(np.asarray(got)).all(), name + " has non-finite values"
    assert e < tol, "%s: rel err %.3e >= %.1e" % (name, e, tol)


def cu(x, dtype=torch.float32):
    return torch.as_tensor(np.ascontiguousarray(x)).to(dtype).to(DEV).contiguous()


def run_block(tr, V, R, d, B, H, w, dOut, mask, keep, relu, norm_mode="canonical", norms=None):
    if norms is None:
        g = ops.Graph(tr, V, R, device=0)
    else:
        g = ops.Graph(tr, V, R, norm_mode="explicit", norm_f=norms[0], norm_b=norms[1], device=0)
    Ht = cu(H).requires_grad_(True)
    Wf, Wb, Ws = (cu(w[k]).requires_grad_(True) for k in ("W_forward", "W_backward", "W_self"))
    m = None if mask is None else cu(mask, torch.uint8)
    out = ops.block_layer(Ht, Wf, Wb, Ws, g, B, m, keep, relu)
    out.backward(cu(dOut))
    torch.cuda.synchronize()
    return out.detach().cpu().numpy(), {"H": Ht.grad.cpu().numpy(), "W_forward": Wf.grad.cpu().numpy(),
                                        "W_backward": Wb.grad.cpu().numpy(), "W_self": Ws.grad.cpu().numpy()}


def run_basis(tr, V, R, d, B, H, w, dOut, mask, keep, relu):
    g = ops.Graph(tr, V, R, device=0)
    Ht = cu(H).requires_grad_(True)
    names = ("W_forward", "W_backward", "C_forward", "C_backward", "W_self")
    ts = [cu(w[k]).requires_grad_(True) for k in names]
    m = None if mask is None else cu(mask, torch.uint8)
    out = ops.basis_layer(Ht, ts[0], ts[1], ts[2], ts[3], ts[4], g, m, keep, relu)
    out.backward(cu(dOut))
    torch.cuda.synchronize()
    grads = {"H": Ht.grad.cpu().numpy()}
    grads.update({k: t.grad.cpu().numpy() for k, t in zip(names, ts)})
    return out.detach().cpu().numpy(), grads


@pytest.mark.parametrize("tag,keep,relu", [("plain", 1.0, True), ("drop", 0.8, False)])
def test_toy_block_layer_matches_committed_golden(layer_golden, tag, keep, relu):
    g = layer_golden
    w = {k[6:]: g[k] for k in g if k.startswith("block_W")}
    mask = g["mask"] if tag == "drop" else None
    out, grads = run_block(g["triples"], 16, 9, 8, 2, g["H"], w, g["dOut"], mask, keep, relu)
    assert_close("out", out, g["block_%s_out" % tag])
    for k, v in grads.items():
        assert_close("d" + k, v, g["block_%s_d%s" % (tag, k)])


@pytest.mark.parametrize("tag,keep,relu", [("plain", 1.0, True), ("drop", 0.8, False)])
def test_toy_basis_layer_matches_committed_golden(layer_golden, tag, keep, relu):
    g = layer_golden
    w = {k[6:]: g[k] for k in g if k.startswith("basis_W") or k.startswith("basis_C")}
    mask = g["mask"] if tag == "drop" else None
    out, grads = run_basis(g["triples"], 16, 9, 8, 2, g["H"], w, g["dOut"], mask, keep, relu)
    assert_close("out", out, g["basis_%s_out" % tag])
    for k, v in grads.items():
        assert_close("d" + k, v, g["basis_%s_d%s" % (tag, k)])


BLOCK_CASES = [
    # V, R, E, d, B, skewed, dropout
    (1500, 23, 12000, 500, 100, True, True),    # FB15k-237 shape: s = 5 (gcn_block.exp)
    (1500, 23, 12000, 512, 64, False, False),   # synthetic shape: s = 8
    (800, 11, 6000, 512, 128, True, False),     # s = 4
    (800, 11, 6000, 512, 32, False, True),      # s = 16
    (600, 7, 5000, 24, 4, True, False),         # generic s = 6
    (600, 7, 5000, 40, 4, False, True),         # generic s = 10
    (600, 7, 5000, 200, 200, True, False),      # s = 1 (diagonal)
    (300, 5, 2500, 8, 1, False, False),         # one dense block
]


@pytest.mark.parametrize("V,R,E,d,B,skewed,drop", BLOCK_CASES)
def test_block_layer_fwd_bwd_vs_oracle(V, R, E, d, B, skewed, drop):
    tr = synthetic_kg(V, R, E, seed=11, skewed=skewed)
    rng = np.random.RandomState(5)
    H = rng.normal(0, 1, (V, d)).astype(np.float32)
    dOut = rng.normal(0, 1, (V, d)).astype(np.float32)
    w = oracle.init_block_layer(rng, R, d, B)
    mask = (rng.uniform(size=(V, d)) < 0.8).astype(np.uint8) if drop else None
    keep = 0.8 if drop else 1.0
    nf, nb = oracle.graph_norms(tr, V)
    ref_out, ref_g = oracle.layer_fwd_bwd("block", H, tr, w, nf, nb, dOut, mask, keep, True, torch.float64)
    out, grads = run_block(tr, V, R, d, B, H, w, dOut, mask, keep, True)
    assert_close("out", out, ref_out.numpy())
    for k in ("H", "W_forward", "W_backward", "W_self"):
        assert_close("d" + k, grads[k], ref_g[k].numpy())


def test_block_layer_supertiles(monkeypatch):
    """Weight-id-major path with several supertiles per view (forced small)."""
    monkeypatch.setenv("RGCN_SUPERTILE_ROWS", "100")
    V, R, E, d, B = 700, 9, 8000, 512, 64
    tr = synthetic_kg(V, R, E, seed=5, skewed=False)
    rng = np.random.RandomState(16)
    H = rng.normal(0, 1, (V, d)).astype(np.float32)
    dOut = rng.normal(0, 1, (V, d)).astype(np.float32)
    w = oracle.init_block_layer(rng, R, d, B)
    mask = (rng.uniform(size=(V, d)) < 0.8).astype(np.uint8)
    nf, nb = oracle.graph_norms(tr, V)
    ref_out, ref_g = oracle.layer_fwd_bwd("block", H, tr, w, nf, nb, dOut, mask, 0.8, True, torch.float64)
    out, grads = run_block(tr, V, R, d, B, H, w, dOut, mask, 0.8, True)
    assert_close("out", out, ref_out.numpy())
    for k in ("H", "W_forward", "W_backward", "W_self"):
        assert_close("d" + k, grads[k], ref_g[k].numpy())


def test_block_layer_split_rows_small_item_max(monkeypatch):
    """Rows longer than item_max are cut into several warp items (L2 vector reductions + last
    arriver epilogue); force that path on every row of a small graph."""
    monkeypatch.setenv("RGCN_ITEM_MAX", "8")
    V, R, E, d, B = 400, 9, 9000, 500, 100
    tr = synthetic_kg(V, R, E, seed=3, skewed=True)
    rng = np.random.RandomState(6)
    H = rng.normal(0, 1, (V, d)).astype(np.float32)
    dOut = rng.normal(0, 1, (V, d)).astype(np.float32)
    w = oracle.init_block_layer(rng, R, d, B)
    mask = (rng.uniform(size=(V, d)) < 0.8).astype(np.uint8)
    nf, nb = oracle.graph_norms(tr, V)
    ref_out, ref_g = oracle.layer_fwd_bwd("block", H, tr, w, nf, nb, dOut, mask, 0.8, True, torch.float64)
    out, grads = run_block(tr, V, R, d, B, H, w, dOut, mask, 0.8, True)
    assert_close("out", out, ref_out.numpy())
    for k in ("H", "W_forward", "W_backward", "W_self"):
        assert_close("d" + k, grads[k], ref_g[k].numpy())


def test_block_layer_tf_unsorted_compat_norms(toy):
    tr = np.array(toy["train"], dtype=np.int32)
    V, R, d, B = 16, 9, 20, 4
    rng = np.random.RandomState(8)
    H = rng.normal(0, 1, (V, d)).astype(np.float32)
    dOut = rng.normal(0, 1, (V, d)).astype(np.float32)
    w = oracle.init_block_layer(rng, R, d, B)
    nf, nb = oracle.graph_norms(tr, V, "tf_unsorted_compat")
    ref_out, ref_g = oracle.layer_fwd_bwd("block", H, tr, w, nf, nb, dOut, None, 1.0, False, torch.float64)
    out, grads = run_block(tr, V, R, d, B, H, w, dOut, None, 1.0, False, norms=(nf, nb))
    assert_close("out", out, ref_out.numpy())
    for k in ("H", "W_forward", "W_backward", "W_self"):
        assert_close("d" + k, grads[k], ref_g[k].numpy())


BASIS_CASES = [
    (1500, 18, 12000, 200, 2, False, True),   # WN18 shape: B = 2, d = 200 (BASELINE configs[2])
    (900, 23, 8000, 500, 5, True, False),     # shipped gcn_basis.exp: B = 5, d = 500
    (500, 7, 4000, 64, 3, True, True),
    (500, 7, 4000, 32, 7, False, False),      # B > 5: two passes of 4 bases
    (300, 5, 2000, 8, 1, False, False),
]


@pytest.mark.parametrize("V,R,E,d,B,skewed,drop", BASIS_CASES)
def test_basis_layer_fwd_bwd_vs_oracle(V, R, E, d, B, skewed, drop):
    tr = synthetic_kg(V, R, E, seed=13, skewed=skewed)
    rng = np.random.RandomState(9)
    H = rng.normal(0, 1, (V, d)).astype(np.float32)
    dOut = rng.normal(0, 1, (V, d)).astype(np.float32)
    w = oracle.init_basis_layer(rng, R, d, B)
    mask = (rng.uniform(size=(V, d)) < 0.8).astype(np.uint8) if drop else None
    keep = 0.8 if drop else 1.0
    nf, nb = oracle.graph_norms(tr, V)
    ref_out, ref_g = oracle.layer_fwd_bwd("basis", H, tr, w, nf, nb, dOut, mask, keep, True, torch.float64)
    out, grads = run_basis(tr, V, R, d, B, H, w, dOut, mask, keep, True)
    assert_close("out", out, ref_out.numpy())
    for k in ("H", "W_forward", "W_backward", "C_forward", "C_backward", "W_self"):
        assert_close("d" + k, grads[k], ref_g[k].numpy())


def test_basis_split_rows(monkeypatch):
    monkeypatch.setenv("RGCN_ITEM_MAX", "8")
    V, R, E, d, B = 300, 6, 6000, 200, 2
    tr = synthetic_kg(V, R, E, seed=17, skewed=True)
    rng = np.random.RandomState(10)
    H = rng.normal(0, 1, (V, d)).astype(np.float32)
    dOut = rng.normal(0, 1, (V, d)).astype(np.float32)
    w = oracle.init_basis_layer(rng, R, d, B)
    nf, nb = oracle.graph_norms(tr, V)
    ref_out, ref_g = oracle.layer_fwd_bwd("basis", H, tr, w, nf, nb, dOut, None, 1.0, False, torch.float64)
    out, grads = run_basis(tr, V, R, d, B, H, w, dOut, None, 1.0, False)
    assert_close("out", out, ref_out.numpy())
    for k in ("H", "W_forward", "W_backward", "C_forward", "C_backward", "W_self"):
        assert_close("d" + k, grads[k], ref_g[k].numpy())


def test_empty_graph_layer_is_self_loop_only():
    V, R, d, B = 40, 3, 16, 4
    rng = np.random.RandomState(1)
    H = rng.normal(0, 1, (V, d)).astype(np.float32)
    w = oracle.init_block_layer(rng, R, d, B)
    out, grads = run_block(np.zeros((0, 3), np.int32), V, R, d, B, H, w, np.ones((V, d), np.float32),
                           None, 1.0, True)
    assert_close("out", out, np.maximum(H @ w["W_self"], 0))
    assert np.abs(grads["W_forward"]).max() == 0 and np.abs(grads["W_backward"]).max() == 0


def test_halo_message_graph_against_dense_reference():
    """1-D node-shard form: V_dst local rows, sources index [local | halo] rows."""
    rng = np.random.RandomState(4)
    V_dst, V_src, R, M, d, B = 60, 100, 4, 900, 16, 4
    s = d // B
    dst = rng.randint(0, V_dst, M).astype(np.int32)
    src = rng.randint(0, V_src, M).astype(np.int32)
    relw = rng.randint(0, 2 * R, M).astype(np.int32)
    norm = rng.rand(M).astype(np.float32)
    g = ops.Graph.from_messages(dst, src, relw, norm, V_dst, V_src, 2 * R, device=0)
    H = rng.normal(size=(V_src, d)).astype(np.float32)
    w = oracle.init_block_layer(rng, R, d, B)
    dOut = rng.normal(size=(V_dst, d)).astype(np.float32)
    Wcat = np.concatenate([w["W_forward"], w["W_backward"]]).astype(np.float64)
    # dense float64 reference with autograd
    Ht = torch.tensor(H, dtype=torch.float64, requires_grad=True)
    Wt = torch.tensor(Wcat, requires_grad=True)
    Ws = torch.tensor(w["W_self"], dtype=torch.float64, requires_grad=True)
    msgs = torch.matmul(Wt[relw.astype(np.int64)], Ht[src.astype(np.int64)].reshape(M, B, s, 1)).reshape(M, d)
    agg = torch.zeros(V_dst, d, dtype=torch.float64).index_add(0, torch.tensor(dst.astype(np.int64)),
                                                                msgs * torch.tensor(norm, dtype=torch.float64)[:, None])
    ref = torch.relu(agg + Ht[:V_dst] @ Ws)
    ref.backward(torch.tensor(dOut, dtype=torch.float64))
    Hc = cu(H).requires_grad_(True)
    Wf, Wb, Wsc = (cu(w[k]).requires_grad_(True) for k in ("W_forward", "W_backward", "W_self"))
    out = ops.block_layer(Hc, Wf, Wb, Wsc, g, B, None, 1.0, True)
    out.backward(cu(dOut))
    assert_close("out", out.detach().cpu().numpy(), ref.detach().numpy())
    assert_close("dH", Hc.grad.cpu().numpy(), Ht.grad.numpy())
    assert_close("dW", np.concatenate([Wf.grad.cpu().numpy(), Wb.grad.cpu().numpy()]), Wt.grad.numpy())
    assert_close("dWself", Wsc.grad.cpu().numpy(), Ws.grad.numpy())


def test_distmult_golden_and_large(layer_golden):
    g = layer_golden
    codes, rel = cu(g["dm_codes"]).requires_grad_(True), cu(g["dm_rel"]).requires_grad_(True)
    en, loss, reg = ops.distmult(codes, rel, cu(g["dm_X"], torch.int32), cu(g["dm_Y"]))
    (loss + 0.01 * reg).backward()
    assert_close("energies", en.detach().cpu().numpy(), g["dm_energies"])
    assert_close("loss", loss.item(), g["dm_loss"])
    assert_close("reg", reg.item(), g["dm_reg"])
    assert_close("dcodes", codes.grad.cpu().numpy(), g["dm_dcodes"])
    assert_close("drel", rel.grad.cpu().numpy(), g["dm_drel"])
    # FB15k-237 decoder shape (d = 500), 11 labels per positive like NegativeSampleRate = 10
    rng = np.random.RandomState(12)
    V, d, N = 3000, 500, 33000
    X = np.stack([rng.randint(0, V, N), rng.randint(0, 237, N), rng.randint(0, V, N)], 1).astype(np.int32)
    Y = np.zeros(N, np.float32)
    Y[:N // 11] = 1
    c, r = rng.normal(0, 0.3, (V, d)).astype(np.float32), rng.normal(0, 0.3, (V, d)).astype(np.float32)
    ct = torch.tensor(c, dtype=torch.float64, requires_grad=True)
    rt = torch.tensor(r, dtype=torch.float64, requires_grad=True)
    l, q, e = oracle.distmult_loss(ct, rt, X, Y, torch.float64)
    (l + 0.01 * q).backward()
    cg, rg = cu(c).requires_grad_(True), cu(r).requires_grad_(True)
    en, loss, reg = ops.distmult(cg, rg, cu(X, torch.int32), cu(Y))
    (loss + 0.01 * reg).backward()
    assert_close("energies", en.detach().cpu().numpy(), e.detach().numpy())
    assert_close("loss", loss.item(), l.item())
    assert_close("reg", reg.item(), q.item())
    assert_close("dcodes", cg.grad.cpu().numpy(), ct.grad.numpy())
    assert_close("drel", rg.grad.cpu().numpy(), rt.grad.numpy())
    # energies-only scoring path (predict): gradient through energies
    cg2 = cu(c).requires_grad_(True)
    en2, _, _ = ops.distmult(cg2, cu(r), cu(X, torch.int32), None)
    en2.sum().backward()
    ct2 = torch.tensor(c, dtype=torch.float64, requires_grad=True)
    e2, _ = oracle.distmult_energies(ct2, torch.tensor(r, dtype=torch.float64), X, torch.float64)
    e2.sum().backward()
    assert_close("dcodes(energy)", cg2.grad.cpu().numpy(), ct2.grad.numpy())


def test_no_cpu_fallback():
    g = ops.Graph(np.array([[0, 0, 1]], np.int32), 4, 1, device=0)
    H = torch.zeros(4, 8)
    W = torch.zeros(1, 2, 4, 4)
    with pytest.raises(_lib.RgcnError, match="CUDA"):
        ops.block_layer(H, W, W, torch.zeros(8, 8), g, 2)


def test_linearity_and_determinism_at_full_width(block_algo):
    """Size-independent properties on a graph too large for the oracle: the layer without ReLU is
    linear in H, and the unsplit-row forward is bit-reproducible run to run."""
    V, R, E, d, B = 20000, 237, 300000, 500, 100
    tr = synthetic_kg(V, R, E, seed=21, skewed=True)
    g = ops.Graph(tr, V, R, device=0)
    gen = torch.Generator(device=DEV).manual_seed(0)
    H1 = torch.randn(V, d, device=DEV, generator=gen)
    H2 = torch.randn(V, d, device=DEV, generator=gen)
    Wf = torch.randn(R, B, 5, 5, device=DEV, generator=gen) * 0.2
    Wb = torch.randn(R, B, 5, 5, device=DEV, generator=gen) * 0.2
    Ws = torch.randn(d, d, device=DEV, generator=gen) * 0.05
    f = lambda h: ops.block_layer(h, Wf, Wb, Ws, g, B, None, 1.0, False)
    o1, o2, o12 = f(H1), f(H2), f(2.0 * H1 - 3.0 * H2)
    ref = 2.0 * o1 - 3.0 * o2
    assert_close("linearity", o12.cpu().numpy(), ref.cpu().numpy(), 5e-5)
    o1b = f(H1)
    info = g.info()
    if info[7] == 0 and block_algo == 0:
        assert torch.equal(o1, o1b)
    else:
        assert_close("rerun", o1b.cpu().numpy(), o1.cpu().numpy(), 1e-6)
    # row-stochastic check: identity blocks, zero self loop, constant features => out = (#dirs with msgs)
    eye = torch.eye(5, device=DEV).repeat(R, B, 1, 1).contiguous()
    ones = torch.ones(V, d, device=DEV)
    o = ops.block_layer(ones, eye, eye, torch.zeros(d, d, device=DEV), g, B, None, 1.0, False)
    s_, o_ = tr[:, 0], tr[:, 2]
    expect = (np.bincount(o_, minlength=V) > 0).astype(np.float32) + (np.bincount(s_, minlength=V) > 0)
    assert_close("row sums", o[:, 0].cpu().numpy(), expect, 1e-5)
