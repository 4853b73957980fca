"""GPU parity of the TMA-staged weight-id-major block kernels (csrc/block_staged.cu, block_algo = 3) against the
float64 oracle: block sizes 4 / 8 / 16, full and narrow last slabs, one and two quads per lane, long runs (few
relations, skewed endpoints), work items longer than one and two index batches, dropout + ReLU, and the
messages-only entry points the node-sharded path uses (V_src != V_dst)."""
import numpy as np
import pytest
import torch

from oracle import rgcn_oracle as oracle
from relationprediction_b200 import _lib, ops
from conftest import synthetic_kg
from test_gpu_parity import assert_close, cu, run_block

pytestmark = pytest.mark.gpu


# (RGCN_STG_FWD, RGCN_STG_BWD): default = cp.async staging; the others select the TMA-bulk-copy and the
# two-quads-per-lane variants of the s = 8 kernels (csrc/block_staged.cu launch_block_stg)
VARIANTS = [("", ""), ("2", "0"), ("0", "3"), ("3", "1")]


@pytest.fixture(autouse=True, params=VARIANTS, ids=["default", "tma-nv1", "tma-nv2", "cpasync-nv1"])
def staged_algo(request, monkeypatch):
    fwd, bwd = request.param
    if fwd:
        monkeypatch.setenv("RGCN_STG_FWD", fwd)
    if bwd:
        monkeypatch.setenv("RGCN_STG_BWD", bwd)
    _lib.set_option("block_algo", 3)
    yield
    _lib.set_option("block_algo", -1)


CASES = [
    # V, R, E, d, B, skewed, dropout
    (1500, 23, 12000, 512, 64, False, False),   # s = 8, two full slabs (the synthetic benchmark shape)
    (1500, 3, 30000, 512, 64, True, True),      # s = 8, few relations: items of 128 messages, long runs
    (700, 5, 9000, 264, 33, True, False),       # s = 8, narrow second slab (8 columns)
    (700, 5, 9000, 128, 16, False, True),       # s = 8, one quad per lane
    (800, 11, 6000, 512, 128, True, False),     # s = 4
    (800, 2, 20000, 260, 65, True, True),       # s = 4, narrow second slab (4 columns), long items
    (800, 11, 6000, 512, 32, False, True),      # s = 16
    (500, 4, 8000, 144, 9, True, False),        # s = 16, narrow second slab (16 columns)
    (300, 1, 4000, 64, 8, False, False),        # one relation: every work item is full
]


@pytest.mark.parametrize("V,R,E,d,B,skewed,drop", CASES)
def test_staged_block_layer_vs_oracle(V, R, E, d, B, skewed, drop):
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


def test_staged_equals_register_path_bitwise_free_of_order():
    """Same work items, same per-run arithmetic: staged and register-path results agree to fp32 summation-order
    noise on a graph large enough to fill every SM (grid = 148 persistent CTAs)."""
    V, R, E, d, B = 30000, 50, 400000, 512, 64
    tr = synthetic_kg(V, R, E, seed=3, skewed=False)
    g = ops.Graph(tr, V, R, device=0)
    gen = torch.Generator(device="cuda:0").manual_seed(0)
    H = torch.randn(V, d, device="cuda:0", generator=gen)
    dOut = torch.randn(V, d, device="cuda:0", generator=gen)
    ws = [(torch.randn(*s, device="cuda:0", generator=gen) * 0.1) for s in ((R, B, 8, 8), (R, B, 8, 8), (d, d))]
    res = {}
    for algo in (1, 3):
        _lib.set_option("block_algo", algo)
        Ht = H.clone().requires_grad_(True)
        wt = [w.clone().requires_grad_(True) for w in ws]
        out = ops.block_layer(Ht, wt[0], wt[1], wt[2], g, B, None, 1.0, True)
        out.backward(dOut)
        torch.cuda.synchronize()
        res[algo] = [out.detach(), Ht.grad] + [w.grad for w in wt]
    for a, b, nm in zip(res[1], res[3], ("out", "dH", "dWf", "dWb", "dWs")):
        err = float((a - b).abs().max() / (a.abs().max() + 1e-30))
        assert err < 2e-5, (nm, err)


def test_staged_messages_only_entry_points():
    """rgcn_block_aggregate / _backward with a separate source row space (halo rows of the node-sharded path)."""
    rng = np.random.RandomState(2)
    V_dst, V_src, R, M, d, B = 400, 650, 6, 7000, 256, 32
    dst = rng.randint(0, V_dst, M).astype(np.int32)
    src = rng.randint(0, V_src, M).astype(np.int32)
    relw = rng.randint(0, 2 * R, M).astype(np.int32)
    norm = rng.uniform(0.1, 1.0, M).astype(np.float32)
    g = ops.Graph.from_messages(dst, src, relw, norm, V_dst, V_src, 2 * R, device=0)
    X = rng.normal(0, 1, (V_src, d)).astype(np.float32)
    G = rng.normal(0, 1, (V_dst, d)).astype(np.float32)
    s = d // B
    Wf = rng.normal(0, 0.3, (R, B, s, s)).astype(np.float32)
    Wb = rng.normal(0, 0.3, (R, B, s, s)).astype(np.float32)
    out = torch.zeros(V_dst, d, device="cuda:0")
    ops.block_aggregate_(out, cu(X), cu(Wf), cu(Wb), g, B)
    dX, dWf, dWb = ops.block_aggregate_backward(cu(X), cu(Wf), cu(Wb), cu(G), g, B)
    torch.cuda.synchronize()
    # float64 restatement of the same messages
    W = np.concatenate([Wf, Wb]).astype(np.float64)
    Xb = X.astype(np.float64).reshape(V_src, B, s)
    msg = np.einsum("mbij,mbj->mbi", W[relw], Xb[src]) * norm[:, None, None]
    ref = np.zeros((V_dst, B, s))
    np.add.at(ref, dst, msg)
    assert_close("aggregate", out.cpu().numpy(), ref.reshape(V_dst, d))
    Gb = G.astype(np.float64).reshape(V_dst, B, s)[dst] * norm[:, None, None]
    ref_dX = np.zeros((V_src, B, s))
    np.add.at(ref_dX, src, np.einsum("mbij,mbi->mbj", W[relw], Gb))
    ref_dW = np.zeros((2 * R, B, s, s))
    np.add.at(ref_dW, relw, np.einsum("mbi,mbj->mbij", Gb, Xb[src]))
    assert_close("dX", dX.cpu().numpy(), ref_dX.reshape(V_src, d))
    assert_close("dWf", dWf.cpu().numpy(), ref_dW[:R])
    assert_close("dWb", dWb.cpu().numpy(), ref_dW[R:])
