"""CPU: numpy emulation of the index algebra of the experimental component-major block path (csrc/block_cm.cu:
k_to_cm, k_relayout_cm, k_block_cm<DW=false/true>, k_cm_add, k_unlayout_cm) against the oracle.  This pins the
layout conventions (which index is transposed where) that the CUDA kernels transcribe; the kernels themselves have
not run on a GPU yet (tests/test_gpu_parity.py::test_component_major_path_opt_in, RGCN_RUN_EXPERIMENTAL=1)."""
import numpy as np
import pytest
import torch

from oracle import rgcn_oracle as oracle
from conftest import synthetic_kg


def to_cm(X, B, S):                      # Xc[row][i*B + b] = X[row][b*S + i]
    return X.reshape(-1, B, S).transpose(0, 2, 1).reshape(-1, B * S)


def from_cm(Xc, B, S):                   # inverse (what k_cm_add adds into the row-major output)
    return Xc.reshape(-1, S, B).transpose(0, 2, 1).reshape(-1, B * S)


def relayout_cm(Wf, Wb, transpose):      # Wc[w][a][c][b] = W[w][b][i][j], (i, j) = (c, a) if transpose else (a, c)
    W = np.concatenate([Wf, Wb], 0)      # [2R, B, S, S] indexed [w][b][i][j]
    return W.transpose(0, 3, 2, 1) if transpose else W.transpose(0, 2, 3, 1)


@pytest.mark.parametrize("B,S", [(8, 5), (100, 5), (4, 5)])
def test_component_major_algebra(B, S):
    V, R, E = 60, 4, 700
    d = B * S
    tr = synthetic_kg(V, R, E, seed=2, skewed=True)
    rng = np.random.RandomState(0)
    H = rng.normal(size=(V, d))
    dOut = rng.normal(size=(V, d))
    w = {k: v.astype(np.float64) for k, v in oracle.init_block_layer(rng, R, d, B).items()}
    nf, nb = oracle.graph_norms(tr, V, "canonical", np.float64)
    ref_out, ref_g = oracle.layer_fwd_bwd("block", H, tr, w, nf, nb, dOut, None, 1.0, True, torch.float64)
    dst, src, relw, norm = oracle.messages_from_triples(tr, R, V)
    norm = np.concatenate([nf, nb])
    Wf, Wb, Ws = w["W_forward"], w["W_backward"], w["W_self"]

    # forward: Mc[dst][i][b] += norm * sum_j Wc[w][i][j][b] * Hc[src][j][b]
    Hc = to_cm(H, B, S).reshape(V, S, B)
    Wc = relayout_cm(Wf, Wb, transpose=False)                       # [w][i][j][b]
    msg = np.einsum("mijb,mjb->mib", Wc[relw], Hc[src]) * norm[:, None, None]
    Mc = np.zeros((V, S, B))
    np.add.at(Mc, dst, msg)
    out = np.maximum(H @ Ws + from_cm(Mc.reshape(V, d), B, S), 0)
    assert np.abs(out - ref_out.numpy()).max() < 1e-10

    # backward: G = dOut * relu'; dHc[src][a][b] += norm * sum_c Wct[w][a][c][b] * Gc[dst][c][b]
    G = dOut * (out > 0)
    Gc = to_cm(G, B, S).reshape(V, S, B)
    Wct = relayout_cm(Wf, Wb, transpose=True)                       # [w][a][c][b] = W[w][b][c][a]
    back = np.einsum("macb,mcb->mab", Wct[relw], Gc[dst]) * norm[:, None, None]
    dHc = np.zeros((V, S, B))
    np.add.at(dHc, src, back)
    dH = G @ Ws.T + from_cm(dHc.reshape(V, d), B, S)
    assert np.abs(dH - ref_g["H"].numpy()).max() < 1e-10

    # weight gradient: dWc[w][i][j][b] += norm * Gc[dst][i][b] * Hc[src][j][b];  dW[w][b][i][j] = dWc[w][i][j][b]
    dWc = np.zeros((2 * R, S, S, B))
    np.add.at(dWc, relw, np.einsum("mib,mjb->mijb", Gc[dst], Hc[src]) * norm[:, None, None, None])
    dW = dWc.transpose(0, 3, 1, 2)
    assert np.abs(dW[:R] - ref_g["W_forward"].numpy()).max() < 1e-10
    assert np.abs(dW[R:] - ref_g["W_backward"].numpy()).max() < 1e-10
    assert np.abs(H.T @ G - ref_g["W_self"].numpy()).max() < 1e-10
