import os, sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np, torch
from oracle import rgcn_oracle as oracle
from relationprediction_b200 import ops, _lib
from conftest import synthetic_kg

def rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))

V, R, E, d, B = 800, 11, 6000, 512, 128
tr = synthetic_kg(V, R, E, seed=11, skewed=True)
rng = np.random.RandomState(5)
H = rng.normal(0, 1, (V, d)).astype(np.float32)
dOut = rng.normal(0, 1, (V, d)).astype(np.float32)
w = oracle.init_block_layer(rng, R, d, B)
nf, nb = oracle.graph_norms(tr, V)
ref_out, ref_g = oracle.layer_fwd_bwd("block", H, tr, w, nf, nb, dOut, None, 1.0, True, torch.float64)
for algo in (0, 1):
    for gm in (0, 2):
        _lib.set_option("block_algo", algo); _lib.set_option("gemm_mode", gm)
        for rep in range(3):
            g = ops.Graph(tr, V, R, device=0)
            Ht = torch.tensor(H, device="cuda", requires_grad=True)
            Wf, Wb, Ws = (torch.tensor(w[k], device="cuda", requires_grad=True) for k in ("W_forward", "W_backward", "W_self"))
            out = ops.block_layer(Ht, Wf, Wb, Ws, g, B, None, 1.0, True)
            out.backward(torch.tensor(dOut, device="cuda"))
            torch.cuda.synchronize()
            print("algo", algo, "gemm", gm, "rep", rep, "out %.2e dH %.2e dWf %.2e dWs %.2e" % (
                rel(out.detach().cpu(), ref_out), rel(Ht.grad.cpu(), ref_g["H"]), rel(Wf.grad.cpu(), ref_g["W_forward"]), rel(Ws.grad.cpu(), ref_g["W_self"])), flush=True)
# the GEMM alone on the same operands
G = torch.tensor(dOut, device="cuda") * (torch.tensor(ref_out.numpy(), device="cuda") > 0)
Wsd = torch.tensor(w["W_self"], device="cuda")
for rep in range(3):
    C = ops.gemm_tf32x3(G.contiguous(), Wsd, b_is_nk=True)
    print("gemm alone rel", rel(C.cpu(), (G.double() @ Wsd.double().T).cpu()), flush=True)
# sparse A (many exact zeros) / tiny values
A = torch.randn(800, 512, device="cuda"); A[A.abs() < 1.0] = 0
print("sparse A", rel(ops.gemm_tf32x3(A, Wsd, b_is_nk=True).cpu(), (A.double() @ Wsd.double().T).cpu()))
