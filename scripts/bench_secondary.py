"""Secondary measurements (not the headline bench): DistMult scorer fwd/bwd bandwidth, basis layer
(WN18 shape, BASELINE configs[2]; shipped gcn_basis.exp shape), block layer train-step graph."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from bench import synthetic_kg  # noqa: E402
from relationprediction_b200 import _lib, ops  # noqa: E402

dev = torch.device("cuda", 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / n


out = {}
# ---- DistMult: FB15k-237 train-step decoder shape: N = 330000 triples, d = 500 ----
V, d, N = 14541, 500, 330000
g = torch.Generator(device=dev).manual_seed(0)
codes = torch.randn(V, d, device=dev, generator=g).requires_grad_(True)
rel = torch.randn(V, d, device=dev, generator=g).requires_grad_(True)
X = torch.stack([torch.randint(0, V, (N,), device=dev, generator=g), torch.randint(0, 237, (N,), device=dev, generator=g),
                 torch.randint(0, V, (N,), device=dev, generator=g)], 1).int().contiguous()
Y = (torch.rand(N, device=dev, generator=g) < 0.09).float()


def dm_fwd():
    with torch.no_grad():
        ops.distmult(codes, rel, X, Y)


def dm_fwd_bwd():
    codes.grad = None
    rel.grad = None
    e, l, r = ops.distmult(codes, rel, X, Y)
    (l + 0.01 * r).backward()


t_f = timeit(dm_fwd)
t_fb = timeit(dm_fwd_bwd)
alg_f = N * (12 * d + 16)
alg_b = N * (12 * d + 16) + N * 12 * d * 2
out["distmult"] = {"N": N, "d": d, "fwd_ms": t_f, "fwd_bwd_ms": t_fb, "fwd_GBps_algorithmic": alg_f / t_f / 1e6,
                   "bwd_GBps_algorithmic": alg_b / max(t_fb - t_f, 1e-6) / 1e6,
                   "note": "codes (29 MB) are L2-resident: gathers run above the HBM roofline"}
# DistMult on a table larger than L2
V2, N2 = 1_000_000, 2_000_000
codes2 = torch.randn(V2, 512, device=dev, generator=g)
rel2 = torch.randn(1000, 512, device=dev, generator=g)
X2 = torch.stack([torch.randint(0, V2, (N2,), device=dev, generator=g), torch.randint(0, 1000, (N2,), device=dev, generator=g),
                  torch.randint(0, V2, (N2,), device=dev, generator=g)], 1).int().contiguous()
Y2 = (torch.rand(N2, device=dev, generator=g) < 0.09).float()
t2 = timeit(lambda: ops.distmult(codes2, rel2, X2, Y2), n=5)
out["distmult_hbm"] = {"V": V2, "N": N2, "d": 512, "fwd_ms": t2, "fwd_GBps_algorithmic": N2 * (12 * 512 + 16) / t2 / 1e6,
                       "frac_of_measured_hbm_6569.6": N2 * (8 * 512 + 16) / t2 / 1e6 / 6569.6,
                       "note": "2 of the 3 rows per triple come from HBM (entity table 2 GB), the relation row from L2"}
del codes2, rel2, X2, Y2


# ---- layers ----
def layer_case(name, V, R, E, d, B, variant, skewed):
    tr = synthetic_kg(V, R, E, seed=1234, skewed=skewed)
    gr = ops.Graph(tr, V, R, device=0)
    H = torch.randn(V, d, device=dev, generator=g).requires_grad_(True)
    dOut = torch.randn(V, d, device=dev, generator=g)
    if variant == "block":
        s = d // B
        std = 3.0 / np.sqrt(R + s)
        ws = [(torch.randn(R, B, s, s, device=dev, generator=g) * std).requires_grad_(True) for _ in range(2)]
        ws.append((torch.randn(d, d, device=dev, generator=g) * std).requires_grad_(True))
        f = lambda: ops.block_layer(H, ws[0], ws[1], ws[2], gr, B, None, 1.0, True)
    else:
        std = 3.0 / np.sqrt(2 * d)
        ws = [(torch.randn(d, B, d, device=dev, generator=g) * std).requires_grad_(True) for _ in range(2)]
        ws += [torch.randn(R, B, device=dev, generator=g).requires_grad_(True) for _ in range(2)]
        ws.append((torch.randn(d, d, device=dev, generator=g) * std).requires_grad_(True))
        f = lambda: ops.basis_layer(H, ws[0], ws[1], ws[2], ws[3], ws[4], gr, None, 1.0, True)

    def step():
        H.grad = None
        for w in ws:
            w.grad = None
        f().backward(dOut)
    ms = timeit(step, n=10)
    with torch.no_grad():
        ms_f = timeit(lambda: f(), n=10)
    _lib.profile_enable(True)
    acc = {}
    for _ in range(5):
        flush.zero_()
        step()
        torch.cuda.synchronize()
        for nm, v in _lib.profile_read():
            acc[nm] = acc.get(nm, 0.0) + v / 5
    _lib.profile_enable(False)
    out[name] = {"V": V, "R": R, "E": E, "d": d, "B": B, "variant": variant, "fwd_ms": ms_f, "fwd_bwd_ms": ms,
                 "M_edges_per_s": E / ms / 1e3, "stages_ms": {k: round(v, 4) for k, v in acc.items()}}


layer_case("wn18_basis_B2_d200 (BASELINE configs[2])", 40943, 18, 141442, 200, 2, "basis", True)
layer_case("fb15k237_basis_B5_d500 (shipped gcn_basis.exp)", 14541, 237, 272115, 500, 5, "basis", True)
layer_case("fb15k237_block_trainstep_E15000", 14541, 237, 15000, 500, 100, "block", True)
layer_case("fb15k_block_B100_d500 (BASELINE configs[3] shape, 1 GPU)", 14951, 1345, 483142, 500, 100, "block", True)
print(json.dumps(out, indent=1))
