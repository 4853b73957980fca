"""GPU: the tcgen05 3xTF32 GEMM (gemm_tf32x3.cu) against a float64 reference -- fp32-level accuracy
(1e-5 relative, an order tighter than the layer bar) at the layer's shapes, ragged edges included."""
import numpy as np
import pytest
import torch

from relationprediction_b200 import ops

pytestmark = pytest.mark.gpu


def rel(a, b):
    return float((a.double() - b).abs().max() / b.abs().max())


@pytest.mark.parametrize("M,N,K", [(128, 128, 32), (128, 128, 128), (256, 256, 64), (14541, 500, 500),
                                   (1000, 512, 512), (77, 8, 8), (300, 200, 200), (129, 132, 36),
                                   (5000, 24, 40),
                                   # many tiles per persistent CTA (one / two / sixteen k-blocks per tile)
                                   (100000, 512, 512), (40000, 132, 32), (30000, 24, 40)])
@pytest.mark.parametrize("b_is_nk", [False, True])
def test_gemm_tf32x3_matches_float64(M, N, K, b_is_nk):
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N)
    A = torch.randn(M, K, device="cuda", generator=g)
    B = torch.randn(*((N, K) if b_is_nk else (K, N)), device="cuda", generator=g)
    ref = A.double() @ (B.double().T if b_is_nk else B.double())
    C = ops.gemm_tf32x3(A, B, b_is_nk=b_is_nk)
    assert torch.isfinite(C).all()
    assert rel(C, ref) < 1e-5, rel(C, ref)
    # accumulate form
    C0 = torch.randn(M, N, device="cuda", generator=g)
    C1 = ops.gemm_tf32x3(A, B, b_is_nk=b_is_nk, out=C0.clone(), accumulate=True)
    assert rel(C1, ref + C0.double()) < 1e-5


def test_gemm_tf32x3_is_tighter_than_single_tf32_and_handles_scales():
    g = torch.Generator(device="cuda").manual_seed(1)
    A = torch.randn(2048, 512, device="cuda", generator=g) * 1e3
    B = torch.randn(512, 512, device="cuda", generator=g) * 1e-3
    ref = A.double() @ B.double()
    assert rel(ops.gemm_tf32x3(A, B), ref) < 1e-5
    # strided A (leading dimension > K)
    big = torch.randn(700, 1000, device="cuda", generator=g)
    Av = big[:, :500]
    Bm = torch.randn(500, 500, device="cuda", generator=g)
    lib_out = ops.gemm_tf32x3(Av.contiguous(), Bm)
    assert rel(lib_out, Av.double() @ Bm.double()) < 1e-5


@pytest.mark.parametrize("K,M,N", [(128, 128, 128), (1000, 128, 128), (14541, 500, 500), (50000, 512, 512),
                                   (33, 8, 8), (4097, 132, 36), (3000, 2500, 500), (100, 200, 24)])
def test_gemm_tn_tf32x3_matches_float64(K, M, N):
    g = torch.Generator(device="cuda").manual_seed(K + M)
    A = torch.randn(K, M, device="cuda", generator=g)
    B = torch.randn(K, N, device="cuda", generator=g)
    ref = A.double().T @ B.double()
    C = ops.gemm_tn_tf32x3(A, B)
    assert torch.isfinite(C).all()
    assert rel(C, ref) < 1e-5, rel(C, ref)
    C0 = torch.randn(M, N, device="cuda", generator=g)
    C1 = ops.gemm_tn_tf32x3(A, B, out=C0.clone(), accumulate=True)
    assert rel(C1, ref + C0.double()) < 1e-5
