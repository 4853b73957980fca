"""GPU (one device is enough): the row helpers of the node-sharded layer -- rgcn_rows_gather (the halo push, here
into local memory), rgcn_rows_add (non-atomic unpack of one peer's returned gradients) and rgcn_relu_backward --
against torch indexing.  All three are exact operations: bit-equal results.  The multi-GPU tests
(test_gpu_parallel.py) use them across devices; the driver's single-GPU test box skips those."""
import pytest
import torch

from relationprediction_b200 import _lib, ops

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n_rows,d,n", [(1000, 512, 700), (333, 500, 333), (50, 8, 1), (64, 36, 0), (4097, 128, 4096)])
def test_row_helpers_match_torch_indexing(n_rows, d, n):
    g = torch.Generator(device="cuda").manual_seed(n_rows + d)
    src = torch.randn(n_rows, d, device="cuda", generator=g)
    rows = torch.randperm(n_rows, device="cuda", generator=g)[:n].contiguous()        # unique, int64
    for max_ctas in (0, 3):                                                           # default grid / bounded grid
        dst = torch.full((max(n, 1), d), float("nan"), device="cuda")
        ops.rows_gather_to(dst.data_ptr(), src, rows, max_ctas)
        torch.cuda.synchronize()
        assert torch.equal(dst[:n], src[rows])
    acc = torch.randn(n_rows, d, device="cuda", generator=g)
    add = torch.randn(n, d, device="cuda", generator=g)
    want = acc.clone()
    want[rows] += add
    ops.rows_add_(acc, rows, add)
    assert torch.equal(acc, want)
    out = torch.randn(n_rows, d, device="cuda", generator=g)
    dOut = torch.randn(n_rows, d, device="cuda", generator=g)
    G = ops.relu_backward(dOut, out)
    assert torch.equal(G, torch.where(out > 0, dOut, torch.zeros_like(dOut)))


def test_row_helpers_reject_bad_arguments():
    src = torch.zeros(8, 6, device="cuda")                    # d % 4 != 0
    rows = torch.arange(4, device="cuda")
    dst = torch.zeros(4, 6, device="cuda")
    with pytest.raises(_lib.RgcnError):
        ops.rows_gather_to(dst.data_ptr(), src, rows)
    with pytest.raises(_lib.RgcnError):
        ops.rows_gather_to(dst.data_ptr(), torch.zeros(8, 8, device="cuda"), rows.to(torch.int32))
    with pytest.raises(_lib.RgcnError):
        ops.rows_add_(torch.zeros(8, 8, device="cuda"), rows, torch.zeros(3, 8, device="cuda"))
