import sys
sys.path.insert(0, ".")
import torch
from relationprediction_b200 import ops
torch.set_printoptions(linewidth=200, precision=2, sci_mode=False)
K, M, N = 32, 128, 128
def probe(kset, mset, name):
    A = torch.zeros(K, M, device="cuda"); B = torch.zeros(K, N, device="cuda")
    for k in kset:
        for m in mset:
            A[k, m] = 1.0
        B[k] = torch.arange(1, N + 1, device="cuda").float() + 1000 * k
    C = ops.gemm_tn_tf32x3(A, B)
    ref = A.T @ B
    nz = C.nonzero()
    print(name, "max err", float((C - ref).abs().max()), "nnz", len(nz), "expected nnz", int((ref != 0).sum()))
    if len(nz):
        rows = sorted(set(nz[:, 0].tolist()))
        print("   nonzero rows:", rows[:20])
        r0 = rows[0]
        print("   C[%d,:12] =" % r0, C[r0, :12].tolist())
        print("   ref[%d,:12] =" % mset[0], ref[mset[0], :12].tolist())
probe([0], [5], "k0 m5")
probe([3], [5], "k3 m5")
probe([9], [40], "k9 m40")
probe([0], [0, 1, 2, 3], "k0 m0-3")
A = torch.randn(K, M, device="cuda"); B = torch.randn(K, N, device="cuda")
C = ops.gemm_tn_tf32x3(A, B); ref = A.double().T @ B.double()
print("random K=32: rel", float((C.double() - ref).abs().max() / ref.abs().max()))
A = torch.randn(64, M, device="cuda"); B = torch.randn(64, N, device="cuda")
C = ops.gemm_tn_tf32x3(A, B); ref = A.double().T @ B.double()
print("random K=64: rel", float((C.double() - ref).abs().max() / ref.abs().max()))
