import sys, torch
sys.path.insert(0, ".")
from relationprediction_b200 import ops
for M, N, K in [(14541, 500, 500), (200000, 512, 512), (1000000, 512, 512)]:
    A = torch.randn(M, K, device="cuda"); B = torch.randn(K, N, device="cuda")
    def t(f, n=10):
        f(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n): f()
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / n
    t_ours = t(lambda: ops.gemm_tf32x3(A, B))
    t_torch = t(lambda: A @ B)
    fl = 2.0 * M * N * K
    print("M=%d N=%d K=%d  tcgen05 3xTF32: %.3f ms (%.1f TFLOP/s fp32-equivalent)   cuBLAS fp32: %.3f ms (%.1f TFLOP/s)"
          % (M, N, K, t_ours, fl / t_ours / 1e9, t_torch, fl / t_torch / 1e9))
for K, M, N in [(14541, 500, 500), (200000, 512, 512)]:
    A = torch.randn(K, M, device="cuda"); B = torch.randn(K, N, device="cuda")
    t_ours = t(lambda: ops.gemm_tn_tf32x3(A, B))
    t_torch = t(lambda: A.T @ B)
    fl = 2.0 * M * N * K
    print("TN K=%d M=%d N=%d  tcgen05 3xTF32: %.3f ms (%.1f TFLOP/s fp32-equivalent)   cuBLAS fp32: %.3f ms (%.1f TFLOP/s)"
          % (K, M, N, t_ours, fl / t_ours / 1e9, t_torch, fl / t_torch / 1e9))
