"""Time the pieces of the tensor-core filter backward (stage 1 kernel, stage-2 reductions) at L=2^20, D=256."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hyena_dna_b200 as H
from hyena_dna_b200 import ops

dev = torch.device("cuda:0")
L, D = 1 << 20, 256
a3 = torch.randn(L, 64, device=dev); a2 = torch.randn(L, 64, device=dev); dp3 = torch.randn(L, 64, device=dev)
dh = torch.randn(D, L, device=dev); z = torch.randn(L, 5, device=dev)

def timeit(name, fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    print(f"{name:40s} {e0.elapsed_time(e1)/n:8.3f} ms")

timeit("dW3 = dh @ a3 (torch sgemm)", lambda: dh @ a3)
dW3 = torch.empty(D, 64, device=dev)
if ops.gemm_mode() == "bf16x9":
    timeit("dW3 bf16x9 gemm", lambda: ops.gemm(0, 0, 64, D, L, a3, 64, 0, dh, L, 0, dW3, 64, 0))
    dW2 = torch.empty(64, 64, device=dev)
    # dW2 (64 i x 64 j) row-major = dp3^T a2 ; col-major dW2^T (64 j x 64 i, ld 64) = a2^T (64 x L, op N, ld 64) dp3 (L x 64: stored (64 x L) ld 64 -> op T)
    timeit("dW2 bf16x9 gemm", lambda: ops.gemm(0, 1, 64, 64, L, a2, 64, 0, dp3, 64, 0, dW2, 64, 0))
    ref = dp3.t().double() @ a2.double()
    print("   dW2 gemm max rel err", float((dW2.double() - ref).abs().max() / ref.abs().max()))
timeit("dW2 = dp3.t() @ a2 (torch)", lambda: dp3.t() @ a2)
timeit("dW0 = dp1.t() @ z (torch)", lambda: dp3.t() @ z)
timeit("colsum (L,64)", lambda: dp3.sum(0))
timeit("dz = dp1 @ W0", lambda: dp3 @ torch.randn(64, 5, device=dev))
