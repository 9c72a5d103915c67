#!/usr/bin/env python
"""Time each projection GEMM shape of large-1m individually (CUDA events), own tcgen05 kernels vs cuBLASLt BF16x9.
   python tools/prof_proj.py [--once]    (--once: a single call of each, for ncu)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import hyena_dna_b200 as H  # noqa: E402

once = "--once" in sys.argv
dev = torch.device("cuda:0")
torch.backends.cuda.matmul.allow_tf32 = False
B, L, D = 1, 1 << 20, 256
g = torch.Generator(device="cpu").manual_seed(0)
u = torch.randn(B, L, D, device=dev)
ych = torch.randn(B, D, L, device=dev)
ds = torch.randn(B, 3 * D, L, device=dev)
Wi = torch.randn(3 * D, D, device=dev) * 0.02
Wo = torch.randn(D, D, device=dev) * 0.02
bo = torch.randn(D, device=dev)
sw = torch.randn(3 * D, 3, device=dev)


def timeit(name, fn, flops):
    n = 1 if once else 5
    if not once:
        fn(); fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"{name:34s} {ms:8.3f} ms   {flops / ms / 1e9:8.1f} TFLOP/s fp32-equivalent", flush=True)


f_in = 2.0 * L * D * 3 * D
f_out = 2.0 * L * D * D
timeit("in_proj    act_row->out_ch N=768", lambda: H.ops.proj_gemm(u, 0, Wi, False, 0), f_in)
timeit("out_proj   act_ch->out_row N=256", lambda: H.ops.proj_gemm(ych, 1, Wo, False, 1, bias=bo), f_out)
timeit("d_pre      act_row->out_ch N=256", lambda: H.ops.proj_gemm(u, 0, Wo, True, 0), f_out)
timeit("du         act_ch->out_row K=768", lambda: H.ops.proj_gemm(ds, 1, Wi, True, 1), f_in)
timeit("du + FIR   act_ch->out_row K=768", lambda: H.ops.proj_gemm(ds, 1, Wi, True, 1, fir=sw), f_in)
timeit("dWi        wgrad M=768 N=256", lambda: H.ops.proj_wgrad(ds, u), f_in)
timeit("dWi + FIR  wgrad M=768 N=256", lambda: H.ops.proj_wgrad(ds, u, fir=sw), f_in)
timeit("dWo        wgrad M=256 N=256", lambda: H.ops.proj_wgrad(ych, u, transposed_out=True), f_out)
if not once and H.ops.gemm_mode() == "bf16x9":
    p = torch.empty(B, 3 * D, L, device=dev)
    timeit("cuBLASLt in_proj (BF16x9)", lambda: H.ops.gemm(1, 0, L, 3 * D, D, u, D, L * D, Wi, D, 0, p, L, 3 * D * L, batch=B), f_in)
    du = torch.empty(B, L, D, device=dev)
    timeit("cuBLASLt du (BF16x9)", lambda: H.ops.gemm(0, 1, D, L, 3 * D, Wi, D, 0, ds, L, 3 * D * L, du, D, L * D, batch=B), f_in)
    dW = torch.empty(3 * D, D, device=dev)
    timeit("cuBLASLt dWi (BF16x9)", lambda: H.ops.gemm(0, 0, D, 3 * D, L, u[0], D, 0, ds[0], L, 0, dW, D, 0, batch=1), f_in)
