#!/usr/bin/env python
"""Accuracy of the tcgen05 projection GEMM against fp64: normwise error and the SIGNED relative bias
mean((got - ref) * sign(ref)) / mean|ref| (truncating accumulators shrink results systematically), next to torch fp32
(cuBLAS SGEMM, the reference's nn.Linear) and cuBLASLt BF16x9."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hyena_dna_b200 as H
from hyena_dna_b200 import ops
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")
L, K, N = 1 << 17, 256, 768
g = torch.Generator().manual_seed(0)
act = torch.randn(1, L, K, generator=g).to(dev)
W = (torch.randn(N, K, generator=g) * 0.06).to(dev)
ref = (act.double() @ W.double().t())
def stats(name, got):
    e = got.double() - ref
    nrm = (e.norm() / ref.norm()).item()
    bias = ((e * ref.sign()).mean() / ref.abs().mean()).item()
    print(f"  {name:14s} normwise {nrm:.3e}   signed bias {bias:+.3e}   max {e.abs().max().item():.3e}")
stats("torch fp32", act @ W.t())
out = ops.proj_gemm(act, 0, W, False, 1)
stats("tc row->row", out)
out = ops.proj_gemm(act, 0, W, False, 0)          # (B, N, L)
stats("tc row->ch", out.transpose(1, 2))
actc = act.transpose(1, 2).contiguous()
out = ops.proj_gemm(actc, 1, W, False, 1)
stats("tc ch->row", out)
