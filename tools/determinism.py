#!/usr/bin/env python
"""Run every forward stage of the operator several times on identical inputs at the headline shape and compare bitwise."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hyena_dna_b200 as H
from hyena_dna_b200 import ops
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")
L, D = 1 << 20, 256
torch.manual_seed(0)
op = H.HyenaOperator(D, L, emb_dim=5, w=10.0, lr_pos_emb=0.0).to(dev)
u = torch.randn(1, L, D, device=dev)
def rep(name, fn, n=4):
    outs = [fn() for _ in range(n)]
    torch.cuda.synchronize()
    bad = [int((o != outs[0]).sum()) for o in outs[1:]]
    mx = [float((o - outs[0]).abs().max()) for o in outs[1:]]
    print(f"{name:28s} differing elements vs run 0: {bad}  max|diff| {mx}")
    return outs[0]
with torch.no_grad():
    k = rep("filter k", lambda: op.filter_fn.filter_channel_major(L))
    ks = rep("filter spectrum", lambda: torch.view_as_real(ops.filter_spectrum(k)).clone())
    p = rep("in_proj (row->ch)", lambda: ops.proj_gemm(u, 0, op.in_proj.weight, False, 0))
    kspec = ops.filter_spectrum(k)
    def core():
        out = ops.HyenaCoreFn.apply(p, op.in_proj.bias, op.short_filter.weight, op.short_filter.bias, k, op.filter_fn.bias, kspec)
        return out.clone()
    yp = rep("core fwd (3 passes)", core)
    y = rep("out_proj (ch->row)", lambda: ops.proj_gemm(yp, 1, op.out_proj.weight, False, 1, bias=op.out_proj.bias))
    rep("whole operator", lambda: op(u).clone(), n=6)
