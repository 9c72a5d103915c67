#!/usr/bin/env python
"""Repeat the projection GEMMs on identical inputs and report where runs differ (bitwise)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hyena_dna_b200 as H
from hyena_dna_b200 import ops
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")
L, D = 1 << 20, 256
torch.manual_seed(0)
u = torch.randn(1, L, D, device=dev)
Wi = torch.randn(3 * D, D, device=dev) * 0.05
Wo = torch.randn(D, D, device=dev) * 0.05
ych = torch.randn(1, D, L, device=dev)
ds = torch.randn(1, 3 * D, L, device=dev)
def rep(name, fn, n=25, layout_ch=False):
    ref = fn().clone()
    tot = 0
    for i in range(n):
        o = fn()
        d = (o != ref)
        nb = int(d.sum())
        tot += nb
        if nb:
            idx = d.nonzero()
            if layout_ch:      # (B, N, L)
                ns = sorted(set(idx[:, 1].tolist())); ls = idx[:, 2]
            else:              # (B, L, N)
                ns = sorted(set(idx[:, 2].tolist())); ls = idx[:, 1]
            tiles = sorted(set((ls // 128).tolist()))
            print(f"  {name} run {i}: {nb} differing; output columns {ns[:12]}{'...' if len(ns) > 12 else ''} ({len(ns)}), "
                  f"position tiles {tiles[:8]} ({len(tiles)}), positions in tile {sorted(set((ls % 128).tolist()))[:6]}... max|d| {float((o-ref).abs().max()):.3g}")
    print(f"{name}: {tot} differing elements over {n} repeats")
rep("in_proj row->ch", lambda: ops.proj_gemm(u, 0, Wi, False, 0), layout_ch=True)
rep("out_proj ch->row", lambda: ops.proj_gemm(ych, 1, Wo, False, 1))
rep("dy_pre row->ch", lambda: ops.proj_gemm(u, 0, Wo, True, 0), layout_ch=True)
rep("du ch->row K=768", lambda: ops.proj_gemm(ds, 1, Wi, True, 1))
