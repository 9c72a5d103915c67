#!/usr/bin/env python
"""Accuracy of the implicit filter k (D, L) at L = 2^20, D = 256: ours vs the reference's fp32 path, both against fp64."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import hyena_oracle as O
import hyena_dna_b200 as H
dev = torch.device("cuda:0")
L, D = 1 << 20, 256
for std in (0.02, None):
    g = torch.Generator().manual_seed(42)
    P = O.init_params(D, L, emb_dim=5, w=10.0, generator=g, init_std=std)
    k64 = O.hyena_filter(L, {k: v.to(dev).double() for k, v in P.items()})[0].t().cpu()
    k32 = O.hyena_filter(L, {k: v.to(dev) for k, v in P.items()})[0].t().double().cpu()
    f = H.HyenaFilter(D, emb_dim=5, order=64, seq_len=L, w=10.0, lr_pos_emb=0.0).to(dev)
    sd = {k[len("filter_fn."):]: v for k, v in P.items() if k.startswith("filter_fn.")}
    for extra in ("implicit_filter.3.freq", "implicit_filter.5.freq"):
        sd[extra] = sd["implicit_filter.1.freq"]
    f.load_state_dict(sd)
    with torch.no_grad():
        ko = f.filter_channel_major(L).double().cpu()
    for name, a in (("ref32", k32), ("ours", ko)):
        e = (a - k64).abs()
        print(f"init_std={std}: {name:6s} k: max|err| {e.max().item():.3e} (max|k| {k64.abs().max().item():.3e})  normwise {(e.norm() / k64.norm()).item():.3e}")
