#!/usr/bin/env python
"""Accuracy of y / du at large-1m (L=2^20, D=256, B=1) against the fp64 oracle on the same GPU, next to the reference's
own fp32 path: max |err|, normwise error, fraction of elements missing 1e-5 + 1e-3|y|."""
import os, sys, gc
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import hyena_oracle as O
import hyena_dna_b200 as H
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")
B, L, D = 1, 1 << 20, 256
g = torch.Generator().manual_seed(42)
P = O.init_params(D, L, emb_dim=5, w=10.0, generator=g, init_std=0.02)
u, _ = O.nucleotide_activations(B, L, D, seed=2222)
dy = torch.randn(B, L, D, generator=torch.Generator().manual_seed(1))
def oracle(dt):
    Pd = {k: v.to(device=dev, dtype=dt) for k, v in P.items()}
    y, du, gr = O.operator_fwd_bwd(u.to(device=dev, dtype=dt), Pd, dy.to(device=dev, dtype=dt))
    out = (y.double().cpu(), du.double().cpu()); del Pd, y, du, gr; gc.collect(); torch.cuda.empty_cache(); return out
y64, du64 = oracle(torch.float64)
y32, du32 = oracle(torch.float32)
sd = dict(P)
for extra in ("filter_fn.implicit_filter.3.freq", "filter_fn.implicit_filter.5.freq"):
    sd[extra] = sd["filter_fn.implicit_filter.1.freq"]
op = H.HyenaOperator(D, L, emb_dim=5, w=10.0, lr_pos_emb=0.0); op.load_state_dict(sd); op = op.to(dev)
ug = u.to(dev).requires_grad_(True); y = op(ug); y.backward(dy.to(dev)); torch.cuda.synchronize()
yo, duo = y.detach().double().cpu(), ug.grad.double().cpu()
def stats(name, a, t):
    e = (a - t).abs(); tol = 1e-5 + 1e-3 * t.abs()
    print(f"  {name:10s} max|err| {e.max().item():.3e}  normwise {(e.norm() / t.norm()).item():.3e}  miss-fraction {(e > tol).double().mean().item():.3e}")
print(f"proj_mode={H.ops.proj_mode()}  max|y| {y64.abs().max().item():.1f}")
stats("ref32 y", y32, y64); stats("ours  y", yo, y64); stats("ref32 du", du32, du64); stats("ours  du", duo, du64)
