#!/usr/bin/env python
"""HBM throughput of the fused residual add + LayerNorm kernels at the headline shape (rows = 2^20, D = 256)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hyena_dna_b200 as H
dev = torch.device("cuda:0")
rows, D = 1 << 20, 256
x = torch.randn(rows, D, device=dev); res = torch.randn(rows, D, device=dev)
w = torch.randn(D, device=dev, requires_grad=True); b = torch.randn(D, device=dev, requires_grad=True)
dy = torch.randn(rows, D, device=dev); dres = torch.randn(rows, D, device=dev)
xx = x.requires_grad_(True); rr = res.requires_grad_(True)
def step():
    y, r = H.ops.add_layer_norm(xx, rr, w, b, 1e-5)
    torch.autograd.backward([y, r], [dy, dres])
for _ in range(3): step()
H._lib.profile_begin()
n = 10
for _ in range(n): step()
p = H._lib.profile_end()
ms = p["add_layer_norm"][0] / n
# fwd: x, res in; y, res_out out (+stats); bwd: dy, dres, r in; dx out  -> 8 tensors of rows*D fp32
gb = 8 * rows * D * 4 / 1e9
print(json.dumps({"add_layer_norm fwd+bwd ms": round(ms, 4), "algorithmic_GB": round(gb, 3), "GB/s": round(gb / (ms * 1e-3), 1),
                  "frac_of_measured_hbm_peak": round(gb / (ms * 1e-3) / 6581.9, 3)}))
