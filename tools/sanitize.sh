#!/bin/bash
# compute-sanitizer passes over small shapes that exercise every kernel family (run under gpurun).
set -u
mkdir -p gpurun_out
cat > /tmp/san_case.py <<'PY'
import sys, torch
sys.path.insert(0, ".")
import hyena_dna_b200 as H
dev = torch.device("cuda:0")
torch.manual_seed(0)
for (B, L, D) in [(2, 1024, 8), (1, 3000, 8), (1, 40000, 4)]:
    op = H.HyenaOperator(D, L, emb_dim=5, w=10.0, lr_pos_emb=0.0).to(dev)
    u = torch.randn(B, L, D, device=dev, requires_grad=True)
    y = op(u); y.backward(torch.randn_like(y))
    uu = torch.randn(B, 3, L, device=dev, requires_grad=True); k = torch.randn(3, L, device=dev, requires_grad=True)
    Dv = torch.randn(3, device=dev, requires_grad=True)
    o = H.fftconv_func(uu, k, Dv, gelu=False, k_rev=k.detach().flip(-1)); o.backward(torch.randn_like(o))
# filter options, order 3, block glue (round 2)
from functools import partial
op = H.HyenaOperator(8, 512, order=3, emb_dim=5, normalized=True, modulation_lr=1e-3, lr_pos_emb=0.0).to(dev)
y = op(torch.randn(2, 512, 8, device=dev, requires_grad=True)); y.sum().backward()
bb = H.Backbone(16, 2, partial(H.HyenaOperator, l_max=256, emb_dim=5, lr_pos_emb=0.0)).to(dev)
x = torch.randn(2, 256, 16, device=dev, requires_grad=True); bb(x).square().sum().backward()
x = torch.randn(300, 50, device=dev, requires_grad=True); w = torch.randn(50, device=dev, requires_grad=True)
yy, rr = H.ops.add_layer_norm(x, None, w, None, 1e-5); (yy.sum() + rr.sum()).backward()
torch.cuda.synchronize(); print("ok")
PY
for tool in memcheck synccheck racecheck; do
  echo "== $tool"
  timeout 420 compute-sanitizer --tool $tool --print-limit 5 python /tmp/san_case.py > gpurun_out/sanitizer_$tool.log 2>&1
  tail -4 gpurun_out/sanitizer_$tool.log
done
