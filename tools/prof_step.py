"""Minimal driver for ncu captures: W warm-up + K fwd+bwd steps of HyenaOperator at the bench shape.

    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
        python tools/prof_step.py --warmup 1 --steps 1
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import hyena_dna_b200 as H  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--seqlen", type=int, default=1 << 20)
ap.add_argument("--d-model", type=int, default=256)
ap.add_argument("--batch", type=int, default=1)
ap.add_argument("--warmup", type=int, default=1)
ap.add_argument("--steps", type=int, default=1)
a = ap.parse_args()
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")
torch.manual_seed(0)
op = H.HyenaOperator(a.d_model, a.seqlen, emb_dim=5, w=10.0, lr_pos_emb=0.0).to(dev)
u = torch.randn(a.batch, a.seqlen, a.d_model, device=dev, requires_grad=True)
dy = torch.randn(a.batch, a.seqlen, a.d_model, device=dev)
for i in range(a.warmup + a.steps):
    if i == a.warmup:
        torch.cuda.synchronize()
        torch.cuda.nvtx.range_push("timed")
    y = op(u)
    y.backward(dy)
torch.cuda.synchronize()
print("done", H.launch_count())
