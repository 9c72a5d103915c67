#!/usr/bin/env python
"""Where does a projection-GEMM CTA wait?  Per-role barrier-wait cycles of CTA 0 (debug counters in proj_gemm_kernel)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hyena_dna_b200 as H
dev = torch.device("cuda:0")
B, L, D = 1, 1 << 20, 256
u = torch.randn(B, L, D, device=dev); ych = torch.randn(B, D, L, device=dev); ds = torch.randn(B, 3 * D, L, device=dev)
Wi = torch.randn(3 * D, D, device=dev) * 0.02; Wo = torch.randn(D, D, device=dev) * 0.02; sw = torch.randn(3 * D, 3, device=dev)
dbg = torch.zeros(16, dtype=torch.int64, device=dev)
H._lib.lib().hyena_b200_proj_debug_buffer(dbg.data_ptr())
cases = [("in_proj act_row->out_ch", lambda: H.ops.proj_gemm(u, 0, Wi, False, 0)),
         ("out_proj act_ch->out_row", lambda: H.ops.proj_gemm(ych, 1, Wo, False, 1)),
         ("du act_ch K=768", lambda: H.ops.proj_gemm(ds, 1, Wi, True, 1)),
         ("du+FIR", lambda: H.ops.proj_gemm(ds, 1, Wi, True, 1, fir=sw))]
for name, fn in cases:
    fn(); torch.cuda.synchronize(); dbg.zero_(); fn(); torch.cuda.synchronize()
    d = dbg.tolist()
    n = max(d[10], 1)
    print(f"{name:28s} chunks/CTA {d[10]:6d}  total {d[8]/n:6.0f}/chunk | conv: wait S_FULL {d[0]/n:5.0f} A_EMPTY {d[1]/n:5.0f} | producer: wait B_EMPTY "
          f"{d[3]/n:5.0f} S_EMPTY {d[4]/n:5.0f} | MMA(0): wait D_EMPTY {d[5]/n:5.0f} B_FULL {d[6]/n:5.0f} A_FULL {d[7]/n:5.0f} | "
          f"epi: wait D_FULL {d[9]/n:5.0f}  (cycles per chunk)")
for name, fn in [("dWi wgrad", lambda: H.ops.proj_wgrad(ds, u)), ("dWi wgrad + FIR", lambda: H.ops.proj_wgrad(ds, u, fir=sw)),
                 ("dWo wgrad", lambda: H.ops.proj_wgrad(ych, u, transposed_out=True))]:
    fn(); torch.cuda.synchronize(); dbg.zero_(); fn(); torch.cuda.synchronize()
    d = dbg.tolist(); n = max(d[15], 1)
    print(f"{name:18s} chunks/CTA {d[15]:5d} total {d[9]/n:6.0f}/chunk | A side: wait A_EMPTY {d[0]/n:5.0f} S_FULL {d[1]/n:5.0f} consume {d[2]/n:5.0f} "
          f"| B side: wait B_EMPTY {d[3]/n:5.0f} S_FULL {d[4]/n:5.0f} convert {d[5]/n:5.0f} | MMA: wait DM_EMPTY {d[6]/n:5.0f} B_FULL {d[7]/n:5.0f} "
          f"A_FULL {d[8]/n:5.0f} | drain: wait DM_FULL {d[10]/n:5.0f} drain {d[11]/n:5.0f}  (cycles per chunk)")
H._lib.lib().hyena_b200_proj_debug_buffer(0)
