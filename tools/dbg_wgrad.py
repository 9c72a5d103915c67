import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hyena_dna_b200 as H
dev = torch.device("cuda:0")
torch.manual_seed(0)
for (B, L, M, N) in [(8, 32768, 768, 256), (4, 160000, 768, 256), (8, 32768, 256, 256)]:
    X = torch.randn(B, M, L, device=dev)
    Y = torch.randn(B, L, N, device=dev)
    taps = torch.randn(M, 3, device=dev)
    for fir in (None, taps):
        Xd = X.double()
        if fir is not None:
            Xp = torch.nn.functional.pad(Xd, (0, 2))
            Xd = fir[:, 2].double()[None, :, None] * Xp[..., :L] + fir[:, 1].double()[None, :, None] * Xp[..., 1:L + 1] + fir[:, 0].double()[None, :, None] * Xp[..., 2:L + 2]
        ref = torch.einsum("bml,bln->mn", Xd, Y.double())
        outs = [H.ops.proj_wgrad(X, Y, fir=fir).double() for _ in range(3)]
        err = (outs[0] - ref).abs()
        bad = (err > 1e-3 * ref.abs() + 1e-5 * ref.abs().max()).nonzero()
        print(f"B={B} L={L} M={M} N={N} fir={fir is not None}: max err {err.max().item():.3e} (ref max {ref.abs().max().item():.1f}), "
              f"bad {bad.shape[0]}, deterministic {all(torch.equal(outs[0], o) for o in outs[1:])}")
        if bad.shape[0]:
            print("   bad rows", sorted(set(bad[:, 0].tolist()))[:20], "cols", sorted(set(bad[:, 1].tolist()))[:20])
