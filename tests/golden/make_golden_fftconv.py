"""Golden vectors for the fftconv variants (k_rev, bidirectional) from the UNMODIFIED reference function
src/models/sequence/hyena.py:59-88 ``fftconv_ref`` (build container only).

    python tests/golden/make_golden_fftconv.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden import REF, OUT, _shim  # noqa: E402


def main():
    sys.path.insert(0, REF)
    _shim()
    from src.models.sequence.hyena import fftconv_ref
    out = {}
    for name, (B, H, L, with_rev, bidir) in {"krev_L100": (2, 4, 100, True, False), "krev_L257": (1, 3, 257, True, False),
                                             "bidir_L128": (2, 4, 128, False, True), "bidir_L101": (1, 3, 101, False, True)}.items():
        g = torch.Generator().manual_seed(len(name) * 131 + L)
        u = torch.randn(B, H, L, generator=g)
        k = torch.randn(H, L, generator=g) / L ** 0.5
        kr = torch.randn(H, L, generator=g) / L ** 0.5 if with_rev else None
        D = torch.randn(H, generator=g)
        dy = torch.randn(B, H, L, generator=g)
        res = {}
        for dt, tag in ((torch.float32, ""), (torch.float64, "64")):
            uu, kk, DD = (x.detach().clone().to(dt).requires_grad_(True) for x in (u, k, D))
            kkr = kr.detach().clone().to(dt).requires_grad_(True) if with_rev else None
            y = fftconv_ref(uu, kk, DD, None, gelu=False, k_rev=kkr, bidirectional=bidir)
            y.backward(dy.to(dt))
            res["y" + tag] = y.detach().numpy(); res["du" + tag] = uu.grad.numpy(); res["dk" + tag] = kk.grad.numpy()
            res["dD" + tag] = DD.grad.numpy()
            if with_rev:
                res["dkrev" + tag] = kkr.grad.numpy()
        for kname, v in dict(u=u, k=k, D=D, dy=dy, **({"krev": kr} if with_rev else {})).items():
            out[f"{name}/{kname}"] = v.numpy()
        for kname, v in res.items():
            out[f"{name}/{kname}"] = v
        out[f"{name}/cfg"] = np.array([B, H, L, int(with_rev), int(bidir)], dtype=np.int64)
    path = os.path.join(OUT, "fftconv_variants.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
