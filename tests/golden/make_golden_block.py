"""Golden vectors for the block glue from the UNMODIFIED reference code (build container only).

    python tests/golden/make_golden_block.py

A two-layer pre-norm backbone exactly as src/models/sequence/long_conv_lm.py:377-396 runs it: the reference's own
flash_attn.modules.block.Block (flash-attention/flash_attn/modules/block.py, imported from /root/reference), mixer =
the reference HyenaOperator (standalone_hyenadna.py), mlp = the reference's Mlp (fc1 -> gelu -> fc2) for the first
case and nn.Identity for the second, followed by the final add -> LayerNorm (ln_f).  Writes tests/golden/block_*.npz:
state_dict, input, upstream grad, output, input grad, parameter grads (fp32) and the fp64 truth of the same modules.
"""
import copy
import importlib.util
import os
import sys
from functools import partial

import numpy as np
import torch
import torch.nn as nn

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))
CASES = {"block_L128_D32_mlp": (2, 128, 32, True), "block_L96_D16_nomlp": (1, 96, 16, False)}


class RefBackbone(nn.Module):
    def __init__(self, Block, Mlp, S, d_model, l_max, n_layer, with_mlp):
        super().__init__()
        mixer = partial(S.HyenaOperator, l_max=l_max, order=2, filter_order=64, emb_dim=5, w=10.0, shift=0.0, lr_pos_emb=0.0)
        mlp = partial(Mlp, hidden_features=2 * d_model, activation=partial(torch.nn.functional.gelu, approximate="tanh")) \
            if with_mlp else nn.Identity
        self.layers = nn.ModuleList([Block(d_model, mixer, mlp, norm_cls=partial(nn.LayerNorm, eps=1e-5), prenorm=True,
                                           resid_dropout1=0.0, resid_dropout2=0.0, fused_dropout_add_ln=False,
                                           residual_in_fp32=True) for _ in range(n_layer)])
        self.ln_f = nn.LayerNorm(d_model, eps=1e-5)

    def forward(self, h):                      # long_conv_lm.py:383-396 (non-fused branch, dropout p = 0)
        residual = None
        for layer in self.layers:
            h, residual = layer(h, residual)
        residual = (h + residual) if residual is not None else h
        return self.ln_f(residual.to(dtype=self.ln_f.weight.dtype))


def main():
    torch.backends.cuda.matmul.allow_tf32 = False
    sys.path.insert(0, REF)
    sys.path.insert(0, os.path.join(REF, "flash-attention"))
    import standalone_hyenadna as S
    spec = importlib.util.spec_from_file_location("ref_block", os.path.join(REF, "flash-attention/flash_attn/modules/block.py"))
    RB = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(RB)
    from flash_attn.modules.mlp import Mlp      # the reference's vendored tree (first on sys.path)
    assert os.path.realpath(sys.modules["flash_attn"].__file__).startswith(REF), "flash_attn did not resolve to the reference tree"
    for case, (B, L, D, with_mlp) in CASES.items():
        torch.manual_seed(4321)
        m = RefBackbone(RB.Block, Mlp, S, D, L, 2, with_mlp)
        m.apply(partial(S._init_weights, n_layer=2, initializer_range=0.02))
        with torch.no_grad():                    # non-trivial LayerNorm parameters
            for mod in m.modules():
                if isinstance(mod, nn.LayerNorm):
                    mod.weight.add_(0.1 * torch.randn_like(mod.weight)); mod.bias.add_(0.1 * torch.randn_like(mod.bias))
        x = torch.randn(B, L, D, generator=torch.Generator().manual_seed(0))
        dy = torch.randn(B, L, D, generator=torch.Generator().manual_seed(1))
        x1 = x.clone().requires_grad_(True)
        y = m(x1)
        y.backward(dy)
        m64 = copy.deepcopy(m).double()
        m64.zero_grad()
        x64 = x.double().requires_grad_(True)
        y64 = m64(x64)
        y64.backward(dy.double())
        out = {"x": x.numpy(), "dy": dy.numpy(), "y": y.detach().numpy(), "dx": x1.grad.numpy(), "y64": y64.detach().numpy(),
               "dx64": x64.grad.numpy(), "meta": np.array([B, L, D, int(with_mlp)], dtype=np.int64)}
        for k, v in m.state_dict().items():
            out["sd/" + k] = v.numpy()
        for k, p in m.named_parameters():
            if p.grad is not None:
                out["grad/" + k] = p.grad.numpy()
        for k, p in m64.named_parameters():
            if p.grad is not None:
                out["grad64/" + k] = p.grad.numpy()
        path = os.path.join(OUT, case + ".npz")
        np.savez_compressed(path, **out)
        print(case, "->", path, os.path.getsize(path) // 1024, "KiB", "max|y-y64|", float((y.double() - y64).abs().max()))


if __name__ == "__main__":
    main()
