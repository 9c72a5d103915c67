"""C-ABI argument checking (no GPU needed: every call below is rejected before any CUDA work)."""
import ctypes

import pytest
from importlib import import_module

_lib = import_module("hyena_dna_b200._lib")


def _err():
    return _lib.lib().hyena_b200_last_error().decode()


def test_sequence_length_limit_is_reported():
    L = _lib.lib()
    one = ctypes.c_void_p(16)       # never dereferenced: the shape check comes first
    rc = L.hyena_b200_core_fwd(one, None, one, one, one, one, one, None, None, 1, 8, (1 << 20) + 2, one, 1 << 30, None)
    assert rc != 0 and "exceeds the supported maximum" in _err()
    rc = L.hyena_b200_fftconv_fwd(one, one, one, one, 1, 2, 0, one, 1 << 20, None)
    assert rc != 0 and "bad shape" in _err()


def test_null_pointers_are_rejected():
    L = _lib.lib()
    rc = L.hyena_b200_core_fwd(None, None, None, None, None, None, None, None, None, 1, 8, 1024, None, 0, None)
    assert rc != 0 and "null pointer" in _err()
    rc = L.hyena_b200_filter_spectrum(None, None, 4, 256, None, 0, None)
    assert rc != 0


def test_filter_shape_limits():
    L = _lib.lib()
    p = ctypes.c_void_p(256)
    args = lambda E, N: (p, E, p, p, p, p, p, p, p, p, p, p, 0.0, 1, 128, E, N, 8, p, None)
    assert L.hyena_b200_filter_fwd(*args(5, 16)) != 0 and "filter_order" in _err()     # HyenaFilter default order=16
    assert L.hyena_b200_filter_fwd(*args(4, 64)) != 0 and "emb_dim" in _err()          # emb_dim must be odd
    assert L.hyena_b200_filter_fwd(*args(17, 64)) != 0 and "emb_dim" in _err()


def test_workspace_sizing_is_consistent():
    L = _lib.lib()
    assert L.hyena_b200_spectrum_elems(1) == 1024
    assert L.hyena_b200_spectrum_elems(1024) == 1024
    assert L.hyena_b200_spectrum_elems(1025) == 2048
    assert L.hyena_b200_spectrum_elems(1 << 20) == 1 << 20
    for B, D, Lq in [(1, 256, 1 << 20), (8, 256, 32768), (4, 256, 160000), (2, 128, 1024)]:
        mn_f = L.hyena_b200_workspace_min_bytes(B, D, Lq, 0)
        mn_b = L.hyena_b200_workspace_min_bytes(B, D, Lq, 1)
        M = L.hyena_b200_spectrum_elems(Lq)
        assert mn_f == B * M * 8 and mn_b == (2 * B + 1) * M * 8
        assert L.hyena_b200_workspace_bytes(B, D, Lq, 0) >= mn_f
        assert L.hyena_b200_workspace_bytes(B, D, Lq, 1) >= mn_b
        assert L.hyena_b200_workspace_bytes(B, D, Lq, 0) % mn_f == 0       # whole channels per launch group


def test_module_rejects_options_outside_the_hot_path():
    import hyena_dna_b200 as H
    for kw in (dict(order=1), dict(num_heads=2), dict(dropout=0.1), dict(activation="gelu"), dict(outer_mixing=True),
               dict(linear_mixer=True), dict(filter_order=16), dict(short_filter_order=4)):
        with pytest.raises(H.HyenaB200Error):
            H.HyenaOperator(8, 64, **kw)
    # accepted-and-ignored factory kwargs (long_conv_lm.py:88-95)
    op = H.HyenaOperator(8, 64, emb_dim=3, layer_idx=3, device=None, dtype=None, fused_fft_conv=True, lr=1e-3, wd=0.0)
    assert op.d_output == 8 and op.filter_fn.fused_fft_conv is True
    assert isinstance(op.filter_fn.pos_emb.z, __import__("torch").nn.Parameter)     # class default lr_pos_emb=1e-5
    op0 = H.HyenaOperator(8, 64, emb_dim=3, lr_pos_emb=0.0)
    assert "filter_fn.pos_emb.z" in dict(op0.named_buffers())
    # filter options that are built (round 2): normalized, bidirectional, trainable modulation deltas
    op1 = H.HyenaOperator(8, 64, emb_dim=3, normalized=True, bidirectional=True, modulation_lr=1e-3)
    assert op1.filter_fn.normalized and op1.filter_fn.bidirectional
    assert op1.filter_fn.modulation.deltas._optim == {"lr": 1e-3, "weight_decay": 0.0}


def test_block_glue_option_guards_and_state_dict_keys():
    """Block / Backbone mirrors: options outside the path raise at construction; the parameter names
    are the reference's (flash_attn Block: mixer / norm1 / mlp / norm2; LMBackbone: layers / ln_f)."""
    from functools import partial
    import torch
    import hyena_dna_b200 as H
    mixer = partial(H.HyenaOperator, l_max=64, emb_dim=5)
    for kw in ({"prenorm": False}, {"resid_dropout1": 0.1}, {"drop_path2": 0.2}, {"sequence_parallel": True},
               {"return_residual": True}):
        with pytest.raises(H.HyenaB200Error):
            H.Block(16, mixer_cls=mixer, **kw)
    with pytest.raises(H.HyenaB200Error):
        H.Block(16, mixer_cls=mixer, norm_cls=torch.nn.BatchNorm1d)
    m = H.Backbone(16, 2, mixer, mlp_cls=None)
    keys = set(m.state_dict().keys())
    assert {"layers.0.norm1.weight", "layers.1.norm1.bias", "ln_f.weight", "layers.0.mixer.in_proj.weight",
            "layers.1.mixer.filter_fn.implicit_filter.5.freq"} <= keys
    assert not any(".norm2." in k for k in keys)          # mlp = Identity: no second norm, as in the reference
    with pytest.raises(H.HyenaB200Error):                  # no CPU fallback
        m(torch.zeros(1, 8, 16))
