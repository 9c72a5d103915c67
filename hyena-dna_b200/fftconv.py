"""The reference's long-convolution op surface on top of the sm_100a library.

Mirrors /root/reference/src/ops/fftconv.py (``fftconv_func``, ``FFTConvFunc``) and the pybind
module ``fftconv`` it imports (csrc/fftconv/fftconv.cpp:238-241: ``fftconv_fwd`` / ``fftconv_bwd``),
for the fp32, gelu=False, no-dropout, head_dim=1, q=v=None case that HyenaFilter.forward uses
(hyena.py:250-259).  Other options raise.  Works for any L up to 2^20 (the reference stops at 8192).
"""
import torch

from . import ops
from ._lib import HyenaB200Error


def _reject(**flags):
    bad = [k for k, v in flags.items() if v]
    if bad:
        raise HyenaB200Error(f"fftconv options not supported by the sm_100a hot path (no fallback): {bad}")


def fftconv_fwd(u, filter, D, v, head_dim, q, dropout_mask, gelu, gelu_inp, gelu_q, fft_size,
                force_fp16_output, output_hbl_layout, fftfp16):
    """Signature of csrc/fftconv/fftconv.cpp:53-61.  ``filter`` is the packed spectrum returned by
    ``filter_spectrum`` (the reference passes rfft(k, fft_size); the packed form carries the same
    information in the kernels' own order)."""
    _reject(v=v is not None, q=q is not None, head_dim=head_dim != 1, dropout_mask=dropout_mask is not None,
            gelu=gelu, gelu_inp=gelu_inp, gelu_q=gelu_q, force_fp16_output=force_fp16_output,
            output_hbl_layout=output_hbl_layout, fftfp16=fftfp16)
    if u.stride(-1) != 1 or not u.is_contiguous():
        u = u.contiguous()
    return ops.fftconv_forward(u, filter, D.contiguous())


def fftconv_bwd(dout, u, filter, D, v, head_dim, q, dropout_mask, gelu, gelu_inp, gelu_q, fft_size,
                output_hbl_layout, fftfp16):
    """Signature of csrc/fftconv/fftconv.cpp:134-143; returns (du, dk, dD, dv, dq) with dk already in the
    time domain (the reference returns dk_f and inverts it in Python, src/ops/fftconv.py:98)."""
    _reject(v=v is not None, q=q is not None, head_dim=head_dim != 1, dropout_mask=dropout_mask is not None,
            gelu=gelu, gelu_inp=gelu_inp, gelu_q=gelu_q, output_hbl_layout=output_hbl_layout, fftfp16=fftfp16)
    du, dk, dD = ops.fftconv_backward(dout.contiguous(), u.contiguous(), filter, D.contiguous())
    return du, dk, dD, None, None


class FFTConvFunc(torch.autograd.Function):
    """src/ops/fftconv.py:58-103."""

    @staticmethod
    def forward(ctx, u, k, D, dropout_mask=None, gelu=True, force_fp16_output=False, output_hbl_layout=False,
                v=None, head_dim=1, q=None, fftfp16=False, k_rev=None):
        _reject(k_rev=k_rev is not None)
        if u.dtype != torch.float32 or k.dtype != torch.float32:
            raise HyenaB200Error("fftconv_func: fp32 inputs only")
        u = u.contiguous()
        D = D.to(torch.float32).contiguous()
        k_f = ops.filter_spectrum(k.contiguous())
        ctx.save_for_backward(u, k_f, D)
        return fftconv_fwd(u, k_f, D, v, head_dim, q, dropout_mask, gelu, False, False, 0, force_fp16_output,
                           output_hbl_layout, fftfp16)

    @staticmethod
    def backward(ctx, dout):
        u, k_f, D = ctx.saved_tensors
        du, dk, dD, _, _ = fftconv_bwd(dout, u, k_f, D, None, 1, None, None, False, False, False, 0, False, False)
        return du, dk, dD, None, None, None, None, None, None, None, None, None


def fftconv_func(u, k, D, dropout_mask=None, gelu=True, force_fp16_output=False, output_hbl_layout=False,
                 v=None, head_dim=1, q=None, fftfp16=False, k_rev=None):
    """u (B, H, L), k (H, L), D (H,) -> (B, H, L); src/ops/fftconv.py:105-108."""
    return FFTConvFunc.apply(u, k, D, dropout_mask, gelu, force_fp16_output, output_hbl_layout, v, head_dim, q,
                             fftfp16, k_rev)
