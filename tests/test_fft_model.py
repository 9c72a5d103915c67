"""CPU checks of the numpy model of the kernels' FFT algebra (oracle/fft_model.py)."""
import numpy as np
import pytest

from oracle import fft_model as fm


def _direct_conv(g, k):
    L = g.shape[-1]
    return np.convolve(g, k)[:L]


def _direct_corr(d, h):
    L = d.shape[-1]
    return np.array([np.dot(d[j:], h[: L - j]) for j in range(L)])


@pytest.mark.parametrize("M1,M2,L", [(1, 16, 16), (2, 16, 30), (4, 8, 32), (8, 8, 37), (16, 4, 64), (4, 32, 100)])
def test_conv_matches_direct(M1, M2, L):
    rng = np.random.default_rng(0)
    g = rng.standard_normal(L)
    k = rng.standard_normal(L)
    y = fm.causal_conv(g, k, M1, M2)
    assert np.allclose(y, _direct_conv(g, k), atol=1e-9)


@pytest.mark.parametrize("M1,M2,L", [(1, 16, 16), (2, 16, 30), (4, 8, 32), (8, 8, 37), (4, 32, 100)])
def test_corr_matches_direct(M1, M2, L):
    rng = np.random.default_rng(1)
    d = rng.standard_normal(L)
    h = rng.standard_normal(L)
    y = fm.causal_corr(d, h, M1, M2)
    assert np.allclose(y, _direct_corr(d, h), atol=1e-9)


def test_matches_reference_rfft_formula():
    # src/models/sequence/hyena.py:59-88 (fftconv_ref) written with numpy
    rng = np.random.default_rng(2)
    L = 48
    g = rng.standard_normal((3, L)); k = rng.standard_normal((3, L))
    n = 2 * L
    ref = np.fft.irfft(np.fft.rfft(g, n=n) * (np.fft.rfft(k, n=n) / n), n=n)[..., :L] * n
    y = fm.causal_conv(g, k, 4, 16)
    assert np.allclose(y, ref, atol=1e-9)


@pytest.mark.parametrize("R1,R2", [(32, 32), (32, 8), (8, 4), (16, 1), (1, 8)])
@pytest.mark.parametrize("inverse", [False, True])
def test_two_stage_block_fft(R1, R2, inverse):
    rng = np.random.default_rng(3)
    N = R1 * R2
    x = rng.standard_normal(N) + 1j * rng.standard_normal(N)
    ref = np.fft.ifft(x) * N if inverse else np.fft.fft(x)
    assert np.allclose(fm.block_fft_two_stage(x, R1, R2, inverse), ref, atol=1e-9)
