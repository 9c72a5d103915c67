"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference HyenaOperator hot path.

This file is the parity oracle.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it; the product
package (``hyena-dna_b200/``) never does and has no CPU fallback.

Every function restates one piece of HazyResearch/hyena-dna in plain torch ops (fp32 or
fp64, CPU or any device torch runs on), written functionally over a flat parameter dict
whose keys are the reference module's ``state_dict`` keys.  Citations are relative to
/root/reference.

Pinning: the reference repo has no tests or golden vectors for this path,
so the oracle is pinned against outputs of the reference code itself, generated in the
build container by ``tests/golden/make_golden.py`` (imports standalone_hyenadna.py and
src/models/sequence/hyena.py) and committed under ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` checks the oracle against them bit-for-bit tolerance
(<= 1e-6 relative, same torch build).
"""
import math

import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------- parameters
def positional_embedding(emb_dim, seq_len):
    """z (1,L,emb_dim), t (1,L,1).  src/models/sequence/hyena.py:109-131."""
    assert emb_dim % 2 == 1 and emb_dim >= 3
    t = torch.linspace(0, 1, seq_len)[None, :, None]
    bands = (emb_dim - 1) // 2
    pos = torch.linspace(0, seq_len - 1, seq_len)[None, :, None]
    w = 2 * math.pi * pos / seq_len
    f = torch.linspace(1e-4, bands - 1, bands)[None, None]
    zc = torch.exp(-1j * f * w)
    return torch.cat([t, zc.real, zc.imag], dim=-1), t


def modulation_deltas(d_model, fast_decay_pct=0.3, slow_decay_pct=1.5, target=1e-2):
    """deltas (1,1,D).  src/models/sequence/hyena.py:134-150."""
    hi = math.log(target) / fast_decay_pct
    lo = math.log(target) / slow_decay_pct
    return torch.linspace(lo, hi, d_model)[None, None]


def init_params(d_model, l_max, order=2, filter_order=64, emb_dim=3, w=1.0, short_filter_order=3,
                generator=None, init_std=None, n_layer=8, dtype=torch.float32):
    """Random parameters with the reference's state_dict keys/shapes.

    init_std=None: unit-ish init (nn.Linear-like uniform); init_std=0.02 mimics
    ``_init_weights`` (standalone_hyenadna.py:612-641): Linear weights N(0, std), biases 0,
    out_proj.weight N(0, std/sqrt(2 n_layer)).
    """
    assert order >= 2
    D, N, E = d_model, filter_order, emb_dim
    g = generator

    def lin(o, i, bias=True):
        if init_std is None:
            bound = 1.0 / math.sqrt(i)
            W = (torch.rand(o, i, generator=g) * 2 - 1) * bound
            b = (torch.rand(o, generator=g) * 2 - 1) * bound if bias else None
        else:
            W = torch.randn(o, i, generator=g) * init_std
            b = torch.zeros(o) if bias else None
        return W, b

    P = {}
    P["in_proj.weight"], P["in_proj.bias"] = lin((order + 1) * D, D)
    P["out_proj.weight"], P["out_proj.bias"] = lin(D, D)
    if init_std is not None:
        P["out_proj.weight"] = torch.randn(D, D, generator=g) * init_std / math.sqrt(2 * n_layer)
    bound = 1.0 / math.sqrt(short_filter_order)
    P["short_filter.weight"] = (torch.rand((order + 1) * D, 1, short_filter_order, generator=g) * 2 - 1) * bound
    P["short_filter.bias"] = (torch.rand((order + 1) * D, generator=g) * 2 - 1) * bound
    P["filter_fn.bias"] = torch.randn(D * (order - 1), generator=g)
    z, t = positional_embedding(E, l_max)
    P["filter_fn.pos_emb.z"], P["filter_fn.pos_emb.t"] = z, t
    P["filter_fn.implicit_filter.0.weight"], P["filter_fn.implicit_filter.0.bias"] = lin(N, E)
    P["filter_fn.implicit_filter.2.weight"], P["filter_fn.implicit_filter.2.bias"] = lin(N, N)
    P["filter_fn.implicit_filter.4.weight"], P["filter_fn.implicit_filter.4.bias"] = lin(N, N)
    P["filter_fn.implicit_filter.6.weight"], _ = lin(D * (order - 1), N, bias=False)
    P["filter_fn.implicit_filter.1.freq"] = w * torch.ones(1, N)
    P["filter_fn.modulation.deltas"] = modulation_deltas(D * (order - 1))
    return {k: v.to(dtype) for k, v in P.items()}


def canonical(P):
    """Accept a reference state_dict (three aliased .freq keys) and return our flat dict."""
    Q = dict(P)
    for k in ("filter_fn.implicit_filter.3.freq", "filter_fn.implicit_filter.5.freq"):
        Q.pop(k, None)
    return Q


# ----------------------------------------------------------------------------- filter
def implicit_filter(z, P):
    """Sin-MLP: Linear -> sin(freq * .) three times, then Linear(no bias).

    src/models/sequence/hyena.py:96-106 (Sin, ONE freq tensor shared by all three
    activations) and :199-215 (layer stack)."""
    freq = P["filter_fn.implicit_filter.1.freq"]
    h = z
    for i in (0, 2, 4):
        h = F.linear(h, P[f"filter_fn.implicit_filter.{i}.weight"], P[f"filter_fn.implicit_filter.{i}.bias"])
        h = torch.sin(freq * h)
    return F.linear(h, P["filter_fn.implicit_filter.6.weight"])


def hyena_filter(L, P, shift=0.0, modulate=True, normalized=False):
    """k (1,L,D).  src/models/sequence/hyena.py:229-238 with :152-155 modulation and the optional L1 normalisation over
    the channel dim (:235-236)."""
    z = P["filter_fn.pos_emb.z"][:, :L]
    t = P["filter_fn.pos_emb.t"][:, :L]
    h = implicit_filter(z, P)
    if modulate:
        h = h * (torch.exp(-t * P["filter_fn.modulation.deltas"].abs()) + shift)
    if normalized:
        h = h / torch.norm(h, dim=-1, p=1, keepdim=True)
    return h


# ----------------------------------------------------------------------------- fftconv
def fftconv_ref(u, k, D, k_rev=None, bidirectional=False):
    """y = irfft(rfft(u, 2L) * rfft(k, 2L)/2L, norm='forward')[:L] + u * D[:, None].

    src/models/sequence/hyena.py:59-88 with gelu=False, dropout_mask=None (the HyenaFilter.forward call at :261);
    without k_rev / bidirectional identical to src/ops/fftconv.py:15-34 and standalone_hyenadna.py:45-60.
    k_rev (:63-65): second filter whose conjugated spectrum is added (an anticausal half).  bidirectional (:67-73): the input
    is padded by ~L/2 on both sides (to exactly the 2L transform points) before the cyclic product.
    u (..., H, L), k (H, L), D (H,).
    """
    L = u.shape[-1]
    n = 2 * L
    k_f = torch.fft.rfft(k, n=n) / n
    if k_rev is not None:
        k_f = k_f + (torch.fft.rfft(k_rev, n=n) / n).conj()
    if bidirectional:
        padded_length = L + 2 * (L // 2)
        pad_before = padded_length // 2 - (L // 2)
        pad_after = padded_length - L - pad_before
        u_f = torch.fft.rfft(F.pad(u.to(k.dtype), (pad_before, pad_after)), n=n)
    else:
        u_f = torch.fft.rfft(u.to(k.dtype), n=n)
    y = torch.fft.irfft(u_f * k_f, n=n, norm="forward")[..., :L]
    return (y + u * D.unsqueeze(-1)).to(u.dtype)


def fftconv_variants_time_domain(u, k, D, k_rev=None, bidirectional=False):
    """The same two options stated in the time domain -- the decomposition hyena_dna_b200.fftconv uses to run them on the
    causal-convolution / correlation kernels (tests/test_variants_model_cpu.py checks it against fftconv_ref above):
      k_rev:          y[t] += sum_{s >= t} u[s] k_rev[s - t]                                   (a correlation)
      bidirectional:  y[t]  = c[t - pad_before] (causal conv c delayed, zeros shifted in)
                              + sum_m u[t + m] r[m],  r[m] = k[2L - pad_before - m] for L - pad_before < m < L   (the wrapped taps)
    O(L^2); float64, small L only."""
    L = u.shape[-1]
    conv = torch.zeros_like(u)
    for j in range(L):
        conv[..., j:] += k[..., j, None] * u[..., : L - j]

    def corr(x, f):
        out = torch.zeros_like(x)
        for m in range(L):
            out[..., : L - m] += f[..., m, None] * x[..., m:]
        return out
    if bidirectional:
        assert k_rev is None
        pad = (L + 2 * (L // 2)) // 2 - L // 2
        y = F.pad(conv[..., : L - pad], (pad, 0)) if pad < L else torch.zeros_like(conv)
        r = torch.zeros_like(k)
        if pad > 1:
            r[..., L - pad + 1:] = k[..., L - pad + 1:].flip(-1)
        y = y + corr(u, r)
    else:
        y = conv
        if k_rev is not None:
            y = y + corr(u, k_rev)
    return y + u * D.unsqueeze(-1)


def fftconv_direct(u, k, D):
    """O(L^2) time-domain statement of the same thing (small L only; float64 advised)."""
    L = u.shape[-1]
    out = torch.zeros_like(u)
    for j in range(L):
        out[..., j:] += k[..., j, None] * u[..., : L - j]
    return out + u * D.unsqueeze(-1)


# ----------------------------------------------------------------------------- operator
def short_filter(p, W, b, L):
    """Depthwise Conv1d(k=3, padding=2, groups=C)(p)[..., :L]  (hyena.py:363-369, :394).

    Equals s[t] = W[c,0] p[t-2] + W[c,1] p[t-1] + W[c,2] p[t] + b[c] with p[<0] = 0."""
    C = p.shape[1]
    return F.conv1d(p, W, b, padding=W.shape[-1] - 1, groups=C)[..., :L]


def hyena_operator(u, P, shift=0.0, modulate=True, return_intermediates=False, normalized=False):
    """HyenaOperator.forward for any order >= 2 (heads=1, blocks=1, activation=id, dropout=0); the order is read off
    in_proj.weight ((order+1)*D rows).

    src/models/sequence/hyena.py:388-444 (== standalone_hyenadna.py:273-293).  u (B, L, D) -> (B, L, D)."""
    B, L, D = u.shape
    order = P["in_proj.weight"].shape[0] // D - 1
    p = F.linear(u, P["in_proj.weight"], P["in_proj.bias"]).transpose(1, 2)       # :391-392
    uc = short_filter(p, P["short_filter.weight"], P["short_filter.bias"], L)      # :394
    *x, v = uc.split(D, dim=1)                                                     # :404
    # filter channels are ordered (v o): channel = v * (order-1) + o                 :405-412
    k = hyena_filter(L, P, shift, modulate, normalized)[0].transpose(0, 1).reshape(D, order - 1, L)
    bias = P["filter_fn.bias"].reshape(D, order - 1)
    g = c = None
    for o, x_i in enumerate(reversed(x[1:])):                                      # :414-423
        g = v * x_i                                                                # :420
        v = c = fftconv_ref(g, k[:, o], bias[:, o])                                # :423
    y_pre = (v * x[0]).transpose(1, 2)                                             # :432-439
    y = F.linear(y_pre, P["out_proj.weight"], P["out_proj.bias"])                  # :440
    if return_intermediates:
        return y, dict(p=p, uc=uc, k=k[:, -1], g=g, c=c, y_pre=y_pre)
    return y


def operator_fwd_bwd(u, P, dy, shift=0.0, grads_for=None, normalized=False):
    """Forward + autograd backward; returns y, du and a dict of parameter grads (grads_for: names, e.g. to include
    "filter_fn.modulation.deltas" when modulation_lr != 0)."""
    names = grads_for or [k for k in P if k not in ("filter_fn.pos_emb.z", "filter_fn.pos_emb.t",
                                                    "filter_fn.modulation.deltas")]
    Q = {k: (v.detach().clone().requires_grad_(k in names)) for k, v in P.items()}
    u = u.detach().clone().requires_grad_(True)
    y = hyena_operator(u, Q, shift, normalized=normalized)
    y.backward(dy)
    return y.detach(), u.grad.detach(), {k: Q[k].grad.detach() for k in names}


def to_dtype(P, dtype):
    return {k: v.to(dtype) for k, v in P.items()}


# ----------------------------------------------------------------------------- synthetic inputs
def nucleotide_activations(B, L, D, seed=2222):
    """ids ~ U{7,8,9,10} (A,C,G,T; hg38_char_tokenizer.py:58-67), embedding
    N(0, 0.02^2) of a 16-row table, LayerNorm -> unit-scale rows drawn from 4 vectors."""
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(7, 11, (B, L), generator=g)
    table = torch.randn(16, D, generator=g) * 0.02
    return F.layer_norm(table[ids], (D,)), ids


# ----------------------------------------------------------------------------- the reference's FFTConvFunc protocol
def reference_fftconv_protocol(fftconv_fwd, fftconv_bwd, u, k, D, dout):
    """Drive an extension exposing ``fftconv_fwd`` / ``fftconv_bwd`` exactly the way the reference's autograd wrapper
    does (src/ops/fftconv.py:61-103, gelu=False, no dropout / v / q / k_rev): fft_size from the sequence length (:64),
    ``k_f = rfft(k, n=fft_size)`` built with torch (:65), ``dk = irfft(dk_f, n=fft_size, norm='forward')[..., :L]``
    (:98).  Returns out, du, dk, dD.  Test infrastructure: lets the extension-level ABI be checked without importing
    /root/reference on the GPU box."""
    seqlen = u.shape[-1]
    fft_size = max(2 * 2 ** int(math.ceil(math.log2(seqlen))), 16)
    k_f = torch.fft.rfft(k, n=fft_size).contiguous()
    out = fftconv_fwd(u, k_f, D.contiguous(), None, 1, None, None, False, False, False, fft_size, False, False, False)
    du, dk_f, dD, _, _ = fftconv_bwd(dout.contiguous(), u, k_f, D.contiguous(), None, 1, None, None, False, False, False,
                                     fft_size, False, False)
    dk = torch.fft.irfft(dk_f, n=fft_size, norm='forward')[..., :seqlen]
    return out, du, dk, dD


# ----------------------------------------------------------------------------- block glue
def prenorm_backbone(x, sd, n_layer, shift=0.0, eps=1e-5, residual_in_fp32=True):
    """Stack of pre-norm Blocks + final norm, restated from flash-attention/flash_attn/modules/block.py:111-148 (prenorm
    branch, dropout p = 0, residual_in_fp32) and src/models/sequence/long_conv_lm.py:383-396:
        residual = hidden + residual;  hidden = mixer(LN1(residual));
        [if the block has an MLP]  residual = hidden + residual;  hidden = fc2(gelu_tanh(fc1(LN2(residual))))
        out = ln_f(hidden + residual)
    sd: state_dict with the reference's keys (layers.N.{mixer.*, norm1.*, mlp.fc1/fc2.*, norm2.*}, ln_f.*)."""
    D = x.shape[-1]
    hidden, residual = x, None
    for i in range(n_layer):
        pre = f"layers.{i}."
        residual = hidden + residual if residual is not None else hidden
        h = F.layer_norm(residual, (D,), sd[pre + "norm1.weight"], sd[pre + "norm1.bias"], eps)
        if residual_in_fp32:                       # block.py:118-119: the residual STREAM is rounded to fp32 (a no-op in fp32 runs)
            residual = residual.to(torch.float32)
        P = canonical({k[len(pre + "mixer."):]: v for k, v in sd.items() if k.startswith(pre + "mixer.")})
        hidden = hyena_operator(h, P, shift)
        if pre + "mlp.fc1.weight" in sd:
            residual = hidden + residual
            h = F.layer_norm(residual, (D,), sd[pre + "norm2.weight"], sd[pre + "norm2.bias"], eps)
            if residual_in_fp32:
                residual = residual.to(torch.float32)
            h = F.gelu(F.linear(h, sd[pre + "mlp.fc1.weight"], sd[pre + "mlp.fc1.bias"]), approximate="tanh")
            hidden = F.linear(h, sd[pre + "mlp.fc2.weight"], sd[pre + "mlp.fc2.bias"])
    return F.layer_norm(hidden + residual, (D,), sd["ln_f.weight"], sd["ln_f.bias"], eps)
