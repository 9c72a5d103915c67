"""Activation-checkpoint-aware driver for a stack of Hyena layers.

The reference trains its 1M-token models with one `torch.utils.checkpoint` region per mixer
(src/models/sequence/long_conv_lm.py:39-45 `checkpoint_mixer`, :196-199; README.md:433-441): only the layer inputs
survive the forward pass and every layer's forward runs a second time during backward.  For the Hyena operator that
recompute would regenerate the implicit filter and its spectrum -- at batch 1 a third of the forward's custom-kernel time
-- although neither depends on the activations.  `CheckpointedHyenaStack` turns on the operators' filter cache
(HyenaFilter.cache_filter: reuse while no parameter version changed) so the recompute skips those kernels, and
`memory_plan` states what stays resident per layer at a given (batch, L, d_model).
"""
import torch
import torch.nn as nn
from torch.utils.checkpoint import checkpoint

from .hyena import HyenaOperator


def enable_filter_cache(module, on=True):
    """Switch the filter / spectrum cache of every HyenaOperator inside `module`."""
    n = 0
    for m in module.modules():
        if isinstance(m, HyenaOperator):
            m.filter_fn.cache_filter = bool(on)
            if not on:
                m.filter_fn._filter_cache = None
                m._kspec_cache = None
            n += 1
    return n


def memory_plan(batch, seqlen, d_model, n_layer, order=2, checkpointed=True, cache_filter=True):
    """Bytes resident on the device for the mixers of an n_layer stack (fp32), by item."""
    a = 4 * batch * seqlen * d_model                      # one (B, L, D) activation
    M = 1 << max(10, (seqlen - 1).bit_length())
    filt = 4 * d_model * (order - 1) * seqlen + 8 * d_model * (order - 1) * M      # k + packed spectrum
    # saved by one operator's autograd node: u, p (order+1 blocks), c, g spectrum, y_pre
    saved = a + (order + 1) * a + a + 8 * batch * d_model * M + a
    plan = {"layer_inputs": n_layer * a if checkpointed else 0,
            "saved_activations": saved if checkpointed else n_layer * saved,
            "filter_cache": n_layer * filt if cache_filter else filt,
            "fft_scratch": 8 * batch * d_model * M * 3}
    plan["total"] = sum(plan.values())
    return plan


class CheckpointedHyenaStack(nn.Module):
    """layers: modules mapping (B, L, D) -> (B, L, D) (HyenaOperator or blocks containing one).  Each layer runs inside
    its own checkpoint region (non-reentrant) with a residual connection when `residual` is set."""

    def __init__(self, layers, residual=True, use_checkpoint=True, cache_filter=True):
        super().__init__()
        self.layers = nn.ModuleList(layers)
        self.residual, self.use_checkpoint = residual, use_checkpoint
        enable_filter_cache(self, cache_filter)

    def _run(self, layer, x):
        y = layer(x)
        y = y[0] if isinstance(y, tuple) else y
        return x + y if self.residual else y

    def forward(self, x):
        for layer in self.layers:
            if self.use_checkpoint and torch.is_grad_enabled():
                x = checkpoint(self._run, layer, x, use_reentrant=False)
            else:
                x = self._run(layer, x)
        return x
