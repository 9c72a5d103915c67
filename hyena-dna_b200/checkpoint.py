"""Reference-checkpoint compatibility for the Hyena mixer.

A HyenaDNA checkpoint (Lightning ``weights.ckpt`` or a bare ``state_dict``) names the mixer of layer ``i``

    [model.]backbone.layers.<i>.mixer[.layer].<operator key>

* the ``model.`` prefix is what the Lightning task wrapper adds (the reference strips it with
  ``consume_prefix_in_state_dict_if_present``, ``src/models/sequence/long_conv_lm.py:586-589``, or adds it when going
  the other way, ``huggingface.py:55-57``);
* ``.layer`` is injected after ``.mixer`` when the model was trained with ``checkpoint_mixer=True``
  (``huggingface.py:29-44``: the mixer is wrapped, so its parameters live one attribute deeper);
* ``<operator key>`` are the ``HyenaOperator.state_dict()`` keys, which this package reproduces one for one
  (``hyena.py:195,203-221,350-369``; tests/test_gpu_parity.py::test_operator_matches_reference_golden loads a
  reference operator state_dict with ``strict=True``).

Pure host logic: no kernels, no device work.  A missing or mis-shaped key raises ``KeyError`` / ``ValueError`` -- the
reference raises ``Exception('key mismatch in the state dicts!')`` (``huggingface.py:62-63``) -- never a silent skip.
"""
import re

import torch

_LAYER_RE = re.compile(r"^(?:model\.)?backbone\.layers\.(\d+)\.mixer\.(?:layer\.)?(.+)$")


def _unwrap(ckpt):
    """Accept a Lightning checkpoint (``{'state_dict': ...}``) or a bare state dict."""
    if isinstance(ckpt, dict) and "state_dict" in ckpt and isinstance(ckpt["state_dict"], dict):
        return ckpt["state_dict"]
    return ckpt


def mixer_layers(ckpt):
    """Sorted layer indices that have mixer weights in ``ckpt``."""
    idx = set()
    for k in _unwrap(ckpt):
        m = _LAYER_RE.match(k)
        if m:
            idx.add(int(m.group(1)))
    return sorted(idx)


def mixer_state_dict(ckpt, layer_idx):
    """The ``HyenaOperator`` state dict of layer ``layer_idx`` cut out of a whole-model checkpoint (tensors are the
    checkpoint's own, not copies)."""
    out = {}
    for k, v in _unwrap(ckpt).items():
        m = _LAYER_RE.match(k)
        if m and int(m.group(1)) == layer_idx:
            if m.group(2) in out:
                raise ValueError(f"checkpoint names mixer key {m.group(2)!r} of layer {layer_idx} twice "
                                 "(with and without '.layer')")
            out[m.group(2)] = v
    if not out:
        raise KeyError(f"no 'backbone.layers.{layer_idx}.mixer.*' keys in the checkpoint "
                       f"(layers present: {mixer_layers(ckpt)})")
    return out


def load_mixer(op, ckpt, layer_idx):
    """Load layer ``layer_idx``'s mixer weights into ``op`` (a ``hyena_dna_b200.hyena.HyenaOperator``), strictly:
    every key of ``op.state_dict()`` must be present with the same shape.  Returns ``op``."""
    sd = mixer_state_dict(ckpt, layer_idx)
    own = op.state_dict()
    missing = [k for k in own if k not in sd]
    extra = [k for k in sd if k not in own]
    if missing or extra:
        raise KeyError(f"mixer key mismatch for layer {layer_idx}: missing {missing}, unexpected {extra}")
    for k, v in own.items():
        if tuple(sd[k].shape) != tuple(v.shape):
            raise ValueError(f"layer {layer_idx} key {k!r}: checkpoint shape {tuple(sd[k].shape)} != module shape "
                             f"{tuple(v.shape)} (d_model / l_max / emb_dim of the module must match the checkpoint)")
    op.load_state_dict({k: torch.as_tensor(v) for k, v in sd.items()}, strict=True)
    return op


def load_mixers(ops, ckpt):
    """``ops``: sequence of operators, one per layer in order.  Loads layer ``i`` into ``ops[i]``."""
    layers = mixer_layers(ckpt)
    if len(ops) > len(layers):
        raise KeyError(f"checkpoint has mixers for layers {layers}, {len(ops)} operators given")
    for i, op in enumerate(ops):
        load_mixer(op, ckpt, i)
    return ops


def export_mixer(op, layer_idx, prefix="model.", checkpointed=False):
    """Inverse of ``mixer_state_dict``: the operator's weights under the reference's whole-model key names, so that a
    model trained here can be handed back to the reference (``load_backbone`` expects the ``model.`` prefix,
    ``long_conv_lm.py:612-614``)."""
    mid = ".mixer.layer." if checkpointed else ".mixer."
    return {f"{prefix}backbone.layers.{layer_idx}{mid}{k}": v for k, v in op.state_dict().items()}
