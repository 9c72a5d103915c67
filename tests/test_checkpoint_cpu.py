"""Reference-checkpoint key mapping, host logic only: the key names and shapes come from the
unmodified reference whole model (tests/golden/ref_model_keys.json, written by tests/golden/make_model_keys.py)."""
import json
import os
from importlib import import_module

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ck = import_module("hyena_dna_b200.checkpoint")
HyenaOperator = import_module("hyena_dna_b200.hyena").HyenaOperator


def _ref_checkpoint(prefix="", checkpointed=False, seed=0):
    """A whole-model state dict with the reference's key names/shapes and seed-pinned random values."""
    meta = json.load(open(os.path.join(HERE, "golden", "ref_model_keys.json")))
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, rec in meta["keys"].items():
        if checkpointed:
            k = k.replace(".mixer.", ".mixer.layer.").replace(".mlp.", ".mlp.layer.")   # huggingface.py:29-44
        sd[prefix + k] = torch.randn(rec["shape"], generator=g)
    for k in list(sd):                                          # one shared `freq` tensor behind three keys per mixer
        if k.endswith(".1.freq"):                               # (hyena.py:199-215): a real checkpoint stores it thrice
            sd[k[:-6] + "3.freq"] = sd[k[:-6] + "5.freq"] = sd[k]
    return meta, sd


def _op(meta):
    return HyenaOperator(d_model=meta["d_model"], l_max=meta["l_max"], order=2, filter_order=64, emb_dim=meta["emb_dim"])


@pytest.mark.parametrize("prefix", ["", "model."])
@pytest.mark.parametrize("checkpointed", [False, True])
@pytest.mark.parametrize("wrapped", [False, True])
def test_every_reference_mixer_key_is_found_and_loaded(prefix, checkpointed, wrapped):
    meta, sd = _ref_checkpoint(prefix, checkpointed)
    ckpt = {"state_dict": sd, "epoch": 3} if wrapped else sd
    assert ck.mixer_layers(ckpt) == list(range(meta["n_layer"]))
    ops = ck.load_mixers([_op(meta) for _ in range(meta["n_layer"])], ckpt)
    mid = ".mixer.layer." if checkpointed else ".mixer."
    n = 0
    for i, op in enumerate(ops):
        own = op.state_dict()
        for k, v in own.items():
            assert torch.equal(v, sd[f"{prefix}backbone.layers.{i}{mid}{k}"]), k      # bit-exact, right layer
            n += 1
    assert n == sum(".mixer." in k for k in meta["keys"])                            # nothing in the checkpoint unused
    # the three Sin modules share ONE frequency tensor in the reference (hyena.py:199-215): after loading, the module
    # still has a single parameter behind the three keys
    f = ops[0].filter_fn.implicit_filter
    assert f[1].freq is f[3].freq is f[5].freq


def test_export_round_trip_matches_reference_names():
    meta, sd = _ref_checkpoint("model.")
    op = ck.load_mixer(_op(meta), sd, 1)
    out = ck.export_mixer(op, 1)
    ref_keys = {k for k in sd if ".layers.1.mixer." in k}
    assert set(out) == ref_keys
    assert all(torch.equal(out[k], sd[k]) for k in ref_keys)
    assert set(ck.export_mixer(op, 1, prefix="", checkpointed=True)) == \
        {k[len("model."):].replace(".mixer.", ".mixer.layer.") for k in ref_keys}


def test_mismatches_raise():
    meta, sd = _ref_checkpoint()
    with pytest.raises(KeyError):                               # no such layer
        ck.mixer_state_dict(sd, 7)
    bad = dict(sd); del bad["backbone.layers.0.mixer.filter_fn.bias"]
    with pytest.raises(KeyError):                               # missing key: never a silent skip
        ck.load_mixer(_op(meta), bad, 0)
    with pytest.raises(ValueError):                             # module built at another width
        ck.load_mixer(HyenaOperator(d_model=8, l_max=meta["l_max"], order=2, filter_order=64, emb_dim=5), sd, 0)
    both = dict(sd); both["backbone.layers.0.mixer.layer.filter_fn.bias"] = sd["backbone.layers.0.mixer.filter_fn.bias"]
    with pytest.raises(ValueError):                             # ambiguous: key present with and without '.layer'
        ck.mixer_state_dict(both, 0)
    with pytest.raises(KeyError):
        ck.load_mixers([_op(meta) for _ in range(3)], sd)       # more operators than layers
