"""Record the state-dict key names and shapes of the UNMODIFIED reference whole model (run in the build container only).

    python tests/golden/make_model_keys.py

Builds /root/reference/standalone_hyenadna.HyenaDNAModel at a tiny size (2 layers, d_model 16, l_max 66 -- the
reference's own `l_max = max_length + 2` convention) and writes tests/golden/ref_model_keys.json: every key of
`model.state_dict()` with its shape, plus a seed-pinned checksum per tensor.  tests/test_checkpoint_cpu.py rebuilds a
checkpoint with exactly these names (with and without the Lightning `model.` prefix and the `.mixer.layer` injection of
huggingface.py:29-44) and checks that hyena_dna_b200.checkpoint finds every mixer tensor.
"""
import json
import os
import sys

import torch

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    sys.path.insert(0, REF)
    import standalone_hyenadna as S
    torch.manual_seed(7)
    layer = dict(emb_dim=5, filter_order=64, local_order=3, l_max=66, modulate=True, w=10, lr=6e-4, wd=0.0,
                 lr_pos_emb=0.0, short_filter_order=3, order=2)
    model = S.HyenaDNAModel(d_model=16, n_layer=2, d_inner=64, vocab_size=12, layer=layer, pad_vocab_size_multiple=8)
    sd = model.state_dict()
    rec = {k: {"shape": list(v.shape), "sum": float(v.double().sum())} for k, v in sd.items()}
    with open(os.path.join(OUT, "ref_model_keys.json"), "w") as f:
        json.dump({"d_model": 16, "n_layer": 2, "l_max": 66, "emb_dim": 5, "keys": rec}, f, indent=0, sort_keys=True)
    print(len(rec), "keys;", sum(".mixer." in k for k in rec), "mixer keys")


if __name__ == "__main__":
    main()
