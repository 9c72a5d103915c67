"""CPU checks of bench.py's pure logic (roofline assembly, algorithmic byte counts, argument defaults)."""
import importlib.util
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_span_bytes_match_survey_formula():
    b = _bench()
    r = b.build_roofline({"row_pass<conv_bwd>": (40.0, 10), "col_fwd<gate>": (12.5, 10)}, 10, 1, 256, 1 << 20, 39.0)
    assert r["algorithmic_bytes_per_step"] == (44 + 16) * 256 * (1 << 20)          # 15,360 B/nt
    assert abs(r["span_ms_per_step"] - 5.25) < 1e-9
    assert r["dominant_kernel"]["name"] == "row_pass<conv_bwd>"
    assert abs(sum(k["share_of_span"] for k in r["kernels"].values()) - 1.0) < 1e-3
    k = r["kernels"]["col_fwd<gate>"]
    assert k["algorithmic_bytes_per_step"] == 8 * 256 * (1 << 20)
    assert abs(k["achieved_gbs"] - 8 * 256 * (1 << 20) / 1.25e-3 / 1e9) < 0.1
    assert 0 < r["frac"] == round(r["achieved"] / r["peak"], 4)
    json.dumps(r)                                                                   # serialisable


def test_roofline_handles_unknown_and_empty_profiles():
    b = _bench()
    r = b.build_roofline({}, 5, 2, 128, 4096, 1.0)
    assert r["dominant_kernel"] is None and r["achieved"] == 0.0 and r["traffic"] is None
    r = b.build_roofline({"twiddle_init": (0.01, 1)}, 1, 1, 8, 1024, 1.0)
    assert "achieved_gbs" not in r["kernels"]["twiddle_init"]


def test_synthetic_inputs_match_the_oracle_recipe():
    import torch
    from oracle import hyena_oracle as O
    b = _bench()
    a = b.nucleotide_activations(2, 100, 16, seed=7)
    ref, ids = O.nucleotide_activations(2, 100, 16, seed=7)
    assert torch.equal(a, ref) and int(ids.min()) >= 7 and int(ids.max()) <= 10
