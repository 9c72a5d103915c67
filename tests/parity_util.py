"""Tolerance policy of the parity tests (BASELINE.json north_star + SURVEY.md S8(c)), with bookkeeping.

Primary check, elementwise:            |ours - ref32| <= 1e-5 + 1e-3 * |ref32|
Escape hatch (S8(c), needed because the reference's own fp32 path misses the elementwise bound on
cancellation-dominated elements of million-point FFTs): if the primary check fails, accept iff the
normwise relative error is <= 1e-5 AND
  * without an fp64 truth: the failing fraction is <= 1e-5 of the elements;
  * with an fp64 truth: max|ours - fp64| <= 2 * max|ref32 - fp64|, and the fraction of elements where OURS misses the
    elementwise bound against the TRUTH is at most 1e-5 + 2x the fraction where the REFERENCE's own fp32 result misses
    it (matched numerics: at L = 2^20 with |y| up to ~2e2 the reference itself misses 1e-5 + 1e-3|y| on ~1e-4 of the
    elements -- the survey's 1e-5 was probed at unit output scale).

Every call records how many elements needed the hatch; `report()` prints the table at session end
(tests/conftest.py), so a run shows how much of the parity rests on it.

Parameter gradients are sums over up to 2^20 positions of products of unit-scale quantities: their natural
absolute scale is max|ref|, not 1, so their absolute term is 1e-5 * max(1, max|ref|);
an element that still misses is accepted only if an fp64 truth exists and ours is no further from it than
twice the reference's own fp32 result (the reference-relative criterion of S8(c)).
"""
import torch

RTOL, ATOL = 1e-3, 1e-5
HATCH_FRAC, HATCH_NORM = 1e-5, 1e-5

_records = []


def _d(x):
    return x.detach().double().cpu()


def check(got, ref32, what, ref64=None, rtol=RTOL, atol=ATOL, param_grad=False):
    got, ref32 = _d(got), _d(ref32)
    assert got.shape == ref32.shape, f"{what}: shape {tuple(got.shape)} vs {tuple(ref32.shape)}"
    assert torch.isfinite(got).all(), f"{what}: non-finite values"
    scale = max(float(ref32.abs().max()), 1e-30)
    a = atol * max(1.0, scale) if param_grad else atol
    err = (got - ref32).abs()
    bad = err > a + rtol * ref32.abs()
    nbad = int(bad.sum())
    frac = nbad / max(bad.numel(), 1)
    nrm = float(err.norm() / max(float(ref32.norm()), 1e-30))
    rec = {"what": what, "n": bad.numel(), "bad": nbad, "frac": frac, "max_err": float(err.max()), "normwise": nrm,
           "scale": scale, "hatch": False, "e_ours64": None, "e_ref64": None, "frac_ours64": None, "frac_ref64": None}
    ok = nbad == 0
    if ref64 is not None:
        ref64 = _d(ref64)
        rec["e_ours64"] = float((got - ref64).abs().max())
        rec["e_ref64"] = float((ref32 - ref64).abs().max())
        tol64 = a + rtol * ref64.abs()
        rec["frac_ours64"] = float(((got - ref64).abs() > tol64).double().mean())
        rec["frac_ref64"] = float(((ref32 - ref64).abs() > tol64).double().mean())
    if not ok:
        rec["hatch"] = True
        vs_truth = True
        if ref64 is not None:
            vs_truth = rec["e_ours64"] <= 2.0 * rec["e_ref64"] + 1e-30
        if param_grad:
            # small tensors: a single element is more than 1e-5 of it; only the reference-relative criterion applies
            ok = ref64 is not None and vs_truth
        else:
            if ref64 is not None:
                ok = nrm <= HATCH_NORM and vs_truth and rec["frac_ours64"] <= HATCH_FRAC + 2.0 * rec["frac_ref64"]
            else:
                ok = frac <= HATCH_FRAC and nrm <= HATCH_NORM
    _records.append(rec)
    assert ok, (f"{what}: {nbad} / {bad.numel()} elements out of tolerance (frac {frac:.2e}), max err "
                f"{rec['max_err']:.3e} (scale {scale:.3e}), normwise rel {nrm:.3e}, "
                f"err vs fp64 ours {rec['e_ours64']} / ref32 {rec['e_ref64']}, fraction missing the bound vs fp64: "
                f"ours {rec['frac_ours64']} / ref32 {rec['frac_ref64']}")
    return rec


def report(write=print):
    if not _records:
        return
    used = [r for r in _records if r["hatch"]]
    write(f"\nparity bookkeeping: {len(_records)} tensor comparisons, {len(used)} needed the S8(c) escape hatch")
    for r in used:
        write(f"  hatch: {r['what']}: {r['bad']}/{r['n']} elements (frac {r['frac']:.2e}), max err {r['max_err']:.2e}, "
              f"normwise {r['normwise']:.2e}, |ours-fp64| {r['e_ours64']}, |ref32-fp64| {r['e_ref64']}, "
              f"miss-fraction vs fp64 ours {r['frac_ours64']} / ref32 {r['frac_ref64']}")
    worst = max(_records, key=lambda r: r["normwise"])
    write(f"  worst normwise error: {worst['what']}: {worst['normwise']:.2e}")
