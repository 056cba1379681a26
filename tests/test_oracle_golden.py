"""The CPU oracle against the committed golden vectors (made from the real
reference by oracle/make_golden.py).  Runs anywhere; no /root/reference needed."""
import pytest
import torch

from conftest import load_golden
from oracle import cases as C

FP32_RTOL = 1e-4   # BASELINE.json north_star: fp32 within rtol 1e-4 of the reference


@pytest.mark.parametrize("case", C.CASES, ids=[c.name for c in C.CASES])
def test_oracle_matches_reference_golden(case):
    g = load_golden(case.name)
    assert g["meta"]["seed"] == case.seed and g["meta"]["B"] == case.B
    outs, gin, gp = C.run_oracle(case, torch.float64)
    for o, ref in zip(outs, g["out"]):
        scale = float(ref.abs().max())
        assert float((o.float() - ref).abs().max()) <= FP32_RTOL * scale
    for gi, fp in zip(gin, g["grad_in"]):
        if fp is not None:
            C.check_fingerprint(gi, fp, FP32_RTOL, f"{case.name} grad_in")
    used = set(case.used_param_keys())
    for key, fp in g["grad_param"].items():
        if key in used:
            C.check_fingerprint(gp[key], fp, FP32_RTOL, f"{case.name} d{key}")
        else:
            assert fp is None and gp[key] is None      # unused HoME params keep grad None


def test_oracle_float32_close_to_float64():
    case = C.CASES_BY_NAME["fuse_b8"]
    o64, _, _ = C.run_oracle(case, torch.float64)
    o32, _, _ = C.run_oracle(case, torch.float32)
    for a, b in zip(o32, o64):
        assert float((a.double() - b).abs().max()) <= 1e-5 * float(b.abs().max())


def test_fully_masked_row_is_nan_like_the_reference():
    """a fully padded row yields NaN (not 'fixed')."""
    from oracle import mmoe_oracle as O
    from oracle import synth
    sd = synth.fill_state_dict(synth.cross_expert_shapes(), 5)
    u, um, i, im = synth.cross_inputs(5, 2)
    im = im.clone()
    im[1] = True
    out = O.cross_expert(sd, u, um, i, im)
    assert torch.isfinite(out[0]).all() and torch.isnan(out[1]).all()


def test_gate_weights_sum_to_one_and_argmax_stable():
    from oracle import mmoe_oracle as O
    case = C.CASES_BY_NAME["head_b256"]
    sd = case.state_dict()
    (ev,) = case.inputs()
    lg, lb, wg, wb = O.two_task_mmoe(sd, ev, return_gates=True)
    assert torch.allclose(wg.sum(-1), torch.ones(case.B), atol=1e-6)
    lg64, lb64, wg64, wb64 = O.two_task_mmoe({k: v.double() for k, v in sd.items()}, ev.double(), return_gates=True)
    margin = wg64.topk(2, -1).values
    safe = (margin[:, 0] - margin[:, 1]) > 1e-5
    assert (wg.argmax(-1)[safe] == wg64.argmax(-1)[safe]).all()
