"""Live pin of the oracle and the state-dict tables against /root/reference.
Skipped where the reference is not mounted (the GPU box)."""
import pytest
import torch

from oracle import cases as C
from oracle.ref_import import reference_available

pytestmark = pytest.mark.skipif(not reference_available(), reason="/root/reference not mounted")


@pytest.mark.parametrize("name", ["head_b16", "fuse_home_b8", "img_pool_mean_b4", "img_proj_b4"])
def test_oracle_vs_live_reference(name):
    from oracle.make_golden import run_reference
    case = C.CASES_BY_NAME[name]
    r_out, r_gin, r_gp = run_reference(case)
    o_out, o_gin, o_gp = C.run_oracle(case, torch.float64)
    for a, b in zip(o_out, r_out):
        assert float((a.float() - b).abs().max()) <= 2e-5 * float(b.abs().max())
    for a, b in zip(o_gin, r_gin):
        if b is not None:
            assert float((a.float() - b).abs().max()) <= 2e-5 * float(b.abs().max())
    for k in case.used_param_keys():
        assert float((o_gp[k].float() - r_gp[k]).abs().max()) <= 2e-5 * max(float(r_gp[k].abs().max()), 1e-30)


def test_golden_files_are_current():
    """The committed golden outputs equal what the reference computes now."""
    from conftest import load_golden
    from oracle.make_golden import run_reference
    case = C.CASES_BY_NAME["head_b256"]
    r_out, _, _ = run_reference(case)
    g = load_golden(case.name)
    for a, b in zip(r_out, g["out"]):
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-6)
