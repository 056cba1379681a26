"""Train-mode arithmetic — the mode the benchmark times — against the oracle.

The CUDA kernels never store dropout masks: every site regenerates its keep-mask from a keyed counter hash of the flat
element index.  The library exports the key derivation (mmoe_site_keys) and the mask itself (mmoe_dropout_mask); the
tests feed exactly those masks to the oracle through its `drop=` hook (oracle/mmoe_oracle.py `_drop`), so outputs AND
gradients of a `.train()` forward/backward are compared, site by site placement and 1/(1-p) scaling included:
attention probabilities, dropout1/dropout2, the FFN dropout folded into the ReLU bit mask, the pooling weights, the MLP
and tower dropouts.  Tolerances as in test_gpu_parity.py (1e-4 fp32, 2e-2 16-bit).
"""
import pytest
import torch

import parity_util as PU
from oracle import cases as C

pytestmark = pytest.mark.gpu

TRAIN_CASES = [
    C.CASES_BY_NAME["cross_b3"], C.CASES_BY_NAME["cross_home_b3"], C.CASES_BY_NAME["fuse_b8"], C.CASES_BY_NAME["fuse_home_b8"],
    C.CASES_BY_NAME["home_head_b8"], C.CASES_BY_NAME["img_pool_mean_b4"],
    C.Case("head_drop_b16", "head", 16, 13, dict(tower_dropout=0.1)),
]
EXPECTED_SITES = {"cross_b3": 2 * 2 * 4 + 4, "cross_home_b3": 2 * 2 * 4 + 2, "fuse_b8": 2 * 4 + 1, "fuse_home_b8": 2 * 4,
                  "home_head_b8": 8 + 2, "img_pool_mean_b4": 1, "head_drop_b16": 4}


def _assert_ok(errs, tol, what):
    bad = {k: v for k, v in errs.items() if not (v <= tol)}
    assert not bad, f"{what}: {len(bad)} of {len(errs)} tensors above {tol}: {sorted(bad.items(), key=lambda kv: -kv[1])[:6]}"


@pytest.mark.parametrize("mode", ["fp32", "bf16", "fp16"])
@pytest.mark.parametrize("case", TRAIN_CASES, ids=[c.name for c in TRAIN_CASES])
def test_train_mode_matches_oracle_with_the_same_masks(case, mode):
    mod = PU.build_module(case).train()
    stats = {}
    errs = PU.compare_with_oracle(case, mode, module=mod, stats=stats, device="cuda", train_seed=1000 + case.seed)
    _assert_ok(errs, PU.TOL[mode], f"{case.name} train {mode}")
    sites = stats["drop_sites"]
    # every dropout call site of the reference module was exercised, each with a keep rate of 1 - p
    assert len({s for s, _, _ in sites}) == EXPECTED_SITES[case.name], sorted({s for s, _, _ in sites})
    for s, keep, n in sites:
        assert abs(keep - 0.9) < 5.0 * (0.09 / n) ** 0.5 + 1e-3, (s, keep, n)


def test_train_mode_differs_from_eval_and_is_seed_dependent():
    """Guards the test above against passing vacuously (masks all-ones, or the module ignoring .train())."""
    case = C.CASES_BY_NAME["fuse_b8"]
    mod = PU.build_module(case)
    out_eval, _, _ = PU.run_cuda(case, "fp32", mod)
    mod.train()
    a, _, _ = PU.run_cuda(case, "fp32", mod, seed=1)
    b, _, _ = PU.run_cuda(case, "fp32", mod, seed=1)
    c, _, _ = PU.run_cuda(case, "fp32", mod, seed=2)
    assert torch.equal(a[0], b[0]) and not torch.equal(a[0], c[0])
    assert PU.nerr(a[0], out_eval[0]) > 1e-2


def test_fp16_inf_reaches_the_gradients():
    """fp16 autocast + GradScaler (train.py:186,241,277-286): an overflowing scaled loss must surface as inf/NaN in the
    parameter gradients so that scaler.step() skips the update — the kernels must not mask it."""
    for name in ("fuse_b8", "cross_b3", "head_b16", "home_head_b8"):
        case = C.CASES_BY_NAME[name]
        mod = PU.build_module(case)
        raw, meta = case.inputs(), case.inputs_meta()
        ins = [t.cuda().clone().requires_grad_(True) if f else t.cuda() for t, f in zip(raw, meta)]
        with torch.autocast("cuda", dtype=torch.float16):
            out = mod(*ins)
        outs = out if isinstance(out, (tuple, list)) else (out,)
        loss = sum(o.float().sum() for o in outs)
        scaler = torch.amp.GradScaler("cuda", init_scale=2.0 ** 24)
        params = [p for p in mod.parameters() if p.requires_grad]
        opt = torch.optim.SGD(params, lr=0.1)
        before = [p.detach().clone() for p in params]
        # an upstream gradient of +inf, as an overflowed fp16 loss scale produces
        torch.autograd.backward([scaler.scale(loss)], [torch.tensor(float("inf"), device="cuda")])
        nonfinite = [not bool(torch.isfinite(p.grad).all()) for p in params if p.grad is not None]
        assert any(nonfinite), f"{name}: an inf upstream gradient left every parameter gradient finite"
        scaler.step(opt)            # must skip: found_inf
        scaler.update()
        assert all(torch.equal(a, p.detach()) for a, p in zip(before, params)), f"{name}: optimizer stepped on inf gradients"
        assert scaler.get_scale() < 2.0 ** 24
