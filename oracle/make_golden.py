"""Generate tests/golden/*.pt from the UNMODIFIED reference (dev container only).

TEST INFRASTRUCTURE ONLY.  Run from the repo root:

    python -m oracle.make_golden

For every case in ``oracle/cases.py`` this
  1. builds the reference module from ``/root/reference/model.py`` /
     ``model_HoME.py`` (imported through ``oracle/ref_import.py``),
  2. checks the reference's ``state_dict()`` keys and shapes against
     ``oracle/synth.py``'s tables and loads the deterministic weights,
  3. runs forward + backward in float32, ``eval()`` mode with autograd on
     (dropout off — SURVEY.md §0 quirk 3), on the deterministic inputs,
  4. runs the CPU oracle in float64 and float32 on the same weights/inputs and
     ASSERTS agreement with the reference (this is what pins the oracle),
  5. writes the reference's outputs (full), input gradients and parameter
     gradients (fingerprints: sum, abs-sum, 257 strided samples) to
     ``tests/golden/<case>.pt``.
The committed files let the GPU box (which has no ``/root/reference``) check both
the oracle and the CUDA path against what the real reference computed.
"""
from __future__ import annotations

import os
import sys
from collections import OrderedDict

import torch

from . import cases as C
from .ref_import import load_reference

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


class _Tokens:
    def __init__(self, t):
        self.last_hidden_state = t


class _FakeBackbone(torch.nn.Module):
    """Stands in for the HF ViT: returns the given token tensor (the backbone is
    not part of the re-implemented path)."""

    class _Cfg:
        hidden_size = 768

    config = _Cfg()

    def forward(self, pixel_values):
        return _Tokens(pixel_values)


def build_reference(case: C.Case):
    m, h = load_reference("model"), load_reference("model_HoME")
    k = case.kind
    if k == "head":
        return m.TwoTaskMMoE(**case.ctor)
    if k == "home_head":
        return h.HOME_MMoE_Complete(expert_dim=768, **case.ctor)
    if k == "cross":
        return m.RobustTextCrossExpert()
    if k == "cross_home":
        return h.RobustTextCrossExpert()
    if k == "fuse":
        return m.EnhancedCrossFuse()
    if k == "fuse_home":
        return h.EnhancedCrossFuse()
    if k == "img_pool":
        return m.ItemImageExpert(_FakeBackbone(), pool_type=case.ctor.get("pool_type", "mean"))
    if k == "img_proj":
        return h.ImageExpertWithProjection(_FakeBackbone())
    raise KeyError(k)


def run_reference(case: C.Case):
    mod = build_reference(case).eval()
    ref_sd = mod.state_dict()
    shapes = case.shapes()
    assert list(ref_sd.keys()) == list(shapes.keys()), (
        f"{case.name}: state_dict keys differ\n ref={list(ref_sd.keys())}\n ours={list(shapes.keys())}")
    for key in shapes:
        assert tuple(ref_sd[key].shape) == tuple(shapes[key]), (case.name, key, ref_sd[key].shape, shapes[key])
    mod.load_state_dict(case.state_dict(), strict=True)
    raw = case.inputs()
    meta = case.inputs_meta()
    ins = [t.clone().requires_grad_(True) if f else t for t, f in zip(raw, meta)]
    if case.kind == "img_pool":
        out = mod(ins[0], trainable=True)
    elif case.kind == "img_proj":
        out = mod(ins[0])[1]
    else:
        out = mod(*ins)
    outs = tuple(out) if isinstance(out, (tuple, list)) else (out,)
    cots = case.cotangents(outs)
    torch.autograd.backward(list(outs), list(cots))
    gin = [t.grad if f else None for t, f in zip(ins, meta)]
    gp = OrderedDict((n, p.grad) for n, p in mod.named_parameters())
    return [o.detach() for o in outs], gin, gp


def _nerr(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.double(), b.double()
    return float((a - b).abs().max()) / max(float(b.abs().max()), 1e-30)


def main():
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count() or 1)
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    worst = 0.0
    for case in C.CASES:
        r_out, r_gin, r_gp = run_reference(case)
        o_out, o_gin, o_gp = C.run_oracle(case, torch.float64)
        f_out, f_gin, f_gp = C.run_oracle(case, torch.float32)
        errs = []
        for a, b in zip(o_out, r_out):
            errs.append(("out", _nerr(a, b)))
        for a, b in zip(f_out, r_out):
            errs.append(("out32", _nerr(a, b)))
        for j, (a, b) in enumerate(zip(o_gin, r_gin)):
            if b is not None:
                errs.append((f"gin{j}", _nerr(a, b)))
        used = set(case.used_param_keys())
        for key, g in r_gp.items():
            if key in used:
                assert g is not None, (case.name, key, "reference grad is None for a used param")
                errs.append((key, _nerr(o_gp[key], g)))
            else:
                assert g is None, (case.name, key, "expected unused param")
                assert o_gp[key] is None, (case.name, key, "oracle touched an unused param")
        m = max(e for _, e in errs)
        worst = max(worst, m)
        bad = [(n, e) for n, e in errs if e > 2e-5]
        print(f"{case.name:20s} max normalised |oracle-reference| = {m:.2e}   ({len(errs)} tensors)")
        assert not bad, f"{case.name}: oracle disagrees with the reference: {bad[:5]}"
        blob = {
            "meta": {"name": case.name, "kind": case.kind, "B": case.B, "seed": case.seed, "ctor": case.ctor,
                     "torch": torch.__version__, "dtype": "float32", "mode": "eval+grad"},
            "out": [o.clone() for o in r_out],
            "grad_in": [None if g is None else C.fingerprint(g) for g in r_gin],
            "grad_param": {k: (None if g is None else C.fingerprint(g)) for k, g in r_gp.items()},
        }
        torch.save(blob, os.path.join(GOLDEN_DIR, case.name + ".pt"))
    print(f"golden vectors written to {GOLDEN_DIR}; worst oracle-vs-reference error {worst:.2e}")


if __name__ == "__main__":
    sys.exit(main())
