"""Parity cases shared by ``oracle/make_golden.py`` and ``tests/``.

TEST INFRASTRUCTURE ONLY (see ``oracle/mmoe_oracle.py``).

A case fixes: the module under test, its constructor arguments, the seed of the
deterministic weights (``synth.fill_state_dict``) and inputs, and the cotangent
used for the backward pass.  The same case is run through
  * the reference module (dev container only, ``make_golden.py``),
  * the CPU oracle (anywhere),
  * the CUDA drop-in module (GPU tests),
and the three are compared.
"""
from __future__ import annotations

from collections import OrderedDict
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Tuple

import numpy as np
import torch

from . import mmoe_oracle as O
from . import synth


@dataclass
class Case:
    name: str
    kind: str                      # head | home_head | cross | cross_home | fuse | fuse_home | img_pool | img_proj
    B: int
    seed: int
    ctor: Dict = field(default_factory=dict)

    # ---- weights -------------------------------------------------------
    def shapes(self) -> "OrderedDict[str, tuple]":
        k = self.kind
        if k == "head":
            return synth.mmoe_head_shapes(768, 6, self.ctor.get("tower_hidden", 256))
        if k == "home_head":
            return synth.home_head_shapes(6, 768, self.ctor.get("n_shared_experts", 4),
                                          self.ctor.get("n_task_experts", 2), self.ctor.get("tower_hidden", 256))
        if k in ("cross", "cross_home"):
            return synth.cross_expert_shapes(768, 2)
        if k in ("fuse", "fuse_home"):
            return synth.cross_fuse_shapes(768, 2)
        if k == "img_pool":
            return synth.image_wrapper_shapes(768)
        if k == "img_proj":
            return synth.image_projection_shapes(768, 768)
        raise KeyError(k)

    def state_dict(self) -> "OrderedDict[str, torch.Tensor]":
        return synth.fill_state_dict(self.shapes(), self.seed)

    # ---- inputs (float32; bool masks) ----------------------------------
    def inputs(self) -> Tuple[torch.Tensor, ...]:
        k, B, s = self.kind, self.B, self.seed
        if k in ("head", "home_head"):
            return (synth.expert_vecs(s, B),)
        if k in ("cross", "cross_home"):
            return synth.cross_inputs(s, B)
        if k in ("fuse", "fuse_home"):
            return (synth.doc_vectors(s, B, stream=0), synth.doc_vectors(s, B, stream=1))
        if k in ("img_pool", "img_proj"):
            return (synth.normal(s, (B, 197, 768), 600),)
        raise KeyError(k)

    def float_input_idx(self) -> List[int]:
        return [i for i, t in enumerate(self.inputs_meta()) if t]

    def inputs_meta(self) -> List[bool]:
        """True where the input is a float tensor that receives a gradient."""
        k = self.kind
        if k in ("cross", "cross_home"):
            return [True, False, True, False]
        if k in ("fuse", "fuse_home"):
            return [True, True]
        return [True]

    # ---- oracle forward -------------------------------------------------
    def oracle_forward(self, sd, inputs, relu_masks=None, trace=None, drop=None) -> Tuple[torch.Tensor, ...]:
        k = self.kind
        kw = dict(relu_masks=relu_masks, trace=trace, drop=drop)
        if k == "head":
            return tuple(O.two_task_mmoe(sd, *inputs, drop=drop))
        if k == "home_head":
            return tuple(O.home_mmoe(sd, *inputs, n_shared=self.ctor.get("n_shared_experts", 4),
                                     n_task=self.ctor.get("n_task_experts", 2), drop=drop))
        if k == "cross":
            return (O.cross_expert(sd, *inputs, **kw),)
        if k == "cross_home":
            return (O.cross_expert(sd, *inputs, home=True, **kw),)
        if k == "fuse":
            return (O.cross_fuse(sd, *inputs, **kw),)
        if k == "fuse_home":
            return (O.cross_fuse(sd, *inputs, home=True, **kw),)
        if k == "img_pool":
            return (O.item_image_pool(sd, *inputs, pool_type=self.ctor.get("pool_type", "mean"), drop=drop),)
        if k == "img_proj":
            return (O.image_projection(sd, *inputs)[1],)
        raise KeyError(k)

    # ---- cotangents -----------------------------------------------------
    def cotangents(self, outs) -> Tuple[torch.Tensor, ...]:
        return tuple(synth.normal(self.seed + 7, tuple(o.shape), 900 + j) for j, o in enumerate(outs))

    def used_param_keys(self) -> List[str]:
        """Parameters that take part in forward (HoME variants keep unused ones,
        SURVEY.md §0 quirk 4: their .grad must stay None)."""
        keys = list(self.shapes().keys())
        if self.kind == "cross_home":
            keys = [k for k in keys if not (k.startswith("norm.") or k.startswith("mlp."))]
        if self.kind == "fuse_home":
            keys = [k for k in keys if not k.startswith("proj.")]
        return keys


def run_oracle(case: Case, dtype=torch.float64, relu_masks=None, trace=None, device="cpu", drop=None):
    """Forward + backward of the oracle.  Returns (outs, input_grads, param_grads) on the CPU.
    ``relu_masks`` / ``trace``: see mmoe_oracle.encoder_layer (activation-pattern injection for low-precision parity).
    ``device``: where the oracle's arithmetic runs — "cuda" lets the GPU tests use benchmark-sized batches (the oracle is
    plain torch, float64 there as here).  ``drop``: keep-mask injection hook (mmoe_oracle._drop) for train-mode parity."""
    sd = OrderedDict((k, v.to(device=device, dtype=dtype).clone().requires_grad_(True)) for k, v in case.state_dict().items())
    raw = case.inputs()
    meta = case.inputs_meta()
    ins = [t.to(device=device, dtype=dtype).clone().requires_grad_(True) if f else t.to(device) for t, f in zip(raw, meta)]
    if relu_masks is not None:
        relu_masks = {k: v.to(device) for k, v in relu_masks.items()}
    outs = case.oracle_forward(sd, ins, relu_masks, trace, drop=drop)
    cots = case.cotangents(outs)
    torch.autograd.backward(list(outs), [c.to(device=device, dtype=dtype) for c in cots])
    gin = [t.grad.cpu() if f else None for t, f in zip(ins, meta)]
    gp = OrderedDict((k, sd[k].grad.cpu() if sd[k].grad is not None else None) for k in sd)
    if trace is not None:
        for k in list(trace):
            trace[k] = trace[k].cpu()
    return [o.detach().cpu() for o in outs], gin, gp


CASES: List[Case] = [
    Case("head_b16", "head", 16, 11),
    Case("head_b256", "head", 256, 12),                       # BASELINE.json configs[0]
    Case("home_head_b8", "home_head", 8, 21, dict(tower_hidden=512)),   # train_HoME.py:176-181
    Case("cross_b3", "cross", 3, 31),
    Case("cross_home_b3", "cross_home", 3, 32),
    Case("fuse_b8", "fuse", 8, 41),
    Case("fuse_home_b8", "fuse_home", 8, 42),
    Case("img_pool_mean_b4", "img_pool", 4, 51, dict(pool_type="mean")),
    Case("img_pool_cls_b4", "img_pool", 4, 52, dict(pool_type="cls")),
    Case("img_proj_b4", "img_proj", 4, 53),
]

CASES_BY_NAME = {c.name: c for c in CASES}


# ----------------------------------------------------------------------------
# compact fingerprints of big tensors for the committed golden files
# ----------------------------------------------------------------------------

def sample_indices(numel: int, n: int = 257) -> np.ndarray:
    if numel <= n:
        return np.arange(numel)
    return np.unique(np.linspace(0, numel - 1, n).astype(np.int64))


def fingerprint(t: torch.Tensor) -> Dict[str, torch.Tensor]:
    f = t.detach().reshape(-1).to(torch.float64)
    idx = torch.from_numpy(sample_indices(f.numel()))
    return {"sum": f.sum().reshape(1), "abs_sum": f.abs().sum().reshape(1),
            "samples": f[idx].to(torch.float32), "numel": torch.tensor([f.numel()])}


def check_fingerprint(t: torch.Tensor, fp: Dict[str, torch.Tensor], rtol: float, what: str = ""):
    """Compare a tensor with a stored fingerprint: sampled elements against the
    tensor's own scale, and the two reductions."""
    f = t.detach().reshape(-1).to(torch.float64)
    assert f.numel() == int(fp["numel"]), f"{what}: numel {f.numel()} != {int(fp['numel'])}"
    idx = torch.from_numpy(sample_indices(f.numel()))
    ref = fp["samples"].to(torch.float64)
    scale = max(float(ref.abs().max()), float(fp["abs_sum"]) / max(f.numel(), 1), 1e-30)
    err = float((f[idx] - ref).abs().max()) / scale
    assert err <= rtol, f"{what}: sampled max err / scale = {err:.3e} > {rtol}"
    abs_sum = float(fp["abs_sum"])
    assert abs(float(f.abs().sum()) - abs_sum) <= rtol * max(abs_sum, 1e-30) * 4, f"{what}: abs_sum"
    assert abs(float(f.sum()) - float(fp["sum"])) <= rtol * max(abs_sum, 1e-30) * 4, f"{what}: sum"
