"""Loss and metric tail of the training / evaluation scripts on the native library.

These live in the reference's *scripts*, not in ``model.py`` (train.py:189-192,253-254; train_HoME.py:43-51,362-364;
inference_and_auc.py:150-178), so they are opt-in: a maintainer swaps

    loss_fn_good = nn.BCEWithLogitsLoss(pos_weight=...); loss_fn_best = ...          ->  TwoTaskBCEWithLogits(pw_good, pw_best)
    calculate_contrastive_loss(a, p) x 3                                             ->  info_nce_losses([(a, p), ...])
    roc_auc_score(np.concatenate(labels), np.concatenate(sigmoid(logits).cpu()))     ->  roc_auc(scores, labels)

No CPU fallback: CPU tensors raise.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Sequence, Tuple

import torch
from torch.autograd.function import once_differentiable

from ._lib import check, lib, ptr_array
from .functional import _bytes, _call, _f32c, _require_cuda, _state, _stream, compute_dtype


class _TwoTaskBCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logit_good, logit_best, y_good, y_best, pw_good: float, pw_best: float):
        _require_cuda(logit_good, logit_best, y_good, y_best)
        B = logit_good.numel()
        logits = torch.stack([_f32c(logit_good).reshape(B), _f32c(logit_best).reshape(B)])
        loss = torch.zeros(1, dtype=torch.float32, device=logits.device)
        dlogits = torch.empty_like(logits)
        check(lib().mmoe_bce2_fwd_bwd(logits.data_ptr(), _f32c(y_good).data_ptr(), _f32c(y_best).data_ptr(), float(pw_good), float(pw_best),
                                      B, loss.data_ptr(), dlogits.data_ptr(), 1.0, _stream()), "bce2")
        ctx.save_for_backward(dlogits)
        ctx.shapes = (logit_good.shape, logit_best.shape, logit_good.dtype, logit_best.dtype)
        return loss.reshape(())

    @staticmethod
    @once_differentiable
    def backward(ctx, dloss):
        (dlogits,) = ctx.saved_tensors
        sg, sb, dg, db = ctx.shapes
        g = dlogits * dloss
        return g[0].reshape(sg).to(dg), g[1].reshape(sb).to(db), None, None, None, None


class TwoTaskBCEWithLogits(torch.nn.Module):
    """loss_fn_good(logit_g, y_good) + loss_fn_best(logit_b, y_best) of train.py:189-192,253-254 (two
    nn.BCEWithLogitsLoss(pos_weight=...), mean reduction) — value and gradient from one kernel launch."""

    def __init__(self, pos_weight_good: float = 858627.0 / 990303.0, pos_weight_best: float = 1328721.0 / 520209.0):
        super().__init__()
        self.pos_weight_good, self.pos_weight_best = float(pos_weight_good), float(pos_weight_best)

    def forward(self, logit_good, logit_best, y_good, y_best):
        return _TwoTaskBCE.apply(logit_good, logit_best, y_good, y_best, self.pos_weight_good, self.pos_weight_best)


class _InfoNCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, temperature: float, n_pairs: int, *tensors):
        _require_cuda(*tensors)
        L = lib()
        dtype = compute_dtype()
        xs = [_f32c(t) for t in tensors]              # a_0, p_0, a_1, p_1, ...
        B, d = xs[0].shape
        dev = xs[0].device
        saved = _bytes(L.mmoe_info_nce_saved_bytes(n_pairs, B, d, dtype), dev)
        work = _bytes(L.mmoe_info_nce_workspace_bytes(n_pairs, B, d, dtype), dev)
        loss = torch.zeros(n_pairs, dtype=torch.float32, device=dev)
        c = _call(dtype, B, False, 0, 0.0, 0, [], None, saved, work)
        a = ptr_array([xs[2 * i].data_ptr() for i in range(n_pairs)])
        p = ptr_array([xs[2 * i + 1].data_ptr() for i in range(n_pairs)])
        check(L.mmoe_info_nce_fwd(C.byref(c), n_pairs, d, a, p, float(temperature), loss.data_ptr()), "info_nce_fwd")
        ctx.state = (dtype, n_pairs, B, d, xs, saved, [t.dtype for t in tensors], [id(t) for t in tensors])
        ctx.needs = [t.requires_grad for t in tensors]
        return loss

    @staticmethod
    @once_differentiable
    def backward(ctx, dloss):
        dtype, n_pairs, B, d, xs, saved, in_dtypes, ids = _state(ctx)
        L = lib()
        dev = xs[0].device
        work = _bytes(L.mmoe_info_nce_workspace_bytes(n_pairs, B, d, dtype), dev)
        # one accumulation buffer per DISTINCT input tensor (i_doc and the projected image vector appear in two pairs)
        bufs = {}
        for k, need in enumerate(ctx.needs):
            if need and ids[k] not in bufs:
                bufs[ids[k]] = torch.zeros((B, d), dtype=torch.float32, device=dev)
        c = _call(dtype, B, False, 0, 0.0, 0, [], None, saved, work)
        a = ptr_array([xs[2 * i].data_ptr() for i in range(n_pairs)])
        p = ptr_array([xs[2 * i + 1].data_ptr() for i in range(n_pairs)])
        da = ptr_array([bufs[ids[2 * i]].data_ptr() if ctx.needs[2 * i] else None for i in range(n_pairs)])
        dp = ptr_array([bufs[ids[2 * i + 1]].data_ptr() if ctx.needs[2 * i + 1] else None for i in range(n_pairs)])
        dl = (C.c_float * n_pairs)(*[float(v) for v in dloss.detach().float().cpu().tolist()])
        check(L.mmoe_info_nce_bwd(C.byref(c), n_pairs, d, a, p, dl, da, dp), "info_nce_bwd")
        ctx.state = None
        # a tensor that appears several times gets its whole gradient at its first position, None at the others
        seen, grads = set(), []
        for k, need in enumerate(ctx.needs):
            if need and ids[k] not in seen:
                seen.add(ids[k])
                grads.append(bufs[ids[k]].to(in_dtypes[k]))
            else:
                grads.append(None)
        return (None, None, *grads)


def info_nce_losses(pairs: Sequence[Tuple[torch.Tensor, torch.Tensor]], temperature: float = 0.07) -> torch.Tensor:
    """calculate_contrastive_loss(anchor, positive, temperature) of train_HoME.py:43-51 for every (anchor, positive) pair
    at once (train_HoME.py:362-364 has three).  Returns a tensor with one loss per pair.  Up to 4 pairs; all [B, d] with
    B and d multiples of 8.  The upstream gradient is read on the host in backward (one tiny sync per step)."""
    flat: List[torch.Tensor] = []
    for a, p in pairs:
        flat += [a, p]
    return _InfoNCE.apply(float(temperature), len(pairs), *flat)


def roc_auc(scores: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
    """sklearn.metrics.roc_auc_score(labels, scores) (inference_and_auc.py:171,178) on the device: returns a 0-dim float64
    CUDA tensor (NaN when one class is empty); no host synchronisation."""
    _require_cuda(scores, labels)
    s, y = _f32c(scores).reshape(-1), _f32c(labels).reshape(-1)
    if s.numel() != y.numel() or s.numel() == 0:
        raise RuntimeError("roc_auc: scores and labels must be non-empty and of equal length")
    L = lib()
    n = s.numel()
    nbytes = L.mmoe_auc_workspace_bytes(n)
    work = torch.empty((nbytes + 7) // 8, dtype=torch.int64, device=s.device)
    out = torch.empty(1, dtype=torch.float64, device=s.device)
    check(L.mmoe_auc(s.data_ptr(), y.data_ptr(), n, work.data_ptr(), work.numel() * 8, out.data_ptr(), _stream()), "auc")
    return out.reshape(())
