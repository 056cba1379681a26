"""HomeExpertWrapper x 6 + torch.stack in one launch.

``HomeExpertWrapper`` is defined in the reference's *script* (train_HoME.py:100-116): ``Dropout(SiLU(BatchNorm1d(x)))`` with
per-rank batch statistics; the script calls six of them and stacks the results into the ``expert_vecs`` of
``HOME_MMoE_Complete`` (train_HoME.py:350-356).  Because the class lives in the script the fusion is opt-in:

    stack6 = FusedHomeExpertStack([u_doc_wrapper, i_doc_wrapper, img_vec_wrapper, ui_vec_wrapper, xui_wrapper, xti_wrapper])
    expert_vecs = stack6(u_doc, i_doc, img_vec, ui_vec, xui, xti)        # replaces train_HoME.py:350-356

``FusedHomeExpertStack`` does not own parameters: it reads ``wrapper.norm.weight / bias / running_mean / running_var`` and
``wrapper.dropout.p`` of the wrappers it is given (the script's own modules, DDP-wrapped or not), updates the running
statistics and ``num_batches_tracked`` exactly as ``nn.BatchNorm1d`` does, and hands gradients for ``norm.weight`` /
``norm.bias`` back through autograd — so the state dict, the optimizer groups and the checkpoints of the script are
untouched.  ``HomeExpertWrapper`` (same constructor, same state-dict keys) is provided for callers that do not import
the script's class.
"""
from __future__ import annotations

import ctypes as C
from typing import Sequence

import torch
import torch.nn as nn
from torch.autograd.function import once_differentiable

from ._lib import F32, check, lib, ptr_array
from .functional import _call, _f32c, _new_seed, _require_cuda, _state


class _BnSiluStack(torch.autograd.Function):
    @staticmethod
    def forward(ctx, training: bool, drop_p: float, momentum: float, eps: float, n: int, running, *tensors):
        xs_in, params = tensors[:n], tensors[n:]
        _require_cuda(*xs_in, *params)
        L = lib()
        xs = [_f32c(x) for x in xs_in]
        B, d = xs[0].shape
        dev = xs[0].device
        for x in xs:
            if tuple(x.shape) != (B, d):
                raise RuntimeError("FusedHomeExpertStack: all inputs must have the same [B, d] shape")
        pt = [_f32c(p) for p in params]
        out = torch.empty((B, n, d), dtype=torch.float32, device=dev)
        save = torch.empty((2, n, d), dtype=torch.float32, device=dev)
        seed = _new_seed(training, drop_p)
        c = _call(F32, B, training, 0, drop_p, seed, pt, None, None, None)
        xp = ptr_array([x.data_ptr() for x in xs])
        rm = ptr_array([r[0].data_ptr() for r in running])
        rv = ptr_array([r[1].data_ptr() for r in running])
        check(L.mmoe_bn_silu_stack_fwd(C.byref(c), n, d, xp, out.data_ptr(), save[0].data_ptr(), save[1].data_ptr(), rm, rv,
                                       float(momentum), float(eps)), "bn_silu_stack_fwd")
        ctx.state = (training, drop_p, seed, eps, n, B, d, xs, pt, save, running, [x.dtype for x in xs_in])
        ctx.needs_x = [x.requires_grad for x in xs_in]
        ctx.param_req = [p.requires_grad for p in params]
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        training, drop_p, seed, eps, n, B, d, xs, pt, save, running, in_dtypes = _state(ctx)
        L = lib()
        dev = dout.device
        do = _f32c(dout)
        grads = torch.zeros((2 * n, d), dtype=torch.float32, device=dev)
        dxs = [torch.empty((B, d), dtype=torch.float32, device=dev) if need else None for need in ctx.needs_x]
        c = _call(F32, B, training, 0, drop_p, seed, pt, [grads[i].data_ptr() for i in range(2 * n)], None, None)
        xp = ptr_array([x.data_ptr() for x in xs])
        rm = ptr_array([r[0].data_ptr() for r in running])
        rv = ptr_array([r[1].data_ptr() for r in running])
        dxp = ptr_array([t.data_ptr() if t is not None else None for t in dxs])
        check(L.mmoe_bn_silu_stack_bwd(C.byref(c), n, d, xp, do.data_ptr(), save[0].data_ptr(), save[1].data_ptr(), rm, rv, dxp,
                                       float(eps)), "bn_silu_stack_bwd")
        ctx.state = None
        gx = [t.to(dt) if t is not None else None for t, dt in zip(dxs, in_dtypes)]
        gp = [grads[i] if req else None for i, req in enumerate(ctx.param_req)]
        return (None, None, None, None, None, None, *gx, *gp)


class HomeExpertWrapper(nn.Module):
    """Same constructor and state-dict keys as train_HoME.py:100-116 (``norm.*`` of an ``nn.BatchNorm1d``).  Called on its
    own it runs the fused kernel with n = 1."""

    def __init__(self, d_model: int, dropout_p: float = 0.1):
        super().__init__()
        self.norm = nn.BatchNorm1d(d_model)
        self.dropout = nn.Dropout(dropout_p)

    def forward(self, x):
        if x.dim() == 3:                                   # (B, L, D) is normalised over B*L, train_HoME.py:110-113
            b, l, d = x.shape
            return FusedHomeExpertStack([self])(x.reshape(b * l, d)).reshape(b, l, d)
        return FusedHomeExpertStack([self])(x)[:, 0]


class FusedHomeExpertStack(nn.Module):
    """expert_vecs[B, n, d] = stack([Dropout(SiLU(BatchNorm1d_e(x_e))) for e in range(n)], dim=1) in one launch."""

    def __init__(self, wrappers: Sequence[nn.Module]):
        super().__init__()
        ws = [w.module if hasattr(w, "module") and not hasattr(w, "norm") else w for w in wrappers]    # unwrap DistributedDataParallel
        for w in ws:
            if not isinstance(getattr(w, "norm", None), nn.BatchNorm1d):
                raise TypeError("FusedHomeExpertStack expects HomeExpertWrapper-like modules with a .norm BatchNorm1d")
        # not registered as sub-modules: the wrappers stay owned by whoever built them (state dict, optimizer, DDP)
        self.__dict__["_wrappers"] = ws

    def forward(self, *xs: torch.Tensor) -> torch.Tensor:
        ws = self.__dict__["_wrappers"]
        if len(xs) != len(ws):
            raise RuntimeError(f"FusedHomeExpertStack: {len(ws)} wrappers, {len(xs)} inputs")
        bn0 = ws[0].norm
        training = ws[0].training
        p = float(ws[0].dropout.p) if hasattr(ws[0], "dropout") else 0.0
        running, params = [], []
        for w in ws:
            bn = w.norm
            if bn.running_mean is None or not bn.affine or bn.momentum is None:
                raise RuntimeError("FusedHomeExpertStack: BatchNorm1d must be affine with running statistics and a fixed momentum")
            if w.training != training or bn.eps != bn0.eps or bn.momentum != bn0.momentum:
                raise RuntimeError("FusedHomeExpertStack: the wrappers must share mode, eps and momentum")
            running.append((bn.running_mean, bn.running_var))
            params += [bn.weight, bn.bias]
        out = _BnSiluStack.apply(training, p, float(bn0.momentum), float(bn0.eps), len(ws), running, *xs, *params)
        if training:
            for w in ws:
                w.norm.num_batches_tracked += 1
        return out
