"""DenseGate called as a stand-alone module (reference model.py:522-524): softmax(fc(x)) through
the library's dense-gate kernels, with autograd support."""
from __future__ import annotations

import torch
from torch.autograd.function import once_differentiable

from ._lib import check, lib
from .functional import _f32c, _require_cuda, _stream


class _DenseGateFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias):
        _require_cuda(x, weight, bias)
        xf, w, b = _f32c(x.reshape(-1, x.shape[-1])), _f32c(weight), _f32c(bias)
        B, d = xf.shape
        n = w.shape[0]
        out = torch.empty((B, n), dtype=torch.float32, device=xf.device)
        check(lib().mmoe_dense_gate_fwd(xf.data_ptr(), w.data_ptr(), b.data_ptr(), out.data_ptr(), B, d, n, _stream()), "dense_gate_fwd")
        ctx.save_for_backward(xf, w, out)
        ctx.shape = x.shape
        ctx.in_dtype = x.dtype
        return out.reshape(*x.shape[:-1], n)

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        xf, w, out = ctx.saved_tensors
        B, d = xf.shape
        n = w.shape[0]
        dw_in = _f32c(dout.reshape(B, n))
        dl = torch.empty_like(out)
        dx = torch.empty_like(xf)
        dwg = torch.zeros_like(w)
        dbg = torch.zeros(n, dtype=torch.float32, device=xf.device)
        check(lib().mmoe_dense_gate_bwd(xf.data_ptr(), w.data_ptr(), out.data_ptr(), dw_in.data_ptr(), dl.data_ptr(), dx.data_ptr(),
                                        dwg.data_ptr(), dbg.data_ptr(), B, d, n, _stream()), "dense_gate_bwd")
        return dx.reshape(ctx.shape).to(ctx.in_dtype), dwg, dbg


def dense_gate_forward(x, weight, bias):
    return _DenseGateFn.apply(x, weight, bias)
