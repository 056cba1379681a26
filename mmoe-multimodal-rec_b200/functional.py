"""torch.autograd.Function wrappers over the C ABI (include/mmoe_b200.h).

PyTorch is used here for device memory (tensors as buffers), streams and autograd
bookkeeping only; every arithmetic step of the path runs in libmmoe_b200.so.
There is no CPU / eager fallback: non-CUDA inputs raise.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import torch
from torch.autograd.function import once_differentiable

from . import _lib
from ._lib import BF16, F16, F32, Call, CrossCfg, FuseCfg, HeadCfg, HomeCfg, check, lib, ptr_array

# test hook: when set to a list, CrossFn / FuseFn append (kind, cfg, home, B, dtype, saved_blob) after forward
DEBUG_SAVED = None
# RobustTextCrossExpert backward as a chain of autograd stages (gradients reach DDP's reducer layer by layer, its all-reduce
# overlaps the rest of backward) or as one node (MMOE_CROSS_STAGED=0: all of the expert's gradients appear at the end)
import os as _os
CROSS_STAGED = _os.environ.get("MMOE_CROSS_STAGED", "1") != "0"
# native exchange: start each collective when its stage is enqueued (overlap) or all of them at the end of backward
NATIVE_DEFER = _os.environ.get("MMOE_NATIVE_DEFER", "0") == "1"
# fused-parameter mode (the default; MMOE_FLAT_PARAMS=0 or set_flat_parameters(False) before construction turns it off):
# modules built while this is on keep all of their (used) parameters in ONE nn.Parameter whose storage the reference-named
# attributes view (modules._Native.fuse_parameters); state_dict keys, order and shapes are unchanged
FLAT_PARAMS = _os.environ.get("MMOE_FLAT_PARAMS", "1") != "0"


def set_flat_parameters(on: bool):
    """Modules constructed after this call are (on=True) / are not (on=False) in fused-parameter mode."""
    global FLAT_PARAMS
    FLAT_PARAMS = bool(on)

_TORCH2MMOE = {torch.float32: F32, torch.bfloat16: BF16, torch.float16: F16}
_MMOE2TORCH = {F32: torch.float32, BF16: torch.bfloat16, F16: torch.float16}


def compute_dtype() -> int:
    """fp32 unless a CUDA autocast region is active, then its 16-bit dtype (the reference
    scripts use torch.cuda.amp.autocast() = fp16; BASELINE's bf16 configs pass dtype=bfloat16)."""
    if torch.is_autocast_enabled("cuda"):
        dt = torch.get_autocast_dtype("cuda")
        if dt in (torch.bfloat16, torch.float16):
            return _TORCH2MMOE[dt]
    return F32


def _require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("mmoe_b200 runs on CUDA (sm_100a) only; got a tensor on %s — there is no CPU fallback" % t.device)


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _f32c(x: torch.Tensor) -> torch.Tensor:
    return x.detach().to(torch.float32).contiguous()


def _new_seed(training: bool, p: float) -> int:
    if not training or p <= 0.0:
        return 0
    # drawn from torch's CPU generator: reproducible under torch.manual_seed, no device sync
    return int(torch.randint(0, 2 ** 62, (1,), dtype=torch.int64).item())


class ParamPack:
    """Device pointers of a module's parameters in state_dict order.

    2-D GEMM weights are handed to the library in the compute dtype; the 16-bit copies
    are produced by the library's cast kernel and cached until the parameter is updated
    in place (optimizer step) — detected through the tensor's version counter.
    """

    def __init__(self, names: Sequence[str], lowp: Sequence[str]):
        self.names = list(names)
        self.lowp = set(lowp)
        self._cache = {}
        self._stable = {}

    def tensors(self, params: Sequence[torch.Tensor], dtype: int, flat: Optional[torch.Tensor] = None) -> List[torch.Tensor]:
        """flat: the fused parameter that owns the storage of `params` (fused-parameter modules).  While its version
        counter and address are unchanged, the tensors handed to the library are the ones of the previous call: the whole
        list — with its ctypes pointer array and gradient-buffer plan — is reused instead of being rebuilt per call."""
        if flat is not None:
            key = (dtype, flat._version, flat.data_ptr(), len(params))
            hit = self._stable.get(dtype)
            if hit is not None and hit.key == key:
                return hit
        out = _Stable() if flat is not None else []
        L = lib()
        for name, p in zip(self.names, params):
            _require_cuda(p)
            if p.dtype != torch.float32:
                raise RuntimeError(f"parameter {name} must be float32 (got {p.dtype})")
            pd = p.detach()
            if not pd.is_contiguous():
                pd = pd.contiguous()
            if dtype == F32 or name not in self.lowp:
                out.append(pd)
                continue
            key_c = (name, dtype)
            ent = self._cache.get(key_c)
            if ent is None or ent[0] != p._version or ent[1] != pd.data_ptr() or ent[2].device != pd.device:
                q = torch.empty(pd.shape, dtype=_MMOE2TORCH[dtype], device=pd.device)
                check(L.mmoe_cast_f32(pd.data_ptr(), q.data_ptr(), pd.numel(), dtype, _stream()), "cast")
                ent = (p._version, pd.data_ptr(), q)
                self._cache[key_c] = ent
            out.append(ent[2])
        if flat is not None:
            out.key = key
            out.arr = ptr_array([t.data_ptr() for t in out])
            out.plans = {}
            self._stable[dtype] = out
        return out


class _Stable(list):
    """Parameter tensor list of a fused-parameter module that stays valid until the parameter changes; carries the
    ctypes pointer array (`arr`) and the gradient-buffer plans (`plans`) derived from it."""
    __slots__ = ("key", "arr", "plans")


_LAYOUTS = {}


def _grad_layout(params: Sequence[torch.Tensor], used: Sequence[bool]):
    """(total floats, [(offset, shape, contiguous strides) or None]) of the flat gradient buffer; cached per shape signature."""
    key = (tuple(tuple(p.shape) for p in params), tuple(bool(u) for u in used))
    lay = _LAYOUTS.get(key)
    if lay is None:
        entries, off = [], 0
        for p, u in zip(params, used):
            if not u:
                entries.append(None)
                continue
            shape = tuple(p.shape)
            strides, acc = [], 1
            for n in reversed(shape):
                strides.append(acc)
                acc *= n
            entries.append((off, shape, tuple(reversed(strides))))
            off += (p.numel() + 63) // 64 * 64
        lay = (max(off, 1), entries)
        _LAYOUTS[key] = lay
    return lay


def _alloc_grads(params: Sequence[torch.Tensor], used: Sequence[bool], want_views: bool = True):
    """One zeroed fp32 buffer holding the gradients of all used parameters (each at a 256-byte aligned offset); returns
    (views, pointer list, flat buffer).  One as_strided per parameter: the view creation is host time on the critical
    path of every backward (≈300 parameters per step).  want_views=False (fused-parameter modules, whose single
    parameter takes the flat buffer itself as its gradient): views is None."""
    if not want_views and isinstance(params, _Stable):
        plan = params.plans.get("grads")
        if plan is None or plan[2] != list(used):
            total, entries = _grad_layout(params, used)
            plan = (total, [None if e is None else 4 * e[0] for e in entries], list(used))
            params.plans["grads"] = plan
        flat = torch.zeros(plan[0], dtype=torch.float32, device=params[0].device)
        base = flat.data_ptr()
        return None, [None if o is None else base + o for o in plan[1]], flat
    total, entries = _grad_layout(params, used)
    flat = torch.zeros(total, dtype=torch.float32, device=params[0].device)
    base = flat.data_ptr()
    views, ptrs = [], []
    for e in entries:
        if e is None:
            views.append(None)
            ptrs.append(None)
        else:
            if want_views:
                views.append(flat.as_strided(e[1], e[2], e[0]))
            ptrs.append(base + 4 * e[0])
    return (views if want_views else None), ptrs, flat


# ----------------------------------------------------------------------------------------------
# Data-parallel gradient exchange on the flat buffers
# ----------------------------------------------------------------------------------------------
class _GradSync:
    """Averages gradients across the ranks of a process group, one collective per module (or encoder-layer stage)
    backward: the backward functions already produce all of a module's parameter gradients in ONE contiguous fp32
    buffer, so the exchange needs no per-parameter hooks, no bucket copies (DistributedDataParallel copies each of the
    ~300 gradients into its buckets with a kernel of its own, ~1 ms per step of launch-bound work) and starts as soon
    as the stage's gradients exist, overlapping the rest of the backward pass.

    Ownership rule (what makes the averaged values the ones that end up in ``.grad``): while the exchange is on, a
    backward function does NOT return its parameter gradients to autograd.  It keeps the flat buffer, starts the
    in-place collective on it, and registers the (parameter, view) pairs here.  ``finish()`` — queued on the autograd
    engine so that it runs when the backward pass ends, exactly where DDP finalises its buckets — first makes the
    compute stream wait for every collective and only then installs the views: ``p.grad = view`` when the parameter
    has no gradient yet (no copy: ``.grad`` aliases the reduced buffer), ``p.grad += view`` when it has one
    (gradient accumulation, ``zero_grad(set_to_none=False)``).  Nothing reads or copies a buffer while NCCL is still
    reducing it."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.world = dist.get_world_size(group)
        self.avg = dist.get_backend(group) == "nccl"
        self.sync = True             # False = accumulate locally (the equivalent of DDP.no_sync())
        self.pending = []            # (work, tensor to divide afterwards or None)
        self.late = []               # NATIVE_DEFER: buffers whose collective starts at the end of the backward pass
        self.deferred = []           # (params, views, flat): installed into .grad by finish()
        self.callback_queued = False

    def reduce(self, flat: torch.Tensor, now: bool = False):
        if self.world == 1 or flat.numel() == 0 or not self.sync:
            return
        if NATIVE_DEFER and not now and self._ensure_callback():
            self.late.append(flat)
            return
        if self.avg:
            w = self.dist.all_reduce(flat, op=self.dist.ReduceOp.AVG, group=self.group, async_op=True)
            self.pending.append((w, None))
        else:                                   # gloo (CPU tests): no AVG
            w = self.dist.all_reduce(flat, op=self.dist.ReduceOp.SUM, group=self.group, async_op=True)
            self.pending.append((w, flat))

    def defer(self, params, views, flat):
        """Hand the gradients of `params` (views of `flat`) over for installation at the end of the backward pass."""
        self.deferred.append((list(params), list(views), flat))
        self._ensure_callback()

    def _ensure_callback(self) -> bool:
        """Queue finish() on the autograd engine (runs when the current backward pass ends); False outside a backward pass."""
        if not self.callback_queued:
            self.callback_queued = True
            try:
                torch.autograd.Variable._execution_engine.queue_callback(self.finish)
            except RuntimeError:
                # not inside a backward pass (direct call in a test): the caller runs wait_grad_allreduce()
                self.callback_queued = False
        return self.callback_queued

    def wait(self):
        for w, t in self.pending:
            w.wait()
            if t is not None:
                t.div_(self.world)
        self.pending.clear()

    def finish(self):
        self.callback_queued = False
        for flat in self.late:
            self.reduce(flat, now=True)
        self.late.clear()
        self.wait()
        for params, views, flat in self.deferred:
            _install_grads(params, views, flat)
        self.deferred.clear()


def _install_grads(params, views, flat):
    """p.grad = view (aliasing `flat`) or p.grad += view; one fused add when every .grad still aliases the previous
    flat buffer of the same layout (the usual gradient-accumulation case)."""
    live = [(p, v) for p, v in zip(params, views) if v is not None and p is not None]
    if not live:
        return
    if all(p.grad is None for p, _ in live):
        for p, v in live:
            p.grad = v
            p._mmoe_grad_flat = flat
        return
    prev = getattr(live[0][0], "_mmoe_grad_flat", None)
    if (prev is not None and prev.shape == flat.shape and prev.device == flat.device and
            all(p.grad is not None and getattr(p, "_mmoe_grad_flat", None) is prev and p.grad.shape == v.shape and
                p.grad.is_contiguous() and p.grad.data_ptr() - prev.data_ptr() == v.data_ptr() - flat.data_ptr()
                for p, v in live)):
        prev.add_(flat)
        return
    for p, v in live:
        if p.grad is None:
            p.grad = v.clone()
        else:
            p.grad.add_(v)
        p._mmoe_grad_flat = None


_GRAD_SYNC = None


def enable_grad_allreduce(group=None):
    """Turn on the gradient all-reduce of every native module's backward (call after init_process_group; instead of
    wrapping the modules in DistributedDataParallel).  The averaged gradients are in ``.grad`` when ``backward()``
    returns; ``wait_grad_allreduce()`` is kept for explicit use outside a backward pass."""
    global _GRAD_SYNC
    _GRAD_SYNC = _GradSync(group)
    return _GRAD_SYNC


def disable_grad_allreduce():
    global _GRAD_SYNC
    _GRAD_SYNC = None


def set_grad_sync(on: bool):
    """on=False: the following backward passes accumulate locally without communicating (DDP.no_sync() equivalent);
    on=True: the next backward all-reduces what it produces.  Because the all-reduce is linear, the usual pattern —
    no_sync on all but the last micro-step — needs the accumulated buffer reduced once: call
    ``allreduce_accumulated(modules)`` after the last micro-step instead of re-enabling the per-stage exchange."""
    if _GRAD_SYNC is not None:
        _GRAD_SYNC.sync = bool(on)


def allreduce_accumulated(modules):
    """Average the accumulated ``.grad`` of the given native modules across ranks: one collective per flat buffer
    whose views are still the parameters' gradients, one per parameter otherwise.  For gradient accumulation with
    set_grad_sync(False) on all micro-steps."""
    gs = _GRAD_SYNC
    if gs is None or gs.world == 1:
        return
    was = gs.sync
    gs.sync = True
    try:
        seen = set()
        for m in modules:
            for p in m.parameters():
                if p.grad is None:
                    continue
                flat = getattr(p, "_mmoe_grad_flat", None)
                inside = flat is not None and 0 <= p.grad.data_ptr() - flat.data_ptr() <= (flat.numel() - p.grad.numel()) * 4
                if inside:
                    if id(flat) not in seen:
                        seen.add(id(flat))
                        gs.reduce(flat)
                else:
                    gs.reduce(p.grad)
        gs.wait()
    finally:
        gs.sync = was


def wait_grad_allreduce():
    """Make the current stream wait for the gradient collectives issued so far and install their results in ``.grad``
    (done automatically at the end of a backward pass)."""
    if _GRAD_SYNC is not None:
        _GRAD_SYNC.finish()


def _sync_grads(flat: torch.Tensor):
    if _GRAD_SYNC is not None:
        _GRAD_SYNC.reduce(flat)


def _hand_over(params, views, req, flat):
    """What a backward function returns for its parameters.  Exchange off: the views themselves (autograd /
    DistributedDataParallel take it from there; no other reference to them is kept, so AccumulateGrad adopts them
    without a copy).  Exchange on: None for every parameter — the views go to _GradSync.defer (see its docstring)."""
    out = [v if (r and v is not None) else None for v, r in zip(views, req)]
    if _GRAD_SYNC is None:
        return out
    _GRAD_SYNC.defer([p if o is not None else None for p, o in zip(params, out)], out, flat)
    return [None] * len(out)


def _split_fused(pack: "ParamPack", params):
    """Modules in fused-parameter mode (modules._Native.fuse_parameters) pass their parameter VIEWS (plain tensors, no
    autograd) followed by the one nn.Parameter that owns the storage; returns (per-name tensors, that parameter or None)."""
    if getattr(pack, "fused", False):
        return params[:-1], params[-1]
    return params, None


def _param_grads(ctx, pt, used):
    """(views, grad pointer list, flat) for a backward whose forward stored ctx.flat_p."""
    return _alloc_grads(pt, used, want_views=ctx.flat_p is None)


def _finish_params(ctx, views, flat):
    """The trailing entries of a backward's return value: one gradient per parameter input of the forward."""
    if ctx.flat_p is None:
        return _hand_over(ctx.params, views, ctx.param_req, flat)
    # fused mode: the flat buffer IS the gradient of the module's single parameter (same layout by construction)
    fp = ctx.flat_p
    if tuple(flat.shape) != tuple(fp.shape):
        raise RuntimeError("mmoe_b200: fused parameter layout changed after fuse_parameters() (got %s, expected %s)"
                           % (tuple(fp.shape), tuple(flat.shape)))
    g = _hand_over([fp], [flat], [fp.requires_grad], flat)
    ctx.flat_p = None
    return [None] * len(ctx.param_req) + list(g)


def _span(flat: torch.Tensor, views, lo: int, hi: int):
    """The contiguous slice of `flat` that holds views[lo:hi] (None entries skipped)."""
    vs = [v for v in views[lo:hi] if v is not None]
    if not vs:
        return flat[0:0]
    start = vs[0].storage_offset() - flat.storage_offset()
    end = vs[-1].storage_offset() - flat.storage_offset() + vs[-1].numel()
    return flat[start:end]


def _state(ctx):
    if ctx.state is None:
        raise RuntimeError("mmoe_b200: backward called twice over the same graph; the saved activations are released "
                           "at the end of the first backward pass (retain_graph is not supported)")
    return ctx.state


def _bytes(n: int, device) -> torch.Tensor:
    return torch.empty(max(int(n), 1), dtype=torch.uint8, device=device)


def _call(dtype, B, training, home, drop_p, seed, ptensors, grad_ptrs, saved, workspace) -> Call:
    c = Call()
    c.dtype, c.B, c.training, c.home = dtype, B, int(training), int(home)
    c.drop_p, c.seed = float(drop_p), int(seed)
    pa = ptensors.arr if isinstance(ptensors, _Stable) else ptr_array([t.data_ptr() for t in ptensors])
    c.params = C.cast(pa, C.POINTER(C.c_void_p))
    c._keep = [pa, ptensors]
    if grad_ptrs is not None:
        ga = ptr_array(grad_ptrs)
        c.grads = C.cast(ga, C.POINTER(C.c_void_p))
        c._keep.append(ga)
    if saved is not None:
        c.saved, c.saved_bytes = saved.data_ptr(), saved.numel()
    if workspace is not None:
        c.workspace, c.workspace_bytes = workspace.data_ptr(), workspace.numel()
    c.stream = _stream()
    return c


# ----------------------------------------------------------------------------------------------
# TwoTaskMMoE / HOME_MMoE_Complete
# ----------------------------------------------------------------------------------------------
class HeadFn(torch.autograd.Function):
    """logits[2,B] = head(expert_vecs[B,n,d]).  kind: 'mmoe' (model.py:562-577) or 'home' (model_HoME.py:590-638)."""

    @staticmethod
    def forward(ctx, pack: ParamPack, kind: str, cfg, training: bool, drop_p: float, want_gates: bool,
                expert_vecs: torch.Tensor, *params: torch.Tensor):
        _require_cuda(expert_vecs)
        L = lib()
        dtype = compute_dtype()
        ev = _f32c(expert_vecs)
        B, dev = ev.shape[0], ev.device
        params, ctx.flat_p = _split_fused(pack, params)
        pt = pack.tensors(params, dtype, ctx.flat_p)
        fwd, bwd, sb, wb = ((L.mmoe_head_fwd, L.mmoe_head_bwd, L.mmoe_head_saved_bytes, L.mmoe_head_workspace_bytes) if kind == "mmoe"
                            else (L.mmoe_home_fwd, L.mmoe_home_bwd, L.mmoe_home_saved_bytes, L.mmoe_home_workspace_bytes))
        saved = _bytes(sb(C.byref(cfg), B, dtype), dev)
        logits = torch.empty((2, B), dtype=torch.float32, device=dev)
        n_gate = cfg.n_expert if kind == "mmoe" else cfg.n_shared + cfg.n_task
        gates = torch.empty((2, B, n_gate), dtype=torch.float32, device=dev) if want_gates else None
        seed = _new_seed(training, drop_p)
        c = _call(dtype, B, training, 0, drop_p, seed, pt, None, saved, None)
        check(fwd(C.byref(c), C.byref(cfg), ev.data_ptr(), logits.data_ptr(), gates.data_ptr() if want_gates else None), kind + "_fwd")
        ctx.state = (pack, kind, cfg, training, drop_p, seed, dtype, ev, pt, saved, bwd, wb)
        ctx.n_params = len(params)
        ctx.params = params
        ctx.param_req = [p.requires_grad for p in params]
        ctx.in_dtype = expert_vecs.dtype
        ctx.mark_non_differentiable(*( [gates] if want_gates else [] ))
        if want_gates:
            return logits, gates
        return logits

    @staticmethod
    @once_differentiable
    def backward(ctx, dlogits, *unused):
        pack, kind, cfg, training, drop_p, seed, dtype, ev, pt, saved, bwd, wb = _state(ctx)
        L = lib()
        B, dev = ev.shape[0], ev.device
        views, gptrs, flat = _param_grads(ctx, pt, [True] * ctx.n_params)
        work = _bytes(wb(C.byref(cfg), B, dtype), dev)
        d_ev = torch.empty_like(ev)
        dl = _f32c(dlogits)
        c = _call(dtype, B, training, 0, drop_p, seed, pt, gptrs, saved, work)
        check(bwd(C.byref(c), C.byref(cfg), ev.data_ptr(), dl.data_ptr(), d_ev.data_ptr()), kind + "_bwd")
        _sync_grads(flat)
        grads = _finish_params(ctx, views, flat)
        ctx.state = ctx.params = None           # release the saved blob / inputs / 16-bit weights with the pass
        return (None, None, None, None, None, None, d_ev.to(ctx.in_dtype), *grads)


# ----------------------------------------------------------------------------------------------
# RobustTextCrossExpert
# ----------------------------------------------------------------------------------------------
class CrossFn(torch.autograd.Function):
    """out[B,d] = cross(user[B,S,d], user_mask[B,S], item[B,S,d], item_mask[B,S]) — model.py:426-451
    (home=True: model_HoME.py:441-466)."""

    @staticmethod
    def forward(ctx, pack: ParamPack, cfg: CrossCfg, home: bool, used: Sequence[bool], training: bool, drop_p: float,
                user, user_mask, item, item_mask, *params):
        _require_cuda(user, user_mask, item, item_mask)
        L = lib()
        dtype = compute_dtype()
        u, it = _f32c(user), _f32c(item)
        um = user_mask.detach().to(torch.bool).contiguous().view(torch.uint8)
        im = item_mask.detach().to(torch.bool).contiguous().view(torch.uint8)
        B, dev = u.shape[0], u.device
        params, ctx.flat_p = _split_fused(pack, params)
        pt = pack.tensors(params, dtype, ctx.flat_p)
        saved = _bytes(L.mmoe_cross_saved_bytes(C.byref(cfg), B, dtype), dev)
        out = torch.empty((B, cfg.d), dtype=torch.float32, device=dev)
        seed = _new_seed(training, drop_p)
        c = _call(dtype, B, training, home, drop_p, seed, pt, None, saved, None)
        check(L.mmoe_cross_fwd(C.byref(c), C.byref(cfg), u.data_ptr(), um.data_ptr(), it.data_ptr(), im.data_ptr(), out.data_ptr()), "cross_fwd")
        if DEBUG_SAVED is not None:
            DEBUG_SAVED.append(("cross", cfg, home, B, dtype, saved))
        ctx.state = (cfg, home, list(used), training, drop_p, seed, dtype, u, um, it, im, pt, saved)
        ctx.params = params
        ctx.param_req = [p.requires_grad for p in params]
        ctx.in_dtypes = (user.dtype, item.dtype)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        cfg, home, used, training, drop_p, seed, dtype, u, um, it, im, pt, saved = _state(ctx)
        L = lib()
        B, dev = u.shape[0], u.device
        views, gptrs, flat = _param_grads(ctx, pt, used)
        work = _bytes(L.mmoe_cross_workspace_bytes(C.byref(cfg), B, dtype), dev)
        d_user, d_item = torch.empty_like(u), torch.empty_like(it)
        do = _f32c(dout)
        c = _call(dtype, B, training, home, drop_p, seed, pt, gptrs, saved, work)
        check(L.mmoe_cross_bwd(C.byref(c), C.byref(cfg), u.data_ptr(), um.data_ptr(), it.data_ptr(), im.data_ptr(), do.data_ptr(),
                               d_user.data_ptr(), d_item.data_ptr()), "cross_bwd")
        _sync_grads(flat)
        grads = _finish_params(ctx, views, flat)
        ctx.state = ctx.params = None
        return (None, None, None, None, None, None, d_user.to(ctx.in_dtypes[0]), None, d_item.to(ctx.in_dtypes[1]), None, *grads)




# ----------------------------------------------------------------------------------------------
# RobustTextCrossExpert with a staged backward.
# Forward is one library call (made by the tail stage); backward is split into autograd nodes — tail, then one per
# encoder layer — so that each node's parameter gradients reach DDP's reducer as soon as that stage has been
# enqueued and their bucketed all-reduce overlaps the stages still to come.  Nodes are created user stack first,
# item stack second, tail last, hence run tail -> item layers -> user layers: DDP's bucket order.
# ----------------------------------------------------------------------------------------------
class _CrossRun:
    """State shared by the stage nodes of one forward call."""
    __slots__ = ("cfg", "home", "training", "drop_p", "seed", "dtype", "u", "um", "it", "im", "pt", "saved",
                 "views", "gptrs", "flat", "work", "d_user", "d_item", "stages_left")


def _run_of(ctx):
    if not ctx.holder:
        raise RuntimeError("mmoe_b200: backward through a RobustTextCrossExpert call whose buffers were already released "
                           "(a second backward over the same graph is not supported)")
    return ctx.holder[0]


def _stage_done(holder, run):
    """The last stage of a backward pass releases the activations, the workspace and the gradient buffer references."""
    run.stages_left -= 1
    if run.stages_left <= 0:
        holder.clear()


def _cross_stage_call(run: "_CrossRun", stage: int, dout_ptr):
    L = lib()
    cfg = run.cfg
    B = run.u.shape[0]
    c = _call(run.dtype, B, run.training, run.home, run.drop_p, run.seed, run.pt, run.gptrs, run.saved, run.work)
    check(L.mmoe_cross_bwd_stage(C.byref(c), C.byref(cfg), stage, run.u.data_ptr(), run.um.data_ptr(), run.it.data_ptr(),
                                 run.im.data_ptr(), dout_ptr, run.d_user.data_ptr(), run.d_item.data_ptr()), "cross_bwd_stage")


class CrossLayerStage(torch.autograd.Function):
    """Autograd node of one encoder layer of the user (stage 100+l) or item (200+l) stack."""

    @staticmethod
    def forward(ctx, holder, stage: int, idx0: int, x, *layer_params):
        ctx.holder, ctx.stage, ctx.idx0 = holder, stage, idx0
        ctx.params = layer_params
        ctx.param_req = [p.requires_grad for p in layer_params]
        ctx.is_first = (stage % 100) == 0
        ctx.x_dtype = x.dtype
        return x.new_empty(0)                      # token: only carries the graph edge to the next stage

    @staticmethod
    @once_differentiable
    def backward(ctx, _tok_grad):
        run = _run_of(ctx)
        _cross_stage_call(run, ctx.stage, None)
        lo, hi = ctx.idx0, ctx.idx0 + len(ctx.param_req)
        span = _span(run.flat, run.views, lo, hi)                     # this layer's 12 gradients
        _sync_grads(span)
        # the views leave `run` with this stage: nothing else may keep a reference to a gradient that autograd (or the
        # exchange) now owns — AccumulateGrad only adopts a tensor nobody else holds
        views = run.views[lo:hi]
        run.views[lo:hi] = [None] * (hi - lo)
        grads = _hand_over(ctx.params, views, ctx.param_req, span)
        del views
        if ctx.is_first:
            dx = (run.d_user if ctx.stage < 200 else run.d_item).to(ctx.x_dtype)
        else:
            dx = _tok_grad.new_empty(0)
        ctx.params = None
        _stage_done(ctx.holder, run)
        return (None, None, None, dx, *grads)


class CrossTailStage(torch.autograd.Function):
    """Runs the whole forward; its backward node covers MLP, LayerNorm, pooling, gate mix and the cross attention."""

    @staticmethod
    def forward(ctx, holder, pack: ParamPack, cfg: CrossCfg, home: bool, used, training: bool, drop_p: float, all_params,
                tail_idx, user, user_mask, item, item_mask, tok_user, tok_item, *tail_params):
        _require_cuda(user, user_mask, item, item_mask)
        L = lib()
        run = _CrossRun()
        run.cfg, run.home, run.training, run.drop_p = cfg, home, training, drop_p
        run.dtype = compute_dtype()
        run.u, run.it = _f32c(user), _f32c(item)
        run.um = user_mask.detach().to(torch.bool).contiguous().view(torch.uint8)
        run.im = item_mask.detach().to(torch.bool).contiguous().view(torch.uint8)
        B, dev = run.u.shape[0], run.u.device
        run.pt = pack.tensors(all_params, run.dtype)
        run.saved = _bytes(L.mmoe_cross_saved_bytes(C.byref(cfg), B, run.dtype), dev)
        run.seed = _new_seed(training, drop_p)
        run.views = run.gptrs = run.flat = run.work = run.d_user = run.d_item = None
        out = torch.empty((B, cfg.d), dtype=torch.float32, device=dev)
        c = _call(run.dtype, B, training, home, drop_p, run.seed, run.pt, None, run.saved, None)
        check(L.mmoe_cross_fwd(C.byref(c), C.byref(cfg), run.u.data_ptr(), run.um.data_ptr(), run.it.data_ptr(), run.im.data_ptr(),
                               out.data_ptr()), "cross_fwd")
        if DEBUG_SAVED is not None:
            DEBUG_SAVED.append(("cross", cfg, home, B, run.dtype, run.saved))
        holder.append(run)
        run.stages_left = 2 * cfg.n_layer + 1
        ctx.holder, ctx.used, ctx.tail_idx = holder, list(used), list(tail_idx)
        ctx.params = tail_params
        ctx.param_req = [p.requires_grad for p in tail_params]
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        run = _run_of(ctx)
        L = lib()
        B, dev = run.u.shape[0], run.u.device
        run.views, run.gptrs, run.flat = _alloc_grads(run.pt, ctx.used)
        run.work = _bytes(L.mmoe_cross_workspace_bytes(C.byref(run.cfg), B, run.dtype), dev)
        run.d_user, run.d_item = torch.empty_like(run.u), torch.empty_like(run.it)
        do = _f32c(dout)
        _cross_stage_call(run, 0, do.data_ptr())
        # gate (first parameter) and everything after the two encoder stacks are final after this stage
        n_enc = 24 * run.cfg.n_layer
        n_all = len(run.views)
        grads = [None] * len(ctx.tail_idx)
        pos = {i: j for j, i in enumerate(ctx.tail_idx)}
        for lo, hi in ((0, 1), (1 + n_enc, n_all)):
            span = _span(run.flat, run.views, lo, hi)
            _sync_grads(span)
            idx = [i for i in range(lo, hi) if i in pos]
            views = [run.views[i] for i in idx]
            run.views[lo:hi] = [None] * (hi - lo)
            out = _hand_over([ctx.params[pos[i]] for i in idx], views, [ctx.param_req[pos[i]] for i in idx], span)
            del views
            for i, g in zip(idx, out):
                grads[pos[i]] = g
        empty = dout.new_empty(0)
        ctx.params = None
        _stage_done(ctx.holder, run)
        return (None,) * 9 + (None, None, None, None, empty, empty, *grads)


def cross_expert_staged(pack: ParamPack, cfg: CrossCfg, home: bool, used, training: bool, drop_p: float,
                        user, user_mask, item, item_mask, params):
    """Builds the stage chain for one forward call.  `params` in state_dict order:
    [gate, self_user.0 (12) ..., self_item.0 (12) ..., cross_attn (4), pool.query, norm (2), mlp (4)]."""
    n = cfg.n_layer
    holder = []
    tok_u = user
    for l in range(n):
        i0 = 1 + 12 * l
        tok_u = CrossLayerStage.apply(holder, 100 + l, i0, tok_u, *params[i0:i0 + 12])
    tok_i = item
    for l in range(n):
        i0 = 1 + 12 * n + 12 * l
        tok_i = CrossLayerStage.apply(holder, 200 + l, i0, tok_i, *params[i0:i0 + 12])
    tail_idx = [0] + list(range(1 + 24 * n, len(params)))
    # unused HoME parameters (norm / mlp) are left out of the node so that their .grad stays None
    tail_idx = [i for i in tail_idx if used[i]]
    return CrossTailStage.apply(holder, pack, cfg, home, used, training, drop_p, list(params), tail_idx,
                                user, user_mask, item, item_mask, tok_u, tok_i, *[params[i] for i in tail_idx])


# ----------------------------------------------------------------------------------------------
# EnhancedCrossFuse
# ----------------------------------------------------------------------------------------------
class FuseFn(torch.autograd.Function):
    """out[B,d] = fuse(v_cls[B,d], t_cls[B,d]) — model.py:491-507 (home=True: model_HoME.py:506-522)."""

    @staticmethod
    def forward(ctx, pack: ParamPack, cfg: FuseCfg, home: bool, used: Sequence[bool], training: bool, drop_p: float,
                v_cls, t_cls, *params):
        _require_cuda(v_cls, t_cls)
        L = lib()
        dtype = compute_dtype()
        v, t = _f32c(v_cls), _f32c(t_cls)
        B, dev = v.shape[0], v.device
        params, ctx.flat_p = _split_fused(pack, params)
        pt = pack.tensors(params, dtype, ctx.flat_p)
        saved = _bytes(L.mmoe_fuse_saved_bytes(C.byref(cfg), B, dtype), dev)
        out = torch.empty((B, cfg.d), dtype=torch.float32, device=dev)
        seed = _new_seed(training, drop_p)
        c = _call(dtype, B, training, home, drop_p, seed, pt, None, saved, None)
        check(L.mmoe_fuse_fwd(C.byref(c), C.byref(cfg), v.data_ptr(), t.data_ptr(), out.data_ptr()), "fuse_fwd")
        if DEBUG_SAVED is not None:
            DEBUG_SAVED.append(("fuse", cfg, home, B, dtype, saved))
        ctx.state = (cfg, home, list(used), training, drop_p, seed, dtype, B, dev, pt, saved)
        ctx.params = params
        ctx.param_req = [p.requires_grad for p in params]
        ctx.in_dtypes = (v_cls.dtype, t_cls.dtype)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        cfg, home, used, training, drop_p, seed, dtype, B, dev, pt, saved = _state(ctx)
        L = lib()
        views, gptrs, flat = _param_grads(ctx, pt, used)
        work = _bytes(L.mmoe_fuse_workspace_bytes(C.byref(cfg), B, dtype), dev)
        d_cat = torch.empty((B, 2, cfg.d), dtype=torch.float32, device=dev)
        do = _f32c(dout)
        c = _call(dtype, B, training, home, drop_p, seed, pt, gptrs, saved, work)
        check(L.mmoe_fuse_bwd(C.byref(c), C.byref(cfg), do.data_ptr(), d_cat.data_ptr()), "fuse_bwd")
        _sync_grads(flat)
        grads = _finish_params(ctx, views, flat)
        ctx.state = ctx.params = None
        return (None, None, None, None, None, None, d_cat[:, 0].to(ctx.in_dtypes[0]), d_cat[:, 1].to(ctx.in_dtypes[1]), *grads)


# ----------------------------------------------------------------------------------------------
# ItemImageExpert tail (pool + LN + dropout) and the HoME projection head
# ----------------------------------------------------------------------------------------------
class ImgPoolFn(torch.autograd.Function):
    """img_vec[B,d] = dropout(LN(pool(tokens[B,n_tok,d]))) — model.py:377-385."""

    @staticmethod
    def forward(ctx, pool_cls: bool, training: bool, drop_p: float, tokens, gamma, beta):
        _require_cuda(tokens, gamma, beta)
        L = lib()
        tk = tokens.detach()
        if tk.dtype not in _TORCH2MMOE:
            tk = tk.float()
        tk = tk.contiguous()
        B, n_tok, d = tk.shape
        dev = tk.device
        g, b = _f32c(gamma), _f32c(beta)
        out = torch.empty((B, d), dtype=torch.float32, device=dev)
        stats = torch.empty((B, 2), dtype=torch.float32, device=dev)
        pooled = torch.empty((B, d), dtype=torch.float32, device=dev)
        seed = _new_seed(training, drop_p)
        c = _call(F32, B, training, 0, drop_p, seed, [g, b], None, None, None)
        check(L.mmoe_img_pool_fwd(C.byref(c), tk.data_ptr(), _TORCH2MMOE[tk.dtype], n_tok, d, int(pool_cls), out.data_ptr(),
                                  stats.data_ptr(), pooled.data_ptr()), "img_pool_fwd")
        ctx.state = (pool_cls, training, drop_p, seed, g, b, stats, pooled, tk.dtype, tuple(tk.shape), tokens.dtype)
        ctx.params = (gamma, beta)
        ctx.needs_tokens = tokens.requires_grad
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        pool_cls, training, drop_p, seed, g, b, stats, pooled, tk_dtype, shape, in_dtype = _state(ctx)
        L = lib()
        B, n_tok, d = shape
        dev = dout.device
        flat = torch.zeros(g.numel() + b.numel(), dtype=torch.float32, device=dev)
        dg, db = flat[:g.numel()].view(g.shape), flat[g.numel():].view(b.shape)
        d_tok = torch.empty(shape, dtype=tk_dtype, device=dev) if ctx.needs_tokens else None
        do = _f32c(dout)
        c = _call(F32, B, training, 0, drop_p, seed, [g, b], [dg.data_ptr(), db.data_ptr()], None, None)
        check(L.mmoe_img_pool_bwd(C.byref(c), n_tok, d, int(pool_cls), stats.data_ptr(), pooled.data_ptr(), do.data_ptr(),
                                  d_tok.data_ptr() if d_tok is not None else None, _TORCH2MMOE[tk_dtype]), "img_pool_bwd")
        _sync_grads(flat)
        dg, db = _hand_over(ctx.params, [dg, db], [p.requires_grad for p in ctx.params], flat)
        ctx.state = ctx.params = None
        return None, None, None, (d_tok.to(in_dtype) if d_tok is not None else None), dg, db


class ImgProjFn(torch.autograd.Function):
    """projected[B,proj] = Linear(GELU(Linear(img_vec))) — model_HoME.py:383-387,397."""

    @staticmethod
    def forward(ctx, pack: ParamPack, d: int, proj: int, img_vec, *params):
        _require_cuda(img_vec)
        L = lib()
        dtype = compute_dtype()
        x = _f32c(img_vec)
        B, dev = x.shape[0], x.device
        pt = pack.tensors(params, dtype)
        saved = _bytes(L.mmoe_img_proj_saved_bytes(B, d, proj, dtype), dev)
        out = torch.empty((B, proj), dtype=torch.float32, device=dev)
        c = _call(dtype, B, False, 0, 0.0, 0, pt, None, saved, None)
        check(L.mmoe_img_proj_fwd(C.byref(c), d, proj, x.data_ptr(), out.data_ptr()), "img_proj_fwd")
        ctx.state = (d, proj, dtype, B, dev, pt, saved)
        ctx.params = params
        ctx.param_req = [p.requires_grad for p in params]
        ctx.in_dtype = img_vec.dtype
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        d, proj, dtype, B, dev, pt, saved = _state(ctx)
        L = lib()
        views, gptrs, flat = _alloc_grads(pt, [True] * len(pt))
        work = _bytes(L.mmoe_img_proj_workspace_bytes(B, d, proj, dtype), dev)
        dx = torch.empty((B, d), dtype=torch.float32, device=dev)
        do = _f32c(dout)
        c = _call(dtype, B, False, 0, 0.0, 0, pt, gptrs, saved, work)
        check(L.mmoe_img_proj_bwd(C.byref(c), d, proj, do.data_ptr(), dx.data_ptr()), "img_proj_bwd")
        _sync_grads(flat)
        grads = _hand_over(ctx.params, views, ctx.param_req, flat)
        ctx.state = ctx.params = None
        return (None, None, None, dx.to(ctx.in_dtype), *grads)
