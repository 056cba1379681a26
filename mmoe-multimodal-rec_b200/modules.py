"""B200-native drop-ins for the fusion-and-head modules of the reference's ``model.py``.

Same class names, constructor arguments, forward signatures and ``state_dict()`` keys as
/root/reference/model.py, so the reference's train.py / inference_and_auc.py
and its published checkpoints work unchanged.  Parameters live in ordinary torch containers
(nn.Linear, nn.LayerNorm, nn.MultiheadAttention, nn.TransformerEncoderLayer) that are never
*called*: every forward hands their tensors to libmmoe_b200.so through functional.py.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import functional as Fn
from ._lib import CrossCfg, FuseCfg, HeadCfg

# when autograd is off (scoring sweeps, config 5) large batches are processed in slices so the
# activation blob stays bounded
NO_GRAD_CHUNK = 4096


def _named(module: nn.Module):
    names, params = [], []
    for n, p in module.named_parameters():
        names.append(n)
        params.append(p)
    return names, params


def _is_gemm_weight(name: str, p: torch.Tensor, exclude=()) -> bool:
    return p.dim() == 2 and not any(name.endswith(e) or name == e for e in exclude)


def _standalone(module, owner: str):
    """Parameter containers that are evaluated inside their owner's fused launch sequence have no launch of their own: a
    direct call fails loudly instead of dropping into torch eager (the package has no eager path)."""
    raise NotImplementedError(
        f"{type(module).__name__} holds parameters of {owner} and is executed inside that module's native kernels; the "
        f"B200-native package has no stand-alone (torch eager) forward for it.  Call the owning module.")


class _NativeMeta(type):
    """Runs the fused-parameter step after the most-derived __init__ has returned (the reference scripts construct the
    modules and hand them straight to .to(device) / DistributedDataParallel, so there is no later hook to do it in)."""

    def __call__(cls, *args, **kwargs):
        obj = super().__call__(*args, **kwargs)
        if cls._fusable and Fn.FLAT_PARAMS:
            obj.fuse_parameters()
        return obj


_FLAT = "_flat_param"


class _Native(nn.Module, metaclass=_NativeMeta):
    """Shared plumbing: parameter pack in state_dict order + chunked no-grad execution + fused-parameter mode.

    Fused-parameter mode (``fuse_parameters()``; automatic for modules built under ``MMOE_FLAT_PARAMS=1`` /
    ``functional.set_flat_parameters(True)``): every parameter the forward uses moves into ONE ``nn.Parameter``
    (``_flat_param``, laid out exactly like the flat gradient buffer the backward kernels write), and the reference-named
    attributes (``mlp[0].weight`` ...) become plain tensor views of it.  ``parameters()`` then yields one tensor per
    module — which is what the reference scripts hand to AdamW / clip_grad_norm_ / DistributedDataParallel
    (train.py:159-167, 136-139) — so DDP copies one gradient per module in and out of its buckets instead of ~60, the
    optimizer updates one tensor, and ``.grad`` is the buffer the kernels wrote, never re-packed.  ``state_dict()`` /
    ``load_state_dict()`` keep the reference's keys, order and shapes (hooks below), so checkpoints are interchangeable
    with the unfused modules and the reference.  What changes for callers: ``named_parameters()`` shows ``_flat_param``
    instead of the per-tensor names (per-name parameter groups need the unfused mode), per-tensor ``.grad`` attributes do
    not exist (slice ``_flat_param.grad`` with ``fused_layout()``), and the 64-float padding between tensors is part of
    the parameter (always zero, zero gradient).  Parameters the HoME variants never use stay ordinary nn.Parameters
    whose ``.grad`` stays None, as in the reference."""

    _lowp_exclude = ()
    _fusable = False

    def _used(self, names):
        return [True] * len(names)

    def _pack(self):
        pk = self.__dict__.get("_mmoe_pack")
        if pk is None:
            names, params = self._names_and_tensors()
            lowp = [n for n, p in zip(names, params) if _is_gemm_weight(n, p, self._lowp_exclude)]
            pk = Fn.ParamPack(names, lowp)
            pk.fused = "_mmoe_fused" in self.__dict__
            self.__dict__["_mmoe_pack"] = pk
        return pk

    def _names_and_tensors(self):
        """(names in state_dict order, the tensor behind each name: nn.Parameter, or view of the fused parameter)."""
        specs = self.__dict__.get("_mmoe_fused")
        if specs is None:
            return _named(self)
        flat = self._parameters[_FLAT]
        if flat.data_ptr() != self.__dict__.get("_mmoe_flat_ptr") or flat.device != self.__dict__.get("_mmoe_flat_dev"):
            self._rebuild_views()            # .to() / .data swap / deepcopy moved the storage
        rest = dict(self.named_parameters())
        fused = {n: owner.__dict__[leaf] for n, owner, leaf, _ in specs}
        names = self.__dict__["_mmoe_names"]
        return names, [fused[n] if n in fused else rest[n] for n in names]

    def _params(self):
        # cached: walking named_parameters() of a 60-tensor module costs ~0.1 ms per call, twice per step and module.
        # Parameter objects keep their identity through .to() / load_state_dict(); _apply drops the cache anyway.
        # Fused mode: the per-name views followed by the fused parameter itself (functional._split_fused).
        ps = self.__dict__.get("_mmoe_params")
        if "_mmoe_fused" in self.__dict__:
            flat = self._parameters[_FLAT]
            if ps is None or flat.data_ptr() != self.__dict__.get("_mmoe_flat_ptr") or ps[-1] is not flat:
                ps = self._names_and_tensors()[1] + [flat]
                self.__dict__["_mmoe_params"] = ps
            return ps
        if ps is None:
            ps = [p for _, p in self.named_parameters()]
            self.__dict__["_mmoe_params"] = ps
        return ps

    def invalidate_weight_cache(self):
        """Forget the cached 16-bit copies of the GEMM weights.  They are refreshed automatically whenever a parameter's
        version counter moves (optimizer steps, load_state_dict, any in-place op on the parameter); writes that bypass the
        counter — ``param.data.mul_(...)``-style updates — need this call."""
        self.__dict__.pop("_mmoe_params", None)
        self.__dict__.pop("_mmoe_pack", None)

    def _apply(self, fn, *args, **kwargs):
        self.__dict__.pop("_mmoe_params", None)
        self.__dict__.pop("_mmoe_pack", None)
        out = super()._apply(fn, *args, **kwargs)
        if "_mmoe_fused" in self.__dict__:
            self._rebuild_views()
        return out

    # ------------------------------------------------------------------ fused-parameter mode
    def fuse_parameters(self):
        """Move every used parameter into one nn.Parameter (see the class docstring).  Idempotent; returns self."""
        if "_mmoe_fused" in self.__dict__:
            return self
        names, params = _named(self)
        used = self._used(names)
        live = [p for p, u in zip(params, used) if u]
        if not live:
            return self
        if any(p.dtype != torch.float32 for p in live) or len({p.device for p in live}) != 1:
            raise RuntimeError("fuse_parameters: parameters must be float32 and on one device")
        if len({p.requires_grad for p in live}) != 1:
            raise RuntimeError("fuse_parameters: parameters must be all trainable or all frozen")
        if len({id(p) for p in live}) != len(live):
            raise RuntimeError("fuse_parameters: tied parameters are not supported")
        sd_order = list(self.state_dict().keys())
        total, entries = Fn._grad_layout(params, used)
        flat = torch.zeros(total, dtype=torch.float32, device=live[0].device)
        specs = []
        with torch.no_grad():
            for n, p, e in zip(names, params, entries):
                if e is None:
                    continue
                flat.as_strided(e[1], e[2], e[0]).copy_(p)
                owner, _, leaf = n.rpartition(".")
                specs.append((n, self.get_submodule(owner) if owner else self, leaf, e))
        for _, owner, leaf, _ in specs:
            del owner._parameters[leaf]
        self.register_parameter(_FLAT, nn.Parameter(flat, requires_grad=live[0].requires_grad))
        self.__dict__["_mmoe_fused"] = specs
        self.__dict__["_mmoe_names"] = names
        self.__dict__["_mmoe_sd_order"] = sd_order
        self._rebuild_views()
        self._register_state_dict_hook(_fused_state_dict_hook)
        self.register_load_state_dict_pre_hook(_fused_load_pre_hook)
        return self

    def _rebuild_views(self):
        flat = self._parameters[_FLAT]
        base = flat.detach()                     # shares storage AND version counter: an optimizer step on the fused
        for _, owner, leaf, e in self.__dict__["_mmoe_fused"]:      # parameter invalidates the 16-bit weight cache
            owner.__dict__[leaf] = base.as_strided(e[1], e[2], e[0])
        self.__dict__["_mmoe_flat_ptr"] = flat.data_ptr()
        self.__dict__["_mmoe_flat_dev"] = flat.device
        self.__dict__.pop("_mmoe_params", None)
        self.__dict__.pop("_mmoe_pack", None)

    def fused_layout(self):
        """{name: (offset, shape)} of the reference-named tensors inside ``_flat_param`` (and its ``.grad``); None when unfused."""
        specs = self.__dict__.get("_mmoe_fused")
        if specs is None:
            return None
        return {n: (e[0], e[1]) for n, _, _, e in specs}

    def named_gradients(self):
        """(reference name, gradient or None) for every parameter, fused or not — the per-name view of ``.grad``."""
        specs = self.__dict__.get("_mmoe_fused")
        if specs is None:
            return [(n, p.grad) for n, p in self.named_parameters()]
        g = self._parameters[_FLAT].grad
        fused = {n: e for n, _, _, e in specs}
        rest = dict(self.named_parameters())
        return [(n, (None if g is None else g.as_strided(fused[n][1], fused[n][2], fused[n][0])) if n in fused else rest[n].grad)
                for n in self.__dict__["_mmoe_names"]]

    _CACHE_KEYS = ("_mmoe_params", "_mmoe_pack", "_mmoe_flat_ptr", "_mmoe_flat_dev")

    def __getstate__(self):
        # torch.save(module) / pickle: the per-call caches hold ctypes pointer arrays and device addresses of this process
        st = self.__dict__.copy()
        for k in self._CACHE_KEYS:
            st.pop(k, None)
        return st

    def __deepcopy__(self, memo):
        # the default deep copy keeps the aliasing of tensors that share a storage, but the per-name views live in the
        # sub-modules' __dict__ next to cached pointers: rebuild them on the copy rather than rely on that
        import copy
        cls = self.__class__
        new = cls.__new__(cls)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            if k not in self._CACHE_KEYS:
                new.__dict__[k] = copy.deepcopy(v, memo)
        if "_mmoe_fused" in new.__dict__:
            new._rebuild_views()
        return new


def _fused_state_dict_hook(module, state_dict, prefix, local_metadata):
    specs = module.__dict__.get("_mmoe_fused")
    if specs is None:
        return
    state_dict.pop(prefix + _FLAT, None)
    if module._parameters[_FLAT].data_ptr() != module.__dict__.get("_mmoe_flat_ptr"):
        module._rebuild_views()
    for n, owner, leaf, _ in specs:
        state_dict[prefix + n] = owner.__dict__[leaf]
    for k in module.__dict__["_mmoe_sd_order"]:          # the reference's key order
        if prefix + k in state_dict:
            state_dict.move_to_end(prefix + k)


def _fused_load_pre_hook(module, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
    specs = module.__dict__.get("_mmoe_fused")
    if specs is None:
        return
    flat = module._parameters[_FLAT]
    if flat.data_ptr() != module.__dict__.get("_mmoe_flat_ptr"):
        module._rebuild_views()
    own = state_dict.pop(prefix + _FLAT, None)           # a raw fused tensor is accepted only with the exact layout
    with torch.no_grad():
        if own is not None and tuple(own.shape) == tuple(flat.shape):
            flat.copy_(own)
        for n, owner, leaf, e in specs:
            key = prefix + n
            if key not in state_dict:
                if own is None:
                    missing_keys.append(key)
                continue
            src = state_dict.pop(key)
            if tuple(src.shape) != tuple(e[1]):
                error_msgs.append(f"size mismatch for {key}: copying a param with shape {tuple(src.shape)} from checkpoint, "
                                  f"the shape in current model is {tuple(e[1])}.")
                continue
            owner.__dict__[leaf].copy_(src)
    state_dict[prefix + _FLAT] = flat.detach()           # satisfies the strict-key check of the module's own loader


# ----------------------------------------------------------------------------------------------
class AttnPool1D(nn.Module):
    """Learned-query attention pooling (reference model.py:192-206).  Inside
    RobustTextCrossExpert it is executed by the fused gate-mix + pooling kernel; it holds the
    ``query`` parameter and the dropout probability."""

    def __init__(self, d, dropout=0.1):
        super().__init__()
        self.query = nn.Parameter(torch.randn(1, 1, d) * (d ** -0.5))
        self.dropout = nn.Dropout(dropout)

    def forward(self, x, mask):
        _standalone(self, "RobustTextCrossExpert (fused gate-mix + pooling kernel)")


class RobustTransformerLayer(nn.TransformerEncoderLayer):
    """Pre-LN encoder layer container (reference model.py:207-212).  Never called as a torch
    module on the hot path: the owning expert runs it through the encoder-layer launch sequence."""


class RobustTextCrossExpert(_Native):
    """Sentence-level user<->item cross-attention expert (reference model.py:386-451)."""

    _home = False
    _lowp_exclude = ()
    _fusable = True

    def __init__(self, d=768, n_layer=2, n_head=8, dropout=0.1):
        super().__init__()
        make = lambda: RobustTransformerLayer(d_model=d, nhead=n_head, dim_feedforward=4 * d, dropout=dropout,
                                              batch_first=True, norm_first=True)
        self.self_user = nn.ModuleList([make() for _ in range(n_layer)])
        self.self_item = nn.ModuleList([make() for _ in range(n_layer)])
        self.cross_attn = nn.MultiheadAttention(d, n_head, dropout=dropout, batch_first=True)
        self.gate = nn.Parameter(torch.tensor([0.5]))
        self.pool = AttnPool1D(d, dropout)
        self.norm = nn.LayerNorm(d)
        self.mlp = nn.Sequential(nn.Linear(d, 4 * d), nn.GELU(), nn.Dropout(dropout), nn.Linear(4 * d, d), nn.Dropout(dropout))
        self._cfg_tuple = (d, n_head, n_layer)
        self._drop_p = float(dropout)

    def _used(self, names):
        if not self._home:
            return [True] * len(names)
        return [not (n.startswith("norm.") or n.startswith("mlp.")) for n in names]

    def forward(self, user_vecs, user_mask, item_vecs, item_mask):
        d, n_head, n_layer = self._cfg_tuple
        B, S = user_vecs.shape[0], user_vecs.shape[1]
        if item_vecs.shape[1] != S:
            raise RuntimeError("user and item must have the same number of sentence slots")
        cfg = CrossCfg(d, S, n_head, n_layer)
        pack = self._pack()
        params = self._params()
        used = self._used(pack.names)

        def run(u, um, i, im):
            if torch.is_grad_enabled() and Fn.CROSS_STAGED and not pack.fused:
                # staged backward: parameter gradients become available layer by layer (DDP overlap)
                return Fn.cross_expert_staged(pack, cfg, self._home, used, self.training, self._drop_p, u, um, i, im, params)
            return Fn.CrossFn.apply(pack, cfg, self._home, used, self.training, self._drop_p, u, um, i, im, *params)

        if not torch.is_grad_enabled() and B > NO_GRAD_CHUNK:
            return torch.cat([run(user_vecs[s:s + NO_GRAD_CHUNK], user_mask[s:s + NO_GRAD_CHUNK],
                                  item_vecs[s:s + NO_GRAD_CHUNK], item_mask[s:s + NO_GRAD_CHUNK])
                              for s in range(0, B, NO_GRAD_CHUNK)], 0)
        return run(user_vecs, user_mask, item_vecs, item_mask)


class EnhancedCrossFuse(_Native):
    """Two-token cross-modal fuse expert (reference model.py:454-507)."""

    _home = False
    _lowp_exclude = ("gate.2.weight",)       # the 384->1 gate row is consumed by a GEMV kernel in fp32
    _fusable = True

    def __init__(self, d=768, n_head=8, depth=2, dropout=0.1):
        super().__init__()
        self.layers = nn.ModuleList([
            nn.TransformerEncoderLayer(d_model=d, nhead=n_head, dim_feedforward=4 * d, dropout=dropout, batch_first=True,
                                       norm_first=True) for _ in range(depth)])
        self.res_proj = nn.Sequential(nn.Linear(2 * d, d), nn.LayerNorm(d))
        self.gate = nn.Sequential(nn.Linear(2 * d, d // 2), nn.GELU(), nn.Linear(d // 2, 1), nn.Sigmoid())
        nn.init.constant_(self.gate[2].bias, 0.5)
        self.proj = nn.Sequential(nn.LayerNorm(d), nn.Linear(d, d), nn.GELU(), nn.Dropout(dropout))
        self._cfg_tuple = (d, n_head, depth)
        self._drop_p = float(dropout)

    def _used(self, names):
        if not self._home:
            return [True] * len(names)
        return [not n.startswith("proj.") for n in names]

    def forward(self, v_cls, t_cls):
        d, n_head, depth = self._cfg_tuple
        cfg = FuseCfg(d, n_head, depth)
        pack = self._pack()
        params = self._params()
        used = self._used(pack.names)
        B = v_cls.shape[0]

        def run(v, t):
            return Fn.FuseFn.apply(pack, cfg, self._home, used, self.training, self._drop_p, v, t, *params)

        if not torch.is_grad_enabled() and B > 8 * NO_GRAD_CHUNK:
            step = 8 * NO_GRAD_CHUNK
            return torch.cat([run(v_cls[s:s + step], t_cls[s:s + step]) for s in range(0, B, step)], 0)
        return run(v_cls, t_cls)


class DenseGate(nn.Module):
    """Dense softmax gate (reference model.py:513-524).  As a sub-module of a head it only holds
    ``fc``; the head's fused gate/mix kernel evaluates it.  Called on its own (e.g. to inspect
    the gate argmax) it runs that same kernel on a single pseudo-expert and returns the weights."""

    def __init__(self, in_dim: int, n_expert: int):
        super().__init__()
        self.fc = nn.Linear(in_dim, n_expert)

    def forward(self, x: torch.Tensor):
        from .gate_only import dense_gate_forward
        return dense_gate_forward(x, self.fc.weight, self.fc.bias)


class TwoTaskMMoE(_Native):
    """Two-task dense-gate MMoE head (reference model.py:527-577)."""

    _lowp_exclude = ("fc.weight", "7.weight")   # gate rows and the 128->1 row stay fp32 (GEMV kernels)
    _fusable = True

    def __init__(self, expert_dim: int = 768, n_expert: int = 6, tower_hidden: int = 256, tower_dropout: float = 0):
        super().__init__()
        self.gate_good = DenseGate(expert_dim, n_expert)
        self.gate_best = DenseGate(expert_dim, n_expert)

        def tower():
            return nn.Sequential(nn.LayerNorm(expert_dim), nn.Linear(expert_dim, tower_hidden), nn.GELU(), nn.Dropout(tower_dropout),
                                 nn.Linear(tower_hidden, tower_hidden // 2), nn.GELU(), nn.Dropout(tower_dropout),
                                 nn.Linear(tower_hidden // 2, 1))

        self.tower_good = tower()
        self.tower_best = tower()
        self._cfg_tuple = (expert_dim, n_expert, tower_hidden, float(tower_dropout))

    def _run(self, expert_vecs, want_gates):
        d, n, h, p = self._cfg_tuple
        if expert_vecs.dim() != 3 or expert_vecs.shape[1] != n or expert_vecs.shape[2] != d:
            raise RuntimeError(f"expert_vecs must be [B,{n},{d}], got {tuple(expert_vecs.shape)}")
        cfg = HeadCfg(d, n, h, p)
        return Fn.HeadFn.apply(self._pack(), "mmoe", cfg, self.training, p, want_gates, expert_vecs, *self._params())

    def forward(self, expert_vecs: torch.Tensor):
        logits = self._run(expert_vecs, False)
        return logits[0], logits[1]

    def gate_weights(self, expert_vecs: torch.Tensor):
        """(w_good, w_best), each [B, n_expert]: the DenseGate outputs the fused kernel used."""
        _, gates = self._run(expert_vecs, True)
        return gates[0], gates[1]


class ItemImageExpert(nn.Module):
    """Image expert wrapper (reference model.py:343-385): HF backbone (unchanged torch module),
    then token mean / CLS -> LayerNorm -> dropout in one native kernel."""

    def __init__(self, base_model: nn.Module, pool_type: str = "mean", dropout_p: float = 0.1):
        super().__init__()
        if pool_type not in ("mean", "cls"):
            raise AssertionError("`pool_type` must be 'mean' or 'cls'")
        self.backbone = base_model
        self.pool_type = pool_type
        self.dropout = nn.Dropout(dropout_p)
        self.norm = nn.LayerNorm(base_model.config.hidden_size)

    def _tokens(self, images):
        if images.dtype == torch.uint8:
            # raw patch bytes [B, 196, 768] as stored on disk (newpatch.py:102-104): no un-patchify / normalise round trip
            from .ingest import vit_tokens_from_patch_bytes
            return vit_tokens_from_patch_bytes(self.backbone, images)
        return self.backbone(pixel_values=images).last_hidden_state

    def forward(self, images: torch.Tensor, trainable: bool = False):
        if trainable:
            tokens = self._tokens(images)
        else:
            with torch.no_grad():
                tokens = self._tokens(images)
        return Fn.ImgPoolFn.apply(self.pool_type == "cls", self.training, float(self.dropout.p), tokens,
                                  self.norm.weight, self.norm.bias)
