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


class _Native(nn.Module):
    """Shared plumbing: parameter pack in state_dict order + chunked no-grad execution."""

    _lowp_exclude = ()

    def _pack(self):
        pk = self.__dict__.get("_mmoe_pack")
        if pk is None:
            names, params = _named(self)
            lowp = [n for n, p in zip(names, params) if _is_gemm_weight(n, p, self._lowp_exclude)]
            pk = Fn.ParamPack(names, lowp)
            self.__dict__["_mmoe_pack"] = pk
        return pk

    def _params(self):
        # cached: walking named_parameters() of a 60-tensor module costs ~0.1 ms per call, twice per step and module.
        # Parameter objects keep their identity through .to() / load_state_dict(); _apply drops the cache anyway.
        ps = self.__dict__.get("_mmoe_params")
        if ps is None:
            ps = [p for _, p in self.named_parameters()]
            self.__dict__["_mmoe_params"] = ps
        return ps

    def _apply(self, fn, *args, **kwargs):
        self.__dict__.pop("_mmoe_params", None)
        self.__dict__.pop("_mmoe_pack", None)
        return super()._apply(fn, *args, **kwargs)


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
            if torch.is_grad_enabled() and Fn.CROSS_STAGED:
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
