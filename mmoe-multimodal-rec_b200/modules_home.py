"""B200-native drop-ins for the fusion-and-head modules of the reference's ``model_HoME.py``
(same names / constructor arguments / forward signatures / state_dict keys)."""
from __future__ import annotations

import torch
import torch.nn as nn

from . import functional as Fn
from . import modules as M
from ._lib import HomeCfg
from .modules import AttnPool1D, DenseGate, RobustTransformerLayer, _Native  # noqa: F401  (re-exported)


def ExpertMLP(expert_dim: int, hidden_dim: int = 1024, dropout_p: float = 0.1):
    """Parameter container of one HoME expert — a factory FUNCTION returning an nn.Sequential, exactly as in the reference
    (model_HoME.py:28-35), so state-dict keys are 0.weight / 0.bias / 3.weight / 3.bias; the eight experts are evaluated
    together by two grouped GEMM launches inside HOME_MMoE_Complete."""
    return nn.Sequential(nn.Linear(expert_dim, hidden_dim), nn.GELU(), nn.Dropout(dropout_p), nn.Linear(hidden_dim, expert_dim))


class FeatureGate(nn.Module):
    """Container of a FeatureGate (reference model_HoME.py:224-234): x[:,None,:] * 2*sigmoid(gate(x))."""

    def __init__(self, d_model: int, n_experts: int):
        super().__init__()
        self.gate = nn.Linear(d_model, d_model * n_experts)
        self.n_experts = n_experts
        self.d_model = d_model

    def forward(self, x):
        M._standalone(self, "HOME_MMoE_Complete (grouped gate GEMM + fg_apply kernel)")


class SelfGate(nn.Module):
    """Container of a SelfGate (reference model_HoME.py:236-243): x + sigmoid(gate(x)) * y."""

    def __init__(self, d_model: int):
        super().__init__()
        self.gate = nn.Sequential(nn.Linear(d_model, d_model), nn.Sigmoid())

    def forward(self, x_original, x_processed):
        M._standalone(self, "HOME_MMoE_Complete (grouped gate GEMM + sg_apply kernel)")


class RobustTextCrossExpert(M.RobustTextCrossExpert):
    """HoME variant (reference model_HoME.py:401-466): returns the pooled vector; ``norm`` and
    ``mlp`` exist for checkpoint compatibility but take no part (their .grad stays None)."""
    _home = True


class EnhancedCrossFuse(M.EnhancedCrossFuse):
    """HoME variant (reference model_HoME.py:469-522): returns fused + identity; ``proj`` is unused."""
    _home = True


class ImageExpertWithProjection(_Native):
    """ViT CLS vector plus a trainable projection head (reference model_HoME.py:373-399).  The ViT
    stays the HF torch module; the Linear-GELU-Linear head runs on the native GEMM engine."""

    def __init__(self, vit_model: nn.Module, expert_dim: int = 768, projection_dim: int = 768):
        super().__init__()
        self.vit_model = vit_model
        self.projection_head = nn.Sequential(nn.Linear(expert_dim, expert_dim * 2), nn.GELU(), nn.Linear(expert_dim * 2, projection_dim))
        self._dims = (expert_dim, projection_dim)

    def forward(self, images: torch.Tensor):
        img_vec = self.vit_model(pixel_values=images).last_hidden_state[:, 0, :]
        pk = self.__dict__.get("_mmoe_pack")
        names = [n for n, _ in self.projection_head.named_parameters()]
        params = [p for _, p in self.projection_head.named_parameters()]
        if pk is None:
            pk = Fn.ParamPack(names, [n for n, p in zip(names, params) if p.dim() == 2])
            self.__dict__["_mmoe_pack"] = pk
        projected = Fn.ImgProjFn.apply(pk, self._dims[0], self._dims[1], img_vec, *params)
        return img_vec, projected


class HOME_MMoE_Complete(_Native):
    """Hierarchical-expert two-task head (reference model_HoME.py:530-638)."""

    _lowp_exclude = ("fc.weight", "tower_good.4.weight", "tower_best.4.weight")
    _fusable = True

    def __init__(self, num_input_experts: int = 6, expert_dim: int = 768, n_shared_experts: int = 4,
                 n_task_experts: int = 2, tower_hidden: int = 256):
        super().__init__()
        self.num_input_experts = num_input_experts
        self.expert_dim = expert_dim
        self.input_projection = nn.Sequential(nn.Linear(num_input_experts * expert_dim, expert_dim), nn.LayerNorm(expert_dim), nn.GELU())
        self.meta_experts = nn.ModuleList([ExpertMLP(expert_dim) for _ in range(n_shared_experts)])
        self.task_experts_good = nn.ModuleList([ExpertMLP(expert_dim) for _ in range(n_task_experts)])
        self.task_experts_best = nn.ModuleList([ExpertMLP(expert_dim) for _ in range(n_task_experts)])
        self.fg_meta = FeatureGate(expert_dim, n_shared_experts)
        self.fg_good = FeatureGate(expert_dim, n_task_experts)
        self.fg_best = FeatureGate(expert_dim, n_task_experts)
        self.sg_meta = SelfGate(expert_dim)
        self.sg_good = SelfGate(expert_dim)
        self.sg_best = SelfGate(expert_dim)
        self.gate_good = DenseGate(expert_dim, n_shared_experts + n_task_experts)
        self.gate_best = DenseGate(expert_dim, n_shared_experts + n_task_experts)
        self.tower_good = self._make_tower(expert_dim, tower_hidden)
        self.tower_best = self._make_tower(expert_dim, tower_hidden)
        self._cfg_tuple = (expert_dim, num_input_experts, n_shared_experts, n_task_experts, tower_hidden, 1024)

    def _make_tower(self, input_dim, hidden_dim):
        return nn.Sequential(nn.LayerNorm(input_dim), nn.Linear(input_dim, hidden_dim), nn.GELU(), nn.Dropout(0.1), nn.Linear(hidden_dim, 1))

    def _run(self, expert_vecs, want_gates):
        d, n_in = self._cfg_tuple[0], self._cfg_tuple[1]
        if expert_vecs.dim() != 3 or expert_vecs.shape[1] != n_in or expert_vecs.shape[2] != d:
            raise RuntimeError(f"expert_vecs must be [B,{n_in},{d}], got {tuple(expert_vecs.shape)}")
        cfg = HomeCfg(*self._cfg_tuple)
        return Fn.HeadFn.apply(self._pack(), "home", cfg, self.training, 0.1, want_gates, expert_vecs, *self._params())

    def forward(self, expert_vecs: torch.Tensor):
        logits = self._run(expert_vecs, False)
        return logits[0], logits[1]

    def gate_weights(self, expert_vecs: torch.Tensor):
        _, gates = self._run(expert_vecs, True)
        return gates[0], gates[1]
