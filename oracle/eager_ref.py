"""The reference's forward passes executed by STOCK torch.nn modules (PyTorch eager) — the "PyTorch-eager-on-B200" bar.

TEST / BENCH INFRASTRUCTURE ONLY (same rule as ``mmoe_oracle.py``): imported by ``tests/`` and by ``bench.py``'s
``eager_b200`` / ``cpu_baseline`` legs, never by the product path.

The reference's hot-path classes are thin compositions of ``nn.TransformerEncoderLayer``, ``nn.MultiheadAttention``,
``nn.LayerNorm``, ``nn.Linear`` (every kernel it runs is an ATen library call).  The drop-in modules keep
their parameters in exactly those containers (state-dict contract) but never *call* them.  The functions
here do call them, in the order the reference's ``forward`` bodies do, so that the same weights can be run through
(a) torch's own kernels — cuBLASLt, SDPA, native LayerNorm/dropout — and (b) the native path, on the same device, and
timed side by side.  Dropout is torch's (``nn.Dropout`` / the containers' own), active in ``.train()`` as in the reference.

Each function cites the reference lines it follows (paths relative to ``/root/reference``).
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


# ----------------------------------------------------------------------------------------------
# containers: stock torch.nn modules laid out like the reference's constructors (same state_dict keys as the reference
# and the drop-ins), so this file needs nothing from the product package
# ----------------------------------------------------------------------------------------------
class _Box(nn.Module):
    pass


def _enc(d, n_head, p):
    return nn.TransformerEncoderLayer(d_model=d, nhead=n_head, dim_feedforward=4 * d, dropout=p, batch_first=True, norm_first=True)


def make_cross_expert(d=768, n_layer=2, n_head=8, dropout=0.1, home=False):
    """RobustTextCrossExpert.__init__ — model.py:387-424."""
    m = _Box()
    m.self_user = nn.ModuleList([_enc(d, n_head, dropout) for _ in range(n_layer)])
    m.self_item = nn.ModuleList([_enc(d, n_head, dropout) for _ in range(n_layer)])
    m.cross_attn = nn.MultiheadAttention(d, n_head, dropout=dropout, batch_first=True)
    m.gate = nn.Parameter(torch.tensor([0.5]))
    m.pool = _Box()
    m.pool.query = nn.Parameter(torch.randn(1, 1, d) * d ** -0.5)
    m.pool.dropout = nn.Dropout(dropout)
    m.norm = nn.LayerNorm(d)
    m.mlp = nn.Sequential(nn.Linear(d, 4 * d), nn.GELU(), nn.Dropout(dropout), nn.Linear(4 * d, d), nn.Dropout(dropout))
    m._home = home
    return m


def make_cross_fuse(d=768, n_head=8, depth=2, dropout=0.1, home=False):
    """EnhancedCrossFuse.__init__ — model.py:456-489."""
    m = _Box()
    m.layers = nn.ModuleList([_enc(d, n_head, dropout) for _ in range(depth)])
    m.res_proj = nn.Sequential(nn.Linear(2 * d, d), nn.LayerNorm(d))
    m.gate = nn.Sequential(nn.Linear(2 * d, d // 2), nn.GELU(), nn.Linear(d // 2, 1), nn.Sigmoid())
    nn.init.constant_(m.gate[2].bias, 0.5)
    m.proj = nn.Sequential(nn.LayerNorm(d), nn.Linear(d, d), nn.GELU(), nn.Dropout(dropout))
    m._home = home
    return m


def make_image_tail(d=768, pool_type="mean", dropout=0.1):
    """ItemImageExpert's own layers — model.py:356-364 (the HF backbone is not part of the timed path)."""
    m = _Box()
    m.pool_type = pool_type
    m.dropout = nn.Dropout(dropout)
    m.norm = nn.LayerNorm(d)
    return m


def make_mmoe_head(d=768, n_expert=6, hidden=256, tower_dropout=0.0):
    """TwoTaskMMoE.__init__ — model.py:532-559."""
    m = _Box()
    for t in ("good", "best"):
        g = _Box()
        g.fc = nn.Linear(d, n_expert)
        setattr(m, "gate_" + t, g)
        setattr(m, "tower_" + t, nn.Sequential(nn.LayerNorm(d), nn.Linear(d, hidden), nn.GELU(), nn.Dropout(tower_dropout),
                                               nn.Linear(hidden, hidden // 2), nn.GELU(), nn.Dropout(tower_dropout),
                                               nn.Linear(hidden // 2, 1)))
    return m


def make_v1_modules(device, train=True, seed=1234):
    """img / cross / concat_ui / concat_ti / head as train.py:118-130 builds them (default constructor arguments)."""
    torch.manual_seed(seed)
    mods = {"img": make_image_tail(), "cross": make_cross_expert(), "concat_ui": make_cross_fuse(), "concat_ti": make_cross_fuse(),
            "head": make_mmoe_head()}
    for m in mods.values():
        m.to(device).train(train)
    return mods


def encoder_layer(layer, x, key_padding_mask=None):
    """RobustTransformerLayer.forward — model.py:208-212: the pre-LN residual form written out (which also keeps torch's
    fused eval fast path out of the way, as the reference's override does).  The stock layers of EnhancedCrossFuse
    (model.py:460-465, norm_first=True) compute the same thing in train mode."""
    x = x + layer._sa_block(layer.norm1(x), None, key_padding_mask)
    return x + layer._ff_block(layer.norm2(x))


def attn_pool(pool, x, mask, home=False):
    """AttnPool1D.forward — model.py:199-206 (HoME: model_HoME.py:205-215)."""
    d = x.shape[-1]
    s = torch.matmul(pool.query, x.transpose(1, 2)).squeeze(1) / d ** 0.5
    s = s.masked_fill(mask, float("-inf"))
    w = torch.softmax(s, dim=-1)
    if home:
        w = torch.where(torch.isfinite(w).any(-1, keepdim=True), w, torch.zeros_like(w))
    w = pool.dropout(w)
    return (w.unsqueeze(-1) * x).sum(dim=1)


def cross_expert(m, user, user_mask, item, item_mask):
    """RobustTextCrossExpert.forward — model.py:426-451 (HoME variant model_HoME.py:441-466)."""
    home = bool(getattr(m, "_home", False))
    for layer in m.self_user:
        user = encoder_layer(layer, user, user_mask)
    for layer in m.self_item:
        item = encoder_layer(layer, item, item_mask)
    cross = m.cross_attn(query=user, key=item, value=item, key_padding_mask=item_mask)[0]
    alpha = torch.sigmoid(m.gate)
    fused = alpha * user + (1 - alpha) * cross
    pooled = attn_pool(m.pool, fused, user_mask, home)
    if home:
        return pooled
    normed = m.norm(pooled)
    return normed + m.mlp(normed)


def cross_fuse(m, v_cls, t_cls):
    """EnhancedCrossFuse.forward — model.py:491-507 (HoME variant model_HoME.py:506-522)."""
    identity = m.res_proj(torch.cat([v_cls, t_cls], dim=-1))
    x = torch.stack([v_cls, t_cls], dim=1)
    for layer in m.layers:
        x = encoder_layer(layer, x)
    v_f, t_f = x[:, 0], x[:, 1]
    g = m.gate(torch.cat([v_f, t_f], dim=-1))
    fused = g * v_f + (1 - g) * t_f
    y = fused + identity
    if getattr(m, "_home", False):
        return y
    return m.proj(y)


def item_image_tail(m, tokens):
    """ItemImageExpert.forward after the backbone — model.py:377-385 (frozen: train.py:244 runs it under no_grad)."""
    with torch.no_grad():
        v = tokens.mean(dim=1) if m.pool_type == "mean" else tokens[:, 0]
    return m.dropout(m.norm(v))


def two_task_mmoe(m, expert_vecs):
    """TwoTaskMMoE.forward — model.py:562-577, DenseGate :522-524."""
    q = expert_vecs.mean(dim=1)
    w_g = F.softmax(m.gate_good.fc(q), dim=-1)
    w_b = F.softmax(m.gate_best.fc(q), dim=-1)
    f_g = (w_g.unsqueeze(-1) * expert_vecs).sum(dim=1)
    f_b = (w_b.unsqueeze(-1) * expert_vecs).sum(dim=1)
    return m.tower_good(f_g).squeeze(-1), m.tower_best(f_b).squeeze(-1)


def home_mmoe(m, expert_vecs):
    """HOME_MMoE_Complete.forward — model_HoME.py:590-638 (FeatureGate :232-234, SelfGate :242-243, ExpertMLP :28-35)."""
    B = expert_vecs.shape[0]
    d = m.expert_dim
    shared = m.input_projection(expert_vecs.reshape(B, -1)) + expert_vecs.mean(dim=1)

    def fg(g, x):
        return x.unsqueeze(1) * (2.0 * torch.sigmoid(g.gate(x))).view(B, g.n_experts, d)

    def sg(g, x, y):
        return x + g.gate(x) * y

    meta_in, good_in, best_in = fg(m.fg_meta, shared), fg(m.fg_good, shared), fg(m.fg_best, shared)
    meta = [sg(m.sg_meta, shared, e(meta_in[:, i])) for i, e in enumerate(m.meta_experts)]
    good = [sg(m.sg_good, shared, e(good_in[:, i])) for i, e in enumerate(m.task_experts_good)]
    best = [sg(m.sg_best, shared, e(best_in[:, i])) for i, e in enumerate(m.task_experts_best)]
    ex_g, ex_b = torch.stack(meta + good, dim=1), torch.stack(meta + best, dim=1)
    w_g = F.softmax(m.gate_good.fc(shared), dim=-1)
    w_b = F.softmax(m.gate_best.fc(shared), dim=-1)
    f_g = (w_g.unsqueeze(-1) * ex_g).sum(dim=1)
    f_b = (w_b.unsqueeze(-1) * ex_b).sum(dim=1)
    return m.tower_good(f_g).squeeze(-1), m.tower_best(f_b).squeeze(-1)


def v1_fusion_step(mods, b, pos_weight_good, pos_weight_best, autocast_dtype=None):
    """One fusion-and-head micro-step of train.py:244-254 in eager PyTorch on the given modules' own torch containers
    (``mods``: dict img / cross / concat_ui / concat_ti / head of drop-in or reference modules — same attribute names).
    Returns the loss (backward not run)."""
    dev_type = b["u_sent"].device.type
    ctx = torch.autocast(dev_type, dtype=autocast_dtype) if autocast_dtype is not None else torch.autocast(dev_type, enabled=False)
    with ctx:
        img_vec = item_image_tail(mods["img"], b["img_tokens"])
        ui = cross_expert(mods["cross"], b["u_sent"], b["u_mask"], b["i_sent"], b["i_mask"])
        xui = cross_fuse(mods["concat_ui"], b["u_doc"], img_vec)
        xti = cross_fuse(mods["concat_ti"], b["i_doc"], img_vec)
        ev = torch.stack([b["u_doc"], b["i_doc"], img_vec, ui, xui, xti], dim=1)
        lg, lb = two_task_mmoe(mods["head"], ev)
        return (F.binary_cross_entropy_with_logits(lg.float(), b["y_good"], pos_weight=pos_weight_good) +
                F.binary_cross_entropy_with_logits(lb.float(), b["y_best"], pos_weight=pos_weight_best))
