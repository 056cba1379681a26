"""CPU oracle for the MMoE / HoME fusion-and-head hot path.

TEST INFRASTRUCTURE ONLY — this is the *checker*, never the product.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it.  The product path (``model.py``,
``model_HoME.py`` and the package ``mmoe-multimodal-rec_b200/``) never does.

What it is: a plain restatement, function by function, of the arithmetic the
reference's modules perform for this path, written with elementary tensor
operations (matmul, exp, erf, sums) on CPU tensors in whatever dtype it is
handed (float32 for the baseline timing, float64 for parity).  It does NOT use
``nn.MultiheadAttention`` / ``nn.TransformerEncoderLayer`` / ``nn.LayerNorm``;
those library semantics are restated explicitly here.
Gradients come from autograd over these elementary operations.

Parity pin: the reference ships no tests or golden vectors, so
the oracle is pinned against *the reference modules themselves*, imported
unmodified from ``/root/reference`` in the development container:
``oracle/make_golden.py`` runs both on identical deterministic weights/inputs
(``oracle/synth.py``), asserts agreement, and writes the reference's outputs to
``tests/golden/*.pt``; ``tests/test_oracle_golden.py`` re-checks the oracle
against those committed vectors anywhere (no ``/root/reference`` needed).

Every function cites the reference lines it follows (paths are relative to
``/root/reference``).  All functions take a ``sd`` mapping with the reference
module's ``state_dict()`` keys.

Dropout: the reference's ``nn.Dropout`` layers are identity in ``eval()``; the
oracle is the eval-mode statement by default.  Passing ``drop=callable``
(``drop(site_name, tensor) -> tensor``) lets a test inject the exact keep-masks
a CUDA kernel used, so train-mode arithmetic can be checked too.
"""
from __future__ import annotations

import math
from typing import Callable, Mapping, Optional

import torch

Tensor = torch.Tensor
Drop = Optional[Callable[[str, Tensor], Tensor]]


def _drop(drop: Drop, site: str, x: Tensor) -> Tensor:
    return x if drop is None else drop(site, x)


# ----------------------------------------------------------------------------
# elementary pieces (torch library semantics restated)
# ----------------------------------------------------------------------------

def layer_norm(x: Tensor, w: Tensor, b: Tensor, eps: float = 1e-5) -> Tensor:
    """nn.LayerNorm over the last dim: biased variance, eps inside the sqrt."""
    mu = x.mean(dim=-1, keepdim=True)
    xc = x - mu
    var = (xc * xc).mean(dim=-1, keepdim=True)
    return xc / torch.sqrt(var + eps) * w + b


def gelu(x: Tensor) -> Tensor:
    """nn.GELU() = exact erf form (approximate='none')."""
    return 0.5 * x * (1.0 + torch.erf(x * (1.0 / math.sqrt(2.0))))


def linear(x: Tensor, w: Tensor, b: Optional[Tensor] = None) -> Tensor:
    y = x @ w.transpose(-1, -2)
    return y if b is None else y + b


def softmax_lastdim(x: Tensor) -> Tensor:
    """softmax(-1) with the usual max subtraction; a row of all -inf gives NaN,
    exactly as in the reference."""
    m = x.max(dim=-1, keepdim=True).values
    e = torch.exp(x - m)
    return e / e.sum(dim=-1, keepdim=True)


def multi_head_attention(sd: Mapping[str, Tensor], prefix: str, q_in: Tensor, kv_in: Tensor,
                         key_padding_mask: Optional[Tensor], n_head: int,
                         drop: Drop = None, site: str = "attn") -> Tensor:
    """nn.MultiheadAttention(batch_first=True) forward, attention output only.

    in_proj_weight = [Wq; Wk; Wv] stacked on dim 0; head h uses columns
    h*hd..h*hd+hd-1; q is scaled by hd**-0.5 before QK^T; padded keys get -inf
    added to their score; dropout acts on the probabilities; out_proj follows.
    (call sites model.py:210 via _sa_block, model.py:435-440.)
    """
    d = q_in.shape[-1]
    hd = d // n_head
    w_in, b_in = sd[prefix + "in_proj_weight"], sd[prefix + "in_proj_bias"]
    q = linear(q_in, w_in[:d], b_in[:d])
    k = linear(kv_in, w_in[d:2 * d], b_in[d:2 * d])
    v = linear(kv_in, w_in[2 * d:], b_in[2 * d:])
    B, Lq, _ = q.shape
    Lk = k.shape[1]
    q = q.reshape(B, Lq, n_head, hd).transpose(1, 2)
    k = k.reshape(B, Lk, n_head, hd).transpose(1, 2)
    v = v.reshape(B, Lk, n_head, hd).transpose(1, 2)
    scores = (q * (hd ** -0.5)) @ k.transpose(-1, -2)            # [B,H,Lq,Lk]
    if key_padding_mask is not None:
        neg = torch.zeros(key_padding_mask.shape, dtype=scores.dtype, device=scores.device)
        neg = neg.masked_fill(key_padding_mask, float("-inf"))
        scores = scores + neg[:, None, None, :]
    p = softmax_lastdim(scores)
    p = _drop(drop, site + ".probs", p)
    ctx = (p @ v).transpose(1, 2).reshape(B, Lq, d)
    return linear(ctx, sd[prefix + "out_proj.weight"], sd[prefix + "out_proj.bias"])


def encoder_layer(sd: Mapping[str, Tensor], prefix: str, x: Tensor, key_padding_mask: Optional[Tensor],
                  n_head: int, drop: Drop = None, relu_masks=None, trace=None) -> Tensor:
    """Pre-LN encoder layer, ReLU feed-forward.

    model.py:207-212 (RobustTransformerLayer) and the stock
    nn.TransformerEncoderLayer(norm_first=True) used at model.py:460-465 compute
    the same thing:  x = x + drop1(SA(LN1(x)));  x = x + drop2(W2 drop(relu(W1 LN2(x)))).
    """
    h = layer_norm(x, sd[prefix + "norm1.weight"], sd[prefix + "norm1.bias"])
    sa = multi_head_attention(sd, prefix + "self_attn.", h, h, key_padding_mask, n_head,
                              drop, prefix + "self_attn")
    x = x + _drop(drop, prefix + "dropout1", sa)
    h = layer_norm(x, sd[prefix + "norm2.weight"], sd[prefix + "norm2.bias"])
    z = linear(h, sd[prefix + "linear1.weight"], sd[prefix + "linear1.bias"])
    if trace is not None:
        trace[prefix + "ffn_pre"] = z.detach()
    if relu_masks is not None and (prefix + "relu") in relu_masks:
        # test hook: use a given activation pattern (e.g. the one a low-precision kernel produced) instead of z > 0
        f = z * relu_masks[prefix + "relu"].to(z.dtype).reshape(z.shape)
    else:
        f = torch.relu(z)
    f = _drop(drop, prefix + "dropout", f)
    f = linear(f, sd[prefix + "linear2.weight"], sd[prefix + "linear2.bias"])
    return x + _drop(drop, prefix + "dropout2", f)


def attn_pool(query: Tensor, x: Tensor, mask: Tensor, home: bool = False,
              drop: Drop = None, site: str = "pool") -> Tensor:
    """AttnPool1D.forward — model.py:199-206; HoME variant model_HoME.py:205-215
    adds the finite-row guard before dropout."""
    d = x.shape[-1]
    s = (x @ query.reshape(d)) / (d ** 0.5)                      # [B,L]
    s = s.masked_fill(mask, float("-inf"))
    w = softmax_lastdim(s)
    if home:
        finite_row = torch.isfinite(w).any(dim=-1, keepdim=True)
        w = torch.where(finite_row, w, torch.zeros_like(w))
    w = _drop(drop, site + ".weights", w)
    return (w.unsqueeze(-1) * x).sum(dim=1)


# ----------------------------------------------------------------------------
# fusion experts
# ----------------------------------------------------------------------------

def cross_expert(sd: Mapping[str, Tensor], user: Tensor, user_mask: Tensor, item: Tensor, item_mask: Tensor,
                 n_layer: int = 2, n_head: int = 8, home: bool = False, drop: Drop = None, relu_masks=None, trace=None) -> Tensor:
    """RobustTextCrossExpert.forward — model.py:426-451; HoME variant
    model_HoME.py:441-466 returns ``pooled`` (norm/mlp unused)."""
    for l in range(n_layer):
        user = encoder_layer(sd, f"self_user.{l}.", user, user_mask, n_head, drop, relu_masks, trace)
    for l in range(n_layer):
        item = encoder_layer(sd, f"self_item.{l}.", item, item_mask, n_head, drop, relu_masks, trace)
    cross = multi_head_attention(sd, "cross_attn.", user, item, item_mask, n_head, drop, "cross_attn")
    alpha = torch.sigmoid(sd["gate"])
    fused = alpha * user + (1.0 - alpha) * cross
    pooled = attn_pool(sd["pool.query"], fused, user_mask, home=home, drop=drop)
    if home:
        return pooled
    normed = layer_norm(pooled, sd["norm.weight"], sd["norm.bias"])
    h = gelu(linear(normed, sd["mlp.0.weight"], sd["mlp.0.bias"]))
    h = _drop(drop, "mlp.2", h)
    h = linear(h, sd["mlp.3.weight"], sd["mlp.3.bias"])
    return normed + _drop(drop, "mlp.4", h)


def cross_fuse(sd: Mapping[str, Tensor], v_cls: Tensor, t_cls: Tensor, depth: int = 2, n_head: int = 8,
               home: bool = False, drop: Drop = None, relu_masks=None, trace=None) -> Tensor:
    """EnhancedCrossFuse.forward — model.py:491-507; HoME variant
    model_HoME.py:506-522 returns ``fused + identity`` (proj unused)."""
    cat = torch.cat([v_cls, t_cls], dim=-1)
    identity = layer_norm(linear(cat, sd["res_proj.0.weight"], sd["res_proj.0.bias"]),
                          sd["res_proj.1.weight"], sd["res_proj.1.bias"])
    x = torch.stack([v_cls, t_cls], dim=1)                       # [B,2,d]
    for l in range(depth):
        x = encoder_layer(sd, f"layers.{l}.", x, None, n_head, drop, relu_masks, trace)
    v_f, t_f = x[:, 0], x[:, 1]
    gi = torch.cat([v_f, t_f], dim=-1)
    g = torch.sigmoid(linear(gelu(linear(gi, sd["gate.0.weight"], sd["gate.0.bias"])),
                             sd["gate.2.weight"], sd["gate.2.bias"]))      # [B,1]
    fused = g * v_f + (1.0 - g) * t_f
    y = fused + identity
    if home:
        return y
    y = layer_norm(y, sd["proj.0.weight"], sd["proj.0.bias"])
    y = gelu(linear(y, sd["proj.1.weight"], sd["proj.1.bias"]))
    return _drop(drop, "proj.3", y)


def item_image_pool(sd: Mapping[str, Tensor], tokens: Tensor, pool_type: str = "mean", drop: Drop = None) -> Tensor:
    """ItemImageExpert.forward after the backbone — model.py:377-385."""
    v = tokens.mean(dim=1) if pool_type == "mean" else tokens[:, 0]
    return _drop(drop, "dropout", layer_norm(v, sd["norm.weight"], sd["norm.bias"]))


def image_projection(sd: Mapping[str, Tensor], tokens: Tensor):
    """ImageExpertWithProjection.forward after the ViT — model_HoME.py:393-399."""
    img_vec = tokens[:, 0, :]
    h = gelu(linear(img_vec, sd["projection_head.0.weight"], sd["projection_head.0.bias"]))
    return img_vec, linear(h, sd["projection_head.2.weight"], sd["projection_head.2.bias"])


# ----------------------------------------------------------------------------
# heads
# ----------------------------------------------------------------------------

def dense_gate(sd: Mapping[str, Tensor], prefix: str, x: Tensor) -> Tensor:
    """DenseGate.forward — model.py:522-524 / model_HoME.py:251-252."""
    return softmax_lastdim(linear(x, sd[prefix + "fc.weight"], sd[prefix + "fc.bias"]))


def _mmoe_tower(sd: Mapping[str, Tensor], p: str, x: Tensor, drop: Drop) -> Tensor:
    """_make_tower — model.py:545-556: LN, 768→h GELU, h→h/2 GELU, h/2→1."""
    x = layer_norm(x, sd[p + "0.weight"], sd[p + "0.bias"])
    x = _drop(drop, p + "3", gelu(linear(x, sd[p + "1.weight"], sd[p + "1.bias"])))
    x = _drop(drop, p + "6", gelu(linear(x, sd[p + "4.weight"], sd[p + "4.bias"])))
    return linear(x, sd[p + "7.weight"], sd[p + "7.bias"]).squeeze(-1)


def two_task_mmoe(sd: Mapping[str, Tensor], expert_vecs: Tensor, drop: Drop = None, return_gates: bool = False):
    """TwoTaskMMoE.forward — model.py:562-577."""
    query = expert_vecs.mean(dim=1)
    w_good = dense_gate(sd, "gate_good.", query)
    w_best = dense_gate(sd, "gate_best.", query)
    fused_good = (w_good.unsqueeze(-1) * expert_vecs).sum(dim=1)
    fused_best = (w_best.unsqueeze(-1) * expert_vecs).sum(dim=1)
    lg = _mmoe_tower(sd, "tower_good.", fused_good, drop)
    lb = _mmoe_tower(sd, "tower_best.", fused_best, drop)
    if return_gates:
        return lg, lb, w_good, w_best
    return lg, lb


def _expert_mlp(sd: Mapping[str, Tensor], p: str, x: Tensor, drop: Drop) -> Tensor:
    """ExpertMLP — model_HoME.py:28-35."""
    h = _drop(drop, p + "2", gelu(linear(x, sd[p + "0.weight"], sd[p + "0.bias"])))
    return linear(h, sd[p + "3.weight"], sd[p + "3.bias"])


def _feature_gate(sd: Mapping[str, Tensor], p: str, x: Tensor, n: int) -> Tensor:
    """FeatureGate.forward — model_HoME.py:232-234."""
    d = x.shape[-1]
    gv = linear(x, sd[p + "gate.weight"], sd[p + "gate.bias"]).reshape(-1, n, d)
    return x.unsqueeze(1) * (2.0 * torch.sigmoid(gv))


def _self_gate(sd: Mapping[str, Tensor], p: str, x_orig: Tensor, x_proc: Tensor) -> Tensor:
    """SelfGate.forward — model_HoME.py:242-243."""
    return x_orig + torch.sigmoid(linear(x_orig, sd[p + "gate.0.weight"], sd[p + "gate.0.bias"])) * x_proc


def home_mmoe(sd: Mapping[str, Tensor], expert_vecs: Tensor, n_shared: int = 4, n_task: int = 2,
              drop: Drop = None, return_gates: bool = False):
    """HOME_MMoE_Complete.forward — model_HoME.py:590-638."""
    B = expert_vecs.shape[0]
    concat = expert_vecs.reshape(B, -1)
    proj = gelu(layer_norm(linear(concat, sd["input_projection.0.weight"], sd["input_projection.0.bias"]),
                           sd["input_projection.1.weight"], sd["input_projection.1.bias"]))
    shared = proj + expert_vecs.mean(dim=1)
    meta_in = _feature_gate(sd, "fg_meta.", shared, n_shared)
    good_in = _feature_gate(sd, "fg_good.", shared, n_task)
    best_in = _feature_gate(sd, "fg_best.", shared, n_task)
    meta = [_expert_mlp(sd, f"meta_experts.{i}.", meta_in[:, i], drop) for i in range(n_shared)]
    good = [_expert_mlp(sd, f"task_experts_good.{i}.", good_in[:, i], drop) for i in range(n_task)]
    best = [_expert_mlp(sd, f"task_experts_best.{i}.", best_in[:, i], drop) for i in range(n_task)]
    meta = [_self_gate(sd, "sg_meta.", shared, o) for o in meta]
    good = [_self_gate(sd, "sg_good.", shared, o) for o in good]
    best = [_self_gate(sd, "sg_best.", shared, o) for o in best]
    ex_good = torch.stack(meta + good, dim=1)
    ex_best = torch.stack(meta + best, dim=1)
    w_good = dense_gate(sd, "gate_good.", shared)
    w_best = dense_gate(sd, "gate_best.", shared)
    fused_good = (w_good.unsqueeze(-1) * ex_good).sum(dim=1)
    fused_best = (w_best.unsqueeze(-1) * ex_best).sum(dim=1)

    def tower(p, x):  # _make_tower — model_HoME.py:581-588
        x = layer_norm(x, sd[p + "0.weight"], sd[p + "0.bias"])
        x = _drop(drop, p + "3", gelu(linear(x, sd[p + "1.weight"], sd[p + "1.bias"])))
        return linear(x, sd[p + "4.weight"], sd[p + "4.bias"]).squeeze(-1)

    lg, lb = tower("tower_good.", fused_good), tower("tower_best.", fused_best)
    if return_gates:
        return lg, lb, w_good, w_best
    return lg, lb


# ----------------------------------------------------------------------------
# the v1 / HoME fusion path as the training scripts compose it
# ----------------------------------------------------------------------------

def bce_with_logits(logit: Tensor, y: Tensor, pos_weight: float) -> Tensor:
    """nn.BCEWithLogitsLoss(pos_weight) mean reduction — train.py:189-192."""
    log_sig = -torch.nn.functional.softplus(-logit)
    log_one_minus = -torch.nn.functional.softplus(logit)
    return -(pos_weight * y * log_sig + (1.0 - y) * log_one_minus).mean()


POS_WEIGHT_GOOD = 858627.0 / 990303.0     # train.py:189-190
POS_WEIGHT_BEST = 1328721.0 / 520209.0    # train.py:191-192


def v1_fusion_path(sds: Mapping[str, Mapping[str, Tensor]], u_sent, u_mask, i_sent, i_mask,
                   u_doc, i_doc, img_tokens, drop: Drop = None):
    """The fusion-and-head part of one train.py micro-step — train.py:244-251.

    ``sds`` holds the state dicts of ``img`` (ItemImageExpert's own norm),
    ``cross``, ``concat_ui``, ``concat_ti`` and ``head``.
    """
    img_vec = item_image_pool(sds["img"], img_tokens, "mean", drop).detach()   # trainable=False → no_grad
    ui_vec = cross_expert(sds["cross"], u_sent, u_mask, i_sent, i_mask, drop=drop)
    xui = cross_fuse(sds["concat_ui"], u_doc, img_vec, drop=drop)
    xti = cross_fuse(sds["concat_ti"], i_doc, img_vec, drop=drop)
    ev = torch.stack([u_doc, i_doc, img_vec, ui_vec, xui, xti], dim=1)
    return two_task_mmoe(sds["head"], ev, drop)
