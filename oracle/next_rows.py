"""CPU oracle for the rows of SURVEY.md §8f — the callers and data formats either side of the fusion path.

TEST INFRASTRUCTURE ONLY (same rule as ``mmoe_oracle.py``): imported by ``tests/`` only.

Plain restatements with elementary tensor operations of
  * ``HomeExpertWrapper`` x n + ``torch.stack``            — train_HoME.py:100-116, 350-356
  * ``nn.BCEWithLogitsLoss(pos_weight)`` x 2               — train.py:189-192, 253-254
  * ``calculate_contrastive_loss``                         — train_HoME.py:43-51
  * ``sklearn.metrics.roc_auc_score``                      — inference_and_auc.py:171,178 (third-party: scikit-learn, unpinned
    in requirements.txt; its published definition — area under the ROC curve by the trapezoidal rule — equals the
    Mann-Whitney statistic with ties counted one half, which is what is restated here)
  * ``decode_sample``'s image branch + HF ``ViTPatchEmbeddings`` — model.py:160-178, 373-376 (third-party: transformers
    ``ViTPatchEmbeddings.forward`` = ``projection(pixel_values).flatten(2).transpose(1, 2)`` with a Conv2d whose kernel equals
    its stride)
  * ``TextExpert.forward`` after the encoder               — model.py:286-338 (HoME: model_HoME.py:328-369)
Pinned by ``oracle/make_golden_next.py`` against the reference's own code (tests/golden/next_*.pt).
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np
import torch

from .mmoe_oracle import layer_norm, softmax_lastdim

IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


def home_wrapper_stack(xs: Sequence[torch.Tensor], weights, biases, running_means, running_vars, training: bool,
                       momentum: float = 0.1, eps: float = 1e-5, drop=None):
    """stack([Dropout(SiLU(BatchNorm1d_e(x_e)))], dim=1) — train_HoME.py:108-116, 350-356.
    Returns (expert_vecs [B,n,d], new_running_means, new_running_vars).  ``drop(site, tensor)`` as in mmoe_oracle."""
    outs, new_rm, new_rv = [], [], []
    for e, x in enumerate(xs):
        if training:
            mean = x.mean(dim=0)
            var = ((x - mean) ** 2).mean(dim=0)                       # biased: normalises the batch
            n = x.shape[0]
            new_rm.append((1 - momentum) * running_means[e] + momentum * mean.detach())
            new_rv.append((1 - momentum) * running_vars[e] + momentum * var.detach() * n / max(n - 1, 1))
        else:
            mean, var = running_means[e], running_vars[e]
            new_rm.append(running_means[e]); new_rv.append(running_vars[e])
        z = (x - mean) / torch.sqrt(var + eps) * weights[e] + biases[e]
        outs.append(z * torch.sigmoid(z))
    out = torch.stack(outs, dim=1)
    if drop is not None:
        out = drop("stack", out)
    return out, new_rm, new_rv


def bce2(logit_good, logit_best, y_good, y_best, pw_good: float, pw_best: float):
    """loss_fn_good(logit_g, y_good) + loss_fn_best(logit_b, y_best) — train.py:253-254, mean reduction."""
    def one(x, y, pw):
        log_sig = -torch.nn.functional.softplus(-x)
        log_one_minus = -torch.nn.functional.softplus(x)
        return -(pw * y * log_sig + (1.0 - y) * log_one_minus).mean()
    return one(logit_good, y_good, pw_good) + one(logit_best, y_best, pw_best)


def info_nce(anchor, positive, temperature: float = 0.07):
    """calculate_contrastive_loss — train_HoME.py:43-51."""
    a = anchor / anchor.norm(dim=1, keepdim=True).clamp_min(1e-12)
    p = positive / positive.norm(dim=1, keepdim=True).clamp_min(1e-12)
    sim = a @ p.t() / temperature
    logp = torch.log(softmax_lastdim(sim))
    return -logp.diagonal().mean()


def roc_auc(scores: np.ndarray, labels: np.ndarray) -> float:
    """Area under the ROC curve = P(score_pos > score_neg) + 0.5 P(equal), by average ranks."""
    scores = np.asarray(scores, dtype=np.float64)
    pos = np.asarray(labels) > 0.5
    order = np.argsort(scores, kind="mergesort")
    s = scores[order]
    ranks = np.empty(len(s), dtype=np.float64)
    i = 0
    while i < len(s):
        j = i
        while j + 1 < len(s) and s[j + 1] == s[i]:
            j += 1
        ranks[i:j + 1] = 0.5 * (i + j) + 1.0
        i = j + 1
    r = np.empty_like(ranks)
    r[order] = ranks
    n1, n0 = pos.sum(), (~pos).sum()
    if n1 == 0 or n0 == 0:
        return float("nan")
    return float((r[pos].sum() - n1 * (n1 + 1) / 2.0) / (n1 * n0))


def unpatchify_normalise(patch_bytes: np.ndarray) -> torch.Tensor:
    """decode_sample's image branch — model.py:160-175: uint8 [196, 3, 16, 16] -> normalised float [3, 224, 224]."""
    t = torch.from_numpy(patch_bytes.reshape(196, 3, 16, 16).copy()).float() / 255.0
    img = t.permute(1, 0, 2, 3).reshape(3, 14, 14, 16, 16).permute(0, 1, 3, 2, 4).reshape(3, 224, 224)
    mean = torch.tensor(IMAGENET_MEAN)[:, None, None]
    std = torch.tensor(IMAGENET_STD)[:, None, None]
    return (img - mean) / std


def patch_embed(images: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor]) -> torch.Tensor:
    """HF ViTPatchEmbeddings.forward for a kernel == stride convolution: [B,C,H,W] -> [B, (H/p)(W/p), hidden]."""
    hidden, C, p, _ = weight.shape
    B, _, H, W = images.shape
    x = images.reshape(B, C, H // p, p, W // p, p).permute(0, 2, 4, 1, 3, 5).reshape(B, (H // p) * (W // p), C * p * p)
    y = x @ weight.reshape(hidden, -1).t()
    return y if bias is None else y + bias


def sentence_gather(h: torch.Tensor, chunk2sample: List[int], sent_pos: List[List[int]], max_sent_count: int,
                    norm_w: Optional[torch.Tensor], norm_b: Optional[torch.Tensor], drop=None):
    """TextExpert.forward after the encoder — model.py:286-338; norm_w None = HoME variant (model_HoME.py:366-367)."""
    n_chunks, seq_len, D = h.shape
    pos = torch.tensor(sent_pos)
    vecs = h[torch.arange(n_chunks)[:, None], pos.clamp(0, seq_len - 1)]
    vecs = vecs * (pos >= 0).unsqueeze(-1).to(h.dtype)
    B = max(chunk2sample) + 1
    rows = []
    for b in range(B):
        bucket = [vecs[i] for i, s in enumerate(chunk2sample) if s == b]
        cat = torch.cat(bucket, dim=0) if bucket else torch.zeros(1, D, dtype=h.dtype)
        cat = cat[:max_sent_count]
        if cat.shape[0] < max_sent_count:
            cat = torch.cat([cat, torch.zeros(max_sent_count - cat.shape[0], D, dtype=h.dtype)], dim=0)
        rows.append(cat)
    padded = torch.stack(rows)
    mask = padded.abs().sum(-1) == 0
    lens = (~mask).sum(dim=1, keepdim=True)
    doc = padded.sum(dim=1) / lens.clamp(min=1)
    if norm_w is not None:
        padded = layer_norm(padded, norm_w, norm_b)
        doc = layer_norm(doc, norm_w, norm_b)
        if drop is not None:
            padded, doc = drop("sent", padded), drop("doc", doc)
    return padded, mask, doc
