"""The data formats either side of the fusion path on the native library.

* ``NativePatchEmbeddings`` — drop-in for HF ``ViTPatchEmbeddings`` (the ``Conv2d(3, 768, 16, stride=16)`` that
  ``ItemImageExpert.backbone(pixel_values=...)`` reaches, reference model.py:373-376): the convolution over non-overlapping
  patches is a GEMM on the tcgen05 engine.  It accepts the normalised float images the unchanged scripts pass
  ([B,3,224,224]) **or** the raw patch bytes as they sit on disk (uint8 [B,196,768] — newpatch.py:102-104,
  data4model.py:254-258), in which case /255, mean and std of ``decode_sample`` (model.py:172-174) are folded into the
  weights and nothing is un-patchified.  ``install_native_patch_embeddings(vit)`` swaps it into an HF ``ViTModel``.
* ``decode_patch_bytes`` — the device-side replacement of ``decode_sample``'s image branch for callers that keep the bytes.
* ``sentence_gather`` — TextExpert's post-encoder gather / bucket / pad / mask / mean / LayerNorm / dropout
  (model.py:286-338) in one launch, with autograd (the encoder is LoRA-trainable).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import torch
import torch.nn as nn
from torch.autograd.function import once_differentiable

from ._lib import BF16, F16, F32, check, lib
from .functional import _MMOE2TORCH, _TORCH2MMOE, _call, _f32c, _new_seed, _require_cuda, _state, _stream, compute_dtype

IMAGENET_MEAN = (0.485, 0.456, 0.406)      # model.py:172
IMAGENET_STD = (0.229, 0.224, 0.225)       # model.py:173


class NativePatchEmbeddings(nn.Module):
    """HF ``ViTPatchEmbeddings`` replacement: same ``projection`` parameter container (state-dict key
    ``projection.weight`` [hidden, C, p, p], ``projection.bias``), same output [B, num_patches, hidden].  Forward only
    (the patch projection is frozen in both training scripts: train.py:244 runs the backbone under no_grad,
    train_HoME.py:230-243 unfreezes only the last two encoder layers)."""

    def __init__(self, conv: nn.Conv2d, image_size=(224, 224)):
        super().__init__()
        if conv.kernel_size != conv.stride or conv.padding not in ((0, 0), 0):
            raise ValueError("patch projection must be a non-overlapping convolution (kernel == stride, no padding)")
        self.projection = conv
        self.patch_size = conv.kernel_size
        self.image_size = tuple(image_size)
        self.num_channels = conv.in_channels
        self.num_patches = (self.image_size[0] // self.patch_size[0]) * (self.image_size[1] // self.patch_size[1])
        self._cache = {}

    def _weights(self, dtype: int, folded: bool):
        """[hidden, C*p*p] weight in the operand dtype and the fp32 bias; folded = for raw byte input:
        W' = W / (255 std_c), b' = b - sum_k W mean_c / std_c  (so that W' u8 + b' == W ((u8/255 - mean)/std) + b)."""
        w, b = self.projection.weight, self.projection.bias
        key = (dtype, folded)
        ent = self._cache.get(key)
        if ent is None or ent[0] != (w._version, w.data_ptr(), None if b is None else b._version):
            hidden = w.shape[0]
            w2 = w.detach().float().reshape(hidden, self.num_channels, -1)
            bias = b.detach().float().clone() if b is not None else torch.zeros(hidden, device=w.device)
            if folded:
                mean = torch.tensor(IMAGENET_MEAN, device=w.device).view(1, -1, 1)
                std = torch.tensor(IMAGENET_STD, device=w.device).view(1, -1, 1)
                bias = bias - (w2 * (mean / std)).sum(dim=(1, 2))
                w2 = w2 / (255.0 * std)
            w2 = w2.reshape(hidden, -1).contiguous()
            wt = w2 if dtype == F32 else w2.to(_MMOE2TORCH[dtype])
            ent = ((w._version, w.data_ptr(), None if b is None else b._version), wt, bias.contiguous())
            self._cache[key] = ent
        return ent[1], ent[2]

    @torch.no_grad()
    def forward(self, pixel_values: torch.Tensor, interpolate_pos_encoding: bool = False):
        _require_cuda(pixel_values)
        L = lib()
        dtype = compute_dtype()
        p = self.patch_size[0]
        K = self.num_channels * p * p
        hidden = self.projection.weight.shape[0]
        dev = pixel_values.device
        if pixel_values.dtype == torch.uint8:
            # raw patch bytes [B, num_patches, C*p*p] (patch.bin as written by newpatch.py:102-104)
            pv = pixel_values.contiguous()
            if pv.dim() != 3 or pv.shape[-1] != K:
                raise RuntimeError(f"uint8 input must be patch-major [B, patches, {K}], got {tuple(pv.shape)}")
            B, n_patch = pv.shape[0], pv.shape[1]
            operand = torch.empty((B * n_patch, K), dtype=_MMOE2TORCH[dtype], device=dev)
            check(L.mmoe_patch_u8_to_operand(pv.data_ptr(), operand.data_ptr(), B * n_patch, K, dtype, _stream()), "patch_u8_to_operand")
            w, bias = self._weights(dtype, folded=True)
        else:
            img = _f32c(pixel_values)
            if img.dim() != 4 or img.shape[1] != self.num_channels:
                raise ValueError("Make sure that the channel dimension of the pixel values match with the one set in the configuration.")
            B, _, H, W = img.shape
            n_patch = (H // p) * (W // p)
            operand = torch.empty((B * n_patch, K), dtype=_MMOE2TORCH[dtype], device=dev)
            check(L.mmoe_patchify(img.data_ptr(), operand.data_ptr(), B, self.num_channels, H, W, p, dtype, _stream()), "patchify")
            w, bias = self._weights(dtype, folded=False)
        out_t = torch.float32 if dtype == F32 else _MMOE2TORCH[dtype]
        out = torch.empty((B, n_patch, hidden), dtype=out_t, device=dev)
        check(L.mmoe_patch_project(operand.data_ptr(), w.data_ptr(), bias.data_ptr(), out.data_ptr(), _TORCH2MMOE[out_t], B * n_patch, hidden, K,
                                   dtype, _stream()), "patch_project")
        return out


def install_native_patch_embeddings(vit: nn.Module) -> nn.Module:
    """Swap the patch projection of an HF ``ViTModel`` (``vit.embeddings.patch_embeddings``) for the native GEMM path, keeping
    its Conv2d parameters (checkpoints keep loading).  Returns the model."""
    emb = vit.embeddings
    old = emb.patch_embeddings
    size = getattr(old, "image_size", (224, 224))
    emb.patch_embeddings = NativePatchEmbeddings(old.projection, size if isinstance(size, (tuple, list)) else (size, size))
    return vit


def vit_tokens_from_patch_bytes(vit: nn.Module, patch_bytes: torch.Tensor) -> torch.Tensor:
    """last_hidden_state [B, 1 + patches, hidden] of an HF ``ViTModel`` fed with raw patch bytes (uint8 [B, patches, C*p*p], the
    patch.bin layout): native patch projection with the normalisation folded into the weights, then the model's own
    cls token / position embeddings / encoder / final LayerNorm.  (HF's ``ViTModel.forward`` insists on [B,C,H,W] pixel
    values, so the byte path enters one level below it.)"""
    emb = vit.embeddings
    if not isinstance(emb.patch_embeddings, NativePatchEmbeddings):
        install_native_patch_embeddings(vit)
    x = emb.patch_embeddings(patch_bytes)
    cls = emb.cls_token.expand(x.shape[0], -1, -1).to(x.dtype)
    x = torch.cat((cls, x), dim=1) + emb.position_embeddings
    x = emb.dropout(x)
    out = vit.encoder(x)
    h = out.last_hidden_state if hasattr(out, "last_hidden_state") else out[0]
    return vit.layernorm(h)


def decode_patch_bytes(patch_bytes: torch.Tensor) -> torch.Tensor:
    """Device-side form of decode_sample's image branch (model.py:160-178) for the v1 scripts' [B,3,224,224] contract:
    uint8 [B,196,768] patch bytes -> normalised float image [B,3,224,224].  Only needed by callers that want the float
    image; ``NativePatchEmbeddings`` consumes the bytes directly."""
    _require_cuda(patch_bytes)
    B = patch_bytes.shape[0]
    x = patch_bytes.view(B, 14, 14, 3, 16, 16).float().div_(255.0)
    x = x.permute(0, 3, 1, 4, 2, 5).reshape(B, 3, 224, 224)
    mean = torch.tensor(IMAGENET_MEAN, device=x.device).view(1, 3, 1, 1)
    std = torch.tensor(IMAGENET_STD, device=x.device).view(1, 3, 1, 1)
    return (x - mean) / std


# ----------------------------------------------------------------------------------------------
def build_slot_table(chunk2sample: Sequence[int], sent_pos: Sequence[Sequence[int]], seq_len: int, max_sent_count: int,
                     n_samples: Optional[int] = None) -> torch.Tensor:
    """int32 [B, max_sent_count]: the row of the flattened hidden states [N_chunks*seq_len, d] that feeds each sentence
    slot, -1 for an empty slot.  Restates the bucketing of model.py:299-323 on the host: the chunks of a sample are
    concatenated in order, each contributing ALL of its (padded) sentence positions, cut at max_sent_count; a position
    < 0 is a zero row (model.py:296), others are clamped to the sequence (model.py:294)."""
    B = (max(chunk2sample) + 1) if n_samples is None else n_samples
    table = [[-1] * max_sent_count for _ in range(B)]
    fill = [0] * B
    for i, s in enumerate(chunk2sample):
        for pos in sent_pos[i]:
            k = fill[s]
            fill[s] += 1
            if k >= max_sent_count:
                continue
            if pos >= 0:
                table[s][k] = i * seq_len + min(max(pos, 0), seq_len - 1)
    return torch.tensor(table, dtype=torch.int32)


class _SentGather(torch.autograd.Function):
    @staticmethod
    def forward(ctx, training: bool, drop_p: float, src, h, gamma, beta):
        _require_cuda(h, src)
        L = lib()
        hd = h.detach()
        if hd.dtype not in _TORCH2MMOE:
            hd = hd.float()
        hd = hd.contiguous()
        d = hd.shape[-1]
        B, S = src.shape
        dev = hd.device
        has_norm = gamma is not None
        pt = [_f32c(gamma), _f32c(beta)] if has_norm else []
        sent = torch.empty((B, S, d), dtype=torch.float32, device=dev)
        mask = torch.empty((B, S), dtype=torch.uint8, device=dev)
        doc = torch.empty((B, d), dtype=torch.float32, device=dev)
        pre_doc = torch.empty((B, d), dtype=torch.float32, device=dev)
        stats = torch.empty((B * (S + 1), 2), dtype=torch.float32, device=dev)
        seed = _new_seed(training and has_norm, drop_p)
        c = _call(F32, B, training, 0, drop_p, seed, pt, None, None, None)
        if not has_norm:
            c.params = None
        check(L.mmoe_sent_gather_fwd(C.byref(c), S, d, hd.data_ptr(), _TORCH2MMOE[hd.dtype], src.data_ptr(), sent.data_ptr(), mask.data_ptr(),
                                     doc.data_ptr(), pre_doc.data_ptr(), stats.data_ptr()), "sent_gather_fwd")
        ctx.state = (training, drop_p, seed, src, hd, pt, mask, pre_doc, stats, h.shape, h.dtype)
        ctx.has_norm = has_norm
        ctx.mark_non_differentiable(mask)
        return sent, mask.view(torch.bool), doc

    @staticmethod
    @once_differentiable
    def backward(ctx, d_sent, _d_mask, d_doc):
        training, drop_p, seed, src, hd, pt, mask, pre_doc, stats, h_shape, h_dtype = _state(ctx)
        L = lib()
        B, S = src.shape
        d = hd.shape[-1]
        dev = hd.device
        dh = torch.zeros(hd.shape, dtype=torch.float32, device=dev)
        g = torch.zeros((2, d), dtype=torch.float32, device=dev) if ctx.has_norm else None
        c = _call(F32, B, training, 0, drop_p, seed, pt, [g[0].data_ptr(), g[1].data_ptr()] if ctx.has_norm else None, None, None)
        if not ctx.has_norm:
            c.params = None
        ds = _f32c(d_sent) if d_sent is not None else None
        dd = _f32c(d_doc) if d_doc is not None else None
        check(L.mmoe_sent_gather_bwd(C.byref(c), S, d, hd.data_ptr(), _TORCH2MMOE[hd.dtype], src.data_ptr(), mask.data_ptr(), pre_doc.data_ptr(),
                                     stats.data_ptr(), ds.data_ptr() if ds is not None else None, dd.data_ptr() if dd is not None else None,
                                     dh.data_ptr()), "sent_gather_bwd")
        ctx.state = None
        return (None, None, None, dh.reshape(h_shape).to(h_dtype), g[0] if ctx.has_norm else None, g[1] if ctx.has_norm else None)


def sentence_gather(h: torch.Tensor, chunk2sample, sent_pos, max_sent_count: int, norm: Optional[nn.LayerNorm], drop_p: float,
                    training: bool, n_samples: Optional[int] = None):
    """(sent_vecs [B,S,d], sent_mask [B,S] bool, doc_vecs [B,d]) of model.py:286-338 from the encoder's last hidden state
    h [N_chunks, seq_len, d].  norm = TextExpert.norm (None for the HoME variant, which skips the final LayerNorm/dropout)."""
    src = build_slot_table(chunk2sample, sent_pos, h.shape[1], max_sent_count, n_samples).to(h.device)
    if norm is not None:
        return _SentGather.apply(training, float(drop_p), src, h, norm.weight, norm.bias)
    return _SentGather.apply(training, 0.0, src, h, None, None)
