"""CPU-side helpers the reference scripts import from ``model`` / ``model_HoME`` that are NOT on the
accelerated path (SURVEY.md §2.1 marks them out of scope): sentence splitting, chunk packing,
WebDataset sample decoding and the BERT-based TextExpert wrapper.  They are re-stated here in
plain Python / torch so that ``from model import preprocess_batch, decode_sample, ...`` keeps
working; behaviour follows the reference line by line (citations below), the code is new.

Third-party packages the reference needs for these helpers (nltk, peft, transformers) are
imported lazily, only by the functions that use them.
"""
from __future__ import annotations

import json
from typing import List

import numpy as np
import torch
import torch.nn as nn

IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


def nltk_sentence_split(text: str) -> list:
    """Punkt sentence split; empty text gives [] (reference model.py:20-26)."""
    if not text:
        return []
    from nltk.tokenize import sent_tokenize
    return sent_tokenize(text)


def preprocess_batch(texts: List[str], tokenizer, max_tok: int, max_chunks_per_sample: int = 4, fixed_sent_count: int = 64,
                     _split=None):
    """Pack ``<SENT>``-prefixed sentences into at most ``max_chunks_per_sample`` chunks of at most
    ``max_tok`` tokens per text (reference model.py:29-117).

    Returns (input_ids [n_chunks][max_chunk_len], chunk2sample [n_chunks],
    sent_pos [n_chunks][max_sents_per_chunk] padded with -1, fixed_sent_count).
    """
    split = _split or nltk_sentence_split
    room = max_tok - 2                              # CLS and SEP take two slots
    sent_id = tokenizer.convert_tokens_to_ids("<SENT>")
    cls_id, sep_id, pad_id = tokenizer.cls_token_id, tokenizer.sep_token_id, tokenizer.pad_token_id

    chunks: List[List[int]] = []
    owners: List[int] = []
    positions: List[List[int]] = []

    def emit(sample_idx, body, marks):
        chunks.append([cls_id] + body + [sep_id])
        owners.append(sample_idx)
        positions.append([m + 1 for m in marks])    # +1 for the leading CLS

    for sample_idx, text in enumerate(texts):
        body: List[int] = []
        marks: List[int] = []
        n_emitted = 0
        for sentence in split(text):
            if n_emitted >= max_chunks_per_sample:
                break
            piece = [sent_id] + tokenizer.encode(sentence, add_special_tokens=False, max_length=room - 1, truncation=True)
            if len(body) + len(piece) > room:
                # the running chunk is closed as it is (even if empty) and the sentence opens the next one
                emit(sample_idx, body, marks)
                n_emitted += 1
                body, marks = list(piece), [0]
            else:
                marks.append(len(body))
                body.extend(piece)
        if n_emitted < max_chunks_per_sample and body:
            emit(sample_idx, body, marks)

    width = max((len(c) for c in chunks), default=0)
    max_marks = max((len(p) for p in positions), default=0)
    vocab = tokenizer.vocab_size
    input_ids = []
    for c in chunks:
        row = c + [pad_id] * (width - len(c))
        # ids beyond the base vocabulary (e.g. the added <SENT>) are rewritten to [PAD] (model.py:101-109)
        input_ids.append([t if t < vocab else pad_id for t in row])
    sent_pos = [p + [-1] * (max_marks - len(p)) for p in positions]
    return input_ids, owners, sent_pos, fixed_sent_count


def safe_float(x, default=0.0):
    """float(x), or ``default`` for anything unparsable / NaN / inf (reference model.py:121-126)."""
    try:
        v = float(x)
    except Exception:
        return default
    return default if (np.isnan(v) or np.isinf(v)) else v


def _unpatchify(raw: bytes, shape) -> torch.Tensor:
    """uint8 patches [196,3,16,16] -> normalised image [3,224,224] (reference model.py:166-176)."""
    patches = torch.from_numpy(np.frombuffer(raw, dtype=np.uint8).copy().reshape(shape)).float() / 255.0
    # patch index = row*14 + col  ->  (channel, row, y, col, x)
    img = patches.reshape(14, 14, 3, 16, 16).permute(2, 0, 3, 1, 4).reshape(3, 224, 224)
    mean = torch.tensor(IMAGENET_MEAN)[:, None, None]
    std = torch.tensor(IMAGENET_STD)[:, None, None]
    img = (img - mean) / std
    if torch.isnan(img).any() or torch.isinf(img).any():
        return torch.zeros(3, 224, 224)
    return img


def decode_sample(sample: dict):
    """WebDataset sample -> dict(user_text, item_text, patch, label_good, label_best) or None for anything
    malformed (reference model.py:127-189)."""
    try:
        user_b, item_b, label_b = sample.get("user.json", b""), sample.get("item.json", b""), sample.get("label.json", b"")
        misc_b = sample.get("misc.json", b"")
        if not user_b or not item_b or not label_b:
            return None
        user_text = user_b.decode("utf-8").strip()
        item_text = item_b.decode("utf-8").strip()
        label = json.loads(label_b)
        misc = json.loads(misc_b) if misc_b else {}
        if not user_text or not item_text:
            return None
        if "label_good" not in label or "label_best" not in label:
            return None
        y_good, y_best = safe_float(label["label_good"]), safe_float(label["label_best"])
        if not (0 <= y_good <= 1) or not (0 <= y_best <= 1):
            return None
        image = torch.zeros(3, 224, 224)
        if misc.get("has_image", 0) and "patch.bin" in sample:
            try:
                image = _unpatchify(sample["patch.bin"], misc["shape"])
            except Exception:
                image = torch.zeros(3, 224, 224)
        return {"user_text": user_text, "item_text": item_text, "patch": image, "label_good": y_good, "label_best": y_best}
    except Exception:
        return None


class TextExpert(nn.Module):
    """BERT(+LoRA) sentence encoder wrapper (reference model.py:214-338; HoME variant
    model_HoME.py:256-369 skips the final LayerNorm/dropout and has no ``trainable`` switch).

    Stays a torch module by design (BASELINE north_star: "the text encoders ... stay reference torch
    modules ... timed separately").  The per-sample Python bucketing loop of the reference is replaced
    by one vectorised scatter with identical results.
    """

    final_norm = True

    def __init__(self, encoder: nn.Module, tokenizer, max_tok=384, d=768):
        super().__init__()
        self.encoder = encoder
        self.max_tok = max_tok
        self.tokenizer = tokenizer
        self.norm = nn.LayerNorm(d)
        self.dropout = nn.Dropout(0.1)

    def forward(self, input_ids, chunk2sample, sent_pos, max_sent_count, trainable=False):
        return self.forward_precomputed(input_ids, chunk2sample, sent_pos, max_sent_count, trainable)

    def forward_precomputed(self, input_ids, chunk2sample, sent_pos, max_sent_count: int, trainable: bool = False):
        device = next(self.encoder.parameters()).device
        if not input_ids:
            B = len(set(chunk2sample)) if chunk2sample else 1
            D = self.encoder.config.hidden_size
            return (torch.zeros(B, max_sent_count, D, device=device),
                    torch.ones(B, max_sent_count, dtype=torch.bool, device=device),
                    torch.zeros(B, D, device=device))
        x = torch.tensor(input_ids, device=device)
        kwargs = dict(input_ids=x, attention_mask=(x != self.tokenizer.pad_token_id).long(), token_type_ids=torch.zeros_like(x),
                      position_ids=torch.arange(x.size(1), dtype=torch.long, device=device).unsqueeze(0).expand_as(x))
        if trainable:
            h = self.encoder(**kwargs).last_hidden_state
        else:
            with torch.no_grad():
                h = self.encoder(**kwargs).last_hidden_state
        if h.is_cuda:
            # one native launch: gather the <SENT> rows, bucket, pad, mask, mean, LayerNorm, dropout (csrc/ingest.cu)
            from .ingest import sentence_gather
            return sentence_gather(h, chunk2sample, sent_pos, max_sent_count, self.norm if self.final_norm else None,
                                   float(self.dropout.p), self.training)
        return self._gather_torch(h, chunk2sample, sent_pos, max_sent_count)

    def _gather_torch(self, h, chunk2sample, sent_pos, max_sent_count: int):
        """The same step with torch ops, for an encoder that lives on the CPU (the text experts are reference torch modules and
        not part of the native path; on CUDA the native kernel above always runs)."""
        from .ingest import build_slot_table
        device = h.device
        D = h.size(-1)
        src = build_slot_table(chunk2sample, sent_pos, h.size(1), max_sent_count).to(device).long()
        rows = torch.cat([h.reshape(-1, D), torch.zeros(1, D, device=device, dtype=h.dtype)], 0)
        padded = rows[torch.where(src >= 0, src, torch.full_like(src, rows.size(0) - 1))]
        sent_mask = padded.abs().sum(-1) == 0
        lens = (~sent_mask).sum(dim=1, keepdim=True)
        doc = padded.sum(dim=1) / lens.clamp(min=1)
        if self.final_norm:
            padded = self.dropout(self.norm(padded))
            doc = self.dropout(self.norm(doc))
        return padded, sent_mask, doc


class TextExpertHoME(TextExpert):
    """model_HoME.py:256-369: encoder always runs with autograd on; no final LayerNorm / dropout."""

    final_norm = False

    def forward(self, input_ids, chunk2sample, sent_pos, max_sent_count):
        return self.forward_precomputed(input_ids, chunk2sample, sent_pos, max_sent_count, trainable=True)


def _lora_text_encoder(model_name: str, lora_r: int, tokenizer, device):
    """AutoModel + resized embeddings + LoRA(r, alpha 32, dropout 0.1) (reference model.py:585-602)."""
    from peft import LoraConfig, TaskType, get_peft_model
    from transformers import AutoModel
    cfg = LoraConfig(task_type=TaskType.FEATURE_EXTRACTION, r=lora_r, lora_alpha=32, lora_dropout=0.1)
    base = AutoModel.from_pretrained(model_name)
    base.resize_token_embeddings(len(tokenizer))
    return get_peft_model(base, cfg).to(device)


def build_text_expert(cls, model_name: str, lora_r: int, max_tok: int, tokenizer, device):
    return cls(_lora_text_encoder(model_name, lora_r, tokenizer, device), tokenizer, max_tok=max_tok).to(device)
