"""Drop-in replacement for the reference's ``model.py`` (JingxiangQU/mmoe-multimodal-rec).

``train.py`` / ``inference_and_auc.py`` do ``from model import (preprocess_batch, decode_sample,
build_text_user_expert, build_text_item_expert, build_img_expert, build_cross_expert, build_concat_ui_expert,
build_concat_ti_expert, TwoTaskMMoE)`` (train.py:20-30) — put this file (and the repo root) ahead of the reference's
on ``sys.path`` and they run unchanged, with the fusion-and-head modules executing hand-written sm_100a CUDA
(``mmoe-multimodal-rec_b200/``) instead of torch eager.  Class names, constructor arguments, forward signatures
and ``state_dict()`` keys equal the reference's, so its checkpoints load with strict=True.
"""
import torch  # noqa: F401

import mmoe_multimodal_rec_b200 as _pkg
from mmoe_multimodal_rec_b200.modules import (AttnPool1D, DenseGate, EnhancedCrossFuse, ItemImageExpert,  # noqa: F401
                                              RobustTextCrossExpert, RobustTransformerLayer, TwoTaskMMoE)
from mmoe_multimodal_rec_b200.text_data import (TextExpert, build_text_expert, decode_sample, nltk_sentence_split,  # noqa: F401
                                                preprocess_batch, safe_float)


def build_text_user_expert(model_name: str, lora_r: int, max_tok: int, tokenizer, device: torch.device) -> TextExpert:
    """reference model.py:585-602"""
    return build_text_expert(TextExpert, model_name, lora_r, max_tok, tokenizer, device)


def build_text_item_expert(model_name: str, lora_r: int, max_tok: int, tokenizer, device: torch.device) -> TextExpert:
    """reference model.py:605-620"""
    return build_text_expert(TextExpert, model_name, lora_r, max_tok, tokenizer, device)


def build_img_expert(model_name: str, pool_type: str, device: torch.device):
    """reference model.py:623-628: pretrained HF ViT backbone inside the native pool+LN+dropout wrapper."""
    from transformers import ViTModel
    return ItemImageExpert(base_model=ViTModel.from_pretrained(model_name), pool_type=pool_type).to(device)


def build_cross_expert(d: int = 768, n_layer: int = 2, n_head: int = 8, dropout: float = 0.1,
                       device: torch.device = None) -> RobustTextCrossExpert:
    """reference model.py:630-638"""
    return RobustTextCrossExpert(d=d, n_layer=n_layer, n_head=n_head, dropout=dropout).to(device)


def build_concat_ui_expert(d: int = 768, n_head: int = 8, depth: int = 2, dropout: float = 0.1,
                           device: torch.device = None) -> EnhancedCrossFuse:
    """reference model.py:641-648"""
    return EnhancedCrossFuse(d=d, n_head=n_head, depth=depth, dropout=dropout).to(device)


def build_concat_ti_expert(d: int = 768, n_head: int = 8, depth: int = 2, dropout: float = 0.1,
                           device: torch.device = None) -> EnhancedCrossFuse:
    """reference model.py:651-658"""
    return EnhancedCrossFuse(d=d, n_head=n_head, depth=depth, dropout=dropout).to(device)
