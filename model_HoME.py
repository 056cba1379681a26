"""Drop-in replacement for the reference's ``model_HoME.py``: same import surface as ``train_HoME.py:22-32`` /
``infer_auc_HoME:16-26`` expect (``preprocess_batch, decode_sample, build_* , HOME_MMoE_Complete`` ...), with the
HoME fusion experts and the hierarchical-expert head running on the native sm_100a library."""
import torch  # noqa: F401

import mmoe_multimodal_rec_b200 as _pkg
from mmoe_multimodal_rec_b200.modules_home import (AttnPool1D, DenseGate, EnhancedCrossFuse, ExpertMLP, FeatureGate,  # noqa: F401
                                                   HOME_MMoE_Complete, ImageExpertWithProjection, RobustTextCrossExpert,
                                                   RobustTransformerLayer, SelfGate)
from mmoe_multimodal_rec_b200.text_data import TextExpertHoME as TextExpert  # noqa: F401
from mmoe_multimodal_rec_b200.text_data import (build_text_expert, decode_sample, nltk_sentence_split,  # noqa: F401
                                                preprocess_batch, safe_float)


def build_text_user_expert(model_name: str, lora_r: int, max_tok: int, tokenizer, device: torch.device) -> TextExpert:
    """reference model_HoME.py:646-664"""
    return build_text_expert(TextExpert, model_name, lora_r, max_tok, tokenizer, device)


def build_text_item_expert(model_name: str, lora_r: int, max_tok: int, tokenizer, device: torch.device) -> TextExpert:
    """reference model_HoME.py:667-682"""
    return build_text_expert(TextExpert, model_name, lora_r, max_tok, tokenizer, device)


def build_img_expert(model_name: str, device: torch.device):
    """reference model_HoME.py:684-698: ViT backbone + trainable projection head (native GEMMs)."""
    from transformers import ViTModel
    return ImageExpertWithProjection(vit_model=ViTModel.from_pretrained(model_name), expert_dim=768, projection_dim=768).to(device)


def build_cross_expert(d: int = 768, n_layer: int = 2, n_head: int = 8, dropout: float = 0.1,
                       device: torch.device = None) -> RobustTextCrossExpert:
    return RobustTextCrossExpert(d=d, n_layer=n_layer, n_head=n_head, dropout=dropout).to(device)


def build_concat_ui_expert(d: int = 768, n_head: int = 8, depth: int = 2, dropout: float = 0.1,
                           device: torch.device = None) -> EnhancedCrossFuse:
    return EnhancedCrossFuse(d=d, n_head=n_head, depth=depth, dropout=dropout).to(device)


def build_concat_ti_expert(d: int = 768, n_head: int = 8, depth: int = 2, dropout: float = 0.1,
                           device: torch.device = None) -> EnhancedCrossFuse:
    return EnhancedCrossFuse(d=d, n_head=n_head, depth=depth, dropout=dropout).to(device)
