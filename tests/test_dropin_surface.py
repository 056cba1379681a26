"""The drop-in ``model.py`` / ``model_HoME.py`` export what the reference scripts import, with constructor
signatures and state_dict keys equal to the reference's (checked live against /root/reference when mounted,
and against the oracle's shape tables everywhere)."""
import inspect

import pytest
import torch

from oracle import synth
from oracle.ref_import import reference_available

V1_NAMES = ["preprocess_batch", "decode_sample", "build_text_user_expert", "build_text_item_expert", "build_img_expert",
            "build_cross_expert", "build_concat_ui_expert", "build_concat_ti_expert", "TwoTaskMMoE"]


def test_import_surface():
    import model
    import model_HoME
    for n in V1_NAMES:
        assert hasattr(model, n), n
        assert hasattr(model_HoME, n) or n == "TwoTaskMMoE", n
    assert hasattr(model_HoME, "HOME_MMoE_Complete")


def test_state_dict_keys_match_shape_tables():
    import model
    import model_HoME
    pairs = [
        (model.TwoTaskMMoE(), synth.mmoe_head_shapes()),
        (model.RobustTextCrossExpert(), synth.cross_expert_shapes()),
        (model.EnhancedCrossFuse(), synth.cross_fuse_shapes()),
        (model_HoME.RobustTextCrossExpert(), synth.cross_expert_shapes()),
        (model_HoME.EnhancedCrossFuse(), synth.cross_fuse_shapes()),
        (model_HoME.HOME_MMoE_Complete(expert_dim=768, n_shared_experts=4, n_task_experts=2, tower_hidden=512),
         synth.home_head_shapes(tower_hidden=512)),
    ]
    for mod, shapes in pairs:
        sd = mod.state_dict()
        assert list(sd.keys()) == list(shapes.keys()), type(mod).__name__
        for k, shp in shapes.items():
            assert tuple(sd[k].shape) == tuple(shp), (type(mod).__name__, k)


@pytest.mark.skipif(not reference_available(), reason="/root/reference not mounted")
def test_signatures_and_state_dicts_match_live_reference():
    import model
    import model_HoME
    from oracle.ref_import import load_reference
    rm, rh = load_reference("model"), load_reference("model_HoME")
    for ours, ref, names in ((model, rm, ["TwoTaskMMoE", "RobustTextCrossExpert", "EnhancedCrossFuse", "AttnPool1D", "DenseGate",
                                          "ItemImageExpert", "build_cross_expert", "build_concat_ui_expert", "build_concat_ti_expert",
                                          "build_img_expert", "build_text_user_expert", "preprocess_batch", "decode_sample"]),
                             (model_HoME, rh, ["HOME_MMoE_Complete", "RobustTextCrossExpert", "EnhancedCrossFuse", "FeatureGate",
                                               "SelfGate", "DenseGate", "ExpertMLP", "ImageExpertWithProjection", "build_img_expert",
                                               "preprocess_batch", "decode_sample"])):
        for n in names:
            a, b = getattr(ours, n), getattr(ref, n)
            fa = a.__init__ if inspect.isclass(a) else a
            fb = b.__init__ if inspect.isclass(b) else b
            pa = [(p.name, p.default) for p in inspect.signature(fa).parameters.values() if not p.name.startswith("_")]
            pb = [(p.name, p.default) for p in inspect.signature(fb).parameters.values()]
            assert pa == pb, (n, pa, pb)
            if inspect.isclass(a) and issubclass(a, torch.nn.Module) and n not in ("ItemImageExpert", "ImageExpertWithProjection"):
                if n in ("AttnPool1D",):
                    sa, sb = a(768).state_dict(), b(768).state_dict()
                elif n in ("DenseGate",):
                    sa, sb = a(768, 6).state_dict(), b(768, 6).state_dict()
                elif n in ("FeatureGate",):
                    sa, sb = a(768, 4).state_dict(), b(768, 4).state_dict()
                elif n in ("SelfGate",):
                    sa, sb = a(768).state_dict(), b(768).state_dict()
                else:
                    sa, sb = a().state_dict(), b().state_dict()
                assert [(k, tuple(v.shape)) for k, v in sa.items()] == [(k, tuple(v.shape)) for k, v in sb.items()], n
    # forward signatures of the hot-path modules
    for ours, ref, n in ((model, rm, "RobustTextCrossExpert"), (model, rm, "EnhancedCrossFuse"), (model, rm, "TwoTaskMMoE"),
                         (model, rm, "ItemImageExpert"), (model_HoME, rh, "HOME_MMoE_Complete"), (model_HoME, rh, "ImageExpertWithProjection")):
        pa = list(inspect.signature(getattr(ours, n).forward).parameters)
        pb = list(inspect.signature(getattr(ref, n).forward).parameters)
        assert pa == pb, (n, pa, pb)


class _Tok:
    """Minimal whitespace tokenizer with the attributes preprocess_batch uses."""
    cls_token_id, sep_token_id, pad_token_id, vocab_size = 101, 102, 0, 30522

    def convert_tokens_to_ids(self, t):
        return 30522

    def encode(self, s, add_special_tokens=False, max_length=None, truncation=True):
        ids = [1000 + (hash(w) % 20000) for w in s.split()]
        return ids[:max_length] if max_length else ids


@pytest.mark.skipif(not reference_available(), reason="/root/reference not mounted")
def test_preprocess_batch_and_decode_sample_equal_reference():
    import json

    import numpy as np

    import model
    from oracle.ref_import import load_reference
    rm = load_reference("model")
    split = lambda t: [s for s in t.split(". ") if s]
    rm.nltk_sentence_split = split
    tok = _Tok()
    texts = ["", "one sentence only", ". ".join("w%d " % i * (3 + i % 40) for i in range(90)), ". ".join(["a b c"] * 300),
             "x " * 500 + ". " + "y " * 500]
    for max_tok, max_chunks in ((384, 4), (32, 2), (16, 4)):
        ours = model.preprocess_batch(texts, tok, max_tok, max_chunks_per_sample=max_chunks, _split=split)
        ref = rm.preprocess_batch(texts, tok, max_tok, max_chunks_per_sample=max_chunks)
        assert ours == ref
    patch = np.random.default_rng(0).integers(0, 256, size=(196, 3, 16, 16), dtype=np.uint8)
    good = {"user.json": b" u ", "item.json": b"i", "label.json": json.dumps({"label_good": 1, "label_best": 0}).encode(),
            "misc.json": json.dumps({"has_image": 1, "shape": [196, 3, 16, 16]}).encode(), "patch.bin": patch.tobytes()}
    for sample in (good, {**good, "label.json": b"{}"}, {**good, "user.json": b""}, {**good, "patch.bin": b"123"},
                   {k: v for k, v in good.items() if k != "misc.json"}, {**good, "label.json": json.dumps({"label_good": 2, "label_best": 0}).encode()}):
        a, b = model.decode_sample(sample), rm.decode_sample(sample)
        assert (a is None) == (b is None)
        if a is not None:
            assert a["user_text"] == b["user_text"] and a["label_good"] == b["label_good"]
            assert torch.equal(a["patch"], b["patch"])
