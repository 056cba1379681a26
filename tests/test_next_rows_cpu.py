"""SURVEY.md §8f rows on the CPU: the oracle restatements (oracle/next_rows.py) against the golden vectors the reference's
own code produced (oracle/make_golden_next.py), and the host-side logic of the native wrappers (slot table of the sentence
gather, TextExpert's CPU path, weight folding of the patch projection)."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import make_golden_next as G
from oracle import next_rows as N


def _close(a, b, tol=1e-5):
    a, b = a.double(), b.double()
    return float((a - b).abs().max()) <= tol * max(float(b.abs().max()), 1e-30)


@pytest.mark.parametrize("mode", ["train", "eval"])
def test_wrapper_stack_oracle_matches_reference_golden(mode):
    g = load_golden("next_wrapper_b16")[mode]
    xs, w, b, rm, rv, cot = G.wrapper_inputs()
    xo = [x.double().requires_grad_(True) for x in xs]
    wo = [t.double().requires_grad_(True) for t in w]
    bo = [t.double().requires_grad_(True) for t in b]
    out, nrm, nrv = N.home_wrapper_stack(xo, wo, bo, [t.double() for t in rm], [t.double() for t in rv], mode == "train")
    out.backward(cot.double())
    assert _close(out, g["out"])
    for e in range(6):
        assert _close(xo[e].grad, g["dx"][e], 1e-4) and _close(wo[e].grad, g["dgamma"][e], 1e-4) and _close(bo[e].grad, g["dbeta"][e], 1e-4)
        assert _close(nrm[e], g["running_mean"][e]) and _close(nrv[e], g["running_var"][e])


def test_loss_oracles_match_reference_golden():
    g = load_golden("next_losses")
    import oracle.synth as synth
    ui, idoc, udoc, proj = (synth.normal(312, (16, 768), k).double().requires_grad_(True) for k in range(4))
    lo = [N.info_nce(ui, idoc), N.info_nce(udoc, proj), N.info_nce(idoc, proj)]
    for i in range(3):
        assert abs(float(lo[i]) - float(g["info_nce"]["loss"][i])) < 1e-5
    (0.5 * lo[0] + 0.7 * lo[1] + 1.3 * lo[2]).backward()
    for k, t in (("ui", ui), ("idoc", idoc), ("udoc", udoc), ("proj", proj)):
        assert _close(t.grad, g["info_nce"]["grads"][k], 1e-4)
    n = 5000
    s = np.round(synth.normal(313, (n,), 1).numpy() * 4.0) / 4.0 + (synth.uniform01(313, n, 2) < 0.5) * 0.5
    y = (synth.uniform01(313, n, 3) < 0.35).astype(np.float32)
    assert abs(N.roc_auc(s, y) - g["auc"]["ties"]) < 1e-12


def test_patch_oracle_and_weight_folding_match_reference_golden():
    g = load_golden("next_patches")
    raw = G.patch_bytes(g["B"])
    w, bias = G.patch_weights()
    images = torch.stack([N.unpatchify_normalise(raw[b]) for b in range(g["B"])])
    assert _close(N.patch_embed(images.double(), w.double(), bias.double()), g["tokens"])
    # the folding the native path uses for raw bytes: W' = W / (255 std), b' = b - sum W mean / std
    w2 = w.double().reshape(768, 3, 256)
    mean = torch.tensor(N.IMAGENET_MEAN, dtype=torch.float64).view(1, 3, 1)
    std = torch.tensor(N.IMAGENET_STD, dtype=torch.float64).view(1, 3, 1)
    b2 = bias.double() - (w2 * (mean / std)).sum(dim=(1, 2))
    w2 = (w2 / (255.0 * std)).reshape(768, 768)
    tokens = torch.from_numpy(raw.astype(np.float64)) @ w2.t() + b2
    assert _close(tokens, g["tokens"])


@pytest.mark.parametrize("which", ["model", "model_HoME"])
def test_sentence_gather_oracle_slot_table_and_textexpert_cpu_path(which):
    g = load_golden("next_gather")[which]
    h, c2s, pos, S = G.gather_inputs()
    nw, nb = (g["norm_w"].double(), g["norm_b"].double()) if which == "model" else (None, None)
    sent, mask, doc = N.sentence_gather(h.double(), c2s, pos, S, nw, nb)
    assert _close(sent, g["sent"]) and _close(doc, g["doc"]) and torch.equal(mask, g["mask"])
    # the host-side slot table of the native kernel reproduces the reference's bucketing
    import mmoe_multimodal_rec_b200 as pkg
    src = pkg.ingest.build_slot_table(c2s, pos, h.shape[1], S)
    rows = torch.cat([h.reshape(-1, 768), torch.zeros(1, 768)], 0)
    padded = rows[torch.where(src >= 0, src.long(), torch.full_like(src, rows.shape[0] - 1).long())]
    ref_rows = N.sentence_gather(h, c2s, pos, S, None, None)[0]
    assert torch.equal(padded, ref_rows)
    # the drop-in TextExpert (CPU path) against what the reference's TextExpert produced
    import types
    from mmoe_multimodal_rec_b200 import text_data as TD
    enc = G._FakeEncoder(h.clone())
    cls = TD.TextExpert if which == "model" else TD.TextExpertHoME
    te = cls(enc, types.SimpleNamespace(pad_token_id=0)).eval()
    if which == "model":
        with torch.no_grad():
            te.norm.weight.copy_(g["norm_w"]); te.norm.bias.copy_(g["norm_b"])
    ids = [[1] * h.shape[1] for _ in c2s]
    out = te(ids, c2s, pos, S, trainable=True) if which == "model" else te(ids, c2s, pos, S)
    assert _close(out[0], g["sent"]) and torch.equal(out[1], g["mask"]) and _close(out[2], g["doc"])
