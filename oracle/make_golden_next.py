"""Pin oracle/next_rows.py against the reference's own code and write tests/golden/next_*.pt (dev container only:
needs /root/reference, transformers and scikit-learn).

    python -m oracle.make_golden_next

For every §8f row: run the reference's code (``train_HoME.HomeExpertWrapper``, ``train_HoME.calculate_contrastive_loss``,
``nn.BCEWithLogitsLoss`` as train.py:189-192 builds it, ``sklearn.metrics.roc_auc_score``, ``model.decode_sample`` + HF
``ViTPatchEmbeddings``, ``model.TextExpert`` / ``model_HoME.TextExpert`` with a stand-in encoder) and the oracle on the
same deterministic inputs (oracle/synth.py), ASSERT agreement, and store inputs' seeds + the reference's outputs.
"""
from __future__ import annotations

import io
import json
import os
import sys
import types

import numpy as np
import torch

from . import next_rows as N
from . import synth
from .ref_import import load_reference

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden")
TOL = 2e-5


def close(a, b, what, tol=TOL):
    a, b = a.detach().double(), b.detach().double()
    err = float((a - b).abs().max()) / max(float(b.abs().max()), 1e-30)
    assert err <= tol, f"{what}: oracle vs reference {err:.3e}"
    return err


def wrapper_inputs(B=16, d=768, n=6, seed=301):
    xs = [synth.normal(seed, (B, d), 10 + e) * (1.0 + 0.3 * e) + 0.2 * e for e in range(n)]
    w = [1.0 + 0.2 * synth.uniform_pm1(seed, (d,), 30 + e) for e in range(n)]
    b = [0.2 * synth.uniform_pm1(seed, (d,), 50 + e) for e in range(n)]
    rm = [0.1 * synth.uniform_pm1(seed, (d,), 70 + e) for e in range(n)]
    rv = [1.0 + 0.5 * synth.uniform01(seed, d, 90 + e).astype(np.float32) for e in range(n)]
    rv = [torch.from_numpy(v) for v in rv]
    cot = synth.normal(seed, (B, n, d), 5)
    return xs, w, b, rm, rv, cot


def golden_wrapper():
    ref = load_reference("train_HoME")
    xs, w, b, rm, rv, cot = wrapper_inputs()
    res = {}
    for mode in ("train", "eval"):
        mods = []
        for e in range(len(xs)):
            m = ref.HomeExpertWrapper(768, dropout_p=0.0)          # dropout checked separately through the keep-mask hook
            with torch.no_grad():
                m.norm.weight.copy_(w[e]); m.norm.bias.copy_(b[e]); m.norm.running_mean.copy_(rm[e]); m.norm.running_var.copy_(rv[e])
            m.train(mode == "train")
            mods.append(m)
        xin = [x.clone().requires_grad_(True) for x in xs]
        out = torch.stack([m(x) for m, x in zip(mods, xin)], dim=1)          # train_HoME.py:350-356
        out.backward(cot)
        # oracle
        xo = [x.double().clone().requires_grad_(True) for x in xs]
        wo = [t.double().clone().requires_grad_(True) for t in w]
        bo = [t.double().clone().requires_grad_(True) for t in b]
        o, nrm, nrv = N.home_wrapper_stack(xo, wo, bo, [t.double() for t in rm], [t.double() for t in rv], mode == "train")
        o.backward(cot.double())
        close(o, out, f"wrapper {mode} out")
        for e in range(len(xs)):
            close(xo[e].grad, xin[e].grad, f"wrapper {mode} dx{e}", 1e-4)
            close(wo[e].grad, mods[e].norm.weight.grad, f"wrapper {mode} dgamma{e}", 1e-4)
            close(bo[e].grad, mods[e].norm.bias.grad, f"wrapper {mode} dbeta{e}", 1e-4)
            close(nrm[e], mods[e].norm.running_mean, f"wrapper {mode} running_mean{e}")
            close(nrv[e], mods[e].norm.running_var, f"wrapper {mode} running_var{e}")
        res[mode] = {"out": out.detach(), "dx": [x.grad for x in xin], "dgamma": [m.norm.weight.grad for m in mods],
                     "dbeta": [m.norm.bias.grad for m in mods], "running_mean": [m.norm.running_mean.clone() for m in mods],
                     "running_var": [m.norm.running_var.clone() for m in mods],
                     "num_batches_tracked": [int(m.norm.num_batches_tracked) for m in mods]}
    torch.save(res, os.path.join(OUT, "next_wrapper_b16.pt"))


def golden_losses():
    ref = load_reference("train_HoME")
    B, d = 64, 768
    lg, lb = synth.normal(311, (B,), 1) * 2.0, synth.normal(311, (B,), 2) * 2.0
    yg = (synth.uniform01(311, B, 3) < 0.5).astype(np.float32)
    yb = (synth.uniform01(311, B, 4) < 0.3).astype(np.float32)
    yg, yb = torch.from_numpy(yg), torch.from_numpy(yb)
    pw_g, pw_b = 858627.0 / 990303.0, 1328721.0 / 520209.0                    # train.py:189-192
    a, b_ = lg.clone().requires_grad_(True), lb.clone().requires_grad_(True)
    loss = torch.nn.BCEWithLogitsLoss(pos_weight=torch.tensor(pw_g))(a, yg) + torch.nn.BCEWithLogitsLoss(pos_weight=torch.tensor(pw_b))(b_, yb)
    loss.backward()
    ao, bo = lg.double().clone().requires_grad_(True), lb.double().clone().requires_grad_(True)
    lo = N.bce2(ao, bo, yg.double(), yb.double(), pw_g, pw_b)
    lo.backward()
    close(lo, loss, "bce2"); close(ao.grad, a.grad, "bce2 dgood"); close(bo.grad, b_.grad, "bce2 dbest")
    res = {"bce2": {"loss": loss.detach(), "dgood": a.grad, "dbest": b_.grad}}
    # InfoNCE, the three pairs of train_HoME.py:362-364
    ui, idoc, udoc, proj = (synth.normal(312, (16, d), k) for k in range(4))
    pairs = [("ui", "idoc"), ("udoc", "proj"), ("idoc", "proj")]
    t = {"ui": ui, "idoc": idoc, "udoc": udoc, "proj": proj}
    tr = {k: v.clone().requires_grad_(True) for k, v in t.items()}
    losses = [ref.calculate_contrastive_loss(tr[x], tr[y]) for x, y in pairs]
    (0.5 * losses[0] + 0.7 * losses[1] + 1.3 * losses[2]).backward()
    to = {k: v.double().clone().requires_grad_(True) for k, v in t.items()}
    lo = [N.info_nce(to[x], to[y]) for x, y in pairs]
    (0.5 * lo[0] + 0.7 * lo[1] + 1.3 * lo[2]).backward()
    for i in range(3):
        close(lo[i], losses[i], f"info_nce {i}")
    for k in t:
        close(to[k].grad, tr[k].grad, f"info_nce d{k}", 1e-4)
    res["info_nce"] = {"loss": torch.stack([l.detach() for l in losses]), "grads": {k: v.grad for k, v in tr.items()}, "weights": [0.5, 0.7, 1.3]}
    # ROC-AUC with ties, against scikit-learn
    from sklearn.metrics import roc_auc_score
    n = 5000
    s = np.round(synth.normal(313, (n,), 1).numpy() * 4.0) / 4.0 + (synth.uniform01(313, n, 2) < 0.5) * 0.5
    y = (synth.uniform01(313, n, 3) < 0.35).astype(np.float32)
    ref_auc = float(roc_auc_score(y, s))
    got = N.roc_auc(s, y)
    assert abs(got - ref_auc) < 1e-12, (got, ref_auc)
    s2 = synth.normal(314, (4099,), 1).numpy()
    y2 = (synth.uniform01(314, 4099, 3) < 0.5).astype(np.float32)
    ref_auc2 = float(roc_auc_score(y2, s2))
    assert abs(N.roc_auc(s2, y2) - ref_auc2) < 1e-12
    res["auc"] = {"ties": ref_auc, "plain": ref_auc2}
    torch.save(res, os.path.join(OUT, "next_losses.pt"))


def patch_weights(seed=322):
    """Conv2d(3, 768, 16, 16) weight / bias of the patch projection, deterministic (regenerated by the tests)."""
    return 0.05 * synth.uniform_pm1(seed, (768, 3, 16, 16), 1), 0.2 * synth.uniform_pm1(seed, (768,), 2)


def patch_bytes(B=2, seed=321):
    return (synth.uniform01(seed, B * 196 * 768, 1) * 256).astype(np.uint8).reshape(B, 196, 768)


def golden_patches():
    ref = load_reference("model")
    from transformers import ViTConfig
    from transformers.models.vit.modeling_vit import ViTPatchEmbeddings
    B = 2
    raw = patch_bytes(B)
    # the reference's decode path: bytes -> decode_sample -> [3,224,224] normalised float
    imgs = []
    for b in range(B):
        sample = {"user.json": b"u", "item.json": b"i", "patch.bin": raw[b].tobytes(),
                  "misc.json": json.dumps({"has_image": 1, "shape": [196, 3, 16, 16]}).encode(),
                  "label.json": json.dumps({"label_good": 1.0, "label_best": 0.0}).encode()}
        out = ref.decode_sample(sample)
        assert out is not None
        imgs.append(out["patch"])
        close(N.unpatchify_normalise(raw[b]), out["patch"], "unpatchify")
    images = torch.stack(imgs)
    pe = ViTPatchEmbeddings(ViTConfig())
    w, bias = patch_weights()
    with torch.no_grad():
        pe.projection.weight.copy_(w); pe.projection.bias.copy_(bias)
        tokens = pe(images)
    o = N.patch_embed(images.double(), pe.projection.weight.double(), pe.projection.bias.double())
    close(o, tokens, "patch_embed", 1e-5)
    torch.save({"tokens": tokens, "seed": 321, "B": B}, os.path.join(OUT, "next_patches.pt"))


class _FakeEncoder(torch.nn.Module):
    """Returns a fixed hidden-state tensor (the gather step does not care how it was produced)."""
    def __init__(self, h):
        super().__init__()
        self.h = torch.nn.Parameter(h)
        self.config = types.SimpleNamespace(hidden_size=h.shape[-1], max_position_embeddings=512)

    def forward(self, **kw):
        return types.SimpleNamespace(last_hidden_state=self.h)


def gather_inputs(seed=331):
    d, seq = 768, 40
    # 5 samples: sample 1 has no chunk at all... (the reference infers B from max(chunk2sample)); sample 3 overflows 8 slots
    chunk2sample = [0, 0, 2, 3, 3, 3, 4]
    sent_pos = [[1, 9, 17, -1], [1, 30, -1, -1], [1, 5, 9, 13], [1, 3, 5, 7], [1, 3, 5, 7], [1, 3, -1, -1], [1, 45, -1, -1]]
    h = synth.normal(seed, (len(chunk2sample), seq, d), 1)
    h[6, 1] = 0.0                                  # a real row that happens to be all zero: masked as padding by the reference
    return h, chunk2sample, sent_pos, 8


def golden_gather():
    res = {}
    h, c2s, pos, S = gather_inputs()
    tok = types.SimpleNamespace(pad_token_id=0)
    ids = [[1] * h.shape[1] for _ in c2s]
    for which in ("model", "model_HoME"):
        ref = load_reference(which)
        te = ref.TextExpert(_FakeEncoder(h.clone()), tok).eval()
        with torch.no_grad():
            te.norm.weight.copy_(1.0 + 0.2 * synth.uniform_pm1(332, (768,), 1)); te.norm.bias.copy_(0.2 * synth.uniform_pm1(332, (768,), 2))
        if which == "model":
            sent, mask, doc = te(ids, c2s, pos, S, trainable=True)
        else:
            sent, mask, doc = te(ids, c2s, pos, S)
        cs, cd = synth.normal(333, tuple(sent.shape), 1), synth.normal(333, tuple(doc.shape), 2)
        torch.autograd.backward([sent, doc], [cs, cd])
        ho = h.double().clone().requires_grad_(True)
        wo = te.norm.weight.detach().double().clone().requires_grad_(True)
        bo = te.norm.bias.detach().double().clone().requires_grad_(True)
        so, mo, do = N.sentence_gather(ho, c2s, pos, S, wo if which == "model" else None, bo if which == "model" else None)
        torch.autograd.backward([so, do], [cs.double(), cd.double()])
        close(so, sent, f"gather {which} sent"); close(do, doc, f"gather {which} doc")
        assert torch.equal(mo, mask)
        close(ho.grad, te.encoder.h.grad, f"gather {which} dh", 1e-4)
        r = {"sent": sent.detach(), "mask": mask, "doc": doc.detach(), "dh": te.encoder.h.grad.clone()}
        if which == "model":
            close(wo.grad, te.norm.weight.grad, "gather dgamma", 1e-4); close(bo.grad, te.norm.bias.grad, "gather dbeta", 1e-4)
            r.update(dgamma=te.norm.weight.grad.clone(), dbeta=te.norm.bias.grad.clone(), norm_w=te.norm.weight.detach().clone(),
                     norm_b=te.norm.bias.detach().clone())
        res[which] = r
    torch.save(res, os.path.join(OUT, "next_gather.pt"))


def main():
    os.makedirs(OUT, exist_ok=True)
    golden_wrapper(); print("wrapper ok")
    golden_losses(); print("losses ok")
    golden_patches(); print("patches ok")
    golden_gather(); print("gather ok")


if __name__ == "__main__":
    main()
