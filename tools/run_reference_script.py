#!/usr/bin/env python
"""Run one of the REFERENCE's own scripts, byte for byte unchanged, on top of the drop-in ``model.py`` / ``model_HoME.py``
(BASELINE north_star: "train.py, train_HoME.py and inference_and_auc.py run unchanged").

    python tools/run_reference_script.py train          [--batches 4] [--batch-size 8] [--grad-accum 2]
    python tools/run_reference_script.py train_HoME     ...
    python tools/run_reference_script.py inference_and_auc

The script file is taken from $MMOE_REFERENCE_DIR, /root/reference, or baseline/_ref/ (a git-ignored staging copy that
``__graft_entry__.build()`` makes in the dev container so that it travels to the GPU box).  Nothing in the script is edited
or monkey-patched; what the harness supplies is the ENVIRONMENT the script expects and this sandbox lacks:

  * ``import model`` / ``import model_HoME`` resolve to the drop-ins at the repo root;
  * the missing third-party packages get minimal stand-ins: ``webdataset`` (a synthetic stream of samples in the shard
    schema of data4model.py:254-258 — user.json / item.json / patch.bin / misc.json / label.json — decoded by the drop-in's
    own ``decode_sample``), ``peft`` (a LoRA-shaped wrapper with trainable ``lora_`` parameters), ``nltk`` (a period
    splitter), ``matplotlib`` (no-op plotting);
  * no network: ``AutoTokenizer / AutoModel / ViTModel.from_pretrained`` return a hashing tokenizer and small random-init
    BERT / ViT models of the right hidden size (768);
  * a one-process NCCL group (RANK=0, WORLD_SIZE=1) so ``dist.init_process_group("nccl")`` and the DDP wrappers work.

The training scripts loop over a hard-coded 5,600 / 7,200 steps per epoch; the synthetic stream ends the run by raising
``HarnessDone`` after ``--batches`` batches, which the harness catches.  Exit code 0 = the script's own loop body ran
that many micro-steps (forward, loss, GradScaler backward, optimizer steps on the sync steps) on the drop-ins.
"""
from __future__ import annotations

import argparse
import glob
import importlib.machinery
import importlib.util
import json
import os
import sys
import tempfile
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402


class HarnessDone(Exception):
    pass


def find_script(name: str) -> str:
    fname = name if name == "infer_auc_HoME" else name + ".py"
    for d in (os.environ.get("MMOE_REFERENCE_DIR"), "/root/reference", os.path.join(ROOT, "baseline", "_ref")):
        if d and os.path.isfile(os.path.join(d, fname)):
            return os.path.join(d, fname)
    raise FileNotFoundError(f"{fname} not found in $MMOE_REFERENCE_DIR, /root/reference or baseline/_ref/")


# ------------------------------------------------------------------------------------------ stand-ins
def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, None)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


WORDS = ("great sturdy cheap broken fast slow blue red small large works failed love hate returned battery screen cable "
         "kitchen garden book album shoes fits tight loose quality value shipping arrived late early gift again").split()


def synth_sample(i: int):
    rng = np.random.RandomState(1000 + i)
    def text(n_sent):
        return " ".join(" ".join(rng.choice(WORDS, size=rng.randint(4, 12))) + "." for _ in range(n_sent))
    patches = rng.randint(0, 256, size=(196, 768), dtype=np.uint8)
    return {"__key__": f"s{i:06d}", "user.json": text(rng.randint(2, 9)).encode(), "item.json": text(rng.randint(2, 7)).encode(),
            "patch.bin": patches.tobytes(), "misc.json": json.dumps({"has_image": 1, "shape": [196, 3, 16, 16]}).encode(),
            "label.json": json.dumps({"label_good": float(rng.rand() < 0.5), "label_best": float(rng.rand() < 0.3)}).encode()}


class FakeWebDataset(torch.utils.data.IterableDataset):
    """The fluent subset of webdataset.WebDataset the scripts use (train.py:46-56, inference_and_auc.py:27-34)."""
    limit_batches = 4
    endless = True

    def __init__(self, urls, **kw):
        self.ops, self.batch, self.collate = [], None, None

    def shuffle(self, n): return self
    def repeat(self): return self
    def map(self, f): self.ops.append(("map", f)); return self
    def select(self, f): self.ops.append(("select", f)); return self
    def batched(self, n, collation_fn=None): self.batch, self.collate = n, collation_fn; return self

    def __iter__(self):
        i, out, n_batches = 0, [], 0
        while True:
            s = synth_sample(i); i += 1
            keep = True
            for kind, f in self.ops:
                if kind == "map":
                    s = f(s)
                elif not f(s):
                    keep = False
                    break
            if not keep:
                continue
            out.append(s)
            if len(out) == self.batch:
                if n_batches == FakeWebDataset.limit_batches:
                    if FakeWebDataset.endless:
                        raise HarnessDone()
                    return
                n_batches += 1
                yield self.collate(out) if self.collate else out
                out = []


class FakeTokenizer:
    pad_token_id, cls_token_id, sep_token_id, vocab_size = 0, 101, 102, 30522

    def __init__(self): self.added = {}
    def add_tokens(self, toks):
        for t in toks:
            self.added.setdefault(t, self.vocab_size + len(self.added))
        return len(toks)
    def convert_tokens_to_ids(self, t): return self.added.get(t, 100)
    def __len__(self): return self.vocab_size + len(self.added)
    def encode(self, text, add_special_tokens=False, max_length=None, truncation=False):
        ids = [1000 + (hash_str(w) % 29000) for w in text.replace(".", " ").split()]
        return ids[:max_length] if (truncation and max_length) else ids


def hash_str(s):
    h = 2166136261
    for ch in s.encode():
        h = ((h ^ ch) * 16777619) & 0xFFFFFFFF
    return h


class FakeLora(torch.nn.Module):
    """Stands in for peft.get_peft_model(base, LoraConfig(...)): the base encoder plus trainable parameters whose names contain
    'lora_' (train.py:151-155 collects them by name); the low-rank update is applied to the last hidden state."""

    def __init__(self, base, r=8):
        super().__init__()
        self.base_model = base
        self.config = base.config
        d = base.config.hidden_size
        self.lora_A = torch.nn.Parameter(torch.randn(d, r) * 0.01)
        self.lora_B = torch.nn.Parameter(torch.zeros(r, d))
        for p in base.parameters():
            p.requires_grad = False

    def forward(self, **kw):
        out = self.base_model(**kw)
        h = out.last_hidden_state
        return types.SimpleNamespace(last_hidden_state=h + (h @ self.lora_A.to(h.dtype)) @ self.lora_B.to(h.dtype))


def install_environment():
    # transformers resolves its classes lazily and probes sys.modules while doing so: import what is needed BEFORE any stand-in
    # module exists
    import transformers
    from transformers import AutoModel, AutoTokenizer, BertConfig, BertModel, ViTConfig, ViTModel, get_linear_schedule_with_warmup  # noqa: F401
    AutoTokenizer.from_pretrained = staticmethod(lambda name, **k: FakeTokenizer())
    AutoModel.from_pretrained = staticmethod(lambda name, **k: BertModel(BertConfig(
        vocab_size=30522, hidden_size=768, num_hidden_layers=1, num_attention_heads=12, intermediate_size=512)))
    ViTModel.from_pretrained = classmethod(lambda cls, name, **k: ViTModel(ViTConfig(num_hidden_layers=2)))

    class _Plt(types.ModuleType):
        def __getattr__(self, k):
            return lambda *a, **kw: None
    _module("webdataset", WebDataset=FakeWebDataset, split_by_node=lambda x: x, split_by_worker=lambda x: x, TarWriter=None)
    _module("peft", get_peft_model=lambda m, cfg: FakeLora(m, getattr(cfg, "r", 8)), LoraConfig=lambda **k: types.SimpleNamespace(**k),
            TaskType=types.SimpleNamespace(FEATURE_EXTRACTION="FEATURE_EXTRACTION"))
    tok = _module("nltk.tokenize", sent_tokenize=lambda t: [s.strip() + "." for s in t.split(".") if s.strip()])
    _module("nltk", tokenize=tok, download=lambda *a, **k: True)
    if importlib.util.find_spec("matplotlib") is None:
        plt = _Plt("matplotlib.pyplot")
        plt.__spec__ = importlib.machinery.ModuleSpec("matplotlib.pyplot", None)
        sys.modules["matplotlib.pyplot"] = plt
        _module("matplotlib", use=lambda *a, **k: None, pyplot=plt)
    os.environ.setdefault("RANK", "0"); os.environ.setdefault("WORLD_SIZE", "1"); os.environ.setdefault("LOCAL_RANK", "0")
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29577")
    return transformers


def load_script(name: str):
    path = find_script(name)
    # the drop-ins must win over a model.py that sits next to the script
    import model  # noqa: F401  (repo root)
    import model_HoME  # noqa: F401
    loader = importlib.machinery.SourceFileLoader("_ref_script_" + name, path)
    spec = importlib.util.spec_from_loader(loader.name, loader)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[loader.name] = mod
    if name == "infer_auc_HoME":
        if "train_HoME" not in sys.modules:          # `from train_HoME import HomeExpertWrapper` (infer_auc_HoME:139)
            sys.modules["train_HoME"] = load_script("train_HoME")[0]
    loader.exec_module(mod)
    return mod, path


def make_checkpoint(path: str, device):
    """A checkpoint with the keys train.py:351-363 writes, from freshly built drop-in modules (random init)."""
    import model as M
    tok = FakeTokenizer(); tok.add_tokens(["<SENT>"])
    mods = {"user": M.build_text_user_expert("x", 8, 384, tok, device), "item": M.build_text_item_expert("x", 8, 384, tok, device),
            "img": M.build_img_expert("x", pool_type="mean", device=device), "cross_ui": M.build_cross_expert(device=device),
            "concat_ui": M.build_concat_ui_expert(device=device), "concat_ti": M.build_concat_ti_expert(device=device),
            "head": M.TwoTaskMMoE().to(device)}
    torch.save({k: m.state_dict() for k, m in mods.items()} | {"epoch": 0}, path)


def make_home_checkpoint(path: str, device, wrapper_cls):
    """A checkpoint with the keys train_HoME.py:432-451 writes (modules + the six BN wrappers), random init."""
    import model_HoME as M
    tok = FakeTokenizer(); tok.add_tokens(["<SENT>"])
    mods = {"user": M.build_text_user_expert("x", 8, 384, tok, device), "item": M.build_text_item_expert("x", 8, 384, tok, device),
            "img": M.build_img_expert("x", device=device), "cross_ui": M.build_cross_expert(device=device),
            "concat_ui": M.build_concat_ui_expert(device=device), "concat_ti": M.build_concat_ti_expert(device=device),
            "head": M.HOME_MMoE_Complete(expert_dim=768, n_shared_experts=4, n_task_experts=2, tower_hidden=512).to(device)}
    for k in ("u_doc_wrapper", "i_doc_wrapper", "img_vec_wrapper", "ui_vec_wrapper", "xui_wrapper", "xti_wrapper"):
        mods[k] = wrapper_cls(d_model=768).to(device)
    torch.save({k: m.state_dict() for k, m in mods.items()} | {"epoch": 0}, path)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("script", choices=["train", "train_HoME", "inference_and_auc", "infer_auc_HoME"])
    ap.add_argument("--batches", type=int, default=4)
    ap.add_argument("--batch-size", type=int, default=8)
    ap.add_argument("--grad-accum", type=int, default=2)
    a = ap.parse_args()
    install_environment()
    work = tempfile.mkdtemp(prefix="mmoe_ref_run_")
    open(os.path.join(work, "shard-000000.tar"), "wb").close()
    FakeWebDataset.limit_batches = a.batches
    FakeWebDataset.endless = a.script in ("train", "train_HoME")
    mod, path = load_script(a.script)
    argv = [path, "--data_pattern", os.path.join(work, "shard-*.tar"), "--batch_size", str(a.batch_size), "--num_workers", "0",
            "--output_dir", os.path.join(work, "out")]
    if a.script == "inference_and_auc":
        ckpt = os.path.join(work, "ckpt.pt")
        make_checkpoint(ckpt, torch.device("cuda"))
        argv += ["--checkpoint_path", ckpt]
    elif a.script == "infer_auc_HoME":
        ckpt = os.path.join(work, "ckpt_home.pt")
        make_home_checkpoint(ckpt, torch.device("cuda"), sys.modules["train_HoME"].HomeExpertWrapper)
        argv += ["--ckpt", ckpt]
    else:
        argv += ["--grad_accum", str(a.grad_accum), "--epochs", "1"]
    sys.argv = argv
    import mmoe_multimodal_rec_b200 as pkg
    L = pkg.lib()
    L.mmoe_launch_count(1)
    try:
        mod.main()
        finished = "script returned"
    except HarnessDone:
        finished = f"stopped by the harness after {a.batches} batches"
    torch.cuda.synchronize()
    n = int(L.mmoe_launch_count(0))
    print(f"[run_reference_script] {os.path.basename(path)} (unchanged, sha1 {sha1(path)}) {finished}; native kernel launches: {n}")
    assert n > 0, "the script ran without a single native launch"
    import torch.distributed as dist
    if dist.is_initialized():
        dist.destroy_process_group()


def sha1(path):
    import hashlib
    return hashlib.sha1(open(path, "rb").read()).hexdigest()[:12]


if __name__ == "__main__":
    main()
