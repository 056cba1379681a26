"""Import the UNMODIFIED reference modules from /root/reference (dev container only).

TEST INFRASTRUCTURE ONLY.  ``/root/reference`` does not exist on the GPU box, so
nothing that runs there (``-m gpu`` tests, ``smoke()``, ``bench.py``) may call
this; it is used by ``oracle/make_golden.py`` and by the CPU tests that pin the
oracle and the drop-in's state-dict contract against the real thing (skipped
when the reference is absent).

The reference's ``model.py`` / ``model_HoME.py`` import ``webdataset``, ``peft``
and ``nltk`` at module top (model.py:7-18), none of which are installed here and
none of which the hot-path classes use; empty stand-ins are pre-seeded in
``sys.modules`` so the files import as they are.
"""
from __future__ import annotations

import importlib.machinery
import importlib.util
import os
import sys
import types

REFERENCE_DIR = os.environ.get("MMOE_REFERENCE_DIR", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_DIR, "model.py"))


def _stub(name: str, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    try:
        return importlib.import_module(name)
    except Exception:
        m = types.ModuleType(name)
        m.__spec__ = importlib.machinery.ModuleSpec(name, None)     # importlib.util.find_spec() on a stub must not raise
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m


def _install_stubs():
    _stub("webdataset")
    _stub("peft", get_peft_model=lambda m, c: m, LoraConfig=lambda **k: None,
          TaskType=types.SimpleNamespace(FEATURE_EXTRACTION="FEATURE_EXTRACTION"))
    nltk = _stub("nltk")
    tok = _stub("nltk.tokenize", sent_tokenize=lambda t: [s for s in t.split(". ") if s])
    if not hasattr(nltk, "tokenize"):
        nltk.tokenize = tok
    _stub("matplotlib")
    _stub("matplotlib.pyplot")


_STUB_NAMES = ("webdataset", "peft", "nltk", "nltk.tokenize", "matplotlib", "matplotlib.pyplot")


def load_reference(which: str = "model"):
    """Return the reference module object (``model``, ``model_HoME``, or a script such as ``train_HoME``) under a private
    name so it never shadows the drop-in ``model.py`` at the repo root.  The stand-ins for the missing third-party modules
    are visible in ``sys.modules`` only while the file is being imported (the reference module keeps its own references)."""
    if not reference_available():
        raise FileNotFoundError(f"reference not mounted at {REFERENCE_DIR}")
    alias = f"_reference_{which}"
    if alias in sys.modules:
        return sys.modules[alias]
    before = {n: sys.modules.get(n) for n in _STUB_NAMES + ("model", "model_HoME")}
    try:
        try:
            # transformers resolves these lazily and probes for `peft` while doing so: let it, before the stub exists
            from transformers import AutoModel, AutoTokenizer, ViTConfig, ViTModel, get_linear_schedule_with_warmup  # noqa: F401
        except Exception:
            pass
        _install_stubs()
        if which in ("train_HoME", "infer_auc_HoME"):            # scripts do `from model_HoME import ...`
            sys.modules["model_HoME"] = load_reference("model_HoME")
            _install_stubs()
        if which in ("train", "inference_and_auc"):
            sys.modules["model"] = load_reference("model")
            _install_stubs()
        mp = sys.modules["matplotlib"]
        if not hasattr(mp, "use"):
            mp.use = lambda *a, **k: None
        path = os.path.join(REFERENCE_DIR, which + ("" if which == "infer_auc_HoME" else ".py"))
        loader = importlib.machinery.SourceFileLoader(alias, path)
        spec = importlib.util.spec_from_loader(alias, loader)
        mod = importlib.util.module_from_spec(spec)
        sys.modules[alias] = mod
        spec.loader.exec_module(mod)
        return mod
    finally:
        for n, m in before.items():
            if m is None:
                sys.modules.pop(n, None)
            else:
                sys.modules[n] = m
