"""CPU-side checks of the C-ABI boundary: the shared object builds for sm_100a, loads, exports every symbol that
include/mmoe_b200.h declares, the ctypes struct mirrors have the compiled sizes, and host-side argument checks fail loudly."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "mmoe_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mmoe_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import mmoe_multimodal_rec_b200 as pkg
    L = pkg.lib()
    names = _declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/mmoe_b200.h but not exported by libmmoe_b200.so"
    from mmoe_multimodal_rec_b200 import _lib
    assert set(names) == set(_lib.PROTOTYPES), set(names) ^ set(_lib.PROTOTYPES)


def test_struct_mirrors_match_compiled_sizes():
    import mmoe_multimodal_rec_b200 as pkg
    from mmoe_multimodal_rec_b200 import _lib
    L = pkg.lib()
    assert L.mmoe_abi_version() == 2
    for i, st in enumerate(_lib.ABI_STRUCTS):
        assert L.mmoe_abi_sizeof(i) == C.sizeof(st), st.__name__


def test_size_queries_and_argument_validation_without_a_gpu():
    import mmoe_multimodal_rec_b200 as pkg
    from mmoe_multimodal_rec_b200 import _lib
    L = pkg.lib()
    cfg = _lib.CrossCfg(768, 64, 8, 2)
    b16 = L.mmoe_cross_saved_bytes(C.byref(cfg), 512, _lib.BF16)
    f32 = L.mmoe_cross_saved_bytes(C.byref(cfg), 512, _lib.F32)
    assert 2e9 < b16 < f32 < 8e9                       # a few GB of activations at B=512
    assert L.mmoe_cross_saved_bytes(C.byref(cfg), 1024, _lib.BF16) > 1.9 * b16
    hc = _lib.HeadCfg(768, 6, 256, 0.0)
    assert L.mmoe_head_saved_bytes(C.byref(hc), 256, _lib.F32) > 256 * 768 * 4
    # invalid configuration: rejected on the host with a message, no device work
    bad = _lib.CrossCfg(768, 65, 8, 2)
    off, nb = C.c_size_t(), C.c_size_t()
    rc = L.mmoe_cross_saved_offset(C.byref(bad), 4, _lib.BF16, 0, 0, 0, 0, C.byref(off), C.byref(nb))
    assert rc != 0 and b"S must be" in L.mmoe_last_error()


def test_no_silent_cpu_path():
    """Without a CUDA tensor the drop-ins raise (there is no eager fallback to fall into)."""
    import torch
    import mmoe_multimodal_rec_b200 as pkg
    head = pkg.modules.TwoTaskMMoE()
    with pytest.raises(RuntimeError, match="CUDA"):
        head(torch.zeros(2, 6, 768))


def test_product_code_never_imports_the_oracle():
    pkg_dir = os.path.join(ROOT, "mmoe-multimodal-rec_b200")
    files = [os.path.join(pkg_dir, f) for f in os.listdir(pkg_dir) if f.endswith(".py")]
    files += [os.path.join(ROOT, f) for f in ("model.py", "model_HoME.py") if os.path.exists(os.path.join(ROOT, f))]
    for f in files:
        src = open(f).read()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
