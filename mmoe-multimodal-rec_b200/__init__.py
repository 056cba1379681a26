"""mmoe-multimodal-rec_b200 — B200-native (sm_100a) MMoE / HoME fusion-and-head path.

Layout: ``csrc/`` hand-written CUDA behind the C ABI of ``include/mmoe_b200.h``;
``_lib.py`` the ctypes binding; ``functional.py`` autograd wrappers; ``modules.py`` /
``modules_home.py`` the torch.nn.Module drop-ins mirroring the reference's model.py /
model_HoME.py; ``losses.py`` / ``home_wrap.py`` / ``ingest.py`` the opt-in native versions of the
script-side steps around the path.  Import as ``mmoe_multimodal_rec_b200`` (the repo-root shim of that name
loads this directory, whose name is not a Python identifier).
"""
from . import _lib, functional, modules, modules_home, losses, home_wrap, ingest  # noqa: F401
from ._lib import lib  # noqa: F401

__all__ = ["lib", "functional", "modules", "modules_home", "losses", "home_wrap", "ingest"]
