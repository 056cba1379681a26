"""Every native entry point once at small shapes — the program to put under `compute-sanitizer --tool memcheck`
(full-size tests are too slow under the sanitizer).  Exits non-zero on any Python-visible failure."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import mmoe_multimodal_rec_b200 as pkg  # noqa: E402
import parity_util as PU  # noqa: E402
from oracle import cases as C  # noqa: E402

pkg.lib().mmoe_init()
dev = torch.device("cuda")
torch.manual_seed(0)
# the drop-in modules: forward + backward, eval and train, fp32 / bf16 / fp16
for name in ("head_b16", "home_head_b8", "cross_b3", "cross_home_b3", "fuse_b8", "fuse_home_b8", "img_pool_mean_b4", "img_pool_cls_b4", "img_proj_b4"):
    case = C.CASES_BY_NAME[name]
    for mode in ("fp32", "bf16", "fp16"):
        for train in (False, True):
            mod = PU.build_module(case)
            mod.train(train)
            PU.run_cuda(case, mode, mod, seed=3)
print("modules ok")
# next rows
HW, LS, IG = pkg.home_wrap, pkg.losses, pkg.ingest
for train in (True, False):
    ws = [HW.HomeExpertWrapper(768).to(dev).train(train) for _ in range(6)]
    xs = [torch.randn(10, 768, device=dev, requires_grad=True) for _ in range(6)]
    out = HW.FusedHomeExpertStack(ws)(*xs)
    out.sum().backward()
lg, lb = torch.randn(37, device=dev, requires_grad=True), torch.randn(37, device=dev, requires_grad=True)
y = (torch.rand(37, device=dev) < 0.5).float()
LS.TwoTaskBCEWithLogits()(lg, lb, y, y).backward()
for mode in ("fp32", "bf16"):
    a, p, q = (torch.randn(24, 768, device=dev, requires_grad=True) for _ in range(3))
    with PU.autocast_ctx(mode):
        l = LS.info_nce_losses([(a, p), (q, p), (a, q)])
    l.sum().backward()
for n in (1, 5, 2049, 5000):
    LS.roc_auc(torch.randn(n, device=dev), (torch.rand(n, device=dev) < 0.4).float())
conv = torch.nn.Conv2d(3, 768, 16, 16).to(dev)
pe = IG.NativePatchEmbeddings(conv)
raw = torch.randint(0, 256, (3, 196, 768), dtype=torch.uint8, device=dev)
for mode in ("fp32", "bf16"):
    with PU.autocast_ctx(mode):
        pe(raw); pe(IG.decode_patch_bytes(raw))
h = torch.randn(5, 40, 768, device=dev, requires_grad=True)
norm = torch.nn.LayerNorm(768).to(dev)
c2s, pos = [0, 0, 2, 3, 3], [[1, 9, -1], [2, 39, 45], [1, 2, 3], [5, -1, -1], [7, 8, -1]]
for nm in (norm, None):
    s, m, dvec = IG.sentence_gather(h, c2s, pos, 8, nm, 0.1, True)
    (s.sum() + dvec.sum()).backward()
torch.cuda.synchronize()
print("next rows ok")
