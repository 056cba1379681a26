"""Time (and, under ncu, profile) the non-GEMM kernels of one encoder layer at the bench shape
(B = 512 samples x 64 sentence slots, d = 768, 8 heads):  python tools/prof_misc.py [reps]

  attention forward / backward on the packed QKV projection [B, S, 3*768]
  LayerNorm backward with residual gradient, dropped 16-bit copy and bias-gradient column sums
Prints the achieved algorithmic GB/s next to the measured HBM peak.
"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import mmoe_multimodal_rec_b200 as pkg  # noqa: E402
from mmoe_multimodal_rec_b200._lib import check  # noqa: E402

L = pkg.lib()
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
dev = "cuda"
st = torch.cuda.current_stream().cuda_stream
try:
    HBM = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6560.0))
except Exception:
    HBM = 6560.0
B, S, H, HD, D = 512, 64, 8, 96, 768


def timed(name, fn, nbytes):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    gbs = nbytes / ms / 1e6
    print(f"{name:28s} {ms * 1e3:8.1f} us  {gbs:7.0f} GB/s algorithmic  ({100 * gbs / HBM:4.1f}% of {HBM:.0f})", flush=True)


qkv = torch.randn(B, S, 3 * D, device=dev).bfloat16()
lens = torch.randint(1, S + 1, (B,), device=dev)
mask = (torch.arange(S, device=dev)[None] >= lens[:, None]).to(torch.uint8).contiguous()
ctx = torch.empty(B, S, D, device=dev, dtype=torch.bfloat16)
dctx = torch.randn(B, S, D, device=dev).bfloat16()
dqkv = torch.empty_like(qkv)
bg = torch.zeros(3 * D, device=dev)
es = 2
q, k, v = qkv.data_ptr(), qkv.data_ptr() + D * es, qkv.data_ptr() + 2 * D * es
dq, dk, dv = dqkv.data_ptr(), dqkv.data_ptr() + D * es, dqkv.data_ptr() + 2 * D * es


def attn_fwd():
    check(L.mmoe_attention_fwd(q, 3 * D, k, 3 * D, v, 3 * D, mask.data_ptr(), ctx.data_ptr(), D, B, S, S, H, HD,
                               0.1, 11, 22, 1, st), "attention_fwd")


def attn_bwd():
    check(L.mmoe_attention_bwd(q, 3 * D, k, 3 * D, v, 3 * D, mask.data_ptr(), dctx.data_ptr(), D, dq, dk, dv,
                               bg.data_ptr(), bg.data_ptr() + D * 4, bg.data_ptr() + 2 * D * 4, B, S, S, H, HD,
                               0.1, 11, 22, 1, st), "attention_bwd")


rows = B * S
timed("attention forward", attn_fwd, rows * (3 * D + D) * es)
timed("attention backward", attn_bwd, rows * (3 * D + D + 3 * D) * es)
