"""TwoTaskMMoE alone, fwd+bwd (BASELINE configs[0] on the GPU): a few steps for an ncu launch list / event timing.
    python tools/prof_head.py [B] [mode: bf16|fp32] [steps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import mmoe_multimodal_rec_b200 as pkg  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
mode = sys.argv[2] if len(sys.argv) > 2 else "bf16"
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dev = torch.device("cuda", 0)
pkg.lib().mmoe_init()
head = pkg.modules.TwoTaskMMoE().to(dev).train()
ev = torch.randn(B, 6, 768, device=dev, requires_grad=True)


def step():
    head.zero_grad(set_to_none=True)
    ev.grad = None
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=(mode == "bf16")):
        lg, lb = head(ev)
    (lg.float().sum() + lb.float().sum()).backward()


for _ in range(2):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
print(f"head fwd+bwd B={B} {mode}: {ms:.3f} ms/step = {B / ms * 1e3 / 1e6:.2f} M samples/s = {55328.0 * B / (ms * 1e-3) / 1e9:.0f} GB/s algorithmic")
