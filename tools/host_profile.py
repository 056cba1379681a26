"""Where the HOST time of one bench step goes (9.95 ms of enqueue per 10.7 ms device step).
cProfile over N steps of bench.py's step (no side streams), foreign calls into libmmoe_b200.so show up as built-ins.

    python tools/host_profile.py [steps] > gpurun_out/host_profile.txt
"""
import cProfile
import io
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

import bench  # noqa: E402
import mmoe_multimodal_rec_b200 as pkg  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
B = int(os.environ.get("B", "512"))
dev = torch.device("cuda", 0)
pkg.lib().mmoe_init()
M = pkg.modules
torch.manual_seed(0)
img = M.ItemImageExpert(bench.Passthrough(), pool_type="mean").to(dev).train()
cross, cui, cti, head = (M.RobustTextCrossExpert().to(dev).train(), M.EnhancedCrossFuse().to(dev).train(),
                         M.EnhancedCrossFuse().to(dev).train(), M.TwoTaskMMoE().to(dev).train())
mods = [img, cross, cui, cti, head]
b = {k: v.to(dev) for k, v in bench.make_host_batch(B, 1, pin=False).items()}


def step():
    for m in mods:
        m.zero_grad(set_to_none=True)
    ins = {k: (v.detach().requires_grad_(True) if k in ("u_sent", "i_sent", "u_doc", "i_doc") else v) for k, v in b.items()}
    with torch.autocast("cuda", dtype=torch.bfloat16):
        img_vec = img(ins["img_tokens"], trainable=False)
        ui = cross(ins["u_sent"], ins["u_mask"], ins["i_sent"], ins["i_mask"])
        xui = cui(ins["u_doc"], img_vec)
        xti = cti(ins["i_doc"], img_vec)
        ev = torch.stack([ins["u_doc"], ins["i_doc"], img_vec, ui, xui, xti], dim=1)
        lg, lb = head(ev)
        loss = F.binary_cross_entropy_with_logits(lg.float(), ins["y_good"]) + F.binary_cross_entropy_with_logits(lb.float(), ins["y_best"])
    loss.backward()


for _ in range(5):
    step()
torch.cuda.synchronize()
# host-only cost: enqueue `steps` steps while the device is kept behind (sync first, time the enqueue loop)
t0 = time.perf_counter()
for _ in range(steps):
    step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"B={B}: enqueue {1e3 * (t1 - t0) / steps:.3f} ms/step, with drain {1e3 * (t2 - t0) / steps:.3f} ms/step, "
      f"launches/step {pkg.lib().mmoe_launch_count(1) / (steps + 5):.0f}")
pr = cProfile.Profile()
pr.enable()
for _ in range(steps):
    step()
pr.disable()
torch.cuda.synchronize()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(45)
print(s.getvalue())
