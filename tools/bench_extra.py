"""Secondary measurements for the other BASELINE.json configurations (the headline line is bench.py):

  head     TwoTaskMMoE fwd+bwd alone, fp32 and bf16, B = 256 ... 65536, against the HBM roofline
           (algorithmic bytes 55,328 B/sample fp32 I/O: read expert_vecs twice, write its gradient once)
  home     HoME fusion path (cross' + 2 fuse' + HOME_MMoE_Complete(768,4,2,512)) fwd+bwd bf16, B = 512   (configs[3], 1 GPU)
  infer    forward-only fp32 scoring of the v1 path, B = 1K ... 16K, torch.no_grad()                          (configs[4])

    python tools/bench_extra.py [head] [home] [infer]  > gpurun_out/extra.jsonl
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import mmoe_multimodal_rec_b200 as pkg  # noqa: E402

M, H = pkg.modules, pkg.modules_home
dev = torch.device("cuda", 0)
PEAKS = {}
try:
    PEAKS = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
except Exception:
    pass
HBM = float(PEAKS.get("hbm_gbs", 6650.0))


def timeit(fn, warm=3, reps=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def bench_head():
    head = M.TwoTaskMMoE().to(dev).train()
    modes = os.environ.get("HEAD_MODES", "fp32,bf16").split(",")
    sizes = [int(x) for x in os.environ.get("HEAD_BS", "256,4096,16384,65536").split(",")]
    for mode in modes:
        for B in sizes:
            ev = torch.randn(B, 6, 768, device=dev, requires_grad=True)

            def step():
                head.zero_grad(set_to_none=True)
                ev.grad = None
                with torch.autocast("cuda", dtype=torch.bfloat16, enabled=(mode == "bf16")):
                    lg, lb = head(ev)
                (lg.float().sum() + lb.float().sum()).backward()
            ms = timeit(step)
            gbs = 55328.0 * B / (ms * 1e-3) / 1e9
            print(json.dumps({"bench": "head_fwd_bwd", "mode": mode, "B": B, "ms": ms, "samples_per_s": B / ms * 1e3,
                              "algorithmic_GBps": gbs, "hbm_peak_GBps": HBM, "frac_of_hbm": gbs / HBM}), flush=True)


class _Pass(torch.nn.Module):
    class _C:
        hidden_size = 768
    config = _C()

    def forward(self, pixel_values):
        class O:
            pass
        o = O()
        o.last_hidden_state = pixel_values
        return o


def bench_home(B=512):
    cross, cui, cti = H.RobustTextCrossExpert().to(dev).train(), H.EnhancedCrossFuse().to(dev).train(), H.EnhancedCrossFuse().to(dev).train()
    head = H.HOME_MMoE_Complete(expert_dim=768, n_shared_experts=4, n_task_experts=2, tower_hidden=512).to(dev).train()
    img = H.ImageExpertWithProjection(_Pass()).to(dev).train()
    mods = [cross, cui, cti, head, img]
    u = torch.randn(B, 64, 768, device=dev, requires_grad=True)
    i = torch.randn(B, 64, 768, device=dev, requires_grad=True)
    lens = torch.randint(1, 65, (B,), device=dev)
    um = torch.arange(64, device=dev)[None] >= lens[:, None]
    im = torch.arange(64, device=dev)[None] >= lens.flip(0)[:, None]
    ud = torch.randn(B, 768, device=dev, requires_grad=True)
    idoc = torch.randn(B, 768, device=dev, requires_grad=True)
    tokens = torch.randn(B, 197, 768, device=dev)
    y = (torch.rand(B, device=dev) < 0.5).float()

    def step():
        for m in mods:
            m.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            img_vec, proj = img(tokens)
            ui = cross(u, um, i, im)
            xui, xti = cui(ud, img_vec), cti(idoc, img_vec)
            ev = torch.stack([ud, idoc, img_vec.float(), ui, xui, xti], 1)      # (HomeExpertWrapper BN stays script-side)
            lg, lb = head(ev)
            loss = torch.nn.functional.binary_cross_entropy_with_logits(lg.float(), y) + \
                torch.nn.functional.binary_cross_entropy_with_logits(lb.float(), y) + 0.01 * proj.float().pow(2).mean()
        loss.backward()
    ms = timeit(step, warm=3, reps=10)
    flop = 12.469e9 * B          # SURVEY §8d HoME fusion path fwd+bwd
    print(json.dumps({"bench": "home_fusion_path_fwd_bwd_bf16", "B": B, "ms": ms, "samples_per_s": B / ms * 1e3,
                      "path_TFLOPs": flop / (ms * 1e-3) / 1e12}), flush=True)


def bench_infer():
    cross, cui, cti, head = (M.RobustTextCrossExpert().to(dev).eval(), M.EnhancedCrossFuse().to(dev).eval(),
                             M.EnhancedCrossFuse().to(dev).eval(), M.TwoTaskMMoE().to(dev).eval())
    img = M.ItemImageExpert(_Pass()).to(dev).eval()
    for B in (1024, 4096, 16384):
        u = torch.randn(B, 64, 768, device=dev)
        i = torch.randn(B, 64, 768, device=dev)
        lens = torch.randint(1, 65, (B,), device=dev)
        um = torch.arange(64, device=dev)[None] >= lens[:, None]
        ud, idoc = torch.randn(B, 768, device=dev), torch.randn(B, 768, device=dev)
        tokens = torch.randn(B, 197, 768, device=dev)
        for mode in ("fp32", "bf16"):
            def step():
                with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=(mode == "bf16")):
                    img_vec = img(tokens)
                    ui = cross(u, um, i, um)
                    ev = torch.stack([ud, idoc, img_vec, ui, cui(ud, img_vec), cti(idoc, img_vec)], 1)
                    lg, lb = head(ev)
                return torch.sigmoid(lg), torch.sigmoid(lb)
            reps = 2 if mode == "fp32" else 5
            ms = timeit(step, warm=1, reps=reps)
            print(json.dumps({"bench": "v1_forward_scoring", "mode": mode, "B": B, "ms": ms, "samples_per_s": B / ms * 1e3,
                              "path_TFLOPs": 4.122e9 * B / (ms * 1e-3) / 1e12}), flush=True)


if __name__ == "__main__":
    which = sys.argv[1:] or ["head", "home", "infer"]
    pkg.lib().mmoe_init()
    if "head" in which:
        bench_head()
    if "home" in which:
        bench_home()
    if "infer" in which:
        bench_infer()
