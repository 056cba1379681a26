#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native MMoE fusion-and-head path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1]; SURVEY.md §8d "v1 fusion path"): one fwd+bwd pass of the v1 fusion-and-head
path of the reference's train.py micro-step (train.py:244-254) under bf16 autocast on a batch of B=512 samples
per GPU: ItemImageExpert tail (token mean + LN + dropout) -> RobustTextCrossExpert -> 2x EnhancedCrossFuse ->
stack -> TwoTaskMMoE -> 2x BCEWithLogits -> backward.  Modules are in train() mode (dropout 0.1 active, as in
training).  Synthetic inputs of the shapes the encoders produce (sentence vectors [B,64,768] + masks, doc
vectors, ViT tokens [B,197,768]) and random-init weights; the text encoders / ViT backbone are the reference's
own torch modules and are not part of the timed path (north_star: "timed separately").

N > 1 (weak scaling, fixed per-GPU batch): the headline is measured the way the reference's train.py runs the modules
— one DistributedDataParallel wrapper per module with default arguments (train.py:136-139), the head called through
`.module` (train.py:251) — so the number is what the UNCHANGED script gets from the drop-ins.  The package's own
flat-buffer gradient exchange (functional.enable_grad_allreduce) is timed next to it and reported under
`native_exchange`.

The headline pass carries no per-launch events; the GEMM roofline is measured in a second, single-stream pass of the
same steps with a CUDA event pair around every GEMM launch.  At N = 1 the line also carries: `eager_b200` (the same
step in eager PyTorch — torch's own cuBLASLt/SDPA kernels — on the same GPU and weights), `cpu_baseline` (the same
step on the host cores), and `extra` (TwoTaskMMoE alone against the HBM roofline, the HoME path, the forward-only
scoring sweep: BASELINE configs[0], [3], [4]).

One JSON line on stdout (rank 0); see README / DESIGN.md for the keys.
"""
from __future__ import annotations

import argparse
import ctypes as C
import gc
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "mmoe_fusion_head_fwd_bwd_samples_per_sec"
UNIT = "samples/s"
S, D, NTOK = 64, 768, 197
# algorithmic GEMM FLOPs of the v1 fusion path per sample, fwd+bwd (SURVEY.md §8d / BASELINE.md §3)
FLOP_PER_SAMPLE = 12.366e9
FLOP_PER_SAMPLE_HOME = 12.469e9
HEAD_BYTES_PER_SAMPLE = 55328.0          # TwoTaskMMoE fwd+bwd, fp32 I/O
POS_W_GOOD, POS_W_BEST = 858627.0 / 990303.0, 1328721.0 / 520209.0      # train.py:189-192


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ----------------------------------------------------------------------------------------------
# synthetic inputs (oracle/synth.py is test infrastructure; the bench makes its own with torch RNG)
# ----------------------------------------------------------------------------------------------
def make_host_batch(B: int, seed: int, pin: bool, lowp_stream: bool = False):
    """lowp_stream: the three big activation tensors as bf16 on the host (the end-to-end leg ships them in 16 bits; the
    modules cast on the device — under autocast their first use is a 16-bit GEMM operand anyway)."""
    g = torch.Generator().manual_seed(seed)
    def t(*shape, lowp=False):
        x = torch.randn(*shape, generator=g)
        if lowp:
            x = x.to(torch.bfloat16)
        return x.pin_memory() if pin else x
    lens_u = torch.randint(1, S + 1, (B,), generator=g)
    lens_i = torch.randint(1, S + 1, (B,), generator=g)
    ar = torch.arange(S)[None]
    batch = {
        "u_sent": t(B, S, D, lowp=lowp_stream), "i_sent": t(B, S, D, lowp=lowp_stream),
        "u_mask": (ar >= lens_u[:, None]), "i_mask": (ar >= lens_i[:, None]),
        "u_doc": t(B, D), "i_doc": t(B, D), "img_tokens": t(B, NTOK, D, lowp=lowp_stream),
        "y_good": (torch.rand(B, generator=g) < 0.5).float(), "y_best": (torch.rand(B, generator=g) < 0.5).float(),
    }
    if pin:
        for k in ("u_mask", "i_mask", "y_good", "y_best"):
            batch[k] = batch[k].pin_memory()
    return batch


def h2d_bytes(batch) -> int:
    return sum(v.numel() * v.element_size() for v in batch.values())


class Passthrough(torch.nn.Module):
    """Stands in for the HF ViT backbone (excluded from the timed path): hands the given tokens on."""
    class _Cfg:
        hidden_size = D
    config = _Cfg()

    def forward(self, pixel_values):
        class _O:
            pass
        o = _O()
        o.last_hidden_state = pixel_values
        return o


# ----------------------------------------------------------------------------------------------
# clocks sampler (NVML during the timed region)
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region through NVML (in-process thread, 20 ms period;
    `nvidia-smi -lms` as a subprocess measurably disturbs a 200 ms timed region, NVML calls do not)."""

    def __init__(self, index: int):
        self.index = index
        self.samples = []
        self._stop = threading.Event()
        self._thr = None
        self._h = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
        except Exception:
            self._h = None

    def _loop(self):
        nv, h = self._nv, self._h
        reasons = nv.nvmlDeviceGetCurrentClocksEventReasons if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
            else nv.nvmlDeviceGetCurrentClocksThrottleReasons
        try:
            mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)      # a constant: asked once
        except Exception:
            mx = 0
        while not self._stop.is_set():
            try:
                self.samples.append((nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM), mx, reasons(h)))
            except Exception:
                pass
            self._stop.wait(0.025)

    def start(self):
        if self._h is None:
            return
        self._thr = threading.Thread(target=self._loop, daemon=True)
        self._thr.start()

    def stop(self):
        self._stop.set()
        if self._thr is not None:
            self._thr.join(timeout=1.0)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        nv = self._nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        sm = sorted(s[0] for s in self.samples)
        reasons = sorted(n for n, bit in names.items() if any(s[2] & bit for s in self.samples))
        return {"sm_mhz": float(sm[len(sm) // 2]), "sm_max_mhz": float(max(s[1] for s in self.samples)), "reasons": reasons,
                "samples": len(sm)}


# ----------------------------------------------------------------------------------------------
# the reference's composition on stock torch modules (oracle/eager_ref.py): CPU arm and eager-on-B200 arm
# ----------------------------------------------------------------------------------------------
def build_modules(dev, train=True, seed=1234, fused=None):
    """fused: None = the package default (MMOE_FLAT_PARAMS), True/False = fused-parameter / per-tensor modules."""
    import mmoe_multimodal_rec_b200 as pkg
    M = pkg.modules
    torch.manual_seed(seed)
    was = pkg.functional.FLAT_PARAMS
    if fused is not None:
        pkg.functional.set_flat_parameters(fused)
    try:
        mods = {"img": M.ItemImageExpert(Passthrough(), pool_type="mean"), "cross": M.RobustTextCrossExpert(),
                "concat_ui": M.EnhancedCrossFuse(), "concat_ti": M.EnhancedCrossFuse(), "head": M.TwoTaskMMoE()}
    finally:
        pkg.functional.set_flat_parameters(was)
    for m in mods.values():
        m.to(dev).train(train)
    return mods


def eager_step_fn(mods, b, dev, autocast_dtype):
    """fwd+bwd of the same micro-step in eager PyTorch: the modules' own nn.TransformerEncoderLayer / nn.MultiheadAttention /
    nn.Linear containers CALLED the way the reference's forward bodies call them (oracle/eager_ref.py), train mode."""
    from oracle import eager_ref as E
    pw_g, pw_b = torch.tensor(POS_W_GOOD, device=dev), torch.tensor(POS_W_BEST, device=dev)
    grad_keys = ("u_sent", "i_sent", "u_doc", "i_doc")

    def step():
        for m in mods.values():
            m.zero_grad(set_to_none=True)
        ins = {k: (v.detach().requires_grad_(True) if k in grad_keys else v) for k, v in b.items()}
        loss = E.v1_fusion_step(mods, ins, pw_g, pw_b, autocast_dtype)
        loss.backward()
        return loss
    return step


def time_cpu_reference(B: int, steps: int, warmup: int, budget_s: float):
    """The reference's CPU path for this workload: its forward composition on stock torch modules (fp32, train mode,
    torch's own dropout), all host threads, B samples per step.  Timed steps are cut short when `budget_s` runs out."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    dev = torch.device("cpu")
    from oracle import eager_ref as E
    mods = E.make_v1_modules(dev, train=True)            # stock torch.nn containers only: nothing of the package on this arm
    b = make_host_batch(B, 99, pin=False)
    step = eager_step_fn(mods, b, dev, None)
    t_begin = time.perf_counter()
    for _ in range(warmup):
        step()
        if time.perf_counter() - t_begin > budget_s / 3:
            break
    done = 0
    t0 = time.perf_counter()
    while done < steps:
        step()
        done += 1
        if time.perf_counter() - t_begin > budget_s and done >= 1:
            break
    dt = (time.perf_counter() - t0) / done
    return B / dt, dt * 1e3, cores, done


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    B = args.ref_batch
    value, ms, cores, done = time_cpu_reference(B, max(args.steps, 1), max(min(args.warmup, 2), 1), args.ref_budget_s)
    sample = (f"reference forward composition on stock torch.nn modules (oracle/eager_ref.py), fp32, train mode, {B}-sample batch per step, "
              f"fwd+bwd, {done} timed steps of the {args.steps} requested ({ms:.0f} ms/step; bounded to ~{args.ref_budget_s:.0f} s)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": done,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.batch, args.gpus) | {"sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(B, n_gpus):
    return {"workload": "v1 MMoE fusion-and-head path fwd+bwd (ItemImageExpert tail + RobustTextCrossExpert + 2x EnhancedCrossFuse "
                        "+ TwoTaskMMoE + BCE), train mode, BASELINE configs[1]",
            "per_gpu_batch": B, "global_batch": B * n_gpus, "sentences": S, "vector_dim": D, "parallelism": f"dp{n_gpus}",
            "autocast": "bf16", "l2": "inputs (~0.5 GB/step) and activations (~3 GB) exceed the 126 MB L2; no flush needed",
            "encoders": "text encoders / ViT backbone excluded (reference torch modules, timed separately)",
            "streams": "the two EnhancedCrossFuse experts run on side streams next to the cross expert"}


# ----------------------------------------------------------------------------------------------
def cuda_time(fn, warm, reps):
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


def extra_measurements(dev, hbm_gbs, peak_tf):
    """BASELINE configs[0] (head alone, HBM roofline), [3] (HoME path, one GPU) and [4] (forward-only scoring sweep)."""
    import torch.nn.functional as F
    import mmoe_multimodal_rec_b200 as pkg
    M, H = pkg.modules, pkg.modules_home
    out = {}
    # ---- TwoTaskMMoE alone, fwd+bwd: bandwidth-bound; 55,328 algorithmic bytes per sample (read expert_vecs twice, write its grad)
    head = M.TwoTaskMMoE().to(dev).train()
    rows = []
    for mode, B in (("bf16", 65536), ("bf16", 256), ("fp32", 65536), ("fp32", 256)):
        ev = torch.randn(B, 6, D, device=dev, requires_grad=True)

        def step():
            head.zero_grad(set_to_none=True)
            ev.grad = None
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=(mode == "bf16")):
                lg, lb = head(ev)
            (lg.float().sum() + lb.float().sum()).backward()
        ms = cuda_time(step, 3, 10)
        gbs = HEAD_BYTES_PER_SAMPLE * B / (ms * 1e-3) / 1e9
        rows.append({"mode": mode, "B": B, "ms": ms, "samples_per_s": B / ms * 1e3, "algorithmic_GBps": gbs, "frac_of_hbm_peak": gbs / hbm_gbs})
        del ev
    out["head_only_fwd_bwd"] = {"bound": "hbm", "peak_GBps": hbm_gbs, "bytes_per_sample": HEAD_BYTES_PER_SAMPLE, "rows": rows}
    # ---- HoME micro-step (train_HoME.py:347-372): cross' + 2 fuse' + image projection head + 6 x HomeExpertWrapper + stack +
    # HOME_MMoE_Complete(768,4,2,512) + two-task BCE + three InfoNCE terms, fwd+bwd bf16, B = 512.  Timed twice: with the
    # script-side steps in torch (what the unchanged script does) and with the opt-in native versions (home_wrap / losses).
    B = 512
    cross, cui, cti = H.RobustTextCrossExpert().to(dev).train(), H.EnhancedCrossFuse().to(dev).train(), H.EnhancedCrossFuse().to(dev).train()
    hhead = H.HOME_MMoE_Complete(expert_dim=D, n_shared_experts=4, n_task_experts=2, tower_hidden=512).to(dev).train()
    img = H.ImageExpertWithProjection(Passthrough()).to(dev).train()
    wraps = [pkg.home_wrap.HomeExpertWrapper(D).to(dev).train() for _ in range(6)]
    fused_wraps = pkg.home_wrap.FusedHomeExpertStack(wraps)
    bce = pkg.losses.TwoTaskBCEWithLogits()
    hm = [cross, cui, cti, hhead, img] + wraps
    b = {k: v.to(dev) for k, v in make_host_batch(B, 5, pin=False).items()}
    u, i, ud, idoc = (b[k].requires_grad_(True) for k in ("u_sent", "i_sent", "u_doc", "i_doc"))
    pw_g, pw_b = torch.tensor(POS_W_GOOD, device=dev), torch.tensor(POS_W_BEST, device=dev)

    def torch_wrapper(w, x):                     # HomeExpertWrapper.forward, train_HoME.py:108-116
        return w.dropout(F.silu(w.norm(x)))

    def torch_nce(a, p, t=0.07):                 # calculate_contrastive_loss, train_HoME.py:43-51
        sim = F.normalize(a, p=2, dim=1) @ F.normalize(p, p=2, dim=1).t() / t
        return F.cross_entropy(sim, torch.arange(sim.size(0), device=sim.device))

    def home_step(native_tail):
        for m in hm:
            m.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            img_vec, proj = img(b["img_tokens"])
            ui = cross(u, b["u_mask"], i, b["i_mask"])
            xui, xti = cui(ud, img_vec), cti(idoc, img_vec)
            six = [ud, idoc, img_vec.float(), ui, xui, xti]
            if native_tail:
                ev = fused_wraps(*six)
                lg, lb = hhead(ev)
                main = bce(lg, lb, b["y_good"], b["y_best"])
                cl = pkg.losses.info_nce_losses([(ui, idoc), (ud, proj), (idoc, proj)]).sum()
            else:
                ev = torch.stack([torch_wrapper(w, x) for w, x in zip(wraps, six)], dim=1)
                lg, lb = hhead(ev)
                main = F.binary_cross_entropy_with_logits(lg.float(), b["y_good"], pos_weight=pw_g) + \
                    F.binary_cross_entropy_with_logits(lb.float(), b["y_best"], pos_weight=pw_b)
                cl = torch_nce(ui, idoc) + torch_nce(ud, proj) + torch_nce(idoc, proj)
            loss = main + 0.1 * cl
        loss.backward()
    res = {}
    for name, flag in (("script_side_torch", False), ("native_tail", True)):
        ms = cuda_time(lambda: home_step(flag), 3, 10)
        tf = FLOP_PER_SAMPLE_HOME * B / (ms * 1e-3) / 1e12
        res[name] = {"ms": ms, "samples_per_s": B / ms * 1e3, "path_tflops": tf, "path_frac_of_peak": tf / peak_tf}
    out["home_step_fwd_bwd_bf16"] = {"B": B, "note": "single stream; train_HoME.py:347-372 on one GPU (BASELINE configs[3] per-GPU work)", **res}
    del hm, cross, cui, cti, hhead, img, wraps
    # ---- forward-only scoring (inference_and_auc.py:130-156), v1 path, no grad
    mods = build_modules(dev, train=False)
    rows = []
    for Bs in (1024, 4096, 16384, 65536):
        bb = {k: v.to(dev) for k, v in make_host_batch(min(Bs, 4096), 6, pin=False).items()}
        if Bs > 4096:
            bb = {k: v.repeat((Bs // 4096,) + (1,) * (v.dim() - 1)) for k, v in bb.items()}
        for mode in ("bf16", "fp32"):
            if mode == "fp32" and Bs > 16384:
                continue

            def score():
                with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=(mode == "bf16")):
                    img_vec = mods["img"](bb["img_tokens"])
                    ui = mods["cross"](bb["u_sent"], bb["u_mask"], bb["i_sent"], bb["i_mask"])
                    ev = torch.stack([bb["u_doc"], bb["i_doc"], img_vec, ui, mods["concat_ui"](bb["u_doc"], img_vec),
                                      mods["concat_ti"](bb["i_doc"], img_vec)], 1)
                    lg, lb = mods["head"](ev)
                return torch.sigmoid(lg), torch.sigmoid(lb)
            ms = cuda_time(score, 1, 2 if mode == "fp32" else 4)
            rows.append({"mode": mode, "B": Bs, "ms": ms, "samples_per_s": Bs / ms * 1e3, "path_tflops": 4.122e9 * Bs / (ms * 1e-3) / 1e12})
        del bb
    out["v1_forward_scoring"] = {"rows": rows}
    return out


def home_ddp_measurement(dev, local_rank, world, steps):
    """BASELINE configs[3]: the HoME micro-step of train_HoME.py:347-404 on N GPUs, wrapped the way that script wraps it —
    DistributedDataParallel(find_unused_parameters=True) around every module incl. the six HomeExpertWrapper instances
    (train_HoME.py:190-202) — per-GPU batch 512, bf16 autocast, train mode.  Returns whole-job samples/s (max over ranks)."""
    import torch.distributed as dist
    import torch.nn.functional as F
    from torch.nn.parallel import DistributedDataParallel as DDP
    import mmoe_multimodal_rec_b200 as pkg
    H = pkg.modules_home
    B = 512
    torch.manual_seed(4321)
    raw = {"cross": H.RobustTextCrossExpert(), "cui": H.EnhancedCrossFuse(), "cti": H.EnhancedCrossFuse(),
           "head": H.HOME_MMoE_Complete(expert_dim=D, n_shared_experts=4, n_task_experts=2, tower_hidden=512),
           "img": H.ImageExpertWithProjection(Passthrough())}
    class ScriptWrapper(torch.nn.Module):        # HomeExpertWrapper as train_HoME.py:100-116 defines it (script-side torch)
        def __init__(self, d, p=0.1):
            super().__init__()
            self.norm, self.dropout = torch.nn.BatchNorm1d(d), torch.nn.Dropout(p)

        def forward(self, x):
            return self.dropout(F.silu(self.norm(x)))
    for i in range(6):
        raw[f"w{i}"] = ScriptWrapper(D)
    mods = {k: DDP(m.to(dev).train(), device_ids=[local_rank], find_unused_parameters=True) for k, m in raw.items()}
    params = [p for m in raw.values() for p in m.parameters()]
    b = {k: v.to(dev) for k, v in make_host_batch(B, 4242 + local_rank, pin=False).items()}
    u, i_, ud, idoc = (b[k].requires_grad_(True) for k in ("u_sent", "i_sent", "u_doc", "i_doc"))
    pw_g, pw_b = torch.tensor(POS_W_GOOD, device=dev), torch.tensor(POS_W_BEST, device=dev)

    def nce(a, p, t=0.07):                       # calculate_contrastive_loss, train_HoME.py:43-51
        sim = F.normalize(a, p=2, dim=1) @ F.normalize(p, p=2, dim=1).t() / t
        return F.cross_entropy(sim, torch.arange(sim.size(0), device=sim.device))

    def step():
        for p in params:
            p.grad = None
        with torch.autocast("cuda", dtype=torch.bfloat16):
            img_vec, proj = mods["img"](b["img_tokens"])
            ui = mods["cross"](u, b["u_mask"], i_, b["i_mask"])
            xui, xti = mods["cui"](ud, img_vec), mods["cti"](idoc, img_vec)
            six = [ud, idoc, img_vec.float(), ui, xui, xti]
            ev = torch.stack([mods[f"w{k}"](x) for k, x in enumerate(six)], dim=1)           # train_HoME.py:350-356
            lg, lb = mods["head"](ev)
            loss = F.binary_cross_entropy_with_logits(lg.float(), b["y_good"], pos_weight=pw_g) + \
                F.binary_cross_entropy_with_logits(lb.float(), b["y_best"], pos_weight=pw_b) + \
                0.1 * (nce(ui, idoc) + nce(ud, proj) + nce(idoc, proj))
        loss.backward()
    for _ in range(4):
        step()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    dist.barrier(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / steps], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    out = {"B_per_gpu": B, "ms_per_step": ms, "samples_per_s": B * world / (ms * 1e-3), "steps": steps,
           "wrapping": "13 DistributedDataParallel(find_unused_parameters=True) wrappers, script-side BN wrappers / InfoNCE in torch"}
    trace = os.environ.get("BENCH_HOME_TRACE")
    if trace:                                     # device timeline of 3 more steps (tools/trace_exchange.py's reduction)
        from torch.profiler import profile, ProfilerActivity
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(3):
                step()
            dist.barrier(); torch.cuda.synchronize()
        if dist.get_rank() == 0:
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import trace_exchange
            prof.export_chrome_trace(trace)
            r = trace_exchange.analyse(trace, 3)
            fams = sorted(r.pop("families").items(), key=lambda kv: -kv[1]["ms_per_step"])[:12]
            r.pop("streams", None)
            out["timeline"] = r | {"top_kernel_families": {k: v for k, v in fams}}
    return out


def small_batch_measurement(dev, step_factory, B_small, ms_per_sample_ref):
    """The README configuration (2 GPUs x batch 128, README.md:253-263): the same step at B = 128 per GPU.  At this size the
    step is launch-latency bound (≈180 launches), so its per-sample cost against the B = 512 line shows what launch overhead
    costs."""
    step = step_factory(B_small)
    cuda_time(step, 5, 20)          # rehearsal: new shapes — let the caching allocator reach its steady state first
    ms = cuda_time(step, 2, 20)
    return {"B": B_small, "ms_per_step": ms, "samples_per_s": B_small / ms * 1e3,
            "per_sample_cost_vs_b512": (ms / B_small) / ms_per_sample_ref}


# ----------------------------------------------------------------------------------------------
_REAL_STDOUT = None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--batch", type=int, default=512, help="per-GPU batch (train.py default 512)")
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--ref-batch", type=int, default=512, help="samples per step of the CPU reference arm")
    ap.add_argument("--ref-budget-s", type=float, default=150.0, help="wall-clock bound of the CPU reference arm")
    ap.add_argument("--cpu-baseline-batch", type=int, default=512)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip eager_b200 / extra configs (N = 1 only)")
    ap.add_argument("--eval-mode", action="store_true", help="dropout off (debug)")
    ap.add_argument("--no-side-stream", action="store_true", help="run both fuse experts on the main stream")
    ap.add_argument("--exchange", default="both", choices=["ddp", "native", "both"],
                    help="N > 1: which gradient exchange to time; the headline is always the DDP wrappers when measured")
    ap.add_argument("--ddp-view", action="store_true", help="gradient_as_bucket_view=True on the DDP wrappers (train.py passes nothing)")
    ap.add_argument("--params", default="default", choices=["default", "tensor", "fused"],
                    help="parameter layout of the modules: per-tensor nn.Parameters, one fused nn.Parameter per module "
                         "(MMOE_FLAT_PARAMS=1), or the package default")
    ap.add_argument("--nccl-ctas", type=int, default=0,
                    help="N>1: if > 0, cap NCCL at this many CTAs and keep as many SMs free of GEMM CTAs")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import torch.distributed as dist
    import torch.nn.functional as F
    import mmoe_multimodal_rec_b200 as pkg
    from mmoe_multimodal_rec_b200._lib import check

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    distributed = world > 1
    if distributed:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")     # keep stdout for the one JSON line
        # NCCL still prints its version banner on fd 1: point fd 1 at stderr for the life of the process and keep a
        # private duplicate of the real stdout for the JSON line
        global _REAL_STDOUT
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)
        if args.nccl_ctas > 0:
            os.environ.setdefault("NCCL_MAX_CTAS", str(args.nccl_ctas))
        dist.init_process_group("nccl", device_id=dev)
    L = pkg.lib()
    check(L.mmoe_init(), "init")
    if distributed and args.nccl_ctas > 0:
        L.mmoe_set_sm_reserve(args.nccl_ctas)

    fused = {"default": pkg.functional.FLAT_PARAMS, "tensor": False, "fused": True}[args.params]

    def make_modules(fused_mode):
        md = build_modules(dev, train=not args.eval_mode, fused=fused_mode)
        ms = [md[k] for k in ("img", "cross", "concat_ui", "concat_ti", "head")]
        if distributed:
            for m in ms:
                for prm in m.parameters():
                    dist.broadcast(prm.data, src=0)
        return ms

    mods = make_modules(fused)
    img, cross, cui, cti, head = mods
    call = {"img": img, "cross": cross, "cui": cui, "cti": cti, "head": head}
    params_note = ("one fused nn.Parameter per module (MMOE_FLAT_PARAMS=1 / functional.set_flat_parameters; state_dict keys unchanged)"
                   if fused else "one nn.Parameter per tensor, as in the reference")
    pw_g = torch.tensor(POS_W_GOOD, device=dev)
    pw_b = torch.tensor(POS_W_BEST, device=dev)

    B = args.batch
    n_host = 2
    host = [make_host_batch(B, 1234 + rank * 17 + j, pin=True, lowp_stream=True) for j in range(n_host)]
    resident = {k: v.to(dev) for k, v in make_host_batch(B, 1234 + rank * 17, pin=False).items()}
    grad_keys = ("u_sent", "i_sent", "u_doc", "i_doc")

    side_stream = None if args.no_side_stream else torch.cuda.Stream(device=dev)
    side2 = torch.cuda.Stream(device=dev) if side_stream is not None else None
    use_side = [side_stream is not None]
    if hasattr(torch.autograd.graph, "set_warn_on_accumulate_grad_stream_mismatch"):
        torch.autograd.graph.set_warn_on_accumulate_grad_stream_mismatch(False)     # intentional: see step()

    all_params = [p for m in mods for p in m.parameters()]

    def ddp_wrap(ms):
        # the reference scripts' way (train.py:136-139): one DistributedDataParallel wrapper per module, default arguments;
        # the head is called through .module (train.py:251), so its gradient is not exchanged
        from torch.nn.parallel import DistributedDataParallel as DDP
        kw = {"gradient_as_bucket_view": True} if args.ddp_view else {}
        w = {k: DDP(m, device_ids=[local_rank], **kw) for k, m in zip(("cross", "cui", "cti", "head"), ms[1:])}
        return {"img": ms[0], "cross": w["cross"], "cui": w["cui"], "cti": w["cti"], "head": w["head"].module}

    def step(b):
        for p in all_params:                 # what optimizer.zero_grad() does (train.py:288): walk a flat parameter list
            p.grad = None
        ins = {k: (v.detach().requires_grad_(True) if k in grad_keys else v) for k, v in b.items()}
        cross_c, cui_c, cti_c, head_c = call["cross"], call["cui"], call["cti"], call["head"]
        with torch.autocast("cuda", dtype=torch.bfloat16):
            img_vec = call["img"](ins["img_tokens"], trainable=False)
            side = side_stream if use_side[0] else None
            if side is not None:
                # the two fuse experts are independent, latency-bound chains of ~100 small launches each (2 tokens per
                # sample): they run on side streams, forward and (through autograd's stream tracking) backward
                main = torch.cuda.current_stream()
                side.wait_stream(main)
                with torch.cuda.stream(side):
                    xti = cti_c(ins["i_doc"], img_vec)
                if side2 is not None:
                    side2.wait_stream(main)
                    with torch.cuda.stream(side2):
                        xui = cui_c(ins["u_doc"], img_vec)
                ui = cross_c(ins["u_sent"], ins["u_mask"], ins["i_sent"], ins["i_mask"])
                if side2 is None:
                    xui = cui_c(ins["u_doc"], img_vec)
                else:
                    main.wait_stream(side2)
                    xui.record_stream(main)
                    img_vec.record_stream(side2)
                    ins["u_doc"].record_stream(side2)
                main.wait_stream(side)
                xti.record_stream(main)
                img_vec.record_stream(side)
                ins["i_doc"].record_stream(side)
            else:
                ui = cross_c(ins["u_sent"], ins["u_mask"], ins["i_sent"], ins["i_mask"])
                xui = cui_c(ins["u_doc"], img_vec)
                xti = cti_c(ins["i_doc"], img_vec)
            ev = torch.stack([ins["u_doc"].float(), ins["i_doc"].float(), img_vec, ui, xui, xti], dim=1)
            lg, lb = head_c(ev)
            loss = F.binary_cross_entropy_with_logits(lg.float(), ins["y_good"], pos_weight=pw_g) + \
                   F.binary_cross_entropy_with_logits(lb.float(), ins["y_best"], pos_weight=pw_b)
        loss.backward()
        return loss

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_loop(n_steps):
        """n steps bracketed by barrier + synchronize; returns (total ms, sorted per-step ms, slowest index, host stats)."""
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        step_ev = [torch.cuda.Event(enable_timing=True) for _ in range(n_steps + 1)]
        for e in step_ev:
            e.record()            # torch creates the CUDA event lazily at the first record: do that outside the timed loop
        barrier()
        ev0.record()
        t_host = time.perf_counter()
        step_ev[0].record()
        host_t = [time.perf_counter()]
        for i in range(n_steps):
            if i >= 2 and not os.environ.get("BENCH_NO_THROTTLE"):
                step_ev[i - 1].synchronize()   # stay at most two steps ahead of the device (bounded launch-queue depth)
            step(resident)
            step_ev[i + 1].record()
            host_t.append(time.perf_counter())
        host_ms = (time.perf_counter() - t_host) * 1e3 / n_steps      # time the host needs to ENQUEUE a step
        ev1.record()
        barrier()
        per = [step_ev[i].elapsed_time(step_ev[i + 1]) for i in range(n_steps)]
        slow = max(range(n_steps), key=lambda i: per[i])
        per.sort()
        return ev0.elapsed_time(ev1), per, slow, host_ms, max((host_t[i + 1] - host_t[i]) * 1e3 for i in range(n_steps))

    def max_over_ranks(ms):
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        if distributed:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    W = max(args.warmup, 3)
    native_result = None
    # ---------------- N > 1: the package's own gradient exchange, timed first (before DDP hooks exist) ----------------
    if distributed and args.exchange in ("native", "both"):
        pkg.functional.enable_grad_allreduce()
        for _ in range(W):
            step(resident)
        barrier()
        timed_loop(args.steps)                     # rehearsal (see below)
        ms_total, per_step, _, host_ms, _ = timed_loop(args.steps)
        ms_native = max_over_ranks(ms_total) / args.steps
        native_result = {"value": B * world / (ms_native * 1e-3), "unit": UNIT, "ms_per_step": ms_native,
                         "host_enqueue_ms_per_step": host_ms,
                         "how": "functional.enable_grad_allreduce(): one in-place NCCL all-reduce (AVG) per module / encoder-layer stage on the "
                                "flat fp32 gradient buffer the backward kernels already wrote, started when the stage is enqueued, installed "
                                "into .grad by an engine callback at the end of backward; no per-parameter hooks or bucket copies"}
        pkg.functional.disable_grad_allreduce()
    exchange = "none"
    other_ddp = None
    if distributed and args.exchange == "both":
        # the DDP wrappers around modules with the OTHER parameter layout, timed next to the headline
        o_mods = make_modules(not fused)
        keep = (dict(call), list(all_params))
        call.update(ddp_wrap(o_mods))
        all_params[:] = [p for m in o_mods for p in m.parameters()]
        for _ in range(W):
            step(resident)
        barrier()
        timed_loop(args.steps)
        ms_o, _, _, _, _ = timed_loop(args.steps)
        ms_o = max_over_ranks(ms_o) / args.steps
        other_ddp = {"parameters": "one nn.Parameter per tensor, as in the reference" if fused else
                                   "one fused nn.Parameter per module (MMOE_FLAT_PARAMS=1)",
                     "value": B * world / (ms_o * 1e-3), "unit": UNIT, "ms_per_step": ms_o,
                     "how": "the same DistributedDataParallel wrappers (default arguments) around modules with the other parameter layout"}
        call.update(keep[0])
        all_params[:] = keep[1]
        del o_mods
        gc.collect()
        torch.cuda.empty_cache()
    if distributed and args.exchange in ("ddp", "both"):
        call.update(ddp_wrap(mods))
        exchange = ("torch DistributedDataParallel, one wrapper per module, default arguments (train.py:136-139); head via .module "
                    "(train.py:251); module parameters: " + params_note)
    elif distributed:
        pkg.functional.enable_grad_allreduce()
        exchange = "native flat-buffer all-reduce (functional.enable_grad_allreduce)"

    # ---------------- device-resident timing ----------------
    # The NVML sampler starts BEFORE the warm-up steps: the first calls of each NVML query take tens of milliseconds and
    # hold a driver lock that kernel launches also need (measured: with the sampler started right before the timed loop its
    # second step took 18-110 ms of host time; started here, the first calls land in the warm-up).
    sampler = ClockSampler(local_rank)
    if rank == 0 and not os.environ.get("BENCH_NO_SAMPLER"):
        sampler.start()
    for _ in range(W):
        step(resident)
    barrier()
    L.mmoe_gemm_timing(128 * args.steps)           # event pairs for the roofline pass are created up front
    L.mmoe_gemm_timing(0)
    # Rehearsal: the same loop, same pacing, untimed.  Tensors that cross to the side streams are returned to torch's caching
    # allocator only when the device has passed them (record_stream), so how far the host runs ahead decides how much
    # memory the allocator needs; a pacing seen for the first time inside the timed loop made it grow there (cudaMalloc,
    # 20-100 ms on the enqueueing thread — the 22 ms / 108 ms "host hiccup" steps of earlier records).
    timed_loop(args.steps)
    barrier()
    sampler.samples.clear()
    gc.collect()
    gc.disable()          # a generation-2 collection inside a timed loop stalls the enqueueing thread for tens of ms

    # pass 1 (the headline): no per-launch events
    L.mmoe_launch_count(1)
    ms_total, per_step, slowest_step, host_ms_step, host_ms_max = timed_loop(args.steps)
    clocks = sampler.stop() if rank == 0 else None
    launches = int(L.mmoe_launch_count(1))
    # pass 2 (the roofline): the same K steps with a CUDA event pair around every GEMM launch.  Kept out of pass 1 because
    # the 2 x 88 event records per step cost ~0.4 ms/step of launch overlap.
    # ... and on one stream, so that no other kernel shares the SMs while a GEMM is being timed.
    use_side[0] = False
    L.mmoe_gemm_timing(1)
    ms_total_ev, _, _, _, _ = timed_loop(args.steps)
    L.mmoe_gemm_timing(0)
    use_side[0] = side_stream is not None
    n_rec = L.mmoe_gemm_timing_dump(None, 0)
    rows = (C.c_double * (10 * max(n_rec, 1)))()
    L.mmoe_gemm_timing_dump(rows, n_rec)
    g_ms, g_fl, g_n = C.c_double(), C.c_double(), C.c_int64()
    L.mmoe_gemm_timing_read(C.byref(g_ms), C.byref(g_fl), C.byref(g_n), 1)
    ms_step = max_over_ranks(ms_total) / args.steps
    value = B * world / (ms_step * 1e-3)

    # ---------------- end-to-end timing: pinned host -> device every step, loss read back ----------------
    copy_stream = torch.cuda.Stream(device=dev)
    slots = [None, None]
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    done = [torch.cuda.Event(), torch.cuda.Event()]      # compute finished with the slot

    def prefetch(j, src):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(done[j])
            slots[j] = {k: v.to(dev, non_blocking=True) for k, v in src.items()}
            ready[j].record(copy_stream)

    cur = torch.cuda.current_stream()
    for j in range(2):
        done[j].record(cur)

    loss_host = torch.empty(max(args.steps, 8), dtype=torch.float32).pin_memory()

    def e2e_loop(n):
        """Every step: its batch comes from pinned host memory (copied on the side stream while the previous step
        computes) and its loss is copied back to pinned host memory and READ by the host — one step late, while the next
        step is already enqueued, the way a training loop logs its loss without draining the device every step."""
        prefetch(0, host[0])
        last = None
        landed = [torch.cuda.Event() for _ in range(n)]
        for i in range(n):
            j = i & 1
            if i + 1 < n:
                prefetch(j ^ 1, host[(i + 1) % n_host])
            cur.wait_event(ready[j])
            for v in slots[j].values():
                v.record_stream(cur)
            l = step(slots[j])
            done[j].record(cur)
            loss_host[i:i + 1].copy_(l.detach().reshape(1), non_blocking=True)      # device -> host read of the step's result
            landed[i].record(cur)
            if i >= 1:
                landed[i - 1].synchronize()
                last = float(loss_host[i - 1])
        landed[n - 1].synchronize()
        last = float(loss_host[n - 1])
        return last

    # warm the host->device path first (the PCIe link and the pinned pages need a few hundred ms of traffic before the
    # copy rate is steady: cold runs measured 32 K samples/s end to end, warm ones 44 K, with the same device time)
    t_warm = time.perf_counter()
    while time.perf_counter() - t_warm < 1.0:
        with torch.cuda.stream(copy_stream):
            for v in host[0].values():
                v.to(dev, non_blocking=True)
        copy_stream.synchronize()
    e2e_loop(5)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    last_loss = e2e_loop(args.steps)
    ev1.record()
    barrier()
    e_ms = max(ev0.elapsed_time(ev1), 0.0)
    e2e_value = B * world / (max_over_ranks(e_ms) / args.steps * 1e-3)
    gc.enable()

    home_ddp = None
    if distributed and not args.no_extras:
        try:
            home_ddp = home_ddp_measurement(dev, local_rank, world, 10)
        except Exception as ex:  # noqa: BLE001
            home_ddp = {"error": repr(ex)}
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        sustained = float(peaks.get("bf16_tflops_sustained", 1400.0))
        burst = float(peaks.get("bf16_tflops", 1590.0))
        hbm = float(peaks.get("hbm_gbs", 6650.0))
        # the burst peak was measured at full clock, the sustained one power-capped at ~1.36 GHz: use the one whose clock
        # regime matches what the sampler saw during the timed region
        full_clock = bool(clocks and clocks.get("sm_mhz") and clocks["sm_mhz"] >= 0.97 * (clocks.get("sm_max_mhz") or 1e9)
                          and "sw_power_cap" not in clocks.get("reasons", []))
        peak_tf = burst if full_clock else sustained
        src = "MEASURED_PEAKS.json" if peaks else "fallback (B200_PROFILING.md)"
        peak_src = (f"{src} bf16_tflops (cuBLAS burst, measured at full clock) — the timed region ran at {clocks['sm_mhz']:.0f} MHz with no power cap"
                    if full_clock else f"{src} bf16_tflops_sustained (cuBLAS, power-capped clock regime)")
        # per-class breakdown of the event-timed GEMM launches
        classes = {}
        pair_ms = pair_fl = 0.0
        for r in range(n_rec):
            ms_, fl_, tc_, bn_, ctas_, npb_, M_, N_, K_, maj_ = (rows[10 * r + c] for c in range(10))
            if not tc_:
                continue
            key = f"{int(M_)}x{int(N_)}x{int(K_)}|{'AB'[int(maj_) & 1]}{'AB'[(int(maj_) >> 1) & 1]}|g{int(npb_)}|bn{int(bn_)}c{int(ctas_)}"
            c = classes.setdefault(key, [0, 0.0, 0.0])
            c[0] += 1; c[1] += ms_; c[2] += fl_
            if int(ctas_) == 2:
                pair_ms += ms_; pair_fl += fl_
        by_class = sorted(({"class": k, "launches_per_step": v[0] / args.steps, "us_per_launch": 1e3 * v[1] / v[0],
                            "ms_per_step": v[1] / args.steps, "tflops": (v[2] / (v[1] * 1e-3) / 1e12) if v[1] > 0 else 0.0}
                           for k, v in classes.items()), key=lambda d: -d["ms_per_step"])
        all_tf = (g_fl.value / (g_ms.value * 1e-3)) / 1e12 if g_ms.value > 0 else 0.0
        pair_tf = (pair_fl / (pair_ms * 1e-3)) / 1e12 if pair_ms > 0 else 0.0
        traffic, traffic_note = None, "no ncu capture recorded for this build"
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json")))
            traffic, traffic_note = tj.get("traffic_bytes_per_launch"), tj.get("note")
        except Exception:
            pass
        path_tf = FLOP_PER_SAMPLE * B * world / (ms_step * 1e-3) / 1e12 / world
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": W,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic", "config": workload_config(B, world) | {"gradient_exchange": exchange},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes(host[0]), "d2h_bytes_per_step": 4,
                    "note": "pinned host batch copied on a side stream (double buffered) every step — sentence vectors and ViT tokens shipped "
                            "as bf16 and cast on the device, the rest fp32; every step's loss copied to pinned host memory and read by the "
                            "host one step later"},
            "gpu_launches": launches,
            "host_enqueue_ms_per_step": host_ms_step,
            "ms_per_step_min_median_max": [per_step[0], per_step[len(per_step) // 2], per_step[-1]], "slowest_step": slowest_step,
            "host_ms_max_step": host_ms_max,
            "clocks": clocks,
            "roofline": {"bound": "tensor",
                         "kernel": "gemm_tc_kernel<256,*,2> — the CTA-pair (cta_group::2) tcgen05 GEMM, all of its launches in the timed region "
                                   "(the dominant kernel: encoder-layer forward / dgrad / wgrad GEMMs)",
                         "achieved": pair_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": pair_tf / peak_tf if peak_tf else None,
                         "traffic": traffic, "traffic_note": traffic_note,
                         "peak_source": peak_src, "frac_of_burst_peak": pair_tf / burst, "frac_of_sustained_peak": pair_tf / sustained,
                         "pair_kernel_ms_per_step": pair_ms / args.steps,
                         "all_gemm_launches": {"achieved": all_tf, "frac": all_tf / peak_tf, "ms_per_step": g_ms.value / args.steps,
                                               "launches_per_step": g_n.value / args.steps},
                         "by_class": by_class[:14],
                         "measured_over": "a second pass of the same K steps, single stream, with an event pair around every GEMM launch",
                         "ms_per_step_with_events": ms_total_ev / args.steps,
                         "gemm_share_of_step": (g_ms.value / ms_total_ev) if ms_total_ev else None,
                         "path_tflops": path_tf, "path_frac_of_peak": path_tf / peak_tf},
            "loss": last_loss,
        }
        if native_result is not None:
            line["native_exchange"] = native_result
        if other_ddp is not None:
            line["ddp_other_parameter_layout"] = other_ddp
        line["config"]["parameters"] = params_note
        if home_ddp is not None:
            line["extra"] = {"home_step_ddp_fwd_bwd_bf16": home_ddp}
        if world == 1 and not args.no_extras:
            # ---- the same step in eager PyTorch on this GPU (torch's own kernels), same weights ----
            try:
                from oracle import eager_ref as E
                eager = {}
                emods = E.make_v1_modules(dev, train=not args.eval_mode)
                native_by_name = dict(zip(("img", "cross", "concat_ui", "concat_ti", "head"), mods))
                for k_, m_ in emods.items():                      # same weights as the native modules (identical state_dict keys)
                    m_.load_state_dict({k: v for k, v in native_by_name[k_].state_dict().items() if not k.startswith("backbone")}, strict=True)
                for name, dt, reps in (("bf16_autocast", torch.bfloat16, 5), ("fp32", None, 2)):
                    stp = eager_step_fn(emods, resident, dev, dt)
                    ms_e = cuda_time(stp, 2, reps)
                    eager[name] = {"ms_per_step": ms_e, "samples_per_s": B / ms_e * 1e3}
                eager["speedup_bf16"] = eager["bf16_autocast"]["ms_per_step"] / ms_step
                eager["what"] = ("the reference's module composition on stock torch.nn modules (oracle/eager_ref.py: nn.TransformerEncoderLayer, "
                                 "nn.MultiheadAttention, nn.Linear, nn.LayerNorm called as the reference's forward bodies call them): "
                                 "PyTorch-eager on this B200, train mode, same weights and batch, fwd+bwd")
                line["eager_b200"] = eager
                del emods, stp
            except Exception as ex:  # noqa: BLE001
                line["eager_b200"] = {"error": repr(ex)}
            try:
                line["extra"] = extra_measurements(dev, hbm, peak_tf)

                def factory(Bs):
                    bs = {k: v.to(dev) for k, v in make_host_batch(Bs, 77, pin=False).items()}
                    return lambda: step(bs)
                line["extra"]["v1_step_b128"] = small_batch_measurement(dev, factory, 128, ms_step / B)
            except Exception as ex:  # noqa: BLE001
                line["extra"] = {"error": repr(ex)}
        if not args.no_cpu_baseline and world == 1:
            try:
                v, ms_cpu, cores, done_ = time_cpu_reference(args.cpu_baseline_batch, 2, 1, 60.0)
                line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                        "sample": f"reference forward composition on stock torch.nn modules (oracle/eager_ref.py), fp32, train "
                                                  f"mode, {args.cpu_baseline_batch}-sample batch, fwd+bwd, {done_} timed steps ({ms_cpu:.0f} ms/step)"}
            except Exception as ex:  # noqa: BLE001
                line["cpu_baseline"] = {"error": repr(ex)}
        print(json.dumps(line), flush=True, file=_REAL_STDOUT or sys.stdout)
    if distributed:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
