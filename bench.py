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
For N > 1 each rank runs the same per-GPU batch (weak scaling) and the gradients are averaged over NCCL, overlapped with
backward: by the package's native flat-buffer exchange (default) or by DistributedDataParallel wrappers (--ddp).
The headline pass carries no per-launch events; the GEMM roofline is measured in a second, single-stream pass of the
same steps with a CUDA event pair around every GEMM launch.

One JSON line on stdout (rank 0); see README / DESIGN.md for the keys.
"""
from __future__ import annotations

import argparse
import ctypes as C
import gc
import json
import os
import subprocess
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


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ----------------------------------------------------------------------------------------------
# synthetic inputs (oracle/synth.py is test infrastructure; the bench makes its own with torch RNG)
# ----------------------------------------------------------------------------------------------
def make_host_batch(B: int, seed: int, pin: bool):
    g = torch.Generator().manual_seed(seed)
    def t(*shape):
        x = torch.randn(*shape, generator=g)
        return x.pin_memory() if pin else x
    lens_u = torch.randint(1, S + 1, (B,), generator=g)
    lens_i = torch.randint(1, S + 1, (B,), generator=g)
    ar = torch.arange(S)[None]
    batch = {
        "u_sent": t(B, S, D), "i_sent": t(B, S, D),
        "u_mask": (ar >= lens_u[:, None]), "i_mask": (ar >= lens_i[:, None]),
        "u_doc": t(B, D), "i_doc": t(B, D), "img_tokens": t(B, NTOK, D),
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
# clocks sampler (nvidia-smi during the timed region)
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
            # CUDA_VISIBLE_DEVICES may remap indices: resolve through the PCI bus id of the torch device
            bus = torch.cuda.get_device_properties(index).pci_bus_id if hasattr(torch.cuda.get_device_properties(index), "pci_bus_id") else None
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index) if bus is None else pynvml.nvmlDeviceGetHandleByIndex(index)
        except Exception:
            self._h = None

    def _loop(self):
        nv, h = self._nv, self._h
        while not self._stop.is_set():
            try:
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
                rs = nv.nvmlDeviceGetCurrentClocksEventReasons(h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.samples.append((sm, mx, rs))
            except Exception:
                pass
            self._stop.wait(0.02)

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
# reference arm / cpu baseline: the oracle port of the reference's modules on the host cores
# ----------------------------------------------------------------------------------------------
def cpu_oracle_step_fn(B: int):
    """fwd+bwd of the same path through oracle/ (CPU restatement of the reference) in fp32, eval-mode arithmetic
    (the oracle has no RNG; dropout costs the reference extra time on CPU, so this favours the baseline)."""
    from oracle import mmoe_oracle as O
    from oracle import synth
    torch.manual_seed(0)
    sds = {
        "img": synth.fill_state_dict(synth.image_wrapper_shapes(), 1),
        "cross": synth.fill_state_dict(synth.cross_expert_shapes(), 2),
        "concat_ui": synth.fill_state_dict(synth.cross_fuse_shapes(), 3),
        "concat_ti": synth.fill_state_dict(synth.cross_fuse_shapes(), 4),
        "head": synth.fill_state_dict(synth.mmoe_head_shapes(), 5),
    }
    for sd in sds.values():
        for v in sd.values():
            v.requires_grad_(True)
    b = make_host_batch(B, 99, pin=False)
    for k in ("u_sent", "i_sent", "u_doc", "i_doc"):
        b[k].requires_grad_(True)

    def step():
        for sd in sds.values():
            for v in sd.values():
                v.grad = None
        lg, lb = O.v1_fusion_path(sds, b["u_sent"], b["u_mask"], b["i_sent"], b["i_mask"], b["u_doc"], b["i_doc"], b["img_tokens"])
        loss = O.bce_with_logits(lg, b["y_good"], O.POS_WEIGHT_GOOD) + O.bce_with_logits(lb, b["y_best"], O.POS_WEIGHT_BEST)
        loss.backward()
        return float(loss.detach())
    return step


def time_cpu_oracle(B: int, steps: int, warmup: int):
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    step = cpu_oracle_step_fn(B)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return B / dt, dt * 1e3, cores


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    B = args.ref_batch
    value, ms, cores = time_cpu_oracle(B, max(args.steps, 1), max(args.warmup, 1))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.batch, args.gpus) | {"sample": f"{B} samples per step on the host CPU"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"oracle/ port of the reference modules, fp32, {B}-sample batch, fwd+bwd, {args.steps} steps"},
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
_REAL_STDOUT = None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--batch", type=int, default=512, help="per-GPU batch (train.py default 512)")
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--ref-batch", type=int, default=16, help="samples per step of the CPU reference arm")
    ap.add_argument("--cpu-baseline-batch", type=int, default=16)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--eval-mode", action="store_true", help="dropout off (debug)")
    ap.add_argument("--no-side-stream", action="store_true", help="run both fuse experts on the main stream")
    ap.add_argument("--bucket-mb", type=int, default=25, help="DDP gradient bucket size (with --ddp)")
    ap.add_argument("--ddp", action="store_true", help="N > 1: wrap the modules in torch DistributedDataParallel (as train.py "
                                                        "does) instead of the native flat-buffer gradient all-reduce")
    ap.add_argument("--nccl-ctas", type=int, default=0,
                    help="N>1: if > 0, cap NCCL at this many CTAs and keep as many SMs free of GEMM CTAs (measured: capping "
                         "lengthens the exposed part of the all-reduce at N=2; default leaves NCCL alone)")
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
        # ~280 MB of gradients per step overlap ~8 ms of backward: a few NCCL CTAs are plenty, and every SM NCCL holds
        # is an SM the persistent GEMM cannot use while it runs
        if args.nccl_ctas > 0:
            os.environ.setdefault("NCCL_MAX_CTAS", str(args.nccl_ctas))
        dist.init_process_group("nccl", device_id=dev)
    L = pkg.lib()
    check(L.mmoe_init(), "init")
    if distributed and args.nccl_ctas > 0:
        L.mmoe_set_sm_reserve(args.nccl_ctas)

    M = pkg.modules
    torch.manual_seed(1234)
    img = M.ItemImageExpert(Passthrough(), pool_type="mean").to(dev)
    cross = M.RobustTextCrossExpert().to(dev)
    cui = M.EnhancedCrossFuse().to(dev)
    cti = M.EnhancedCrossFuse().to(dev)
    head = M.TwoTaskMMoE().to(dev)
    mods = [img, cross, cui, cti, head]
    for m in mods:
        m.train(not args.eval_mode)
    cross_c, cui_c, cti_c, head_c = cross, cui, cti, head
    if distributed and args.ddp:
        # the reference scripts' way (train.py:133-139): one DistributedDataParallel wrapper per module
        from torch.nn.parallel import DistributedDataParallel as DDP
        cross_c, cui_c, cti_c, head_c = (DDP(m, device_ids=[local_rank], gradient_as_bucket_view=True, bucket_cap_mb=args.bucket_mb)
                                         for m in (cross, cui, cti, head))
    elif distributed:
        # native exchange: every module backward all-reduces (averages) its ONE flat gradient buffer as soon as it is
        # complete — no per-parameter hooks or bucket copies (DDP spends ~1 ms/step on ~300 per-parameter copy kernels)
        for m in mods:
            for prm in m.parameters():
                dist.broadcast(prm.data, src=0)
        pkg.functional.enable_grad_allreduce()
    pw_g = torch.tensor(858627.0 / 990303.0, device=dev)      # train.py:189-192
    pw_b = torch.tensor(1328721.0 / 520209.0, device=dev)

    B = args.batch
    n_host = 2
    host = [make_host_batch(B, 1234 + rank * 17 + j, pin=True) for j in range(n_host)]
    resident = {k: v.to(dev) for k, v in host[0].items()}
    grad_keys = ("u_sent", "i_sent", "u_doc", "i_doc")

    side_stream = None if args.no_side_stream else torch.cuda.Stream(device=dev)
    side2 = torch.cuda.Stream(device=dev) if side_stream is not None else None
    use_side = [side_stream is not None]
    if hasattr(torch.autograd.graph, "set_warn_on_accumulate_grad_stream_mismatch"):
        torch.autograd.graph.set_warn_on_accumulate_grad_stream_mismatch(False)     # intentional: see step()

    def step(b):
        for m in mods:
            m.zero_grad(set_to_none=True)
        ins = {k: (v.detach().requires_grad_(True) if k in grad_keys else v) for k, v in b.items()}
        with torch.autocast("cuda", dtype=torch.bfloat16):
            img_vec = img(ins["img_tokens"], trainable=False)
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
            ev = torch.stack([ins["u_doc"], ins["i_doc"], img_vec, ui, xui, xti], dim=1)
            lg, lb = head_c(ev)
            loss = F.binary_cross_entropy_with_logits(lg.float(), ins["y_good"], pos_weight=pw_g) + \
                   F.binary_cross_entropy_with_logits(lb.float(), ins["y_best"], pos_weight=pw_b)
        loss.backward()
        pkg.functional.wait_grad_allreduce()       # no-op unless the native gradient exchange is on
        return loss

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident timing ----------------
    for _ in range(max(args.warmup, 3)):
        step(resident)
    barrier()
    L.mmoe_gemm_timing(128 * args.steps)           # event pairs for the roofline pass are created up front
    L.mmoe_gemm_timing(0)
    sampler = ClockSampler(local_rank)
    if rank == 0 and not os.environ.get("BENCH_NO_SAMPLER"):
        sampler.start()                            # first NVML calls happen during the extra warm-up step below
    step(resident)
    barrier()
    sampler.samples.clear()
    gc.collect()
    gc.disable()          # a generation-2 collection inside a timed loop stalls the enqueueing thread for tens of ms

    def timed_loop():
        """K steps bracketed by barrier + synchronize; returns (total ms, sorted per-step ms, slowest index, host stats)."""
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        step_ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
        for e in step_ev:
            e.record()            # torch creates the CUDA event lazily at the first record: do that outside the timed loop
        barrier()
        ev0.record()
        t_host = time.perf_counter()
        step_ev[0].record()
        host_t = [time.perf_counter()]
        for i in range(args.steps):
            if i >= 2 and not os.environ.get("BENCH_NO_THROTTLE"):
                step_ev[i - 1].synchronize()   # stay at most two steps ahead of the device (bounded launch-queue depth)
            step(resident)
            step_ev[i + 1].record()
            host_t.append(time.perf_counter())
        host_ms = (time.perf_counter() - t_host) * 1e3 / args.steps      # time the host needs to ENQUEUE a step
        ev1.record()
        barrier()
        per = [step_ev[i].elapsed_time(step_ev[i + 1]) for i in range(args.steps)]
        slow = max(range(args.steps), key=lambda i: per[i])
        per.sort()
        return ev0.elapsed_time(ev1), per, slow, host_ms, max((host_t[i + 1] - host_t[i]) * 1e3 for i in range(args.steps))

    # pass 1 (the headline): no per-launch events
    L.mmoe_launch_count(1)
    ms_total, per_step, slowest_step, host_ms_step, host_ms_max = timed_loop()
    clocks = sampler.stop() if rank == 0 else None
    launches = int(L.mmoe_launch_count(1))
    # pass 2 (the roofline): the same K steps with a CUDA event pair around every GEMM launch.  Kept out of pass 1 because
    # the 2 x 88 event records per step cost ~0.4 ms/step of launch overlap.
    # ... and on one stream, so that no other kernel shares the SMs while a GEMM is being timed.
    use_side[0] = False
    L.mmoe_gemm_timing(1)
    ms_total_ev, _, _, _, _ = timed_loop()
    L.mmoe_gemm_timing(0)
    use_side[0] = side_stream is not None
    g_ms, g_fl, g_n = C.c_double(), C.c_double(), C.c_int64()
    L.mmoe_gemm_timing_read(C.byref(g_ms), C.byref(g_fl), C.byref(g_n), 1)
    t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
    if distributed:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
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
    t0 = time.perf_counter()
    ev0.record()
    last_loss = e2e_loop(args.steps)
    ev1.record()
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    e_ms = max(ev0.elapsed_time(ev1), 0.0)
    t = torch.tensor([e_ms], device=dev, dtype=torch.float64)
    if distributed:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = B * world / (float(t.item()) / args.steps * 1e-3)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak_tf = float(peaks.get("bf16_tflops_sustained", 1400.0))
        peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (cuBLAS, measured)" if peaks else "fallback 1400 TFLOP/s sustained (B200_PROFILING.md)"
        achieved = (g_fl.value / (g_ms.value * 1e-3)) / 1e12 if g_ms.value > 0 else 0.0
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic", "config": workload_config(B, world),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes(host[0]), "d2h_bytes_per_step": 4,
                    "note": "pinned host batch copied on a side stream (double buffered) every step; every step's loss copied to pinned host memory and read by the host one step later"},
            "gpu_launches": launches,
            "host_enqueue_ms_per_step": host_ms_step,
            "ms_per_step_min_median_max": [per_step[0], per_step[len(per_step) // 2], per_step[-1]], "slowest_step": slowest_step,
            "host_ms_max_step": host_ms_max,
            "clocks": clocks,
            "roofline": {"bound": "tensor", "kernel": "gemm_tc_kernel (tcgen05 grouped GEMM, all launches of the timed region)",
                         "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf if peak_tf else None,
                         "traffic": None,
                         "traffic_note": "aggregate over 88 launches, so no single per-launch figure; ncu --set full of the FFN1 "
                                         "768->3072 launch: 55 MB read + 147 MB written vs 255 MB algorithmic "
                                         "(profiles/r01_gemm_final_ncu_full.md)",
                         "peak_source": peak_src,
                         "gemm_ms_per_step": g_ms.value / args.steps, "gemm_launches_per_step": g_n.value / args.steps,
                         "measured_over": "a second pass of the same K steps, single stream, with an event pair around every GEMM launch",
                         "ms_per_step_with_events": ms_total_ev / args.steps,
                         "gemm_share_of_step": (g_ms.value / ms_total_ev) if ms_total_ev else None,
                         "path_tflops": FLOP_PER_SAMPLE * B / (ms_step * 1e-3) / 1e12,
                         "path_frac_of_peak": FLOP_PER_SAMPLE * B / (ms_step * 1e-3) / 1e12 / peak_tf},
            "loss": last_loss,
        }
        if not args.no_cpu_baseline and world == 1:
            try:
                v, ms_cpu, cores = time_cpu_oracle(args.cpu_baseline_batch, 2, 1)
                line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                        "sample": f"oracle/ port of the reference modules, fp32, {args.cpu_baseline_batch}-sample batch, "
                                                  f"fwd+bwd, 2 timed steps ({ms_cpu:.0f} ms/step)"}
            except Exception as ex:  # noqa: BLE001
                line["cpu_baseline"] = {"error": repr(ex)}
        print(json.dumps(line), flush=True, file=_REAL_STDOUT or sys.stdout)
    if distributed:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
