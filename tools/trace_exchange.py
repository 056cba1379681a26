#!/usr/bin/env python
"""Device timeline of the benchmark step under each gradient exchange, to name what the N > 1 step pays for.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
        tools/trace_exchange.py --out gpurun_out/r2_trace

For each mode (none / native / ddp) rank 0 records K steps with torch.profiler (CUPTI kernel activity only) and prints:
the span per step, the time in which no non-NCCL kernel runs, the NCCL kernels (count, time, share that overlaps
compute), the kernels DDP adds (bucket copies), and per kernel family how much its total time grew against mode "none"
(SM / HBM contention from the collective).  nsys is not in the image; the chrome trace of each mode is kept next to
the summary.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import torch.nn.functional as F  # noqa: E402

import bench  # noqa: E402


def union_len(iv):
    iv = sorted(iv)
    tot, cur_s, cur_e = 0.0, None, None
    for s, e in iv:
        if cur_e is None or s > cur_e:
            if cur_e is not None:
                tot += cur_e - cur_s
            cur_s, cur_e = s, e
        else:
            cur_e = max(cur_e, e)
    if cur_e is not None:
        tot += cur_e - cur_s
    return tot


def overlap_len(a, b):
    """total length of (union of a) ∩ (union of b)."""
    return union_len(a) + union_len(b) - union_len(a + b)


def family(name: str) -> str:
    n = name
    if "nccl" in n.lower():
        return "nccl"
    if n.startswith("mmoe::") or "mmoe::" in n:
        n = n.split("mmoe::", 1)[1]
        return "mmoe::" + n.split("<", 1)[0].split("(", 1)[0]
    if "at::" in n or "void at" in n:
        return "torch elementwise/copy"
    return n[:40]


def analyse(trace_path, steps):
    ev = json.load(open(trace_path))["traceEvents"]
    ks = [e for e in ev if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "dur" in e]
    t0 = min(e["ts"] for e in ks)
    t1 = max(e["ts"] + e["dur"] for e in ks)
    comp = [(e["ts"], e["ts"] + e["dur"]) for e in ks if "nccl" not in e["name"].lower()]
    nccl = [(e["ts"], e["ts"] + e["dur"]) for e in ks if "nccl" in e["name"].lower()]
    fam = {}
    for e in ks:
        f = family(e["name"])
        c = fam.setdefault(f, [0, 0.0])
        c[0] += 1
        c[1] += e["dur"]
    streams = {}
    for e in ks:
        s = e.get("args", {}).get("stream", -1)
        c = streams.setdefault(s, [0, 0.0])
        c[0] += 1
        c[1] += e["dur"]
    span = (t1 - t0) / 1e3
    return {"span_ms_per_step": span / steps,
            "compute_busy_ms_per_step": union_len(comp) / 1e3 / steps,
            "no_compute_kernel_ms_per_step": (span - union_len(comp) / 1e3) / steps,
            "nccl_launches_per_step": len(nccl) / steps,
            "nccl_ms_per_step": sum(e - s for s, e in nccl) / 1e3 / steps,
            "nccl_overlapping_compute_ms_per_step": overlap_len(nccl, comp) / 1e3 / steps if nccl else 0.0,
            "nccl_exposed_ms_per_step": (union_len(nccl) - overlap_len(nccl, comp)) / 1e3 / steps if nccl else 0.0,
            "families": {k: {"launches_per_step": v[0] / steps, "ms_per_step": v[1] / 1e3 / steps} for k, v in fam.items()},
            "streams": {str(k): {"launches_per_step": v[0] / steps, "ms_per_step": v[1] / 1e3 / steps} for k, v in streams.items()}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/r2_trace")
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--batch", type=int, default=512)
    ap.add_argument("--modes", default="none,native,ddp,ddpflat")
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import mmoe_multimodal_rec_b200 as pkg
    from mmoe_multimodal_rec_b200._lib import check
    check(pkg.lib().mmoe_init(), "init")

    mods_d = bench.build_modules(dev)
    img, cross, cui, cti, head = (mods_d[k] for k in ("img", "cross", "concat_ui", "concat_ti", "head"))
    mods = [img, cross, cui, cti, head]
    if world > 1:
        for m in mods:
            for p in m.parameters():
                dist.broadcast(p.data, src=0)
    call = {"img": img, "cross": cross, "cui": cui, "cti": cti, "head": head}
    b = {k: v.to(dev) for k, v in bench.make_host_batch(args.batch, 1234 + rank * 17, pin=False).items()}
    pw_g, pw_b = torch.tensor(bench.POS_W_GOOD, device=dev), torch.tensor(bench.POS_W_BEST, device=dev)
    side, side2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    if hasattr(torch.autograd.graph, "set_warn_on_accumulate_grad_stream_mismatch"):
        torch.autograd.graph.set_warn_on_accumulate_grad_stream_mismatch(False)
    params = [p for m in mods for p in m.parameters()]
    gk = ("u_sent", "i_sent", "u_doc", "i_doc")

    def step():                                   # bench.py main().step, same stream layout
        for p in params:
            p.grad = None
        ins = {k: (v.detach().requires_grad_(True) if k in gk else v) for k, v in b.items()}
        with torch.autocast("cuda", dtype=torch.bfloat16):
            img_vec = call["img"](ins["img_tokens"], trainable=False)
            main_s = torch.cuda.current_stream()
            side.wait_stream(main_s)
            with torch.cuda.stream(side):
                xti = call["cti"](ins["i_doc"], img_vec)
            side2.wait_stream(main_s)
            with torch.cuda.stream(side2):
                xui = call["cui"](ins["u_doc"], img_vec)
            ui = call["cross"](ins["u_sent"], ins["u_mask"], ins["i_sent"], ins["i_mask"])
            main_s.wait_stream(side2)
            main_s.wait_stream(side)
            for t, s in ((xui, main_s), (xti, main_s), (img_vec, side), (img_vec, side2), (ins["u_doc"], side2), (ins["i_doc"], side)):
                t.record_stream(s)
            ev = torch.stack([ins["u_doc"].float(), ins["i_doc"].float(), img_vec, ui, xui, xti], dim=1)
            lg, lb = call["head"](ev)
            loss = F.binary_cross_entropy_with_logits(lg.float(), ins["y_good"], pos_weight=pw_g) + \
                   F.binary_cross_entropy_with_logits(lb.float(), ins["y_best"], pos_weight=pw_b)
        loss.backward()

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    from torch.profiler import profile, ProfilerActivity
    results = {}
    for mode in args.modes.split(","):
        if mode != "none" and world == 1:
            continue
        if mode == "native":
            pkg.functional.enable_grad_allreduce()
        elif mode in ("ddp", "ddpflat"):
            from torch.nn.parallel import DistributedDataParallel as DDP
            ms = mods
            if mode == "ddpflat":                 # fused-parameter modules (one nn.Parameter each)
                md = bench.build_modules(dev, fused=True)
                ms = [md[k] for k in ("img", "cross", "concat_ui", "concat_ti", "head")]
                for m in ms:
                    for p in m.parameters():
                        dist.broadcast(p.data, src=0)
                params[:] = [p for m in ms for p in m.parameters()]
            w = {k: DDP(m, device_ids=[local]) for k, m in zip(("cross", "cui", "cti", "head"), ms[1:])}
            call.update(img=ms[0], cross=w["cross"], cui=w["cui"], cti=w["cti"], head=w["head"].module)
        for _ in range(12):
            step()
        sync()
        # un-profiled timing of the same loop for reference
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            step()
        e1.record()
        sync()
        plain_ms = e0.elapsed_time(e1) / 10
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(args.steps):
                step()
            sync()
        if mode == "native":
            pkg.functional.disable_grad_allreduce()
        if rank == 0:
            path = f"{args.out}_{mode}_n{world}.json"
            prof.export_chrome_trace(path)
            r = analyse(path, args.steps)
            r["plain_ms_per_step"] = plain_ms
            results[mode] = r
        sync()
    if rank == 0:
        base = results.get("none")
        for mode, r in results.items():
            print(f"== {mode} (N={world}): un-profiled {r['plain_ms_per_step']:.3f} ms/step; traced span {r['span_ms_per_step']:.3f}, "
                  f"compute busy {r['compute_busy_ms_per_step']:.3f}, no-compute-kernel time {r['no_compute_kernel_ms_per_step']:.3f}")
            print(f"   nccl: {r['nccl_launches_per_step']:.1f} launches, {r['nccl_ms_per_step']:.3f} ms, overlapping compute "
                  f"{r['nccl_overlapping_compute_ms_per_step']:.3f}, exposed {r['nccl_exposed_ms_per_step']:.3f}")
            fams = sorted(r["families"].items(), key=lambda kv: -kv[1]["ms_per_step"])
            for k, v in fams[:14]:
                d = ""
                if base is not None and mode != "none":
                    b0 = base["families"].get(k, {"ms_per_step": 0.0, "launches_per_step": 0.0})
                    d = f"   (vs none: {v['ms_per_step'] - b0['ms_per_step']:+.3f} ms, {v['launches_per_step'] - b0['launches_per_step']:+.1f} launches)"
                print(f"   {k:44s} {v['launches_per_step']:7.1f} x  {v['ms_per_step']:7.3f} ms{d}")
        json.dump(results, open(f"{args.out}_summary_n{world}.json", "w"), indent=1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
