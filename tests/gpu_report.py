"""One-shot GPU diagnostic: runs every operator / module parity check without stopping at the first
failure and writes gpurun_out/report.json.  (The pytest files assert on the same checks.)

    python tests/gpu_report.py [--quick]
"""
from __future__ import annotations

import ctypes as C
import json
import os
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402


def _gemm_case(L, lib_mod, dtype_t, M, N, K, a_major, b_major, engine, variant, ks=1):
    """D = epilogue(A B^T) against torch.  Returns the normalised max error(s)."""
    from mmoe_multimodal_rec_b200._lib import GemmProblem, check
    dev = "cuda"
    g = torch.Generator(device="cpu").manual_seed(1234 + M + N + K)
    A = torch.randn(M, K, generator=g).to(dev)
    Bm = torch.randn(N, K, generator=g).to(dev)
    mm = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}[dtype_t]
    At, Bt = A.to(dtype_t), Bm.to(dtype_t)
    a_mem = At.contiguous() if a_major == 0 else At.t().contiguous()     # [M,K] or [K,M]
    b_mem = Bt.contiguous() if b_major == 0 else Bt.t().contiguous()
    ref = At.double() @ Bt.double().t()
    p = GemmProblem()
    p.a, p.lda, p.a_major = a_mem.data_ptr(), a_mem.shape[1], a_major
    p.b, p.ldb, p.b_major = b_mem.data_ptr(), b_mem.shape[1], b_major
    p.M, p.N, p.K, p.k_splits = M, N, K, ks
    e = p.epi
    e.alpha = 1.0
    res = {}
    keep = []
    if variant == "plain_f32":
        out = torch.zeros(M, N, device=dev)
        e.out, e.out_dtype, e.ldo = out.data_ptr(), 0, N
        if ks > 1:
            e.accumulate = 1
    elif variant == "bias_relu_t":
        bias = torch.randn(N, generator=g).to(dev)
        out = torch.zeros(M, N, device=dev, dtype=dtype_t)
        e.out, e.out_dtype, e.ldo, e.bias, e.act = out.data_ptr(), mm, N, bias.data_ptr(), 1
        ref = torch.relu(ref + bias.double())
        keep.append(bias)
    elif variant == "bias_drop_t":
        bias = torch.randn(N, generator=g).to(dev)
        out = torch.zeros(M, N, device=dev, dtype=dtype_t)
        e.out, e.out_dtype, e.ldo, e.bias = out.data_ptr(), mm, N, bias.data_ptr()
        e.drop_p, e.drop_key0, e.drop_key1 = 0.25, 1234567, 7654321
        keepm = torch.empty(M * N, dtype=torch.uint8, device=dev)
        check(L.mmoe_dropout_mask(1234567, 7654321, 0.25, M * N, keepm.data_ptr(), torch.cuda.current_stream().cuda_stream), "mask")
        ref = (ref + bias.double()) * keepm.reshape(M, N).double() / 0.75
        res["stat_keep_rate"] = float(keepm.float().mean())
        keep += [bias, keepm]
    elif variant == "gelu_preact_colsum":
        bias = torch.randn(N, generator=g).to(dev)
        out = torch.zeros(M, N, device=dev, dtype=dtype_t)
        pre = torch.zeros(M, N, device=dev, dtype=dtype_t)
        cs = torch.zeros(N, device=dev)
        e.out, e.out_dtype, e.ldo, e.bias, e.act = out.data_ptr(), mm, N, bias.data_ptr(), 2
        e.preact, e.colsum = pre.data_ptr(), cs.data_ptr()
        z = ref + bias.double()
        ref = torch.nn.functional.gelu(z)
        keep += [bias, pre, cs]
    elif variant == "residual_f32":
        bias = torch.randn(N, generator=g).to(dev)
        resid = torch.randn(M, N, generator=g).to(dev)
        out = torch.zeros(M, N, device=dev)
        e.out, e.out_dtype, e.ldo, e.bias, e.residual, e.ld_res = out.data_ptr(), 0, N, bias.data_ptr(), resid.data_ptr(), N
        ref = ref + bias.double() + resid.double()
        keep += [bias, resid]
    else:
        raise KeyError(variant)
    check(L.mmoe_gemm_grouped(C.byref(p), 1, mm, engine, torch.cuda.current_stream().cuda_stream), "gemm")
    torch.cuda.synchronize()
    scale = float(ref.abs().max())
    res["out"] = float((out.double() - ref).abs().max()) / scale
    if variant == "gelu_preact_colsum":
        res["preact"] = float((pre.double() - z).abs().max()) / float(z.abs().max())
        res["colsum"] = float((cs.double() - out.double().sum(0)).abs().max()) / max(float(out.double().sum(0).abs().max()), 1e-9)
    return res


def gemm_checks(report, quick):
    import mmoe_multimodal_rec_b200 as pkg
    L = pkg.lib()
    shapes = [(256, 256, 128), (300, 392, 200), (128, 768, 768), (1000, 3072, 768)]
    if not quick:
        shapes.append((8192, 2304, 768))
    for dtype_t, name in ((torch.bfloat16, "bf16"), (torch.float16, "fp16"), (torch.float32, "fp32")):
        for (M, N, K) in shapes:
            for (am, bm) in ((0, 0), (0, 1), (1, 1), (1, 0)):
                if (am == 1 and M % 8) or (bm == 1 and N % 8) or ((am == 0 or bm == 0) and K % 8):
                    continue
                for variant in ("plain_f32", "bias_relu_t", "bias_drop_t", "gelu_preact_colsum", "residual_f32"):
                    if dtype_t is torch.float16 and variant != "plain_f32":
                        continue
                    for engine in ((0,) if dtype_t is torch.float32 else (0, 1)):
                        key = f"gemm/{name}/eng{engine}/{M}x{N}x{K}/a{am}b{bm}/{variant}"
                        try:
                            report[key] = _gemm_case(L, pkg, dtype_t, M, N, K, am, bm, engine, variant)
                        except Exception as ex:  # noqa: BLE001
                            report[key] = {"error": repr(ex)}
        # split-K accumulation (wgrad shape)
        for engine in ((0,) if dtype_t is torch.float32 else (0, 1)):
            key = f"gemm/{name}/eng{engine}/768x3072x4096/a1b1/splitk8"
            try:
                report[key] = _gemm_case(L, pkg, dtype_t, 768, 3072, 4096, 1, 1, engine, "plain_f32", ks=8)
            except Exception as ex:  # noqa: BLE001
                report[key] = {"error": repr(ex)}


def attention_checks(report):
    import mmoe_multimodal_rec_b200 as pkg
    from mmoe_multimodal_rec_b200._lib import check
    L = pkg.lib()
    for dtype_t, mm, name in ((torch.float32, 0, "fp32"), (torch.bfloat16, 1, "bf16")):
        for (B, S, H, hd) in ((3, 64, 8, 96), (5, 2, 8, 96), (2, 37, 4, 64)):
            d = H * hd
            g = torch.Generator().manual_seed(7)
            qkv = torch.randn(B, S, 3 * d, generator=g).cuda().to(dtype_t)
            lens = torch.randint(1, S + 1, (B,), generator=g)
            mask = (torch.arange(S)[None] >= lens[:, None]).cuda()
            dctx = torch.randn(B, S, d, generator=g).cuda().to(dtype_t)
            q, k, v = [t.double().reshape(B, S, H, hd).transpose(1, 2).requires_grad_(True) for t in qkv.split(d, -1)]
            sc = (q * hd ** -0.5) @ k.transpose(-1, -2)
            sc = sc.masked_fill(mask[:, None, None, :], float("-inf"))
            ctx_ref = (sc.softmax(-1) @ v).transpose(1, 2).reshape(B, S, d)
            ctx_ref.backward(dctx.double())
            dq_ref = torch.cat([t.grad.transpose(1, 2).reshape(B, S, d) for t in (q, k, v)], -1)
            ctx = torch.empty(B, S, d, device="cuda", dtype=dtype_t)
            st = torch.cuda.current_stream().cuda_stream
            es = qkv.element_size()
            m8 = mask.view(torch.uint8)
            key = f"attn/{name}/B{B}S{S}H{H}hd{hd}"
            try:
                check(L.mmoe_attention_fwd(qkv.data_ptr(), 3 * d, qkv.data_ptr() + d * es, 3 * d, qkv.data_ptr() + 2 * d * es, 3 * d,
                                           m8.data_ptr(), ctx.data_ptr(), d, B, S, S, H, hd, 0.0, 0, 0, mm, st), "attn_fwd")
                dqkv = torch.empty_like(qkv)
                bg = torch.zeros(3 * d, device="cuda")
                check(L.mmoe_attention_bwd(qkv.data_ptr(), 3 * d, qkv.data_ptr() + d * es, 3 * d, qkv.data_ptr() + 2 * d * es, 3 * d,
                                           m8.data_ptr(), dctx.data_ptr(), d, dqkv.data_ptr(), dqkv.data_ptr() + d * es,
                                           dqkv.data_ptr() + 2 * d * es, bg.data_ptr(), bg.data_ptr() + 4 * d, bg.data_ptr() + 8 * d,
                                           B, S, S, H, hd, 0.0, 0, 0, mm, st), "attn_bwd")
                torch.cuda.synchronize()
                report[key] = {
                    "ctx": float((ctx.double() - ctx_ref.detach()).abs().max()) / float(ctx_ref.abs().max()),
                    "dqkv": float((dqkv.double() - dq_ref).abs().max()) / float(dq_ref.abs().max()),
                    "bias_grad": float((bg.double() - dqkv.double().sum((0, 1))).abs().max()) / float(dqkv.double().sum((0, 1)).abs().max()),
                }
            except Exception as ex:  # noqa: BLE001
                report[key] = {"error": repr(ex)}


def module_checks(report, quick):
    import parity_util as PU
    from conftest import load_golden
    from oracle import cases as Cs
    for case in Cs.CASES:
        for mode in ("fp32", "bf16", "fp16"):
            if quick and mode == "fp16":
                continue
            key = f"module/{case.name}/{mode}"
            t0 = time.time()
            try:
                stats = {}
                errs = PU.compare_with_oracle(case, mode, stats=stats)
                worst = sorted(errs.items(), key=lambda kv: -kv[1])[:6]
                report[key] = {"max": max(errs.values()), "worst": worst, "n": len(errs), "sec": round(time.time() - t0, 2), **stats}
            except Exception as ex:  # noqa: BLE001
                report[key] = {"error": repr(ex), "trace": traceback.format_exc()[-1500:]}
        try:
            g = load_golden(case.name)
            errs = PU.check_against_golden(case, g, "fp32")
            report[f"golden/{case.name}/fp32"] = {"max": max(errs.values()), "bad": [k for k, v in errs.items() if v > 1e-4][:6]}
        except Exception as ex:  # noqa: BLE001
            report[f"golden/{case.name}/fp32"] = {"error": repr(ex)}


def main():
    quick = "--quick" in sys.argv
    only = None
    out_name = "report.json"
    for i, a in enumerate(sys.argv):
        if a == "--only":
            only = set(sys.argv[i + 1].split(","))
        if a == "--out":
            out_name = sys.argv[i + 1]
    out_dir = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    report = {"device": torch.cuda.get_device_name(0), "torch": torch.__version__}
    import mmoe_multimodal_rec_b200 as pkg
    from mmoe_multimodal_rec_b200._lib import check
    check(pkg.lib().mmoe_init(), "init")
    for fn, args in ((gemm_checks, (report, quick)), (attention_checks, (report,)), (module_checks, (report, quick))):
        if only is not None and fn.__name__.split("_")[0] not in only:
            continue
        try:
            fn(*args)
        except Exception as ex:  # noqa: BLE001
            report[fn.__name__] = {"error": repr(ex), "trace": traceback.format_exc()[-2000:]}
        with open(os.path.join(out_dir, out_name), "w") as f:
            json.dump(report, f, indent=1, default=str)
    # console summary
    bad = 0
    for k, v in report.items():
        if not isinstance(v, dict):
            continue
        if "error" in v:
            print("ERR ", k, v["error"][:200])
            bad += 1
            continue
        mx = v.get("max", max([x for k2, x in v.items() if isinstance(x, float) and not k2.startswith("stat_")], default=0.0))
        tol = 2e-2 if ("bf16" in k or "fp16" in k) else 1e-4
        flag = "ok  " if mx <= tol else "BAD "
        bad += mx > tol
        if mx > tol or k.startswith("module") or k.startswith("golden"):
            print(flag, k, f"{mx:.3e}", v.get("worst", "")[:3] if isinstance(v.get("worst"), list) else "")
    print(f"{bad} problems; report in gpurun_out/{out_name}")


if __name__ == "__main__":
    main()
