"""GPU parity tests (run with -m gpu on a B200): the CUDA drop-in modules, called through the C ABI, against the
CPU oracle on the same deterministic weights/inputs, and against the golden vectors the real reference produced.

Tolerances (BASELINE.json north_star): max|a-b|/max|ref| <= 1e-4 in fp32, <= 2e-2 under bf16 / fp16 autocast."""
import pytest
import torch

import parity_util as PU
from conftest import load_golden
from oracle import cases as C

pytestmark = pytest.mark.gpu
IDS = [c.name for c in C.CASES]


def _assert_ok(errs, tol, what):
    bad = {k: v for k, v in errs.items() if not (v <= tol)}
    assert not bad, f"{what}: {len(bad)} of {len(errs)} tensors above {tol}: {sorted(bad.items(), key=lambda kv: -kv[1])[:6]}"


@pytest.mark.parametrize("case", C.CASES, ids=IDS)
def test_fp32_matches_oracle(case):
    stats = {}
    _assert_ok(PU.compare_with_oracle(case, "fp32", stats=stats), PU.TOL["fp32"], f"{case.name} fp32")
    if stats:
        # ReLU pattern of the kernels vs the fp64 one: may differ only in entries that are zero to fp32 accuracy
        assert stats["flip_frac"] < 1e-5 and stats["flip_max_rel_z"] < 1e-5, stats


@pytest.mark.parametrize("case", C.CASES, ids=IDS)
def test_fp32_matches_reference_golden(case):
    _assert_ok(PU.check_against_golden(case, load_golden(case.name), "fp32"), PU.TOL["fp32"], f"{case.name} golden")


@pytest.mark.parametrize("mode", ["bf16", "fp16"])
@pytest.mark.parametrize("case", C.CASES, ids=IDS)
def test_16bit_autocast_matches_oracle(case, mode):
    stats = {}
    _assert_ok(PU.compare_with_oracle(case, mode, stats=stats), PU.TOL[mode], f"{case.name} {mode}")
    if stats:
        # the kernels' ReLU activation pattern may differ from the fp64 one only within rounding noise of zero
        assert stats["flip_frac"] < 0.01 and stats["flip_max_rel_z"] < 0.05, stats


# Benchmark-sized batches: at B = 128 (README config, 8192 sentence rows) and B = 512 (train.py default,
# the batch bench.py times) the GEMM launcher resolves to the CTA-pair 256x256 kernel with the bit-mask FFN epilogues — the
# kernels the benchmark actually runs.  The float64 oracle runs on the GPU here (plain torch, same code as on the CPU).
BIG_CASES = [
    C.Case("cross_b128", "cross", 128, 131), C.Case("cross_home_b128", "cross_home", 128, 132),
    C.Case("fuse_b128", "fuse", 128, 141), C.Case("fuse_home_b128", "fuse_home", 128, 142),
    C.Case("home_head_b512", "home_head", 512, 121, dict(tower_hidden=512)),
    C.Case("head_b4096", "head", 4096, 112),
]


@pytest.mark.parametrize("mode", ["fp32", "bf16", "fp16"])
@pytest.mark.parametrize("case", BIG_CASES, ids=[c.name for c in BIG_CASES])
def test_benchmark_sized_batches_match_oracle(case, mode):
    import mmoe_multimodal_rec_b200 as pkg
    L = pkg.lib()
    L.mmoe_launch_trace(1)
    stats = {}
    errs = PU.compare_with_oracle(case, mode, stats=stats, device="cuda")
    n = L.mmoe_launch_trace_read(None, 0)
    L.mmoe_launch_trace(0)
    _assert_ok(errs, PU.TOL[mode], f"{case.name} {mode}")
    if stats:
        if mode == "fp32":     # fp32: the kernels' ReLU pattern may differ from the fp64 one only in entries that are ~0 in fp32
            assert stats["flip_frac"] < 1e-5 and stats["flip_max_rel_z"] < 1e-5, stats
        else:
            assert stats["flip_frac"] < 0.01 and stats["flip_max_rel_z"] < 0.05, stats
    if mode != "fp32":
        assert n > 0           # the tensor-core engine ran


def test_cross_expert_at_the_benchmark_batch_uses_the_pair_kernel_and_matches():
    """B = 512 under bf16 autocast: exactly bench.py's cross-expert call (eval arithmetic), M = 32768 rows."""
    import ctypes as Ct
    import mmoe_multimodal_rec_b200 as pkg
    L = pkg.lib()
    case = C.Case("cross_b512", "cross", 512, 231)
    L.mmoe_launch_trace(1)
    stats = {}
    errs = PU.compare_with_oracle(case, "bf16", stats=stats, device="cuda")
    n = L.mmoe_launch_trace_read(None, 0)
    buf = (Ct.c_int32 * (4 * n))()
    L.mmoe_launch_trace_read(buf, n)
    L.mmoe_launch_trace(0)
    pair = sum(1 for i in range(n) if buf[4 * i + 1] == 2)
    # 4 encoder layers x (4 forward + 4 backward) GEMM launches + the cross-attention projections (the [B, .] MLP stays small)
    assert pair >= 32, f"only {pair} of {n} GEMM launches used the CTA-pair kernel"
    _assert_ok(errs, PU.TOL["bf16"], "cross_b512 bf16")
    assert stats["flip_frac"] < 0.01 and stats["flip_max_rel_z"] < 0.05, stats


def test_16bit_forward_matches_fp64_oracle_without_any_injection():
    """Forward outputs have no discontinuity: compare straight against the fp64 oracle."""
    for name in ("cross_b3", "fuse_b8", "cross_home_b3", "fuse_home_b8"):
        case = C.CASES_BY_NAME[name]
        o_out, _, _ = C.run_oracle(case, torch.float64)
        c_out, _, _ = PU.run_cuda(case, "bf16")
        for a, b in zip(c_out, o_out):
            assert PU.nerr(a, b) <= PU.TOL["bf16"], name


@pytest.mark.parametrize("name", ["cross_b3", "fuse_b8", "head_b16"])
def test_bf16_is_as_close_to_fp64_as_eager_autocast(name):
    """besides "within 2e-2 of fp64", say how the kernels compare with what the reference itself
    computes under the same autocast — the stock torch.nn modules (oracle/eager_ref.py: cuBLASLt / SDPA / native
    LayerNorm) on this GPU, same weights and inputs, eval mode.  Outputs and input gradients of both are measured against
    the un-injected fp64 oracle; the native path may be at most 2.5x further from fp64 than eager (with a floor of a
    quarter of the 16-bit tolerance, below which both are at rounding noise)."""
    from oracle import eager_ref as E
    case = C.CASES_BY_NAME[name]
    make, fwd = {"cross": (E.make_cross_expert, E.cross_expert), "fuse": (E.make_cross_fuse, E.cross_fuse),
                 "head": (E.make_mmoe_head, lambda m, x: E.two_task_mmoe(m, x))}[case.kind]
    eager = make().cuda().eval()
    eager.load_state_dict(case.state_dict(), strict=True)
    meta = case.inputs_meta()
    ins = [t.cuda().clone().requires_grad_(True) if f else t.cuda() for t, f in zip(case.inputs(), meta)]
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = fwd(eager, *ins)
    outs = tuple(out) if isinstance(out, (tuple, list)) else (out,)
    cots = case.cotangents([o.detach().cpu() for o in outs])
    torch.autograd.backward(list(outs), [c.to(o.device, o.dtype) for c, o in zip(cots, outs)])
    e_out = [o.detach().float().cpu() for o in outs]
    e_gin = [t.grad.detach().float().cpu() if f else None for t, f in zip(ins, meta)]
    o_out, o_gin, _ = C.run_oracle(case, torch.float64)
    c_out, c_gin, _ = PU.run_cuda(case, "bf16")
    floor = 0.25 * PU.TOL["bf16"]
    report = {}
    for tag, ours, theirs, ref in ([(f"out{j}", a, b, r) for j, (a, b, r) in enumerate(zip(c_out, e_out, o_out))] +
                                   [(f"grad_in{j}", a, b, r) for j, (a, b, r) in enumerate(zip(c_gin, e_gin, o_gin)) if r is not None]):
        report[tag] = (PU.nerr(ours, ref), PU.nerr(theirs, ref))
    print(f"[eager-vs-native] {name}: " + ", ".join(f"{k} native {v[0]:.2e} / eager {v[1]:.2e}" for k, v in report.items()))
    bad = {k: v for k, v in report.items() if not v[0] <= max(2.5 * v[1], floor)}
    assert not bad, f"{name}: (native err, eager err) vs fp64: {report}"


def test_gate_argmax_identical_in_fp32():
    """north_star: gate argmax bit-identical on the fp32 path (wherever the fp64 top-2 margin exceeds 1e-5,
    i.e. beyond what any fp32 summation order can resolve — SURVEY.md §7.2)."""
    from oracle import mmoe_oracle as O
    case = C.CASES_BY_NAME["head_b256"]
    mod = PU.build_module(case)
    (ev,) = case.inputs()
    wg, wb = mod.gate_weights(ev.cuda())
    sd64 = {k: v.double() for k, v in case.state_dict().items()}
    _, _, rg, rb = O.two_task_mmoe(sd64, ev.double(), return_gates=True)
    for w, r in ((wg, rg), (wb, rb)):
        top = r.topk(2, -1).values
        safe = (top[:, 0] - top[:, 1]) > 1e-5
        assert safe.float().mean() > 0.99
        assert (w.cpu().argmax(-1)[safe] == r.argmax(-1)[safe]).all()
        assert float((w.cpu().double() - r).abs().max()) < 1e-6
    # the stand-alone DenseGate module agrees with the fused kernel
    w2 = mod.gate_good(ev.cuda().mean(1))
    assert float((w2 - wg).abs().max()) < 1e-6


def test_scoring_order_matches_in_fp32():
    """AUC ordering on the fp32 path: ranks of the logits equal the oracle's except for pairs closer than fp32 noise."""
    from oracle import mmoe_oracle as O
    case = C.Case("head_b4096", "head", 4096, 77)
    mod = PU.build_module(case)
    (ev,) = case.inputs()
    with torch.no_grad():
        lg, lb = mod(ev.cuda())
    sd64 = {k: v.double() for k, v in case.state_dict().items()}
    rg, rb = O.two_task_mmoe(sd64, ev.double())
    for a, r in ((lg, rg), (lb, rb)):
        a = a.cpu().double()
        eps = 1e-4 * float(r.abs().max())              # the fp32 tolerance of the north star
        assert float((a - r).abs().max()) <= eps
        # two samples can only change places if their true logits are closer than 2*eps: check it pairwise
        order_a = a.argsort()
        r_in_a_order = r[order_a]
        inversions = (r_in_a_order[1:] < r_in_a_order[:-1])
        if inversions.any():
            assert float((r_in_a_order[:-1] - r_in_a_order[1:])[inversions].max()) <= 2 * eps
        # and the AUC of the two scorings against any labelling agrees to 1e-6 (rank statistic)
        y = (torch.arange(r.numel()) % 3 == 0).double()
        def auc(s):
            ranks = torch.empty_like(s); ranks[s.argsort()] = torch.arange(1, s.numel() + 1, dtype=s.dtype)
            n1 = y.sum(); n0 = y.numel() - n1
            return float((ranks[y == 1].sum() - n1 * (n1 + 1) / 2) / (n0 * n1))
        assert abs(auc(a) - auc(r)) < 1e-6


@pytest.mark.parametrize("name", ["head_b16", "home_head_b8", "fuse_b8", "fuse_home_b8", "cross_b3", "cross_home_b3"])
@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_parameter_modes_agree(name, mode):
    """Per-tensor parameters (staged cross-expert backward, ~60 gradients handed to autograd) and the fused parameter
    (one backward node, the flat buffer is the gradient) run the same kernels: same outputs, same per-name gradients."""
    case = C.CASES_BY_NAME[name]
    per_tensor = PU.build_module(case, fused=False)
    fused = PU.build_module(case, fused=True)
    assert len(list(fused.parameters())) == 1 + len(set(case.shapes()) - set(case.used_param_keys()))
    assert list(fused.state_dict().keys()) == list(per_tensor.state_dict().keys())
    oa, gia, gpa = PU.run_cuda(case, mode, module=per_tensor)
    ob, gib, gpb = PU.run_cuda(case, mode, module=fused)
    for a, b in zip(oa, ob):
        assert torch.equal(a, b)
    for a, b in zip(gia, gib):
        assert (a is None) == (b is None) and (a is None or PU.nerr(b, a) <= 1e-6)
    assert list(gpa.keys()) == list(gpb.keys())
    for k in gpa:
        assert (gpa[k] is None) == (gpb[k] is None), k
        if gpa[k] is not None:
            assert PU.nerr(gpb[k], gpa[k]) <= 1e-5, k


def test_unused_home_parameters_keep_grad_none():
    for name in ("cross_home_b3", "fuse_home_b8"):
        case = C.CASES_BY_NAME[name]
        _, _, gp = PU.run_cuda(case, "fp32")
        unused = set(case.shapes()) - set(case.used_param_keys())
        assert unused
        for k in unused:
            assert gp[k] is None
        for k in case.used_param_keys():
            assert gp[k] is not None


def test_fully_masked_row_gives_nan_like_the_reference():
    case = C.CASES_BY_NAME["cross_b3"]
    mod = PU.build_module(case)
    u, um, i, im = case.inputs()
    im = im.clone()
    im[2] = True
    with torch.no_grad():
        out = mod(u.cuda(), um.cuda(), i.cuda(), im.cuda()).cpu()
    assert torch.isfinite(out[:2]).all() and torch.isnan(out[2]).all()


def test_no_cpu_fallback():
    case = C.CASES_BY_NAME["head_b16"]
    mod = PU.build_module(case, device="cuda")
    with pytest.raises(RuntimeError, match="CUDA"):
        mod(case.inputs()[0])          # CPU tensor -> loud failure, never a silent eager path


def test_empty_batch():
    case = C.CASES_BY_NAME["head_b16"]
    mod = PU.build_module(case)
    lg, lb = mod(torch.zeros(0, 6, 768, device="cuda"))
    assert lg.shape == (0,) and lb.shape == (0,)


def test_train_mode_dropout_statistics_and_determinism():
    """Train mode: dropout masks come from a keyed counter hash; same torch seed -> same result, the keep rate is 1-p,
    and the backward pass regenerates the same mask (gradient of a dropped unit is exactly zero)."""
    import ctypes as Ct
    import mmoe_multimodal_rec_b200 as pkg
    L = pkg.lib()
    n = 1 << 20
    out = torch.empty(n, dtype=torch.uint8, device="cuda")
    assert L.mmoe_dropout_mask(123, 456, 0.1, n, out.data_ptr(), torch.cuda.current_stream().cuda_stream) == 0
    keep = out.float().mean().item()
    assert abs(keep - 0.9) < 2e-3
    case = C.CASES_BY_NAME["fuse_b8"]
    mod = PU.build_module(case).train()
    v, t = [x.cuda() for x in case.inputs()]
    torch.manual_seed(5); a = mod(v, t)
    torch.manual_seed(5); b = mod(v, t)
    torch.manual_seed(6); c = mod(v, t)
    assert torch.equal(a, b) and not torch.equal(a, c)
    frac_zero = (a == 0).float().mean().item()          # final Dropout(0.1) of proj
    assert 0.05 < frac_zero < 0.16
    v.requires_grad_(True)
    torch.manual_seed(5)
    out2 = mod(v, t)
    out2.sum().backward()
    assert torch.isfinite(v.grad).all()
