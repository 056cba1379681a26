"""Launch a few representative GEMM configurations of the encoder layer (for ncu): python tools/prof_gemm.py [reps]"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import mmoe_multimodal_rec_b200 as pkg  # noqa: E402
from mmoe_multimodal_rec_b200._lib import GemmProblem, check  # noqa: E402

L = pkg.lib()
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
M = int(os.environ.get("PROF_M", "32768"))
dev = "cuda"
st = torch.cuda.current_stream().cuda_stream


def run(name, N, K, variant):
    A = torch.randn(M, K, device=dev).bfloat16()
    W = torch.randn(N, K, device=dev).bfloat16()
    bias = torch.randn(N, device=dev)
    p = GemmProblem()
    p.a, p.lda, p.a_major = A.data_ptr(), K, 0
    p.b, p.ldb, p.b_major = W.data_ptr(), K, 0
    p.M, p.N, p.K, p.k_splits = M, N, K, 1
    e = p.epi
    e.alpha = 1.0
    e.bias = bias.data_ptr()
    keep = [A, W, bias]
    if variant == "plain":
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        e.out, e.out_dtype, e.ldo = out.data_ptr(), 1, N
    elif variant == "relu_drop":
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        e.out, e.out_dtype, e.ldo, e.act = out.data_ptr(), 1, N, 1
        e.drop_p, e.drop_key0, e.drop_key1 = 0.1, 123, 456
    elif variant == "relu_drop_mask":
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        bits = torch.empty(M, N // 64, device=dev, dtype=torch.int64)
        e.out, e.out_dtype, e.ldo, e.act, e.mask_out = out.data_ptr(), 1, N, 1, bits.data_ptr()
        e.drop_p, e.drop_key0, e.drop_key1 = 0.1, 123, 456
        keep.append(bits)
    elif variant == "resid_drop":
        out = torch.empty(M, N, device=dev)
        res = torch.randn(M, N, device=dev)
        e.out, e.out_dtype, e.ldo, e.residual, e.ld_res = out.data_ptr(), 0, N, res.data_ptr(), N
        e.drop_p, e.drop_key0, e.drop_key1 = 0.1, 123, 456
        keep.append(res)
    keep.append(out)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    check(L.mmoe_gemm_grouped(C.byref(p), 1, 1, 0, st), "gemm")
    torch.cuda.synchronize()
    ev0.record()
    for _ in range(reps):
        check(L.mmoe_gemm_grouped(C.byref(p), 1, 1, 0, st), "gemm")
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / reps
    print(f"{name:12s} N={N} K={K} {variant:10s} {ms*1e3:8.1f} us  {2.0*M*N*K/ms/1e9:7.1f} TFLOP/s", flush=True)


def run_bwd_pair(name, N_out, K_in):
    """{dgrad: dX[M,K_in] = dY[M,N_out] W[N_out,K_in] | wgrad: dW[N_out,K_in] += dY^T X} as ONE grouped launch (encoder.cuh)"""
    dY = torch.randn(M, N_out, device=dev).bfloat16()
    W = torch.randn(N_out, K_in, device=dev).bfloat16()
    X = torch.randn(M, K_in, device=dev).bfloat16()
    dX = torch.empty(M, K_in, device=dev, dtype=torch.bfloat16)
    dW = torch.zeros(N_out, K_in, device=dev)
    P = (GemmProblem * 2)()
    d, w = P[0], P[1]
    d.a, d.lda, d.a_major = dY.data_ptr(), N_out, 0
    d.b, d.ldb, d.b_major = W.data_ptr(), K_in, 1
    d.M, d.N, d.K, d.k_splits = M, K_in, N_out, 1
    d.epi.alpha, d.epi.out, d.epi.out_dtype, d.epi.ldo = 1.0, dX.data_ptr(), 1, K_in
    w.a, w.lda, w.a_major = dY.data_ptr(), N_out, 1
    w.b, w.ldb, w.b_major = X.data_ptr(), K_in, 1
    w.M, w.N, w.K = N_out, K_in, M
    w.k_splits = int(os.environ.get("PROF_SPLITS", "0"))                    # 0 = chosen by the launcher
    w.epi.alpha, w.epi.out, w.epi.out_dtype, w.epi.ldo, w.epi.accumulate = 1.0, dW.data_ptr(), 0, K_in, 1
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    check(L.mmoe_gemm_grouped(P, 2, 1, 0, st), "gemm")
    torch.cuda.synchronize()
    ev0.record()
    for _ in range(reps):
        check(L.mmoe_gemm_grouped(P, 2, 1, 0, st), "gemm")
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / reps
    print(f"{name:12s} dgrad|wgrad N_out={N_out} K_in={K_in} splits={w.k_splits} {ms*1e3:8.1f} us  {4.0*M*N_out*K_in/ms/1e9:7.1f} TFLOP/s", flush=True)


run("qkv", 2304, 768, "plain")
run("out_proj", 768, 768, "plain")
run("ffn1", 3072, 768, "relu_drop")
run("ffn1+bits", 3072, 768, "relu_drop_mask")
run("plain_k3072", 768, 3072, "plain")
run_bwd_pair("ffn2_bwd", 768, 3072)
run_bwd_pair("ffn1_bwd", 3072, 768)
run_bwd_pair("out_bwd", 768, 768)
