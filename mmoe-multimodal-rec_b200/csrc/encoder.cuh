// Pre-LN transformer encoder layer (RobustTransformerLayer, model.py:207-212; the stock
// nn.TransformerEncoderLayer(norm_first=True) of EnhancedCrossFuse, model.py:460-465),
// forward and backward, as a sequence of fused launches:
//
//   forward  : [residual add +] LN1 -> [QKV GEMM +bias] -> attention -> [out-proj GEMM +bias +dropout]
//              -> [residual add + LN2] -> [FFN1 GEMM +bias +ReLU +dropout] -> [FFN2 GEMM +bias +dropout]
//              Every large GEMM writes a 16-bit tile through the TMA-store epilogue; the residual adds ride on the
//              LayerNorm kernels (x = x + delta is exactly what autocast computes: a 16-bit linear output added to the
//              fp32 stream).  The layer's output x1 + y2 is materialised by its consumer (next LN1 / add_cast).
//   backward : cast/drop(+db2) -> {dgrad FFN2 | wgrad W2} -> [ReLU/drop mask + db1] -> {dgrad FFN1 | wgrad W1}
//              -> LN2' (+residual, +dropout1 mask, +d b_out) -> {dgrad out-proj | wgrad Wo}
//              -> attention' (+d b_in) -> {dgrad QKV | wgrad W_in} -> LN1' (+residual)
//   "{a | b}" = one grouped GEMM launch.
#pragma once
#include "kernels.cuh"

namespace mmoe {

// Bump allocator over a caller-owned blob.  With a null base it only measures: the "pointers" it returns are then
// the byte offsets themselves (used by the *_bytes size queries and the *_saved_offset layout queries).
struct Arena {
  uintptr_t base;
  size_t off;
  explicit Arena(void* b) : base(reinterpret_cast<uintptr_t>(b)), off(0) {}
  void* take(size_t bytes) {
    off = (off + 255) & ~(size_t)255;
    void* p = reinterpret_cast<void*>(base + off);
    off += bytes;
    return p;
  }
};

// parameters / gradients of one encoder layer, in state_dict order (12 tensors)
struct EncW {
  const void* w_in; const float* b_in; const void* w_out; const float* b_out;
  const void* w1; const float* b1; const void* w2; const float* b2;
  const float *ln1_w, *ln1_b, *ln2_w, *ln2_b;
};
struct EncG {
  float *w_in, *b_in, *w_out, *b_out, *w1, *b1, *w2, *b2, *ln1_w, *ln1_b, *ln2_w, *ln2_b;
};
inline EncW enc_w(const void* const* p) {
  EncW w;
  w.w_in = p[0]; w.b_in = (const float*)p[1]; w.w_out = p[2]; w.b_out = (const float*)p[3];
  w.w1 = p[4]; w.b1 = (const float*)p[5]; w.w2 = p[6]; w.b2 = (const float*)p[7];
  w.ln1_w = (const float*)p[8]; w.ln1_b = (const float*)p[9]; w.ln2_w = (const float*)p[10]; w.ln2_b = (const float*)p[11];
  return w;
}
inline EncG enc_g(void* const* p) {
  EncG g;
  g.w_in = (float*)p[0]; g.b_in = (float*)p[1]; g.w_out = (float*)p[2]; g.b_out = (float*)p[3];
  g.w1 = (float*)p[4]; g.b1 = (float*)p[5]; g.w2 = (float*)p[6]; g.b2 = (float*)p[7];
  g.ln1_w = (float*)p[8]; g.ln1_b = (float*)p[9]; g.ln2_w = (float*)p[10]; g.ln2_b = (float*)p[11];
  return g;
}

// activations kept for backward
struct EncSaved {
  float* x_sum;   // layer input when it arrived as (x_prev + delta_prev); unused for a layer fed directly
  void* xn1; float* st1; void* qkv; void* ctx; void* y1; float* x1; void* xn2; float* st2; void* h; void* y2;
  uint64_t* hbits;   // [M, ff/64]: bit = (h != 0), the ReLU/dropout pattern of the FFN hidden layer (16-bit modes)
};
inline EncSaved enc_layout(Arena& A, int64_t M, int d, int ff, size_t es) {
  EncSaved s;
  s.x_sum = (float*)A.take(M * d * sizeof(float));
  s.xn1 = A.take(M * d * es);
  s.st1 = (float*)A.take(M * 2 * sizeof(float));
  s.qkv = A.take(M * 3 * d * es);
  s.ctx = A.take(M * d * es);
  s.y1 = A.take(M * d * es);
  s.x1 = (float*)A.take(M * d * sizeof(float));
  s.xn2 = A.take(M * d * es);
  s.st2 = (float*)A.take(M * 2 * sizeof(float));
  s.h = A.take(M * ff * es);
  s.y2 = A.take(M * d * es);
  s.hbits = (uint64_t*)A.take(M * ((ff + 63) / 64) * sizeof(uint64_t));
  return s;
}
// scratch of the backward pass
struct EncScratch {
  void* g; void* dh; void* dxn; float* dx1; void* dqkv; void* dctx;
};
inline EncScratch enc_scratch_layout(Arena& A, int64_t M, int d, int ff, size_t es) {
  EncScratch s;
  s.g = A.take(M * d * es);
  s.dh = A.take(M * ff * es);
  s.dxn = A.take(M * d * es);
  s.dx1 = (float*)A.take(M * d * sizeof(float));
  s.dqkv = A.take(M * 3 * d * es);
  s.dctx = A.take(M * d * es);
  return s;
}

struct EncCtx {
  int dtype; int64_t M; int Bseq, S, d, ff, H;
  const uint8_t* mask;        // [Bseq,S] or null
  float drop_p;               // 0 in eval
  uint64_t seed; uint32_t site0;   // 4 dropout sites: site0 + {0 probs, 1 dropout1, 2 ffn, 3 dropout2}
  cudaStream_t stream;
};

// y = epilogue(x W^T + b): nn.Linear forward
inline mmoe_gemm_problem linear_fwd(const void* x, int64_t ldx, const void* w, int M, int N, int K, const mmoe_epilogue& e) {
  return gemm_problem(x, ldx, 0, w, K, 0, M, N, K, e);
}
// dx[M,K] = epilogue(dy[M,N] W[N,K])
inline mmoe_gemm_problem linear_dgrad(const void* dy, int64_t lddy, const void* w, int M, int N, int K, const mmoe_epilogue& e) {
  return gemm_problem(dy, lddy, 0, w, K, 1, M, K, N, e);
}
// dW[N,K] += dy[M,N]^T x[M,K]   (fp32 atomic accumulation, split over M)
inline mmoe_gemm_problem linear_wgrad(const void* dy, int64_t lddy, const void* x, int64_t ldx, float* dw, int M, int N, int K) {
  mmoe_epilogue e = epi_none();
  e.out = dw; e.out_dtype = MMOE_F32; e.ldo = K; e.accumulate = 1;
  return gemm_problem(dy, lddy, 1, x, ldx, 1, N, K, M, e, 0 /* split count chosen by the GEMM launcher */);
}

// x_prev (+ delta_prev, 16-bit, may be null) is the layer input.  On return the layer output is s.x1 + s.y2 (pending).
inline int enc_fwd(const EncCtx& c, const EncW& w, const float* x_prev, const void* delta_prev, const EncSaved& s) {
  const int d = c.d, ff = c.ff; const int M = (int)c.M;
  const size_t es = dtype_size(c.dtype);
  uint32_t k0, k1;
  const float* x_in = x_prev;
  if (delta_prev != nullptr) {
    MMOE_TRY(layernorm_fwd_add(x_prev, delta_prev, s.x_sum, w.ln1_w, w.ln1_b, s.xn1, s.st1, M, d, c.dtype, c.stream));
    x_in = s.x_sum;
  } else {
    MMOE_TRY(layernorm_fwd(x_prev, MMOE_F32, w.ln1_w, w.ln1_b, s.xn1, nullptr, s.st1, M, d, c.dtype, c.stream));
  }
  {
    mmoe_epilogue e = epi_none();
    e.out = s.qkv; e.out_dtype = c.dtype; e.ldo = 3 * d; e.bias = w.b_in;
    mmoe_gemm_problem p = linear_fwd(s.xn1, d, w.w_in, M, 3 * d, d, e);
    MMOE_TRY(gemm_grouped(&p, 1, c.dtype, 0, c.stream));
  }
  {
    AttnArgs a{};
    a.q = s.qkv; a.k = (const char*)s.qkv + (size_t)d * es; a.v = (const char*)s.qkv + (size_t)2 * d * es;
    a.ldq = a.ldk = a.ldv = 3 * d; a.mask = c.mask; a.ctx = s.ctx; a.ldc = d;
    a.B = c.Bseq; a.Sq = c.S; a.Sk = c.S; a.H = c.H; a.hd = d / c.H;
    site_keys(c.seed, c.site0 + 0, &k0, &k1);
    a.drop_p = c.drop_p; a.k0 = k0; a.k1 = k1; a.dtype = c.dtype;
    MMOE_TRY(attention_fwd(a, c.stream));
  }
  {
    mmoe_epilogue e = epi_none();       // y1 = drop1(ctx Wo^T + bo)
    e.out = s.y1; e.out_dtype = c.dtype; e.ldo = d; e.bias = w.b_out;
    site_keys(c.seed, c.site0 + 1, &k0, &k1);
    e.drop_p = c.drop_p; e.drop_key0 = k0; e.drop_key1 = k1;
    mmoe_gemm_problem p = linear_fwd(s.ctx, d, w.w_out, M, d, d, e);
    MMOE_TRY(gemm_grouped(&p, 1, c.dtype, 0, c.stream));
  }
  // x1 = x_in + y1 ; xn2 = LN2(x1)
  MMOE_TRY(layernorm_fwd_add(x_in, s.y1, s.x1, w.ln2_w, w.ln2_b, s.xn2, s.st2, M, d, c.dtype, c.stream));
  {
    mmoe_epilogue e = epi_none();
    e.out = s.h; e.out_dtype = c.dtype; e.ldo = ff; e.bias = w.b1; e.act = 1;
    if (gemm_bitmask_supported(c.dtype, M, ff)) e.mask_out = s.hbits;      // 1 bit per element for the backward
    site_keys(c.seed, c.site0 + 2, &k0, &k1);
    e.drop_p = c.drop_p; e.drop_key0 = k0; e.drop_key1 = k1;
    mmoe_gemm_problem p = linear_fwd(s.xn2, d, w.w1, M, ff, d, e);
    MMOE_TRY(gemm_grouped(&p, 1, c.dtype, 0, c.stream));
  }
  {
    mmoe_epilogue e = epi_none();       // y2 = drop2(h W2^T + b2)
    e.out = s.y2; e.out_dtype = c.dtype; e.ldo = d; e.bias = w.b2;
    site_keys(c.seed, c.site0 + 3, &k0, &k1);
    e.drop_p = c.drop_p; e.drop_key0 = k0; e.drop_key1 = k1;
    mmoe_gemm_problem p = linear_fwd(s.h, ff, w.w2, M, d, ff, e);
    MMOE_TRY(gemm_grouped(&p, 1, c.dtype, 0, c.stream));
  }
  return 0;
}

// Hand-off between stacked layers: the layer ABOVE finishes with a LayerNorm backward that produces this layer's output
// gradient; it can write the 16-bit, dropout-masked copy the FFN2 backward needs (and d b2) in the same pass.
struct EncHandoff {
  void* g16;            // [M,d] 16-bit: drop2-masked output gradient of the layer below
  float* b2_grad;       // its FFN2 bias gradient (column sums of g16)
  uint32_t site0;       // dropout site base of the layer below
};

// dy (+ dy_t): gradient w.r.t. the layer output — fp32 [M,d] and/or a 16-bit [M,d] summand (a dgrad GEMM's output taken
// straight from its fast 16-bit epilogue instead of an fp32 read-modify-write); dx: gradient w.r.t. the layer input (fp32,
// may alias dy).  dy16: if non-null, the layer above already produced cast_drop(dy) and d b2 (EncHandoff).  below: if
// non-null, do the same for the layer below.
inline int enc_bwd(const EncCtx& c, const EncW& w, const EncG& g, const float* x_in, const EncSaved& s, const EncScratch& t,
                   const float* dy, float* dx, const void* dy16 = nullptr, const EncHandoff* below = nullptr,
                   const void* dy_t = nullptr) {
  const int d = c.d, ff = c.ff; const int M = (int)c.M;
  const size_t es = dtype_size(c.dtype);
  uint32_t k0, k1;
  // FFN2: x2 = x1 + drop2(h W2^T + b2)
  const void* g2 = dy16;
  if (g2 == nullptr) {
    site_keys(c.seed, c.site0 + 3, &k0, &k1);
    MMOE_TRY(cast_drop_colsum(dy, t.g, g.b2, M, d, c.drop_p, k0, k1, c.dtype, c.stream, dy_t));
    g2 = t.g;
  }
  const bool bits = gemm_bitmask_supported(c.dtype, M, ff);
  {
    // ReLU (+dropout) backward: the kept/active pattern is h != 0.  16-bit modes: the dgrad epilogue applies the bit
    // pattern the forward FFN1 epilogue wrote and sums d b1 (K = 3072 leaves that epilogue plenty of slack).
    mmoe_epilogue e = epi_none();
    e.out = t.dh; e.out_dtype = c.dtype; e.ldo = ff;
    if (bits) { e.bwd_mode = 4; e.aux = s.hbits; e.drop_p = c.drop_p; e.colsum = g.b1; }
    mmoe_gemm_problem p[2] = {linear_dgrad(g2, d, w.w2, M, d, ff, e), linear_wgrad(g2, d, s.h, ff, g.w2, M, d, ff)};
    MMOE_TRY(gemm_grouped(p, 2, c.dtype, 0, c.stream));
  }
  if (!bits)
    MMOE_TRY(relu_mask_colsum(t.dh, s.h, g.b1, M, ff, c.drop_p > 0.f ? 1.f / (1.f - c.drop_p) : 1.f, c.dtype, c.stream));
  // FFN1: h = drop(relu(xn2 W1^T + b1))
  {
    mmoe_epilogue e = epi_none();
    e.out = t.dxn; e.out_dtype = c.dtype; e.ldo = d;
    mmoe_gemm_problem p[2] = {linear_dgrad(t.dh, ff, w.w1, M, ff, d, e), linear_wgrad(t.dh, ff, s.xn2, d, g.w1, M, ff, d)};
    MMOE_TRY(gemm_grouped(p, 2, c.dtype, 0, c.stream));
  }
  // LN2 and the residual: dx1 = dy + LN2'(dxn);  g = drop1-mask(dx1) for the out-proj backward (+ d b_out)
  {
    LnBwdArgs a{};
    a.dy = t.dxn; a.dy_dtype = c.dtype; a.x = s.x1; a.x_dtype = MMOE_F32; a.stats = s.st2; a.gamma = w.ln2_w;
    a.dres = dy; a.dres_t = dy_t; a.dx = t.dx1; a.dgamma = g.ln2_w; a.dbeta = g.ln2_b; a.g_out = t.g; a.g_colsum = g.b_out;
    site_keys(c.seed, c.site0 + 1, &k0, &k1);
    a.drop_p = c.drop_p; a.k0 = k0; a.k1 = k1; a.rows = M; a.d = d; a.dtype = c.dtype;
    MMOE_TRY(layernorm_bwd(a, c.stream));
  }
  // out-proj: x1 = x_in + drop1(ctx Wo^T + bo)
  {
    mmoe_epilogue e = epi_none();
    e.out = t.dctx; e.out_dtype = c.dtype; e.ldo = d;
    mmoe_gemm_problem p[2] = {linear_dgrad(t.g, d, w.w_out, M, d, d, e), linear_wgrad(t.g, d, s.ctx, d, g.w_out, M, d, d)};
    MMOE_TRY(gemm_grouped(p, 2, c.dtype, 0, c.stream));
  }
  {
    AttnArgs a{};
    a.q = s.qkv; a.k = (const char*)s.qkv + (size_t)d * es; a.v = (const char*)s.qkv + (size_t)2 * d * es;
    a.ldq = a.ldk = a.ldv = 3 * d; a.mask = c.mask; a.ctx = t.dctx; a.ldc = d;
    a.dq = t.dqkv; a.dk = (char*)t.dqkv + (size_t)d * es; a.dv = (char*)t.dqkv + (size_t)2 * d * es;
    a.bgq = g.b_in; a.bgk = g.b_in + d; a.bgv = g.b_in + 2 * d;
    a.B = c.Bseq; a.Sq = c.S; a.Sk = c.S; a.H = c.H; a.hd = d / c.H;
    site_keys(c.seed, c.site0 + 0, &k0, &k1);
    a.drop_p = c.drop_p; a.k0 = k0; a.k1 = k1; a.dtype = c.dtype;
    MMOE_TRY(attention_bwd(a, c.stream));
  }
  {
    mmoe_epilogue e = epi_none();
    e.out = t.dxn; e.out_dtype = c.dtype; e.ldo = d;
    mmoe_gemm_problem p[2] = {linear_dgrad(t.dqkv, 3 * d, w.w_in, M, 3 * d, d, e),
                              linear_wgrad(t.dqkv, 3 * d, s.xn1, d, g.w_in, M, 3 * d, d)};
    MMOE_TRY(gemm_grouped(p, 2, c.dtype, 0, c.stream));
  }
  {
    LnBwdArgs a{};
    a.dy = t.dxn; a.dy_dtype = c.dtype; a.x = x_in; a.x_dtype = MMOE_F32; a.stats = s.st1; a.gamma = w.ln1_w;
    a.dres = t.dx1; a.dx = dx; a.dgamma = g.ln1_w; a.dbeta = g.ln1_b;
    if (below != nullptr) {
      a.g_out = below->g16; a.g_colsum = below->b2_grad;
      site_keys(c.seed, below->site0 + 3, &k0, &k1);
      a.drop_p = c.drop_p; a.k0 = k0; a.k1 = k1;
    }
    a.rows = M; a.d = d; a.dtype = c.dtype;
    MMOE_TRY(layernorm_bwd(a, c.stream));
  }
  return 0;
}

}  // namespace mmoe
