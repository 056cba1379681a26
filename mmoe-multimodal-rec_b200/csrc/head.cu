// TwoTaskMMoE — model.py:527-577.
// forward : [mean query + 2 gate GEMVs + softmax + 2 weighted expert sums] (one kernel, one pass over
//           expert_vecs) -> LN x2 -> {tower1_good | tower1_best} GEMM +GELU -> {tower2_good | tower2_best}
//           GEMM +GELU -> final 128->1 layer.
// backward: final layer' (+GELU') -> {dgrad | wgrad} x2 tasks per tower layer -> LN' -> gate/mix' -> gate wgrad.
#include "head_kernels.cuh"

namespace mmoe {


// ---------------------------------------------------------------- stand-alone DenseGate (model.py:522-524)
// w = softmax(x Wg^T + bg): one warp per row.  Backward: dl = w*(dw - <w,dw>), dx = dl Wg.
__global__ void __launch_bounds__(256) dense_gate_fwd_kernel(const float* __restrict__ x, const float* __restrict__ wg,
                                                             const float* __restrict__ bg, float* __restrict__ out, int64_t B, int d, int n) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t b = warp0; b < B; b += nw) {
    float logit[MIX_MAXN];
    float m = -INFINITY;
#pragma unroll
    for (int k = 0; k < MIX_MAXN; ++k) {
      logit[k] = -INFINITY;
      if (k < n) {
        float acc = 0.f;
        for (int c = lane; c < d; c += 32) acc = fmaf(x[b * d + c], wg[k * d + c], acc);
        logit[k] = warp_sum(acc) + bg[k];
        m = fmaxf(m, logit[k]);
      }
    }
    float ssum = 0.f;
#pragma unroll
    for (int k = 0; k < MIX_MAXN; ++k) { logit[k] = k < n ? expf(logit[k] - m) : 0.f; ssum += logit[k]; }
    if (lane == 0)
      for (int k = 0; k < n; ++k) out[b * n + k] = logit[k] / ssum;
  }
}
__global__ void __launch_bounds__(256) dense_gate_bwd_kernel(const float* __restrict__ w, const float* __restrict__ dw,
                                                             const float* __restrict__ wg, float* __restrict__ dl,
                                                             float* __restrict__ dx, int64_t B, int d, int n) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t b = warp0; b < B; b += nw) {
    float g[MIX_MAXN];
    float dot = 0.f;
    for (int k = 0; k < n; ++k) dot += w[b * n + k] * dw[b * n + k];
#pragma unroll
    for (int k = 0; k < MIX_MAXN; ++k) {
      g[k] = 0.f;
      if (k < n) { g[k] = w[b * n + k] * (dw[b * n + k] - dot); if (lane == 0) dl[b * n + k] = g[k]; }
    }
    for (int c = lane; c < d; c += 32) {
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < MIX_MAXN; ++k) if (k < n) acc = fmaf(g[k], wg[k * d + c], acc);
      dx[b * d + c] = acc;
    }
  }
}

struct HeadIdx { int gw[2], gb[2], ln_w[2], ln_b[2], w1[2], b1[2], w2[2], b2[2], w3[2], b3[2]; };
static HeadIdx head_idx() {
  HeadIdx i;
  i.gw[0] = 0; i.gb[0] = 1; i.gw[1] = 2; i.gb[1] = 3;
  for (int t = 0; t < 2; ++t) {
    const int p = 4 + 8 * t;
    i.ln_w[t] = p; i.ln_b[t] = p + 1; i.w1[t] = p + 2; i.b1[t] = p + 3; i.w2[t] = p + 4; i.b2[t] = p + 5; i.w3[t] = p + 6; i.b3[t] = p + 7;
  }
  return i;
}

struct HeadSaved { float *fused, *query, *w, *st; void *xn, *z1, *a1, *z2, *a2; };
static HeadSaved head_layout(Arena& A, const mmoe_head_cfg& cfg, int B, int dtype) {
  HeadSaved s{};
  const size_t es = dtype_size(dtype), Bz = (size_t)B; const int d = cfg.d, h1 = cfg.hidden, h2 = cfg.hidden / 2;
  s.fused = (float*)A.take(2 * Bz * d * 4); s.query = (float*)A.take(Bz * d * 4); s.w = (float*)A.take(2 * Bz * cfg.n_expert * 4);
  s.st = (float*)A.take(2 * Bz * 2 * 4); s.xn = A.take(2 * Bz * d * es);
  s.z1 = A.take(2 * Bz * h1 * es); s.a1 = A.take(2 * Bz * h1 * es); s.z2 = A.take(2 * Bz * h2 * es); s.a2 = A.take(2 * Bz * h2 * es);
  return s;
}
struct HeadScratch { void *dz2, *dz1, *dxn; float *dfused, *dl; };
static HeadScratch head_scratch_layout(Arena& A, const mmoe_head_cfg& cfg, int B, int dtype) {
  HeadScratch t{};
  const size_t es = dtype_size(dtype), Bz = (size_t)B; const int d = cfg.d, h1 = cfg.hidden, h2 = cfg.hidden / 2;
  t.dz2 = A.take(2 * Bz * h2 * es); t.dz1 = A.take(2 * Bz * h1 * es); t.dxn = A.take(2 * Bz * d * es);
  t.dfused = (float*)A.take(2 * Bz * d * 4); t.dl = (float*)A.take(2 * Bz * cfg.n_expert * 4);
  return t;
}
static int check_head(const mmoe_head_cfg* cfg) {
  MMOE_CHECK(cfg->d % 8 == 0 && cfg->d <= 1024, "head: unsupported expert_dim %d", cfg->d);
  MMOE_CHECK(cfg->n_expert >= 1 && cfg->n_expert <= MIX_MAXN, "head: n_expert must be in [1,%d]", MIX_MAXN);
  MMOE_CHECK(cfg->hidden % 16 == 0 && cfg->hidden <= 1024, "head: unsupported tower_hidden %d", cfg->hidden);
  return 0;
}

template <typename T>
static int head_fwd_t(const mmoe_call* c, const mmoe_head_cfg* cfg, const float* ev, float* logits, float* gate_w) {
  const int B = c->B, d = cfg->d, n = cfg->n_expert, h1 = cfg->hidden, h2 = cfg->hidden / 2, dtype = c->dtype;
  const size_t es = dtype_size(dtype);
  cudaStream_t st = (cudaStream_t)c->stream;
  Arena A(c->saved);
  HeadSaved s = head_layout(A, *cfg, B, dtype);
  const HeadIdx ix = head_idx();
  const void* const* P = c->params;
  const float drop_p = c->training ? cfg->tower_drop_p : 0.f;
  uint32_t k0, k1;
  {
    MixDev a{};
    a.experts = ev; a.expert_stride = d; a.row_stride = (int64_t)n * d;
    for (int t = 0; t < 2; ++t) for (int k = 0; k < n; ++k) a.sel[t][k] = k;
    a.wg[0] = (const float*)P[ix.gw[0]]; a.wg[1] = (const float*)P[ix.gw[1]];
    a.bg[0] = (const float*)P[ix.gb[0]]; a.bg[1] = (const float*)P[ix.gb[1]];
    a.fused = s.fused; a.query_out = s.query; a.w = s.w; a.B = B; a.d = d; a.n = n;
    mix_fwd_kernel<<<rows_grid(B, 8), 256, 0, st>>>(a);
    MMOE_LAUNCH_OK("mix_fwd_kernel");
    if (gate_w != nullptr) MMOE_CUDA(cudaMemcpyAsync(gate_w, s.w, (size_t)2 * B * n * 4, cudaMemcpyDeviceToDevice, st));
  }
  for (int t = 0; t < 2; ++t)
    MMOE_TRY(layernorm_fwd(s.fused + (size_t)t * B * d, MMOE_F32, (const float*)P[ix.ln_w[t]], (const float*)P[ix.ln_b[t]],
                           (char*)s.xn + (size_t)t * B * d * es, nullptr, s.st + (size_t)t * B * 2, B, d, dtype, st));
  {
    mmoe_gemm_problem p[2];
    for (int t = 0; t < 2; ++t) {
      mmoe_epilogue e = epi_none();
      e.out = (char*)s.a1 + (size_t)t * B * h1 * es; e.preact = (char*)s.z1 + (size_t)t * B * h1 * es;
      e.out_dtype = dtype; e.ldo = h1; e.bias = (const float*)P[ix.b1[t]]; e.act = 2;
      site_keys(c->seed, 10 + t, &k0, &k1);
      e.drop_p = drop_p; e.drop_key0 = k0; e.drop_key1 = k1;
      p[t] = linear_fwd((char*)s.xn + (size_t)t * B * d * es, d, P[ix.w1[t]], B, h1, d, e);
    }
    MMOE_TRY(gemm_grouped(p, 2, dtype, 0, st));
  }
  {
    mmoe_gemm_problem p[2];
    for (int t = 0; t < 2; ++t) {
      mmoe_epilogue e = epi_none();
      e.out = (char*)s.a2 + (size_t)t * B * h2 * es; e.preact = (char*)s.z2 + (size_t)t * B * h2 * es;
      e.out_dtype = dtype; e.ldo = h2; e.bias = (const float*)P[ix.b2[t]]; e.act = 2;
      site_keys(c->seed, 20 + t, &k0, &k1);
      e.drop_p = drop_p; e.drop_key0 = k0; e.drop_key1 = k1;
      p[t] = linear_fwd((char*)s.a1 + (size_t)t * B * h1 * es, h1, P[ix.w2[t]], B, h2, h1, e);
    }
    MMOE_TRY(gemm_grouped(p, 2, dtype, 0, st));
  }
  for (int t = 0; t < 2; ++t) {
    gemv_fwd_kernel<T><<<rows_grid(B, 8), 256, 0, st>>>((const T*)s.a2 + (size_t)t * B * h2, (const float*)P[ix.w3[t]],
                                                        (const float*)P[ix.b3[t]], logits + (size_t)t * B, B, h2);
    MMOE_LAUNCH_OK("gemv_fwd_kernel");
  }
  return 0;
}

template <typename T>
static int head_bwd_t(const mmoe_call* c, const mmoe_head_cfg* cfg, const float* ev, const float* dlogits, float* d_ev) {
  const int B = c->B, d = cfg->d, n = cfg->n_expert, h1 = cfg->hidden, h2 = cfg->hidden / 2, dtype = c->dtype;
  const size_t es = dtype_size(dtype);
  cudaStream_t st = (cudaStream_t)c->stream;
  Arena A(c->saved);
  HeadSaved s = head_layout(A, *cfg, B, dtype);
  Arena W(c->workspace);
  HeadScratch t = head_scratch_layout(W, *cfg, B, dtype);
  const HeadIdx ix = head_idx();
  const void* const* P = c->params;
  void* const* G = c->grads;
  const float drop_p = c->training ? cfg->tower_drop_p : 0.f;
  const uint32_t thresh = drop_p > 0.f ? drop_threshold(drop_p) : 0u;
  const float scale = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  uint32_t k0, k1;
  for (int k = 0; k < 2; ++k) {
    site_keys(c->seed, 20 + k, &k0, &k1);
    gemv_bwd_kernel<T><<<rows_grid(B, 2), 256, 0, st>>>(dlogits + (size_t)k * B, (const T*)s.a2 + (size_t)k * B * h2,
                                                        (const T*)s.z2 + (size_t)k * B * h2, (const float*)P[ix.w3[k]],
                                                        (T*)t.dz2 + (size_t)k * B * h2, (float*)G[ix.w3[k]], (float*)G[ix.b3[k]],
                                                        (float*)G[ix.b2[k]], B, h2, thresh, scale, k0, k1);
    MMOE_LAUNCH_OK("gemv_bwd_kernel");
  }
  {
    mmoe_gemm_problem p[4];
    for (int k = 0; k < 2; ++k) {
      mmoe_epilogue e = epi_none();
      e.out = (char*)t.dz1 + (size_t)k * B * h1 * es; e.out_dtype = dtype; e.ldo = h1; e.bwd_mode = 2;
      e.aux = (char*)s.z1 + (size_t)k * B * h1 * es; e.ld_aux = h1; e.colsum = (float*)G[ix.b1[k]];
      site_keys(c->seed, 10 + k, &k0, &k1);
      e.drop_p = drop_p; e.drop_key0 = k0; e.drop_key1 = k1;
      const char* dz2 = (const char*)t.dz2 + (size_t)k * B * h2 * es;
      p[2 * k] = linear_dgrad(dz2, h2, P[ix.w2[k]], B, h2, h1, e);
      p[2 * k + 1] = linear_wgrad(dz2, h2, (char*)s.a1 + (size_t)k * B * h1 * es, h1, (float*)G[ix.w2[k]], B, h2, h1);
    }
    MMOE_TRY(gemm_grouped(p, 4, dtype, 0, st));
  }
  {
    mmoe_gemm_problem p[4];
    for (int k = 0; k < 2; ++k) {
      mmoe_epilogue e = epi_none();
      e.out = (char*)t.dxn + (size_t)k * B * d * es; e.out_dtype = dtype; e.ldo = d;
      const char* dz1 = (const char*)t.dz1 + (size_t)k * B * h1 * es;
      p[2 * k] = linear_dgrad(dz1, h1, P[ix.w1[k]], B, h1, d, e);
      p[2 * k + 1] = linear_wgrad(dz1, h1, (char*)s.xn + (size_t)k * B * d * es, d, (float*)G[ix.w1[k]], B, h1, d);
    }
    MMOE_TRY(gemm_grouped(p, 4, dtype, 0, st));
  }
  for (int k = 0; k < 2; ++k) {
    LnBwdArgs a{};
    a.dy = (char*)t.dxn + (size_t)k * B * d * es; a.dy_dtype = dtype; a.x = s.fused + (size_t)k * B * d; a.x_dtype = MMOE_F32;
    a.stats = s.st + (size_t)k * B * 2; a.gamma = (const float*)P[ix.ln_w[k]]; a.dx = t.dfused + (size_t)k * B * d;
    a.dgamma = (float*)G[ix.ln_w[k]]; a.dbeta = (float*)G[ix.ln_b[k]]; a.rows = B; a.d = d; a.dtype = dtype;
    MMOE_TRY(layernorm_bwd(a, st));
  }
  {
    MixDev a{};
    a.experts = ev; a.expert_stride = d; a.row_stride = (int64_t)n * d;
    for (int k2 = 0; k2 < 2; ++k2) for (int k = 0; k < n; ++k) a.sel[k2][k] = k;
    a.wg[0] = (const float*)P[ix.gw[0]]; a.wg[1] = (const float*)P[ix.gw[1]];
    a.w = s.w; a.dfused = t.dfused; a.dl = t.dl; a.dexperts = d_ev; a.accumulate_dexperts = 0; a.dquery = nullptr;
    a.B = B; a.d = d; a.n = n;
    mix_bwd_kernel<<<rows_grid(B, 8), 256, 0, st>>>(a, n);
    MMOE_LAUNCH_OK("mix_bwd_kernel");
  }
  for (int k = 0; k < 2; ++k)
    MMOE_TRY(small_wgrad(t.dl + (size_t)k * B * n, s.query, (float*)G[ix.gw[k]], (float*)G[ix.gb[k]], B, d, n, st));
  return 0;
}

}  // namespace mmoe

using namespace mmoe;

extern "C" size_t mmoe_head_saved_bytes(const mmoe_head_cfg* cfg, int32_t B, int dtype) {
  Arena A(nullptr);
  head_layout(A, *cfg, B, dtype);
  return A.off + 256;
}
extern "C" size_t mmoe_head_workspace_bytes(const mmoe_head_cfg* cfg, int32_t B, int dtype) {
  Arena A(nullptr);
  head_scratch_layout(A, *cfg, B, dtype);
  return A.off + 256;
}
extern "C" int mmoe_head_fwd(const mmoe_call* c, const mmoe_head_cfg* cfg, const float* expert_vecs, float* logits, float* gate_w) {
  MMOE_TRY(check_head(cfg));
  if (c->B == 0) return 0;
  MMOE_CHECK(c->saved != nullptr && c->saved_bytes >= mmoe_head_saved_bytes(cfg, c->B, c->dtype), "head_fwd: saved blob too small");
  if (c->dtype == MMOE_BF16) return head_fwd_t<__nv_bfloat16>(c, cfg, expert_vecs, logits, gate_w);
  if (c->dtype == MMOE_F16) return head_fwd_t<__half>(c, cfg, expert_vecs, logits, gate_w);
  return head_fwd_t<float>(c, cfg, expert_vecs, logits, gate_w);
}
extern "C" int mmoe_head_bwd(const mmoe_call* c, const mmoe_head_cfg* cfg, const float* expert_vecs, const float* dlogits, float* d_expert_vecs) {
  MMOE_TRY(check_head(cfg));
  if (c->B == 0) return 0;
  MMOE_CHECK(c->saved != nullptr && c->saved_bytes >= mmoe_head_saved_bytes(cfg, c->B, c->dtype), "head_bwd: saved blob too small");
  MMOE_CHECK(c->workspace != nullptr && c->workspace_bytes >= mmoe_head_workspace_bytes(cfg, c->B, c->dtype), "head_bwd: workspace too small");
  if (c->dtype == MMOE_BF16) return head_bwd_t<__nv_bfloat16>(c, cfg, expert_vecs, dlogits, d_expert_vecs);
  if (c->dtype == MMOE_F16) return head_bwd_t<__half>(c, cfg, expert_vecs, dlogits, d_expert_vecs);
  return head_bwd_t<float>(c, cfg, expert_vecs, dlogits, d_expert_vecs);
}

extern "C" int mmoe_dense_gate_fwd(const float* x, const float* wg, const float* bg, float* out, int64_t B, int32_t d, int32_t n, void* stream) {
  MMOE_CHECK(n >= 1 && n <= MIX_MAXN, "dense gate: n_expert must be in [1,%d]", MIX_MAXN);
  if (B == 0) return 0;
  dense_gate_fwd_kernel<<<rows_grid(B, 8), 256, 0, (cudaStream_t)stream>>>(x, wg, bg, out, B, d, n);
  MMOE_LAUNCH_OK("dense_gate_fwd_kernel");
  return 0;
}
/* dl: scratch fp32 [B,n]; dwg [n,d] and dbg [n] are accumulated into (zero them first). */
extern "C" int mmoe_dense_gate_bwd(const float* x, const float* wg, const float* w, const float* dw, float* dl, float* dx,
                                   float* dwg, float* dbg, int64_t B, int32_t d, int32_t n, void* stream) {
  MMOE_CHECK(n >= 1 && n <= MIX_MAXN, "dense gate: n_expert must be in [1,%d]", MIX_MAXN);
  if (B == 0) return 0;
  dense_gate_bwd_kernel<<<rows_grid(B, 8), 256, 0, (cudaStream_t)stream>>>(w, dw, wg, dl, dx, B, d, n);
  MMOE_LAUNCH_OK("dense_gate_bwd_kernel");
  return small_wgrad(dl, x, dwg, dbg, B, d, n, (cudaStream_t)stream);
}
