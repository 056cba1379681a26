// EnhancedCrossFuse — model.py:454-507 (HoME variant model_HoME.py:469-522).
//
// [v_cls, t_cls] is at the same time the concatenated [B,2d] operand of res_proj / gate and the
// [B,2,d] two-token sequence of the encoder layers (same memory), so no cat/stack copies exist.
// forward : stack -> res_proj GEMM -> LN -> 2 encoder layers (S=2) -> cast -> [gate GEMM +GELU] ->
//           fused [gate GEMV + sigmoid + mix + identity add] kernel -> (v1) LN -> [proj GEMM +GELU +dropout]
#include "encoder.cuh"

namespace mmoe {

template <typename T>
__global__ void stack2_kernel(const float* __restrict__ v, const float* __restrict__ t, float* __restrict__ x0,
                              T* __restrict__ xt, int64_t B, int d) {
  const int64_t n = B * 2 * d;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / (2 * d);
    const int r = (int)(i - b * 2 * d);
    const float val = r < d ? v[b * d + r] : t[b * d + (r - d)];
    x0[i] = val;
    xt[i] = from_f<T>(val);
  }
}

// y = g*x0 + (1-g)*x1 + identity,  g = sigmoid(<ga, w2> + b2)            (model.py:501-507)
template <typename T>
__global__ void __launch_bounds__(256) fuse_gate_fwd_kernel(const T* __restrict__ ga, const float* __restrict__ w2,
                                                            const float* __restrict__ b2, const float* __restrict__ x,
                                                            const float* __restrict__ identity, float* __restrict__ y,
                                                            float* __restrict__ g_saved, int64_t B, int d, int dh) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t b = warp0; b < B; b += nw) {
    float acc = 0.f;
    for (int k = lane; k < dh; k += 32) acc = fmaf(to_f<T>(ga[b * dh + k]), w2[k], acc);
    const float g = sigmoid_f(warp_sum(acc) + b2[0]);
    if (lane == 0) g_saved[b] = g;
    const float* x0 = x + b * 2 * d;
    const float* x1 = x0 + d;
    for (int c = lane; c < d; c += 32) y[b * d + c] = g * x0[c] + (1.f - g) * x1[c] + identity[b * d + c];
  }
}

template <typename T>
__global__ void __launch_bounds__(256) fuse_gate_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                            const float* __restrict__ g_saved, const T* __restrict__ ga,
                                                            const T* __restrict__ zg, const float* __restrict__ w2,
                                                            float* __restrict__ dX, T* __restrict__ dzg, float* __restrict__ dw2,
                                                            float* __restrict__ db2, float* __restrict__ db0, int64_t B, int d, int dh) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  constexpr int MAXK = 16;    // dh <= 512
  float acc_w2[MAXK], acc_b0[MAXK];
#pragma unroll
  for (int i = 0; i < MAXK; ++i) acc_w2[i] = acc_b0[i] = 0.f;
  float acc_b2 = 0.f;
  for (int64_t b = warp0; b < B; b += nw) {
    const float g = g_saved[b];
    const float* x0 = x + b * 2 * d;
    const float* x1 = x0 + d;
    float dg = 0.f;
    for (int c = lane; c < d; c += 32) {
      const float v = dy[b * d + c];
      dg = fmaf(v, x0[c] - x1[c], dg);
      dX[b * 2 * d + c] = g * v;
      dX[b * 2 * d + d + c] = (1.f - g) * v;
    }
    const float dgl = warp_sum(dg) * g * (1.f - g);
    acc_b2 += dgl;
#pragma unroll
    for (int i = 0; i < MAXK; ++i) {
      const int k = lane + 32 * i;
      if (k < dh) {
        const T o = from_f<T>(dgl * w2[k] * gelu_grad_f(to_f<T>(zg[b * dh + k])));
        dzg[b * dh + k] = o;
        acc_b0[i] += to_f<T>(o);
        acc_w2[i] = fmaf(dgl, to_f<T>(ga[b * dh + k]), acc_w2[i]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < MAXK; ++i) {
    const int k = lane + 32 * i;
    if (k < dh) { atomicAdd(dw2 + k, acc_w2[i]); atomicAdd(db0 + k, acc_b0[i]); }
  }
  if (lane == 0) atomicAdd(db2, acc_b2);
}

// dz = T(dout * dropmask * gelu'(z)) (+ column sums): backward of  out = drop(gelu(z))  when dout is not a GEMM result
template <typename T>
__global__ void act_bwd_kernel(const float* __restrict__ dout, const T* __restrict__ z, T* __restrict__ dz,
                               float* __restrict__ colsum, int64_t rows, int cols, uint32_t thresh, float scale, uint32_t k0, uint32_t k1,
                               int rpb) {
  const int64_t r0 = (int64_t)blockIdx.x * rpb, r1 = min(rows, r0 + rpb);
  for (int c = threadIdx.x; c < cols; c += blockDim.x) {
    float acc = 0.f;
    for (int64_t r = r0; r < r1; ++r) {
      const float zz = to_f<T>(z[r * cols + c]);
      // 16-bit modes: the branch-free form (|err| ~6e-7, three orders below the rounding of the stored value)
      float v = dout[r * cols + c] * (sizeof(T) == 2 ? gelu_grad_fast_f(zz) : gelu_grad_f(zz));
      if (thresh != 0) v = drop_keep(k0, k1, (uint64_t)r * cols + c, thresh) ? v * scale : 0.f;
      const T o = from_f<T>(v);
      dz[r * cols + c] = o;
      acc += to_f<T>(o);
    }
    atomicAdd(colsum + c, acc);
  }
}

struct FuseIdx { int layer0, r_w, r_b, rln_w, rln_b, g0_w, g0_b, g2_w, g2_b, pln_w, pln_b, p_w, p_b; };
static FuseIdx fuse_idx(int depth) {
  FuseIdx i; int p = 12 * depth;
  i.layer0 = 0; i.r_w = p++; i.r_b = p++; i.rln_w = p++; i.rln_b = p++; i.g0_w = p++; i.g0_b = p++; i.g2_w = p++; i.g2_b = p++;
  i.pln_w = p++; i.pln_b = p++; i.p_w = p++; i.p_b = p++;
  return i;
}

struct FuseSaved {
  float* x0; void* xt0; float* r; float* st_r; float* identity; EncSaved layer[4];
  float* xl; void* xt; void* zg; void* ga; float* g; float* y; float* st_y; void* yn; void* zp;
};
static FuseSaved fuse_layout(Arena& A, const mmoe_fuse_cfg& cfg, int B, int dtype, int home) {
  FuseSaved s{};
  const int d = cfg.d; const size_t es = dtype_size(dtype); const size_t Bz = (size_t)B;
  s.x0 = (float*)A.take(Bz * 2 * d * 4); s.xt0 = A.take(Bz * 2 * d * es);
  s.r = (float*)A.take(Bz * d * 4); s.st_r = (float*)A.take(Bz * 2 * 4); s.identity = (float*)A.take(Bz * d * 4);
  for (int l = 0; l < cfg.depth; ++l) s.layer[l] = enc_layout(A, (int64_t)2 * B, d, 4 * d, es);
  s.xl = (float*)A.take(Bz * 2 * d * 4); s.xt = A.take(Bz * 2 * d * es); s.zg = A.take(Bz * (d / 2) * es); s.ga = A.take(Bz * (d / 2) * es);
  s.g = (float*)A.take(Bz * 4); s.y = (float*)A.take(Bz * d * 4);
  if (!home) { s.st_y = (float*)A.take(Bz * 2 * 4); s.yn = A.take(Bz * d * es); s.zp = A.take(Bz * d * es); }
  return s;
}
struct FuseScratch { EncScratch enc; float* dy; void* dzp; void* dyn; float* dX; void* dzg; float* dr; void* gr; void* hand; };
static FuseScratch fuse_scratch_layout(Arena& A, const mmoe_fuse_cfg& cfg, int B, int dtype) {
  FuseScratch t{};
  const int d = cfg.d; const size_t es = dtype_size(dtype); const size_t Bz = (size_t)B;
  t.enc = enc_scratch_layout(A, (int64_t)2 * B, d, 4 * d, es);
  t.dy = (float*)A.take(Bz * d * 4); t.dzp = A.take(Bz * d * es); t.dyn = A.take(Bz * d * es);
  t.dX = (float*)A.take(Bz * 2 * d * 4); t.dzg = A.take(Bz * (d / 2) * es);
  t.dr = (float*)A.take(Bz * d * 4); t.gr = A.take(Bz * d * es);
  t.hand = A.take(Bz * 2 * d * es);
  return t;
}
static int check_fuse(const mmoe_fuse_cfg* cfg) {
  MMOE_CHECK(cfg->depth >= 1 && cfg->depth <= 4, "fuse expert: depth must be in [1,4]");
  MMOE_CHECK(cfg->d % 16 == 0 && cfg->d <= 1024 && cfg->d % cfg->n_head == 0, "fuse expert: unsupported d=%d n_head=%d", cfg->d, cfg->n_head);
  return 0;
}

template <typename T>
static int fuse_fwd_t(const mmoe_call* c, const mmoe_fuse_cfg* cfg, const float* v_cls, const float* t_cls, float* out) {
  const int B = c->B, d = cfg->d, dh = d / 2, dtype = c->dtype;
  cudaStream_t st = (cudaStream_t)c->stream;
  Arena A(c->saved);
  FuseSaved s = fuse_layout(A, *cfg, B, dtype, c->home);
  const FuseIdx ix = fuse_idx(cfg->depth);
  const void* const* P = c->params;
  const float drop_p = c->training ? c->drop_p : 0.f;
  uint32_t k0, k1;
  {
    const int64_t n = (int64_t)B * 2 * d;
    int blocks = (int)((n + 255) / 256); if (blocks > sm_count() * 8) blocks = sm_count() * 8;
    stack2_kernel<T><<<blocks, 256, 0, st>>>(v_cls, t_cls, s.x0, (T*)s.xt0, B, d);
    MMOE_LAUNCH_OK("stack2_kernel");
  }
  {
    mmoe_epilogue e = epi_none();
    e.out = s.r; e.out_dtype = MMOE_F32; e.ldo = d; e.bias = (const float*)P[ix.r_b];
    mmoe_gemm_problem p = linear_fwd(s.xt0, 2 * d, P[ix.r_w], B, d, 2 * d, e);
    MMOE_TRY(gemm_grouped(&p, 1, dtype, 0, st));
  }
  MMOE_TRY(layernorm_fwd(s.r, MMOE_F32, (const float*)P[ix.rln_w], (const float*)P[ix.rln_b], nullptr, s.identity, s.st_r, B, d, dtype, st));
  EncCtx ec{};
  ec.dtype = dtype; ec.M = 2 * (int64_t)B; ec.Bseq = B; ec.S = 2; ec.d = d; ec.ff = 4 * d; ec.H = cfg->n_head;
  ec.mask = nullptr; ec.drop_p = drop_p; ec.seed = c->seed; ec.stream = st;
  const float* x = s.x0;
  const void* delta = nullptr;
  for (int l = 0; l < cfg->depth; ++l) {
    ec.site0 = 16 * l;
    MMOE_TRY(enc_fwd(ec, enc_w(P + ix.layer0 + 12 * l), x, delta, s.layer[l]));
    x = s.layer[l].x1; delta = s.layer[l].y2;
  }
  MMOE_TRY(add_cast(x, delta, s.xl, s.xt, (int64_t)B * 2 * d, dtype, st));
  x = s.xl;
  {
    mmoe_epilogue e = epi_none();
    e.out = s.ga; e.out_dtype = dtype; e.ldo = dh; e.bias = (const float*)P[ix.g0_b]; e.act = 2; e.preact = s.zg;
    mmoe_gemm_problem p = linear_fwd(s.xt, 2 * d, P[ix.g0_w], B, dh, 2 * d, e);
    MMOE_TRY(gemm_grouped(&p, 1, dtype, 0, st));
  }
  {
    int blocks = (B + 7) / 8; if (blocks > sm_count() * 4) blocks = sm_count() * 4;
    fuse_gate_fwd_kernel<T><<<blocks, 256, 0, st>>>((const T*)s.ga, (const float*)P[ix.g2_w], (const float*)P[ix.g2_b], x, s.identity,
                                                    c->home ? out : s.y, s.g, B, d, dh);
    MMOE_LAUNCH_OK("fuse_gate_fwd_kernel");
  }
  if (c->home) return 0;
  MMOE_TRY(layernorm_fwd(s.y, MMOE_F32, (const float*)P[ix.pln_w], (const float*)P[ix.pln_b], s.yn, nullptr, s.st_y, B, d, dtype, st));
  {
    mmoe_epilogue e = epi_none();
    e.out = out; e.out_dtype = MMOE_F32; e.ldo = d; e.bias = (const float*)P[ix.p_b]; e.act = 2; e.preact = s.zp;
    site_keys(c->seed, 100, &k0, &k1);
    e.drop_p = drop_p; e.drop_key0 = k0; e.drop_key1 = k1;
    mmoe_gemm_problem p = linear_fwd(s.yn, d, P[ix.p_w], B, d, d, e);
    MMOE_TRY(gemm_grouped(&p, 1, dtype, 0, st));
  }
  return 0;
}

template <typename T>
static int fuse_bwd_t(const mmoe_call* c, const mmoe_fuse_cfg* cfg, const float* dout, float* d_cat) {
  const int B = c->B, d = cfg->d, dh = d / 2, dtype = c->dtype;
  cudaStream_t st = (cudaStream_t)c->stream;
  Arena A(c->saved);
  FuseSaved s = fuse_layout(A, *cfg, B, dtype, c->home);
  Arena W(c->workspace);
  FuseScratch t = fuse_scratch_layout(W, *cfg, B, dtype);
  const FuseIdx ix = fuse_idx(cfg->depth);
  const void* const* P = c->params;
  void* const* G = c->grads;
  const float drop_p = c->training ? c->drop_p : 0.f;
  uint32_t k0, k1;
  const float* x_last = s.xl;
  const float* dy = dout;
  if (!c->home) {
    site_keys(c->seed, 100, &k0, &k1);
    // rows per block: enough blocks to cover the chip (32 rows per block left a B = 512 call on 16 SMs: 55 us for 0.4 M elements)
    int rpb = (int)(((int64_t)B + 2 * sm_count() - 1) / (2 * sm_count()));
    if (rpb > 32) rpb = 32;
    if (rpb < 1) rpb = 1;
    act_bwd_kernel<T><<<(B + rpb - 1) / rpb, 256, 0, st>>>(dout, (const T*)s.zp, (T*)t.dzp, (float*)G[ix.p_b], B, d,
                                                     drop_p > 0.f ? drop_threshold(drop_p) : 0u, drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f, k0, k1, rpb);
    MMOE_LAUNCH_OK("act_bwd_kernel");
    {
      mmoe_epilogue e = epi_none();
      e.out = t.dyn; e.out_dtype = dtype; e.ldo = d;
      mmoe_gemm_problem p[2] = {linear_dgrad(t.dzp, d, P[ix.p_w], B, d, d, e), linear_wgrad(t.dzp, d, s.yn, d, (float*)G[ix.p_w], B, d, d)};
      MMOE_TRY(gemm_grouped(p, 2, dtype, 0, st));
    }
    {
      LnBwdArgs a{};
      a.dy = t.dyn; a.dy_dtype = dtype; a.x = s.y; a.x_dtype = MMOE_F32; a.stats = s.st_y; a.gamma = (const float*)P[ix.pln_w];
      a.dx = t.dy; a.dgamma = (float*)G[ix.pln_w]; a.dbeta = (float*)G[ix.pln_b]; a.rows = B; a.d = d; a.dtype = dtype;
      MMOE_TRY(layernorm_bwd(a, st));
    }
    dy = t.dy;
  }
  {
    int blocks = (B + 7) / 8; if (blocks > sm_count() * 2) blocks = sm_count() * 2;
    fuse_gate_bwd_kernel<T><<<blocks, 256, 0, st>>>(dy, x_last, s.g, (const T*)s.ga, (const T*)s.zg, (const float*)P[ix.g2_w], t.dX,
                                                    (T*)t.dzg, (float*)G[ix.g2_w], (float*)G[ix.g2_b], (float*)G[ix.g0_b], B, d, dh);
    MMOE_LAUNCH_OK("fuse_gate_bwd_kernel");
  }
  {
    mmoe_epilogue e = epi_none();   // dX += dzg Wg0
    e.out = t.dX; e.out_dtype = MMOE_F32; e.ldo = 2 * d; e.residual = t.dX; e.ld_res = 2 * d;
    mmoe_gemm_problem p[2] = {linear_dgrad(t.dzg, dh, P[ix.g0_w], B, dh, 2 * d, e),
                              linear_wgrad(t.dzg, dh, s.xt, 2 * d, (float*)G[ix.g0_w], B, dh, 2 * d)};
    MMOE_TRY(gemm_grouped(p, 2, dtype, 0, st));
  }
  EncCtx ec{};
  ec.dtype = dtype; ec.M = 2 * (int64_t)B; ec.Bseq = B; ec.S = 2; ec.d = d; ec.ff = 4 * d; ec.H = cfg->n_head;
  ec.mask = nullptr; ec.drop_p = drop_p; ec.seed = c->seed; ec.stream = st;
  for (int l = cfg->depth - 1; l >= 0; --l) {
    ec.site0 = 16 * l;
    const float* x_in = l == 0 ? s.x0 : s.layer[l].x_sum;
    EncHandoff below{t.hand, l > 0 ? enc_g(G + ix.layer0 + 12 * (l - 1)).b2 : nullptr, (uint32_t)(16 * (l - 1))};
    MMOE_TRY(enc_bwd(ec, enc_w(P + ix.layer0 + 12 * l), enc_g(G + ix.layer0 + 12 * l), x_in, s.layer[l], t.enc, t.dX, t.dX,
                     l < cfg->depth - 1 ? t.hand : nullptr, l > 0 ? &below : nullptr));
  }
  // identity = LN(res_proj(cat)); its gradient is dy
  {
    LnBwdArgs a{};
    a.dy = dy; a.dy_dtype = MMOE_F32; a.x = s.r; a.x_dtype = MMOE_F32; a.stats = s.st_r; a.gamma = (const float*)P[ix.rln_w];
    a.dx = t.dr; a.dgamma = (float*)G[ix.rln_w]; a.dbeta = (float*)G[ix.rln_b];
    a.g_out = t.gr; a.g_colsum = (float*)G[ix.r_b]; a.rows = B; a.d = d; a.dtype = dtype;
    MMOE_TRY(layernorm_bwd(a, st));
  }
  {
    mmoe_epilogue e = epi_none();   // d_cat = dX0 + gr Wr
    e.out = d_cat; e.out_dtype = MMOE_F32; e.ldo = 2 * d; e.residual = t.dX; e.ld_res = 2 * d;
    mmoe_gemm_problem p[2] = {linear_dgrad(t.gr, d, P[ix.r_w], B, d, 2 * d, e),
                              linear_wgrad(t.gr, d, s.xt0, 2 * d, (float*)G[ix.r_w], B, d, 2 * d)};
    MMOE_TRY(gemm_grouped(p, 2, dtype, 0, st));
  }
  return 0;
}

}  // namespace mmoe

using namespace mmoe;

extern "C" size_t mmoe_fuse_saved_bytes(const mmoe_fuse_cfg* cfg, int32_t B, int dtype) {
  Arena A(nullptr);
  fuse_layout(A, *cfg, B, dtype, 0);
  return A.off + 256;
}
// layout query (tests / debugging), see mmoe_cross_saved_offset.  which: 0 = h of encoder layer `layer` ([2B, 4d] T)
extern "C" int mmoe_fuse_saved_offset(const mmoe_fuse_cfg* cfg, int32_t B, int dtype, int home, int layer, int which,
                                      size_t* offset, size_t* bytes) {
  MMOE_TRY(check_fuse(cfg));
  MMOE_CHECK(layer >= 0 && layer < cfg->depth && which == 0, "saved_offset: bad layer/buffer");
  Arena A(nullptr);
  FuseSaved s = fuse_layout(A, *cfg, B, dtype, home);
  *offset = (size_t)(char*)s.layer[layer].h;
  *bytes = (size_t)2 * B * 4 * cfg->d * dtype_size(dtype);
  return 0;
}

extern "C" size_t mmoe_fuse_workspace_bytes(const mmoe_fuse_cfg* cfg, int32_t B, int dtype) {
  Arena A(nullptr);
  fuse_scratch_layout(A, *cfg, B, dtype);
  return A.off + 256;
}
extern "C" int mmoe_fuse_fwd(const mmoe_call* c, const mmoe_fuse_cfg* cfg, const float* v_cls, const float* t_cls, float* out) {
  MMOE_TRY(check_fuse(cfg));
  if (c->B == 0) return 0;
  MMOE_CHECK(c->saved != nullptr && c->saved_bytes >= mmoe_fuse_saved_bytes(cfg, c->B, c->dtype), "fuse_fwd: saved blob too small");
  if (c->dtype == MMOE_BF16) return fuse_fwd_t<__nv_bfloat16>(c, cfg, v_cls, t_cls, out);
  if (c->dtype == MMOE_F16) return fuse_fwd_t<__half>(c, cfg, v_cls, t_cls, out);
  return fuse_fwd_t<float>(c, cfg, v_cls, t_cls, out);
}
extern "C" int mmoe_fuse_bwd(const mmoe_call* c, const mmoe_fuse_cfg* cfg, const float* dout, float* d_cat) {
  MMOE_TRY(check_fuse(cfg));
  if (c->B == 0) return 0;
  MMOE_CHECK(c->saved != nullptr && c->saved_bytes >= mmoe_fuse_saved_bytes(cfg, c->B, c->dtype), "fuse_bwd: saved blob too small");
  MMOE_CHECK(c->workspace != nullptr && c->workspace_bytes >= mmoe_fuse_workspace_bytes(cfg, c->B, c->dtype), "fuse_bwd: workspace too small");
  if (c->dtype == MMOE_BF16) return fuse_bwd_t<__nv_bfloat16>(c, cfg, dout, d_cat);
  if (c->dtype == MMOE_F16) return fuse_bwd_t<__half>(c, cfg, dout, d_cat);
  return fuse_bwd_t<float>(c, cfg, dout, d_cat);
}
