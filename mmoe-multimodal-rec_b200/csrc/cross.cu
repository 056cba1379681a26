// RobustTextCrossExpert — model.py:386-451 (HoME variant model_HoME.py:401-466).
//
// forward : 2+2 encoder layers (encoder.cuh) -> cast -> {Q proj | KV proj} (one grouped GEMM) ->
//           cross attention -> out-proj GEMM -> fused [gate mix + AttnPool1D] kernel
//           -> (v1 only) LN -> [MLP1 GEMM +GELU +dropout] -> [MLP2 GEMM +dropout +residual]
// backward: mirrors it; every bias / LayerNorm / query / gate gradient is produced inside the kernel that
//           already touches the data (no separate reduction passes).
#include "encoder.cuh"

namespace mmoe {

// ------------------------------------------------------------------------------------------
// fused = a*U + (1-a)*C ; AttnPool1D(fused, user_mask)         (model.py:443-447, 199-206)
// one CTA per sample; 256 threads
// ------------------------------------------------------------------------------------------
struct PoolDev {
  const float* U; const void* C; const float* gate; const float* query; const uint8_t* mask;
  float* pooled; float* w_saved;
  // backward
  const float* dpooled; float* dU; void* dC; float* dquery; float* dgate; float* dbo;
  int S, d, home;
  uint32_t thresh, k0, k1; float drop_scale;
};

// 4 consecutive elements of a row as fp32 (16-byte / 8-byte accesses; d % 8 == 0 keeps them aligned)
template <typename T> __device__ __forceinline__ float4 pool_ld4(const T* p);
template <> __device__ __forceinline__ float4 pool_ld4<float>(const float* p) { return *reinterpret_cast<const float4*>(p); }
template <> __device__ __forceinline__ float4 pool_ld4<__nv_bfloat16>(const __nv_bfloat16* p) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  return make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xFFFF0000u), __uint_as_float(u.y << 16), __uint_as_float(u.y & 0xFFFF0000u));
}
template <> __device__ __forceinline__ float4 pool_ld4<__half>(const __half* p) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
  return make_float4(a.x, a.y, b.x, b.y);
}
template <typename T> __device__ __forceinline__ void pool_st4(T* p, float4 v, float4& stored);
template <> __device__ __forceinline__ void pool_st4<float>(float* p, float4 v, float4& stored) { *reinterpret_cast<float4*>(p) = v; stored = v; }
template <> __device__ __forceinline__ void pool_st4<__nv_bfloat16>(__nv_bfloat16* p, float4 v, float4& stored) {
  const __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
  uint2 u; u.x = *reinterpret_cast<const uint32_t*>(&lo); u.y = *reinterpret_cast<const uint32_t*>(&hi);
  *reinterpret_cast<uint2*>(p) = u;
  stored = make_float4(__low2float(lo), __high2float(lo), __low2float(hi), __high2float(hi));
}
template <> __device__ __forceinline__ void pool_st4<__half>(__half* p, float4 v, float4& stored) {
  const __half2 lo = __floats2half2_rn(v.x, v.y), hi = __floats2half2_rn(v.z, v.w);
  uint2 u; u.x = *reinterpret_cast<const uint32_t*>(&lo); u.y = *reinterpret_cast<const uint32_t*>(&hi);
  *reinterpret_cast<uint2*>(p) = u;
  stored = make_float4(__low2float(lo), __high2float(lo), __low2float(hi), __high2float(hi));
}
__device__ __forceinline__ float4 pool_mix(float alpha, float4 u, float4 c) {
  const float be = 1.f - alpha;
  return make_float4(alpha * u.x + be * c.x, alpha * u.y + be * c.y, alpha * u.z + be * c.z, alpha * u.w + be * c.w);
}
__device__ __forceinline__ float pool_dot(float4 a, float4 b) { return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w))); }

// One CTA per sample.  Pass 1 (a warp per sentence row, 16-byte loads, the whole row in flight at once): scores of the
// gate-mixed rows against the pooling query.  Pass 2 (a thread per 4 columns, rows from L2): the weighted sum.
template <typename T>
__global__ void __launch_bounds__(256) pool_fwd_kernel(const PoolDev a) {
  __shared__ float sc[64];
  const int b = blockIdx.x, S = a.S, d = a.d;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float alpha = sigmoid_f(a.gate[0]);
  const float* U = a.U + (int64_t)b * S * d;
  const T* C = (const T*)a.C + (int64_t)b * S * d;
  const float inv_sqrt_d = rsqrtf((float)d);
  for (int i = warp; i < S; i += 8) {
    float acc = 0.f;
#pragma unroll 8
    for (int c = lane * 4; c < d; c += 128)
      acc += pool_dot(pool_mix(alpha, pool_ld4<float>(U + i * d + c), pool_ld4<T>(C + i * d + c)), pool_ld4<float>(a.query + c));
    acc = warp_sum(acc) * inv_sqrt_d;
    if (lane == 0) sc[i] = a.mask[(int64_t)b * S + i] ? -INFINITY : acc;
  }
  __syncthreads();
  if (warp == 0) {
    float m = -INFINITY;
    for (int i = lane; i < S; i += 32) m = fmaxf(m, sc[i]);
    m = warp_max(m);
    float s = 0.f;
    for (int i = lane; i < S; i += 32) { const float e = expf(sc[i] - m); sc[i] = e; s += e; }
    s = warp_sum(s);
    const float inv = 1.f / s;
    bool any_finite = false;
    for (int i = lane; i < S; i += 32) { sc[i] *= inv; any_finite |= isfinite(sc[i]); }
    any_finite = __any_sync(0xffffffffu, any_finite);
    for (int i = lane; i < S; i += 32) {
      float w = sc[i];
      if (a.home && !any_finite) w = 0.f;                    // model_HoME.py:210-211
      a.w_saved[(int64_t)b * S + i] = w;
      if (a.thresh != 0) w = drop_keep(a.k0, a.k1, (uint64_t)b * S + i, a.thresh) ? w * a.drop_scale : 0.f;
      sc[i] = w;
    }
  }
  __syncthreads();
  for (int c = threadIdx.x * 4; c < d; c += 1024) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
    for (int i = 0; i < S; ++i) {
      const float4 f = pool_mix(alpha, pool_ld4<float>(U + i * d + c), pool_ld4<T>(C + i * d + c));
      const float w = sc[i];
      acc.x = fmaf(w, f.x, acc.x); acc.y = fmaf(w, f.y, acc.y); acc.z = fmaf(w, f.z, acc.z); acc.w = fmaf(w, f.w, acc.w);
    }
    *reinterpret_cast<float4*>(a.pooled + (int64_t)b * d + c) = acc;
  }
}

template <typename T>
__global__ void __launch_bounds__(256) pool_bwd_kernel(const PoolDev a) {
  __shared__ float wd[64];    // weights after dropout
  __shared__ float ds[64];    // d(score) / sqrt(d)
  __shared__ float red[8];
  const int b = blockIdx.x, S = a.S, d = a.d;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float alpha = sigmoid_f(a.gate[0]);
  const float* U = a.U + (int64_t)b * S * d;
  const T* C = (const T*)a.C + (int64_t)b * S * d;
  const float* dp = a.dpooled + (int64_t)b * d;
  const float inv_sqrt_d = rsqrtf((float)d);
  // dwd[i] = <dpooled, fused_i>
  for (int i = warp; i < S; i += 8) {
    float acc = 0.f;
#pragma unroll 8
    for (int c = lane * 4; c < d; c += 128)
      acc += pool_dot(pool_mix(alpha, pool_ld4<float>(U + i * d + c), pool_ld4<T>(C + i * d + c)), pool_ld4<float>(dp + c));
    acc = warp_sum(acc);
    if (lane == 0) ds[i] = acc;
  }
  __syncthreads();
  if (warp == 0) {
    float dot = 0.f;
    for (int i = lane; i < S; i += 32) {
      const float w = a.w_saved[(int64_t)b * S + i];
      float dw = ds[i];
      float wdrop = w;
      if (a.thresh != 0) {
        const bool keep = drop_keep(a.k0, a.k1, (uint64_t)b * S + i, a.thresh);
        dw = keep ? dw * a.drop_scale : 0.f;
        wdrop = keep ? w * a.drop_scale : 0.f;
      }
      wd[i] = wdrop;
      ds[i] = dw;
      dot += w * dw;
    }
    dot = warp_sum(dot);
    for (int i = lane; i < S; i += 32) {
      const float w = a.w_saved[(int64_t)b * S + i];
      ds[i] = w * (ds[i] - dot) * inv_sqrt_d;   // masked keys have w = 0 -> 0
    }
  }
  __syncthreads();
  float dalpha = 0.f;
  float* dU = a.dU + (int64_t)b * S * d;
  T* dC = (T*)a.dC + (int64_t)b * S * d;
  const float be = 1.f - alpha;
  for (int c = threadIdx.x * 4; c < d; c += 1024) {
    const float4 q = pool_ld4<float>(a.query + c), dpc = pool_ld4<float>(dp + c);
    float4 dq = make_float4(0.f, 0.f, 0.f, 0.f), dbo = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
    for (int i = 0; i < S; ++i) {
      const float4 u = pool_ld4<float>(U + i * d + c), cc = pool_ld4<T>(C + i * d + c);
      const float4 f = pool_mix(alpha, u, cc);
      const float w = wd[i], s_ = ds[i];
      const float4 df = make_float4(fmaf(w, dpc.x, s_ * q.x), fmaf(w, dpc.y, s_ * q.y), fmaf(w, dpc.z, s_ * q.z), fmaf(w, dpc.w, s_ * q.w));
      dq.x = fmaf(s_, f.x, dq.x); dq.y = fmaf(s_, f.y, dq.y); dq.z = fmaf(s_, f.z, dq.z); dq.w = fmaf(s_, f.w, dq.w);
      dalpha += df.x * (u.x - cc.x) + df.y * (u.y - cc.y) + df.z * (u.z - cc.z) + df.w * (u.w - cc.w);
      *reinterpret_cast<float4*>(dU + i * d + c) = make_float4(alpha * df.x, alpha * df.y, alpha * df.z, alpha * df.w);
      float4 o;
      pool_st4<T>(dC + i * d + c, make_float4(be * df.x, be * df.y, be * df.z, be * df.w), o);
      dbo.x += o.x; dbo.y += o.y; dbo.z += o.z; dbo.w += o.w;
    }
    atomicAdd(a.dquery + c + 0, dq.x); atomicAdd(a.dquery + c + 1, dq.y); atomicAdd(a.dquery + c + 2, dq.z); atomicAdd(a.dquery + c + 3, dq.w);
    atomicAdd(a.dbo + c + 0, dbo.x); atomicAdd(a.dbo + c + 1, dbo.y); atomicAdd(a.dbo + c + 2, dbo.z); atomicAdd(a.dbo + c + 3, dbo.w);
  }
  dalpha = warp_sum(dalpha);
  if (lane == 0) red[warp] = dalpha;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w];
    atomicAdd(a.dgate, t * alpha * (1.f - alpha));
  }
}

template <typename T> static void launch_pool_fwd(const PoolDev& a, int B, cudaStream_t s) { pool_fwd_kernel<T><<<B, 256, 0, s>>>(a); }
template <typename T> static void launch_pool_bwd(const PoolDev& a, int B, cudaStream_t s) { pool_bwd_kernel<T><<<B, 256, 0, s>>>(a); }

// ------------------------------------------------------------------------------------------
// parameter table (state_dict order)
// ------------------------------------------------------------------------------------------
struct CrossIdx {
  int gate, user0, item0, ca_w, ca_b, ca_ow, ca_ob, query, norm_w, norm_b, m0_w, m0_b, m3_w, m3_b, count;
};
static CrossIdx cross_idx(int n_layer) {
  CrossIdx i;
  i.gate = 0; i.user0 = 1; i.item0 = 1 + 12 * n_layer;
  int p = 1 + 24 * n_layer;
  i.ca_w = p++; i.ca_b = p++; i.ca_ow = p++; i.ca_ob = p++; i.query = p++;
  i.norm_w = p++; i.norm_b = p++; i.m0_w = p++; i.m0_b = p++; i.m3_w = p++; i.m3_b = p++;
  i.count = p;
  return i;
}

struct CrossSaved {
  EncSaved user[4], item[4];
  float *U, *I;
  void *Ut, *It, *q, *kv, *ctx, *c;
  float* w_pool; float* pooled;
  float *normed_f, *st; void *normed_t, *z, *h;
};
static CrossSaved cross_layout(Arena& A, const mmoe_cross_cfg& cfg, int B, int dtype, int home) {
  CrossSaved s{};
  const int64_t M = (int64_t)B * cfg.S; const int d = cfg.d, ff = 4 * cfg.d; const size_t es = dtype_size(dtype);
  for (int l = 0; l < cfg.n_layer; ++l) s.user[l] = enc_layout(A, M, d, ff, es);
  for (int l = 0; l < cfg.n_layer; ++l) s.item[l] = enc_layout(A, M, d, ff, es);
  s.U = (float*)A.take(M * d * 4); s.I = nullptr;   // only the 16-bit copy of the item stream is needed downstream
  s.Ut = A.take(M * d * es); s.It = A.take(M * d * es);
  s.q = A.take(M * d * es); s.kv = A.take(M * 2 * d * es);
  s.ctx = A.take(M * d * es); s.c = A.take(M * d * es);
  s.w_pool = (float*)A.take((size_t)B * cfg.S * 4);
  s.pooled = (float*)A.take((size_t)B * d * 4);
  if (!home) {
    s.normed_f = (float*)A.take((size_t)B * d * 4); s.st = (float*)A.take((size_t)B * 2 * 4);
    s.normed_t = A.take((size_t)B * d * es); s.z = A.take((size_t)B * ff * es); s.h = A.take((size_t)B * ff * es);
  }
  return s;
}
struct CrossScratch {
  EncScratch enc;
  float *dU, *dI, *dpooled, *dnormed; void *dC, *dctx, *dq, *dkv, *g, *dz;
  void *hand_u, *hand_i;       // EncHandoff buffers of the user / item stack (separate: the two stacks' stage calls may interleave)
  void *dUq, *dIt;             // 16-bit dgrads of the cross-attention projections: dq Wq (added to dU) and dkv Wkv (= dI)
};
static CrossScratch cross_scratch_layout(Arena& A, const mmoe_cross_cfg& cfg, int B, int dtype) {
  CrossScratch t{};
  const int64_t M = (int64_t)B * cfg.S; const int d = cfg.d, ff = 4 * cfg.d; const size_t es = dtype_size(dtype);
  t.enc = enc_scratch_layout(A, M, d, ff, es);
  t.dU = (float*)A.take(M * d * 4); t.dI = (float*)A.take(M * d * 4);
  t.dpooled = (float*)A.take((size_t)B * d * 4); t.dnormed = (float*)A.take((size_t)B * d * 4);
  t.dC = A.take(M * d * es); t.dctx = A.take(M * d * es); t.dq = A.take(M * d * es); t.dkv = A.take(M * 2 * d * es);
  t.g = A.take((size_t)B * d * es); t.dz = A.take((size_t)B * ff * es);
  t.hand_u = A.take(M * d * es); t.hand_i = A.take(M * d * es);
  t.dUq = A.take(M * d * es); t.dIt = A.take(M * d * es);
  return t;
}

static int check_cfg(const mmoe_cross_cfg* cfg, int B) {
  MMOE_CHECK(cfg->n_layer >= 1 && cfg->n_layer <= 4, "cross expert: n_layer must be in [1,4]");
  MMOE_CHECK(cfg->S >= 1 && cfg->S <= 64, "cross expert: S must be in [1,64]");
  MMOE_CHECK(cfg->d % 8 == 0 && cfg->d <= 1024 && cfg->d % cfg->n_head == 0, "cross expert: unsupported d=%d n_head=%d", cfg->d, cfg->n_head);
  MMOE_CHECK(B >= 0, "negative batch");
  return 0;
}

}  // namespace mmoe

using namespace mmoe;

extern "C" size_t mmoe_cross_saved_bytes(const mmoe_cross_cfg* cfg, int32_t B, int dtype) {
  Arena A(nullptr);
  cross_layout(A, *cfg, B, dtype, 0);
  return A.off + 256;
}
// layout query (tests / debugging): byte range of a saved activation inside the blob.
// stream_id 0 = user stack, 1 = item stack; which: 0 = h (post ReLU/dropout FFN activation [B*S, 4d] T),
// 1 = x1 (stream after the attention block, fp32 [B*S, d]).
extern "C" int mmoe_cross_saved_offset(const mmoe_cross_cfg* cfg, int32_t B, int dtype, int home, int stream_id, int layer, int which,
                                       size_t* offset, size_t* bytes) {
  MMOE_TRY(check_cfg(cfg, B));
  MMOE_CHECK(layer >= 0 && layer < cfg->n_layer && (stream_id == 0 || stream_id == 1), "saved_offset: bad layer/stream");
  Arena A(nullptr);
  CrossSaved s = cross_layout(A, *cfg, B, dtype, home);
  const EncSaved& e = stream_id == 0 ? s.user[layer] : s.item[layer];
  const size_t M = (size_t)B * cfg->S;
  if (which == 0) { *offset = (size_t)(char*)e.h; *bytes = M * 4 * cfg->d * dtype_size(dtype); }
  else if (which == 1) { *offset = (size_t)(char*)e.x1; *bytes = M * cfg->d * 4; }
  else { set_error("saved_offset: unknown buffer %d", which); return -1; }
  return 0;
}

extern "C" size_t mmoe_cross_workspace_bytes(const mmoe_cross_cfg* cfg, int32_t B, int dtype) {
  Arena A(nullptr);
  cross_scratch_layout(A, *cfg, B, dtype);
  return A.off + 256;
}

extern "C" int mmoe_cross_fwd(const mmoe_call* c, const mmoe_cross_cfg* cfg, const float* user, const uint8_t* user_mask,
                              const float* item, const uint8_t* item_mask, float* out) {
  MMOE_TRY(check_cfg(cfg, c->B));
  if (c->B == 0) return 0;
  const int B = c->B, S = cfg->S, d = cfg->d, ff = 4 * d, dtype = c->dtype;
  const int M = B * S;
  const size_t es = dtype_size(dtype);
  cudaStream_t st = (cudaStream_t)c->stream;
  MMOE_CHECK(c->saved != nullptr && c->saved_bytes >= mmoe_cross_saved_bytes(cfg, B, dtype), "cross_fwd: saved blob too small");
  Arena A(c->saved);
  CrossSaved s = cross_layout(A, *cfg, B, dtype, c->home);
  const CrossIdx ix = cross_idx(cfg->n_layer);
  const void* const* P = c->params;
  const float drop_p = c->training ? c->drop_p : 0.f;
  uint32_t k0, k1;

  EncCtx ec{};
  ec.dtype = dtype; ec.M = M; ec.Bseq = B; ec.S = S; ec.d = d; ec.ff = ff; ec.H = cfg->n_head;
  ec.drop_p = drop_p; ec.seed = c->seed; ec.stream = st;
  const float* x = user;
  const void* delta = nullptr;
  for (int l = 0; l < cfg->n_layer; ++l) {
    ec.mask = user_mask; ec.site0 = 16 * l;
    MMOE_TRY(enc_fwd(ec, enc_w(P + ix.user0 + 12 * l), x, delta, s.user[l]));
    x = s.user[l].x1; delta = s.user[l].y2;
  }
  // stream outputs: U = x1 + y2 (fp32 for the gate mix / pooling) and its 16-bit copy (Q-projection operand)
  MMOE_TRY(add_cast(x, delta, s.U, s.Ut, (int64_t)M * d, dtype, st));
  const float* U = s.U;
  x = item; delta = nullptr;
  for (int l = 0; l < cfg->n_layer; ++l) {
    ec.mask = item_mask; ec.site0 = 16 * l + 8;
    MMOE_TRY(enc_fwd(ec, enc_w(P + ix.item0 + 12 * l), x, delta, s.item[l]));
    x = s.item[l].x1; delta = s.item[l].y2;
  }
  // cross attention: q from the user stream, k/v from the item stream (model.py:435-440); only the 16-bit copy of I is needed
  MMOE_TRY(add_cast(x, delta, nullptr, s.It, (int64_t)M * d, dtype, st));
  const char* w_in = (const char*)P[ix.ca_w];
  const float* b_in = (const float*)P[ix.ca_b];
  {
    mmoe_epilogue eq = epi_none(), ek = epi_none();
    eq.out = s.q; eq.out_dtype = dtype; eq.ldo = d; eq.bias = b_in;
    ek.out = s.kv; ek.out_dtype = dtype; ek.ldo = 2 * d; ek.bias = b_in + d;
    mmoe_gemm_problem p[2] = {linear_fwd(s.Ut, d, w_in, M, d, d, eq),
                              linear_fwd(s.It, d, w_in + (size_t)d * d * es, M, 2 * d, d, ek)};
    MMOE_TRY(gemm_grouped(p, 2, dtype, 0, st));
  }
  {
    AttnArgs a{};
    a.q = s.q; a.ldq = d; a.k = s.kv; a.v = (const char*)s.kv + (size_t)d * es; a.ldk = a.ldv = 2 * d;
    a.mask = item_mask; a.ctx = s.ctx; a.ldc = d; a.B = B; a.Sq = S; a.Sk = S; a.H = cfg->n_head; a.hd = d / cfg->n_head;
    site_keys(c->seed, 100, &k0, &k1);
    a.drop_p = drop_p; a.k0 = k0; a.k1 = k1; a.dtype = dtype;
    MMOE_TRY(attention_fwd(a, st));
  }
  {
    mmoe_epilogue e = epi_none();
    e.out = s.c; e.out_dtype = dtype; e.ldo = d; e.bias = (const float*)P[ix.ca_ob];
    mmoe_gemm_problem p = linear_fwd(s.ctx, d, P[ix.ca_ow], M, d, d, e);
    MMOE_TRY(gemm_grouped(&p, 1, dtype, 0, st));
  }
  {
    PoolDev a{};
    a.U = U; a.C = s.c; a.gate = (const float*)P[ix.gate]; a.query = (const float*)P[ix.query]; a.mask = user_mask;
    a.pooled = c->home ? out : s.pooled; a.w_saved = s.w_pool; a.S = S; a.d = d; a.home = c->home;
    site_keys(c->seed, 101, &k0, &k1);
    a.thresh = drop_p > 0.f ? drop_threshold(drop_p) : 0u; a.drop_scale = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
    a.k0 = k0; a.k1 = k1;
    if (dtype == MMOE_BF16) launch_pool_fwd<__nv_bfloat16>(a, B, st);
    else if (dtype == MMOE_F16) launch_pool_fwd<__half>(a, B, st);
    else launch_pool_fwd<float>(a, B, st);
    MMOE_LAUNCH_OK("pool_fwd_kernel");
  }
  if (c->home) return 0;
  // normed = LN(pooled); out = normed + drop(W2 drop(gelu(W1 normed)))     (model.py:450-451)
  MMOE_TRY(layernorm_fwd(s.pooled, MMOE_F32, (const float*)P[ix.norm_w], (const float*)P[ix.norm_b], s.normed_t, s.normed_f,
                         s.st, B, d, dtype, st));
  {
    mmoe_epilogue e = epi_none();
    e.out = s.h; e.out_dtype = dtype; e.ldo = ff; e.bias = (const float*)P[ix.m0_b]; e.act = 2; e.preact = s.z;
    site_keys(c->seed, 102, &k0, &k1);
    e.drop_p = drop_p; e.drop_key0 = k0; e.drop_key1 = k1;
    mmoe_gemm_problem p = linear_fwd(s.normed_t, d, P[ix.m0_w], B, ff, d, e);
    MMOE_TRY(gemm_grouped(&p, 1, dtype, 0, st));
  }
  {
    mmoe_epilogue e = epi_none();
    e.out = out; e.out_dtype = MMOE_F32; e.ldo = d; e.bias = (const float*)P[ix.m3_b]; e.residual = s.normed_f; e.ld_res = d;
    site_keys(c->seed, 103, &k0, &k1);
    e.drop_p = drop_p; e.drop_key0 = k0; e.drop_key1 = k1;
    mmoe_gemm_problem p = linear_fwd(s.h, ff, P[ix.m3_w], B, d, ff, e);
    MMOE_TRY(gemm_grouped(&p, 1, dtype, 0, st));
  }
  return 0;
}

// stage: -1 = whole backward; 0 = tail (MLP, LN, pooling, cross attention: leaves dU / dI in the workspace);
// 100 + l = encoder layer l of the user stack; 200 + l = encoder layer l of the item stack.  Stages must run tail
// first, then each stack from its last layer down; the same workspace (and zeroed grads) must be passed to all of them.
// Splitting lets the caller hand finished parameter gradients to DDP's bucketed all-reduce while later stages still run.
extern "C" int mmoe_cross_bwd_stage(const mmoe_call* c, const mmoe_cross_cfg* cfg, int stage, const float* user, const uint8_t* user_mask,
                                    const float* item, const uint8_t* item_mask, const float* dout, float* d_user, float* d_item) {
  MMOE_TRY(check_cfg(cfg, c->B));
  if (c->B == 0) return 0;
  const int B = c->B, S = cfg->S, d = cfg->d, ff = 4 * d, dtype = c->dtype;
  const int M = B * S;
  const size_t es = dtype_size(dtype);
  cudaStream_t st = (cudaStream_t)c->stream;
  MMOE_CHECK(c->saved != nullptr && c->saved_bytes >= mmoe_cross_saved_bytes(cfg, B, dtype), "cross_bwd: saved blob too small");
  MMOE_CHECK(c->workspace != nullptr && c->workspace_bytes >= mmoe_cross_workspace_bytes(cfg, B, dtype), "cross_bwd: workspace too small");
  Arena A(c->saved);
  CrossSaved s = cross_layout(A, *cfg, B, dtype, c->home);
  Arena W(c->workspace);
  CrossScratch t = cross_scratch_layout(W, *cfg, B, dtype);
  const CrossIdx ix = cross_idx(cfg->n_layer);
  const void* const* P = c->params;
  void* const* G = c->grads;
  const float drop_p = c->training ? c->drop_p : 0.f;
  uint32_t k0, k1;
  const float* U = s.U;

  const bool do_tail = stage == -1 || stage == 0;
  const float* dpooled = dout;
  if (do_tail && !c->home) {
    // out = normed + drop(h W2^T + b2)
    site_keys(c->seed, 103, &k0, &k1);
    MMOE_TRY(cast_drop_colsum(dout, t.g, (float*)G[ix.m3_b], B, d, drop_p, k0, k1, dtype, st));
    {
      mmoe_epilogue e = epi_none();
      e.out = t.dz; e.out_dtype = dtype; e.ldo = ff; e.bwd_mode = 2; e.aux = s.z; e.ld_aux = ff; e.colsum = (float*)G[ix.m0_b];
      site_keys(c->seed, 102, &k0, &k1);
      e.drop_p = drop_p; e.drop_key0 = k0; e.drop_key1 = k1;
      mmoe_gemm_problem p[2] = {linear_dgrad(t.g, d, P[ix.m3_w], B, d, ff, e), linear_wgrad(t.g, d, s.h, ff, (float*)G[ix.m3_w], B, d, ff)};
      MMOE_TRY(gemm_grouped(p, 2, dtype, 0, st));
    }
    {
      mmoe_epilogue e = epi_none();   // dnormed = dout + dz W1
      e.out = t.dnormed; e.out_dtype = MMOE_F32; e.ldo = d; e.residual = dout; e.ld_res = d;
      mmoe_gemm_problem p[2] = {linear_dgrad(t.dz, ff, P[ix.m0_w], B, ff, d, e),
                                linear_wgrad(t.dz, ff, s.normed_t, d, (float*)G[ix.m0_w], B, ff, d)};
      MMOE_TRY(gemm_grouped(p, 2, dtype, 0, st));
    }
    {
      LnBwdArgs a{};
      a.dy = t.dnormed; a.dy_dtype = MMOE_F32; a.x = s.pooled; a.x_dtype = MMOE_F32; a.stats = s.st;
      a.gamma = (const float*)P[ix.norm_w]; a.dx = t.dpooled; a.dgamma = (float*)G[ix.norm_w]; a.dbeta = (float*)G[ix.norm_b];
      a.rows = B; a.d = d; a.dtype = dtype;
      MMOE_TRY(layernorm_bwd(a, st));
    }
    dpooled = t.dpooled;
  }
  if (do_tail) {
  {
    PoolDev a{};
    a.U = U; a.C = s.c; a.gate = (const float*)P[ix.gate]; a.query = (const float*)P[ix.query]; a.mask = user_mask;
    a.w_saved = s.w_pool; a.S = S; a.d = d; a.home = c->home;
    a.dpooled = dpooled; a.dU = t.dU; a.dC = t.dC; a.dquery = (float*)G[ix.query]; a.dgate = (float*)G[ix.gate];
    a.dbo = (float*)G[ix.ca_ob];
    site_keys(c->seed, 101, &k0, &k1);
    a.thresh = drop_p > 0.f ? drop_threshold(drop_p) : 0u; a.drop_scale = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
    a.k0 = k0; a.k1 = k1;
    if (dtype == MMOE_BF16) launch_pool_bwd<__nv_bfloat16>(a, B, st);
    else if (dtype == MMOE_F16) launch_pool_bwd<__half>(a, B, st);
    else launch_pool_bwd<float>(a, B, st);
    MMOE_LAUNCH_OK("pool_bwd_kernel");
  }
  const char* w_in = (const char*)P[ix.ca_w];
  float* g_w_in = (float*)G[ix.ca_w];
  float* g_b_in = (float*)G[ix.ca_b];
  {
    mmoe_epilogue e = epi_none();
    e.out = t.dctx; e.out_dtype = dtype; e.ldo = d;
    mmoe_gemm_problem p[2] = {linear_dgrad(t.dC, d, P[ix.ca_ow], M, d, d, e), linear_wgrad(t.dC, d, s.ctx, d, (float*)G[ix.ca_ow], M, d, d)};
    MMOE_TRY(gemm_grouped(p, 2, dtype, 0, st));
  }
  {
    AttnArgs a{};
    a.q = s.q; a.ldq = d; a.k = s.kv; a.v = (const char*)s.kv + (size_t)d * es; a.ldk = a.ldv = 2 * d;
    a.mask = item_mask; a.ctx = t.dctx; a.ldc = d;
    a.dq = t.dq; a.dk = t.dkv; a.dv = (char*)t.dkv + (size_t)d * es;
    a.bgq = g_b_in; a.bgk = g_b_in + d; a.bgv = g_b_in + 2 * d;
    a.B = B; a.Sq = S; a.Sk = S; a.H = cfg->n_head; a.hd = d / cfg->n_head;
    site_keys(c->seed, 100, &k0, &k1);
    a.drop_p = drop_p; a.k0 = k0; a.k1 = k1; a.dtype = dtype;
    MMOE_TRY(attention_bwd(a, st));
  }
  {
    // 16-bit outputs through the TMA-store epilogue (the fp32 read-modify-write epilogue made this launch epilogue-bound:
    // 402 us for 232 GFLOP); the encoder stacks take them as an extra summand of their output gradient
    mmoe_epilogue eu = epi_none(), ei = epi_none();
    eu.out = t.dUq; eu.out_dtype = dtype; eu.ldo = d;       // dq Wq   (dU = pooling gradient + this)
    ei.out = t.dIt; ei.out_dtype = dtype; ei.ldo = d;       // dkv Wkv (= dI)
    mmoe_gemm_problem p[4] = {
        linear_dgrad(t.dq, d, w_in, M, d, d, eu),
        linear_dgrad(t.dkv, 2 * d, w_in + (size_t)d * d * es, M, 2 * d, d, ei),
        linear_wgrad(t.dq, d, s.Ut, d, g_w_in, M, d, d),
        linear_wgrad(t.dkv, 2 * d, s.It, d, g_w_in + (size_t)d * d, M, 2 * d, d)};
    MMOE_TRY(gemm_grouped(p, 4, dtype, 0, st));
  }
  }
  EncCtx ec{};
  ec.dtype = dtype; ec.M = M; ec.Bseq = B; ec.S = S; ec.d = d; ec.ff = ff; ec.H = cfg->n_head;
  ec.drop_p = drop_p; ec.seed = c->seed; ec.stream = st;
  // item stack first: DDP reduces buckets in reverse registration order (mlp/cross_attn, self_item, self_user)
  for (int l = cfg->n_layer - 1; l >= 0; --l) {
    if (!(stage == -1 || stage == 200 + l)) continue;
    ec.mask = item_mask; ec.site0 = 16 * l + 8;
    const float* x_in = l == 0 ? item : s.item[l].x_sum;
    EncHandoff below{t.hand_i, l > 0 ? enc_g(G + ix.item0 + 12 * (l - 1)).b2 : nullptr, (uint32_t)(16 * (l - 1) + 8)};
    const bool top = l == cfg->n_layer - 1;     // the top layer's output gradient is the 16-bit dkv Wkv alone
    MMOE_TRY(enc_bwd(ec, enc_w(P + ix.item0 + 12 * l), enc_g(G + ix.item0 + 12 * l), x_in, s.item[l], t.enc, top ? nullptr : t.dI,
                     l == 0 ? d_item : t.dI, top ? nullptr : t.hand_i, l > 0 ? &below : nullptr, top ? t.dIt : nullptr));
  }
  for (int l = cfg->n_layer - 1; l >= 0; --l) {
    if (!(stage == -1 || stage == 100 + l)) continue;
    ec.mask = user_mask; ec.site0 = 16 * l;
    const float* x_in = l == 0 ? user : s.user[l].x_sum;
    EncHandoff below{t.hand_u, l > 0 ? enc_g(G + ix.user0 + 12 * (l - 1)).b2 : nullptr, (uint32_t)(16 * (l - 1))};
    const bool top = l == cfg->n_layer - 1;     // top layer: pooling gradient (fp32) + 16-bit dq Wq
    MMOE_TRY(enc_bwd(ec, enc_w(P + ix.user0 + 12 * l), enc_g(G + ix.user0 + 12 * l), x_in, s.user[l], t.enc, t.dU,
                     l == 0 ? d_user : t.dU, top ? nullptr : t.hand_u, l > 0 ? &below : nullptr, top ? t.dUq : nullptr));
  }
  return 0;
}

extern "C" int mmoe_cross_bwd(const mmoe_call* c, const mmoe_cross_cfg* cfg, const float* user, const uint8_t* user_mask,
                              const float* item, const uint8_t* item_mask, const float* dout, float* d_user, float* d_item) {
  return mmoe_cross_bwd_stage(c, cfg, -1, user, user_mask, item, item_mask, dout, d_user, d_item);
}
