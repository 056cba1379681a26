// Fused masked-softmax attention core, forward and backward, one CTA per (batch, head).
// Q/K/V tiles (S <= 64 sentence vectors x head_dim <= 128) are staged in shared memory,
// scores, mask, softmax, dropout and P.V never leave the SM.  Backward recomputes the
// probabilities from Q,K (nothing S x S is ever stored in HBM) and regenerates the dropout
// mask from its key.  Semantics: nn.MultiheadAttention: scores =
// (q * hd^-0.5) k^T, -inf on padded keys, softmax, dropout on the probabilities, P v.
// A row whose keys are all padded yields NaN, as in the reference.
//
// This is the exact-arithmetic (fp32 FFMA) implementation, used for MMOE_F32 and as the
// cross-check of the tensor-core variant.
#include <cstdlib>

#include "kernels.cuh"

namespace mmoe {

constexpr int ATT_THREADS = 256;

struct AttnDev {
  const void *q, *k, *v; int64_t ldq, ldk, ldv;
  const uint8_t* mask;
  void* ctx; int64_t ldc;
  void *dq, *dk, *dv;
  float *bgq, *bgk, *bgv;
  int B, Sq, Sk, H, hd;
  float qscale, drop_scale; uint32_t thresh, k0, k1;
};

template <typename T>
__device__ __forceinline__ void load_tile(float* dst, int pitch, const T* src, int64_t ld, int rows, int hd) {
  for (int e = threadIdx.x; e < rows * hd; e += ATT_THREADS) {
    const int r = e / hd, c = e - r * hd;
    dst[r * pitch + c] = to_f<T>(src[(int64_t)r * ld + c]);
  }
}

// scores + softmax into P (pre-dropout probabilities); returns nothing, P[Sq][pp]
__device__ __forceinline__ void scores_softmax(const float* Qs, const float* Ks, float* P, const uint8_t* mrow, int Sq, int Sk,
                                               int hd, int pt, int pp, float qscale) {
  for (int e = threadIdx.x; e < Sq * Sk; e += ATT_THREADS) {
    const int i = e / Sk, j = e - i * Sk;
    float acc = 0.f;
    const float* qi = Qs + i * pt;
    const float* kj = Ks + j * pt;
#pragma unroll 8
    for (int c = 0; c < hd; ++c) acc = fmaf(qi[c], kj[c], acc);
    acc *= qscale;
    if (mrow != nullptr && mrow[j]) acc = -INFINITY;
    P[i * pp + j] = acc;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = warp; i < Sq; i += ATT_THREADS / 32) {
    float m = -INFINITY;
    for (int j = lane; j < Sk; j += 32) m = fmaxf(m, P[i * pp + j]);
    m = warp_max(m);
    float s = 0.f;
    for (int j = lane; j < Sk; j += 32) {
      const float e = expf(P[i * pp + j] - m);   // all -inf row: (-inf) - (-inf) = NaN, like torch
      P[i * pp + j] = e;
      s += e;
    }
    s = warp_sum(s);
    const float inv = 1.0f / s;
    for (int j = lane; j < Sk; j += 32) P[i * pp + j] *= inv;
  }
  __syncthreads();
}

template <typename T>
__global__ void __launch_bounds__(ATT_THREADS) attn_fwd_kernel(const AttnDev a) {
  extern __shared__ float sm[];
  const int b = blockIdx.x / a.H, h = blockIdx.x - b * a.H;
  const int hd = a.hd, Sq = a.Sq, Sk = a.Sk, pt = hd + 1, pp = Sk + 1;
  float* Qs = sm;
  float* Ks = Qs + Sq * pt;
  float* Vs = Ks + Sk * pt;
  float* P = Vs + Sk * pt;
  load_tile<T>(Qs, pt, (const T*)a.q + (int64_t)b * Sq * a.ldq + h * hd, a.ldq, Sq, hd);
  load_tile<T>(Ks, pt, (const T*)a.k + (int64_t)b * Sk * a.ldk + h * hd, a.ldk, Sk, hd);
  load_tile<T>(Vs, pt, (const T*)a.v + (int64_t)b * Sk * a.ldv + h * hd, a.ldv, Sk, hd);
  __syncthreads();
  scores_softmax(Qs, Ks, P, a.mask ? a.mask + (int64_t)b * Sk : nullptr, Sq, Sk, hd, pt, pp, a.qscale);
  if (a.thresh != 0) {
    const uint64_t base = (uint64_t)blockIdx.x * Sq * Sk;
    for (int e = threadIdx.x; e < Sq * Sk; e += ATT_THREADS) {
      const int i = e / Sk, j = e - i * Sk;
      const float p = P[i * pp + j];
      P[i * pp + j] = drop_keep(a.k0, a.k1, base + e, a.thresh) ? p * a.drop_scale : 0.f;
    }
    __syncthreads();
  }
  T* out = (T*)a.ctx + (int64_t)b * Sq * a.ldc + h * hd;
  for (int e = threadIdx.x; e < Sq * hd; e += ATT_THREADS) {
    const int i = e / hd, c = e - i * hd;
    float acc = 0.f;
    for (int j = 0; j < Sk; ++j) acc = fmaf(P[i * pp + j], Vs[j * pt + c], acc);
    out[(int64_t)i * a.ldc + c] = from_f<T>(acc);
  }
}

template <typename T>
__global__ void __launch_bounds__(ATT_THREADS) attn_bwd_kernel(const AttnDev a) {
  extern __shared__ float sm[];
  const int b = blockIdx.x / a.H, h = blockIdx.x - b * a.H;
  const int hd = a.hd, Sq = a.Sq, Sk = a.Sk, pt = hd + 1, pp = Sk + 1;
  float* Qs = sm;
  float* Ks = Qs + Sq * pt;
  float* Vs = Ks + Sk * pt;
  float* dOs = Vs + Sk * pt;
  float* P = dOs + Sq * pt;       // softmax probabilities (pre-dropout)
  float* Pd = P + Sq * pp;        // probabilities after dropout
  float* dS = Pd + Sq * pp;       // dP, then dS
  float* bg = dS + Sq * pp;       // [3][hd] bias-grad partials
  load_tile<T>(Qs, pt, (const T*)a.q + (int64_t)b * Sq * a.ldq + h * hd, a.ldq, Sq, hd);
  load_tile<T>(Ks, pt, (const T*)a.k + (int64_t)b * Sk * a.ldk + h * hd, a.ldk, Sk, hd);
  load_tile<T>(Vs, pt, (const T*)a.v + (int64_t)b * Sk * a.ldv + h * hd, a.ldv, Sk, hd);
  load_tile<T>(dOs, pt, (const T*)a.ctx + (int64_t)b * Sq * a.ldc + h * hd, a.ldc, Sq, hd);
  for (int e = threadIdx.x; e < 3 * hd; e += ATT_THREADS) bg[e] = 0.f;
  __syncthreads();
  scores_softmax(Qs, Ks, P, a.mask ? a.mask + (int64_t)b * Sk : nullptr, Sq, Sk, hd, pt, pp, a.qscale);
  // dP = dropmask * (dO V^T); Pd = dropmask * P
  {
    const uint64_t base = (uint64_t)blockIdx.x * Sq * Sk;
    for (int e = threadIdx.x; e < Sq * Sk; e += ATT_THREADS) {
      const int i = e / Sk, j = e - i * Sk;
      float acc = 0.f;
      const float* di = dOs + i * pt;
      const float* vj = Vs + j * pt;
#pragma unroll 8
      for (int c = 0; c < hd; ++c) acc = fmaf(di[c], vj[c], acc);
      float p = P[i * pp + j];
      if (a.thresh != 0) {
        const bool keep = drop_keep(a.k0, a.k1, base + e, a.thresh);
        acc = keep ? acc * a.drop_scale : 0.f;
        p = keep ? p * a.drop_scale : 0.f;
      }
      dS[i * pp + j] = acc;
      Pd[i * pp + j] = p;
    }
  }
  __syncthreads();
  // dS = P * (dP - sum_j P dP)
  {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = warp; i < Sq; i += ATT_THREADS / 32) {
      float dl = 0.f;
      for (int j = lane; j < Sk; j += 32) dl += P[i * pp + j] * dS[i * pp + j];
      dl = warp_sum(dl);
      for (int j = lane; j < Sk; j += 32) dS[i * pp + j] = P[i * pp + j] * (dS[i * pp + j] - dl);
    }
  }
  __syncthreads();
  T* dq = (T*)a.dq + (int64_t)b * Sq * a.ldq + h * hd;
  T* dk = (T*)a.dk + (int64_t)b * Sk * a.ldk + h * hd;
  T* dv = (T*)a.dv + (int64_t)b * Sk * a.ldv + h * hd;
  for (int e = threadIdx.x; e < Sq * hd; e += ATT_THREADS) {      // dQ = qscale * dS K
    const int i = e / hd, c = e - i * hd;
    float acc = 0.f;
    for (int j = 0; j < Sk; ++j) acc = fmaf(dS[i * pp + j], Ks[j * pt + c], acc);
    const T o = from_f<T>(acc * a.qscale);
    dq[(int64_t)i * a.ldq + c] = o;
    if (a.bgq != nullptr) atomicAdd(&bg[c], to_f<T>(o));
  }
  for (int e = threadIdx.x; e < Sk * hd; e += ATT_THREADS) {      // dK = qscale * dS^T Q ; dV = Pd^T dO
    const int j = e / hd, c = e - j * hd;
    float ak = 0.f, av = 0.f;
    for (int i = 0; i < Sq; ++i) {
      ak = fmaf(dS[i * pp + j], Qs[i * pt + c], ak);
      av = fmaf(Pd[i * pp + j], dOs[i * pt + c], av);
    }
    const T ok = from_f<T>(ak * a.qscale), ov = from_f<T>(av);
    dk[(int64_t)j * a.ldk + c] = ok;
    dv[(int64_t)j * a.ldv + c] = ov;
    if (a.bgk != nullptr) atomicAdd(&bg[hd + c], to_f<T>(ok));
    if (a.bgv != nullptr) atomicAdd(&bg[2 * hd + c], to_f<T>(ov));
  }
  __syncthreads();
  for (int c = threadIdx.x; c < hd; c += ATT_THREADS) {
    if (a.bgq != nullptr) atomicAdd(a.bgq + h * hd + c, bg[c]);
    if (a.bgk != nullptr) atomicAdd(a.bgk + h * hd + c, bg[hd + c]);
    if (a.bgv != nullptr) atomicAdd(a.bgv + h * hd + c, bg[2 * hd + c]);
  }
}


// ------------------------------------------------------------------------------------------
// Short sequences (S <= 4: the two-token sequences of EnhancedCrossFuse, model.py:496-499): one WARP per
// (sample, head), lanes across head_dim, everything in registers, fp32 math for every dtype.  A block takes
// 8 samples of one head so the bias-gradient column sums reduce in shared memory before touching global memory.
// ------------------------------------------------------------------------------------------
constexpr int SMALL_S = 4, SMALL_C = 4;   // S <= 4, head_dim <= 128

template <typename T>
__global__ void __launch_bounds__(256) attn_small_kernel(const AttnDev a, int bwd) {
  __shared__ float red[8][3 * 128];
  pdl_entry();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.x * 8 + warp, h = blockIdx.y;
  const int hd = a.hd, Sq = a.Sq, Sk = a.Sk;
  const bool live = b < a.B;
  float q[SMALL_S][SMALL_C], k[SMALL_S][SMALL_C], v[SMALL_S][SMALL_C], d_o[SMALL_S][SMALL_C];
#pragma unroll
  for (int i = 0; i < SMALL_S; ++i)
#pragma unroll
    for (int c = 0; c < SMALL_C; ++c) {
      const int col = lane + 32 * c;
      const bool okc = live && col < hd;
      q[i][c] = (okc && i < Sq) ? to_f<T>(((const T*)a.q)[((int64_t)b * Sq + i) * a.ldq + h * hd + col]) : 0.f;
      k[i][c] = (okc && i < Sk) ? to_f<T>(((const T*)a.k)[((int64_t)b * Sk + i) * a.ldk + h * hd + col]) : 0.f;
      v[i][c] = (okc && i < Sk) ? to_f<T>(((const T*)a.v)[((int64_t)b * Sk + i) * a.ldv + h * hd + col]) : 0.f;
      d_o[i][c] = (bwd && okc && i < Sq) ? to_f<T>(((const T*)a.ctx)[((int64_t)b * Sq + i) * a.ldc + h * hd + col]) : 0.f;
    }
  float p[SMALL_S][SMALL_S], pd[SMALL_S][SMALL_S];
  const uint64_t base = ((uint64_t)b * a.H + h) * Sq * Sk;
#pragma unroll
  for (int i = 0; i < SMALL_S; ++i) {
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < SMALL_S; ++j) {
      float acc = 0.f;
#pragma unroll
      for (int c = 0; c < SMALL_C; ++c) acc = fmaf(q[i][c], k[j][c], acc);
      acc = warp_sum(acc) * a.qscale;
      if (j >= Sk || (live && a.mask != nullptr && a.mask[(int64_t)b * Sk + j])) acc = -INFINITY;
      p[i][j] = acc;
      mx = fmaxf(mx, acc);
    }
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < SMALL_S; ++j) { p[i][j] = expf(p[i][j] - mx); sum += p[i][j]; }   // all-masked row -> NaN, as torch
    const float inv = 1.f / sum;
#pragma unroll
    for (int j = 0; j < SMALL_S; ++j) {
      p[i][j] *= inv;
      pd[i][j] = p[i][j];
      if (a.thresh != 0) pd[i][j] = drop_keep(a.k0, a.k1, base + (uint64_t)i * Sk + j, a.thresh) ? p[i][j] * a.drop_scale : 0.f;
    }
  }
  if (!bwd) {
#pragma unroll
    for (int i = 0; i < SMALL_S; ++i)
#pragma unroll
      for (int c = 0; c < SMALL_C; ++c) {
        const int col = lane + 32 * c;
        if (live && i < Sq && col < hd) {
          float acc = 0.f;
#pragma unroll
          for (int j = 0; j < SMALL_S; ++j) acc = fmaf(pd[i][j], v[j][c], acc);
          ((T*)a.ctx)[((int64_t)b * Sq + i) * a.ldc + h * hd + col] = from_f<T>(acc);
        }
      }
    return;
  }
  // backward
  float ds[SMALL_S][SMALL_S];
#pragma unroll
  for (int i = 0; i < SMALL_S; ++i) {
    float dp[SMALL_S], dl = 0.f;
#pragma unroll
    for (int j = 0; j < SMALL_S; ++j) {
      float acc = 0.f;
#pragma unroll
      for (int c = 0; c < SMALL_C; ++c) acc = fmaf(d_o[i][c], v[j][c], acc);
      acc = warp_sum(acc);
      if (a.thresh != 0) acc = drop_keep(a.k0, a.k1, base + (uint64_t)i * Sk + j, a.thresh) ? acc * a.drop_scale : 0.f;
      dp[j] = acc;
      dl += p[i][j] * acc;
    }
#pragma unroll
    for (int j = 0; j < SMALL_S; ++j) ds[i][j] = (i < Sq && j < Sk) ? p[i][j] * (dp[j] - dl) * a.qscale : 0.f;
  }
  float sq[SMALL_C], sk[SMALL_C], sv[SMALL_C];
#pragma unroll
  for (int c = 0; c < SMALL_C; ++c) sq[c] = sk[c] = sv[c] = 0.f;
#pragma unroll
  for (int i = 0; i < SMALL_S; ++i)
#pragma unroll
    for (int c = 0; c < SMALL_C; ++c) {
      const int col = lane + 32 * c;
      if (!(live && col < hd)) continue;
      if (i < Sq) {
        float acc = 0.f;
#pragma unroll
        for (int j = 0; j < SMALL_S; ++j) acc = fmaf(ds[i][j], k[j][c], acc);
        const T o = from_f<T>(acc);
        ((T*)a.dq)[((int64_t)b * Sq + i) * a.ldq + h * hd + col] = o;
        sq[c] += to_f<T>(o);
      }
      if (i < Sk) {           // here i indexes a key
        float ak = 0.f, av = 0.f;
#pragma unroll
        for (int r = 0; r < SMALL_S; ++r) { ak = fmaf(ds[r][i], q[r][c], ak); av = fmaf(pd[r][i], d_o[r][c], av); }
        const T ok = from_f<T>(ak), ov = from_f<T>(av);
        ((T*)a.dk)[((int64_t)b * Sk + i) * a.ldk + h * hd + col] = ok;
        ((T*)a.dv)[((int64_t)b * Sk + i) * a.ldv + h * hd + col] = ov;
        sk[c] += to_f<T>(ok); sv[c] += to_f<T>(ov);
      }
    }
  if (a.bgq == nullptr && a.bgk == nullptr && a.bgv == nullptr) return;
#pragma unroll
  for (int c = 0; c < SMALL_C; ++c) {
    const int col = lane + 32 * c;
    red[warp][col] = sq[c]; red[warp][128 + col] = sk[c]; red[warp][256 + col] = sv[c];
  }
  __syncthreads();
  for (int e = threadIdx.x; e < 3 * 128; e += 256) {
    const int which = e >> 7, col = e & 127;
    if (col >= hd) continue;
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w][e];
    float* dst = which == 0 ? a.bgq : (which == 1 ? a.bgk : a.bgv);
    if (dst != nullptr) atomicAdd(dst + h * hd + col, t);
  }
}

template <typename T>
static int launch_small(const AttnDev& d, int bwd, cudaStream_t s) {
  dim3 grid((unsigned)((d.B + 7) / 8), (unsigned)d.H);
  MMOE_CUDA(launch_pdl(attn_small_kernel<T>, grid, 256, 0, s, d, bwd));
  MMOE_LAUNCH_OK("attn_small_kernel");
  return 0;
}
static bool small_ok(const AttnArgs& a) { return a.Sq <= SMALL_S && a.Sk <= SMALL_S && a.hd <= 32 * SMALL_C; }

static int fill(AttnDev* d, const AttnArgs& a) {
  MMOE_CHECK(a.Sq >= 1 && a.Sq <= 64 && a.Sk >= 1 && a.Sk <= 64, "attention: sequence lengths must be in [1,64] (got %d,%d)", a.Sq, a.Sk);
  MMOE_CHECK(a.hd >= 1 && a.hd <= 128, "attention: head_dim must be <= 128 (got %d)", a.hd);
  d->q = a.q; d->k = a.k; d->v = a.v; d->ldq = a.ldq; d->ldk = a.ldk; d->ldv = a.ldv;
  d->mask = a.mask; d->ctx = a.ctx; d->ldc = a.ldc; d->dq = a.dq; d->dk = a.dk; d->dv = a.dv;
  d->bgq = a.bgq; d->bgk = a.bgk; d->bgv = a.bgv;
  d->B = a.B; d->Sq = a.Sq; d->Sk = a.Sk; d->H = a.H; d->hd = a.hd;
  d->qscale = 1.0f / sqrtf((float)a.hd);
  d->thresh = a.drop_p > 0.f ? drop_threshold(a.drop_p) : 0u;
  d->drop_scale = a.drop_p > 0.f ? 1.f / (1.f - a.drop_p) : 1.f;
  d->k0 = a.k0; d->k1 = a.k1;
  return 0;
}

template <typename T>
static int launch_fwd(const AttnDev& d, cudaStream_t s) {
  const size_t smem = sizeof(float) * ((size_t)(d.Sq + 2 * d.Sk) * (d.hd + 1) + (size_t)d.Sq * (d.Sk + 1));
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    MMOE_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    configured = 160 * 1024;
  }
  attn_fwd_kernel<T><<<d.B * d.H, ATT_THREADS, smem, s>>>(d);
  MMOE_LAUNCH_OK("attn_fwd_kernel");
  return 0;
}
template <typename T>
static int launch_bwd(const AttnDev& d, cudaStream_t s) {
  const size_t smem = sizeof(float) * ((size_t)(2 * d.Sq + 2 * d.Sk) * (d.hd + 1) + 3 * (size_t)d.Sq * (d.Sk + 1) + 3 * d.hd);
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    MMOE_CUDA(cudaFuncSetAttribute(attn_bwd_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    configured = 200 * 1024;
  }
  attn_bwd_kernel<T><<<d.B * d.H, ATT_THREADS, smem, s>>>(d);
  MMOE_LAUNCH_OK("attn_bwd_kernel");
  return 0;
}

static bool force_exact() {
  static const bool f = getenv("MMOE_DEBUG_FORCE_SIMT") != nullptr;   // test-only cross-check switch
  return f;
}

int attention_fwd(const AttnArgs& a, cudaStream_t s) {
  if (a.B == 0) return 0;
  if (small_ok(a)) {
    AttnDev d;
    MMOE_TRY(fill(&d, a));
    if (a.dtype == MMOE_BF16) return launch_small<__nv_bfloat16>(d, 0, s);
    if (a.dtype == MMOE_F16) return launch_small<__half>(d, 0, s);
    return launch_small<float>(d, 0, s);
  }
  if (!force_exact() && attention_tc_supported(a)) return attention_tc(a, false, s);
  AttnDev d;
  MMOE_TRY(fill(&d, a));
  if (a.dtype == MMOE_BF16) return launch_fwd<__nv_bfloat16>(d, s);
  if (a.dtype == MMOE_F16) return launch_fwd<__half>(d, s);
  return launch_fwd<float>(d, s);
}
int attention_bwd(const AttnArgs& a, cudaStream_t s) {
  if (a.B == 0) return 0;
  if (small_ok(a)) {
    AttnDev d;
    MMOE_TRY(fill(&d, a));
    if (a.dtype == MMOE_BF16) return launch_small<__nv_bfloat16>(d, 1, s);
    if (a.dtype == MMOE_F16) return launch_small<__half>(d, 1, s);
    return launch_small<float>(d, 1, s);
  }
  if (!force_exact() && attention_tc_supported(a)) return attention_tc(a, true, s);
  AttnDev d;
  MMOE_TRY(fill(&d, a));
  if (a.dtype == MMOE_BF16) return launch_bwd<__nv_bfloat16>(d, s);
  if (a.dtype == MMOE_F16) return launch_bwd<__half>(d, s);
  return launch_bwd<float>(d, s);
}

}  // namespace mmoe

using namespace mmoe;

extern "C" int mmoe_attention_fwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                                  const uint8_t* key_padding_mask, void* ctx, int64_t ldc, int32_t B, int32_t Sq, int32_t Sk,
                                  int32_t n_head, int32_t hd, float drop_p, uint32_t key0, uint32_t key1, int dtype, void* stream) {
  AttnArgs a{};
  a.q = q; a.k = k; a.v = v; a.ldq = ldq; a.ldk = ldk; a.ldv = ldv; a.mask = key_padding_mask; a.ctx = ctx; a.ldc = ldc;
  a.B = B; a.Sq = Sq; a.Sk = Sk; a.H = n_head; a.hd = hd; a.drop_p = drop_p; a.k0 = key0; a.k1 = key1; a.dtype = dtype;
  return attention_fwd(a, (cudaStream_t)stream);
}
extern "C" int mmoe_attention_bwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                                  const uint8_t* key_padding_mask, const void* dctx, int64_t ldc, void* dq, void* dk, void* dv,
                                  float* bias_grad_q, float* bias_grad_k, float* bias_grad_v, int32_t B, int32_t Sq, int32_t Sk,
                                  int32_t n_head, int32_t hd, float drop_p, uint32_t key0, uint32_t key1, int dtype, void* stream) {
  AttnArgs a{};
  a.q = q; a.k = k; a.v = v; a.ldq = ldq; a.ldk = ldk; a.ldv = ldv; a.mask = key_padding_mask;
  a.ctx = const_cast<void*>(dctx); a.ldc = ldc; a.dq = dq; a.dk = dk; a.dv = dv;
  a.bgq = bias_grad_q; a.bgk = bias_grad_k; a.bgv = bias_grad_v;
  a.B = B; a.Sq = Sq; a.Sk = Sk; a.H = n_head; a.hd = hd; a.drop_p = drop_p; a.k0 = key0; a.k1 = key1; a.dtype = dtype;
  return attention_bwd(a, (cudaStream_t)stream);
}
