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

// ------------------------------------------------------------------------------------------
// Fused head kernels (TwoTaskMMoE, model.py:562-577).  One warp per sample, float4 lanes, warp-shuffle reductions.
//   forward : expert_vecs is read from HBM ONCE: mean query -> 2 gate GEMVs -> 2 softmaxes -> 2 weighted expert sums ->
//             2 tower LayerNorms; only the normalised 16-bit tower inputs (3 KB/sample), the query and a few scalars
//             are written.  Nothing fp32 of size [B,768] is materialised.
//   backward: LayerNorm backward of both towers, gate-weight / softmax backward, query backward and the expert-vector
//             gradient in one pass: reads d(xn) and expert_vecs, writes d(expert_vecs).
// Algorithmic bytes per sample: forward 18.4 KB in + 3 KB out; backward 18.4 + 3 KB in, 18.4 KB out.
// ------------------------------------------------------------------------------------------
struct HeadFusedDev {
  const float* ev; const float* wg[2]; const float* bg[2]; const float* ln_w[2]; const float* ln_b[2];
  void* query_t; float* w; float* stats; void* xn;
  const void* dxn; float* dl; float* d_ev; float* dgamma[2]; float* dbeta[2];
  int64_t B; int d, n;
};

template <typename T, int NV>
__global__ void __launch_bounds__(256) head_mix_ln_fwd_kernel(const HeadFusedDev a) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int d = a.d, n = a.n;
  const float inv_d = 1.f / (float)d, inv_n = 1.f / (float)n;
  for (int64_t b = warp0; b < a.B; b += nw) {
    const float* evb = a.ev + b * (int64_t)n * d;
    float4 q[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) q[i] = make_float4(0, 0, 0, 0);
    for (int k = 0; k < n; ++k) {
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c = (i * 32 + lane) * 4;
        if (c < d) { const float4 v = ld4(evb + k * d + c); q[i].x += v.x; q[i].y += v.y; q[i].z += v.z; q[i].w += v.w; }
      }
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      q[i].x *= inv_n; q[i].y *= inv_n; q[i].z *= inv_n; q[i].w *= inv_n;
      const int c = (i * 32 + lane) * 4;
      if (c < d) {
        T* qo = (T*)a.query_t + b * d + c;
        qo[0] = from_f<T>(q[i].x); qo[1] = from_f<T>(q[i].y); qo[2] = from_f<T>(q[i].z); qo[3] = from_f<T>(q[i].w);
      }
    }
    for (int t = 0; t < 2; ++t) {
      float wk[MIX_MAXN];
      float m = -INFINITY;
#pragma unroll
      for (int k = 0; k < MIX_MAXN; ++k) {
        wk[k] = -INFINITY;
        if (k < n) {
          float acc = 0.f;
#pragma unroll
          for (int i = 0; i < NV; ++i) {
            const int c = (i * 32 + lane) * 4;
            if (c < d) { const float4 w = ld4(a.wg[t] + k * d + c); acc += q[i].x * w.x + q[i].y * w.y + q[i].z * w.z + q[i].w * w.w; }
          }
          wk[k] = warp_sum(acc) + a.bg[t][k];
          m = fmaxf(m, wk[k]);
        }
      }
      float ssum = 0.f;
#pragma unroll
      for (int k = 0; k < MIX_MAXN; ++k) { wk[k] = k < n ? expf(wk[k] - m) : 0.f; ssum += wk[k]; }
      const float inv = 1.f / ssum;
      float4 f[NV];
#pragma unroll
      for (int i = 0; i < NV; ++i) f[i] = make_float4(0, 0, 0, 0);
      float sum = 0.f;
#pragma unroll
      for (int k = 0; k < MIX_MAXN; ++k) {
        if (k < n) {
          wk[k] *= inv;
          if (lane == 0) a.w[((int64_t)t * a.B + b) * n + k] = wk[k];
#pragma unroll
          for (int i = 0; i < NV; ++i) {
            const int c = (i * 32 + lane) * 4;
            if (c < d) { const float4 v = ld4(evb + k * d + c); f[i].x += wk[k] * v.x; f[i].y += wk[k] * v.y; f[i].z += wk[k] * v.z; f[i].w += wk[k] * v.w; }
          }
        }
      }
#pragma unroll
      for (int i = 0; i < NV; ++i) sum += f[i].x + f[i].y + f[i].z + f[i].w;      // lanes beyond d hold zeros
      const float mean = warp_sum(sum) * inv_d;
      float sq = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c = (i * 32 + lane) * 4;
        if (c < d) { const float x0 = f[i].x - mean, x1 = f[i].y - mean, x2 = f[i].z - mean, x3 = f[i].w - mean; sq += x0 * x0 + x1 * x1 + x2 * x2 + x3 * x3; }
      }
      const float rstd = rsqrtf(warp_sum(sq) * inv_d + 1e-5f);
      if (lane == 0) { a.stats[((int64_t)t * a.B + b) * 2] = mean; a.stats[((int64_t)t * a.B + b) * 2 + 1] = rstd; }
      T* xo = (T*)a.xn + ((int64_t)t * a.B + b) * d;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c = (i * 32 + lane) * 4;
        if (c < d) {
          const float4 g = ld4(a.ln_w[t] + c), be = ld4(a.ln_b[t] + c);
          xo[c + 0] = from_f<T>((f[i].x - mean) * rstd * g.x + be.x);
          xo[c + 1] = from_f<T>((f[i].y - mean) * rstd * g.y + be.y);
          xo[c + 2] = from_f<T>((f[i].z - mean) * rstd * g.z + be.z);
          xo[c + 3] = from_f<T>((f[i].w - mean) * rstd * g.w + be.w);
        }
      }
    }
  }
}

template <typename T, int NV>
__global__ void __launch_bounds__(128, 2) head_ln_mix_bwd_kernel(const HeadFusedDev a) {
  constexpr int WARPS = 4;
  __shared__ float red[WARPS][NV * 128];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t warp0 = (int64_t)blockIdx.x * WARPS + wib, nw = (int64_t)gridDim.x * WARPS;
  const int d = a.d, n = a.n;
  const float inv_d = 1.f / (float)d, inv_n = 1.f / (float)n;
  float4 acc_g[2][NV], acc_b[2][NV];
#pragma unroll
  for (int t = 0; t < 2; ++t)
#pragma unroll
    for (int i = 0; i < NV; ++i) acc_g[t][i] = acc_b[t][i] = make_float4(0, 0, 0, 0);
  for (int64_t b = warp0; b < a.B; b += nw) {
    const float* evb = a.ev + b * (int64_t)n * d;
    float4 df[2][NV], dq[NV];
    float wt[2][MIX_MAXN];
#pragma unroll
    for (int i = 0; i < NV; ++i) dq[i] = make_float4(0, 0, 0, 0);
#pragma unroll
    for (int t = 0; t < 2; ++t) {
#pragma unroll
      for (int k = 0; k < MIX_MAXN; ++k) wt[t][k] = k < n ? a.w[((int64_t)t * a.B + b) * n + k] : 0.f;
      // fused_t = sum_k w_t[k] E_k  -> xhat
      float4 f[NV];
#pragma unroll
      for (int i = 0; i < NV; ++i) f[i] = make_float4(0, 0, 0, 0);
#pragma unroll
      for (int k = 0; k < MIX_MAXN; ++k) {
        if (k < n) {
#pragma unroll
          for (int i = 0; i < NV; ++i) {
            const int c = (i * 32 + lane) * 4;
            if (c < d) { const float4 v = ld4(evb + k * d + c); f[i].x += wt[t][k] * v.x; f[i].y += wt[t][k] * v.y; f[i].z += wt[t][k] * v.z; f[i].w += wt[t][k] * v.w; }
          }
        }
      }
      const float mean = a.stats[((int64_t)t * a.B + b) * 2], rstd = a.stats[((int64_t)t * a.B + b) * 2 + 1];
      const T* dy = (const T*)a.dxn + ((int64_t)t * a.B + b) * d;
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c = (i * 32 + lane) * 4;
        df[t][i] = make_float4(0, 0, 0, 0);
        if (c < d) {
          const float4 g = ld4(a.ln_w[t] + c);
          const float y0 = to_f<T>(dy[c]), y1 = to_f<T>(dy[c + 1]), y2 = to_f<T>(dy[c + 2]), y3 = to_f<T>(dy[c + 3]);
          const float4 xh = make_float4((f[i].x - mean) * rstd, (f[i].y - mean) * rstd, (f[i].z - mean) * rstd, (f[i].w - mean) * rstd);
          acc_g[t][i].x += y0 * xh.x; acc_g[t][i].y += y1 * xh.y; acc_g[t][i].z += y2 * xh.z; acc_g[t][i].w += y3 * xh.w;
          acc_b[t][i].x += y0; acc_b[t][i].y += y1; acc_b[t][i].z += y2; acc_b[t][i].w += y3;
          const float4 dg = make_float4(y0 * g.x, y1 * g.y, y2 * g.z, y3 * g.w);
          s1 += dg.x + dg.y + dg.z + dg.w;
          s2 += dg.x * xh.x + dg.y * xh.y + dg.z * xh.z + dg.w * xh.w;
          df[t][i] = dg; f[i] = xh;          // keep dg in df, xhat in f until the row sums are known
        }
      }
      s1 = warp_sum(s1) * inv_d; s2 = warp_sum(s2) * inv_d;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c = (i * 32 + lane) * 4;
        if (c < d) {
          df[t][i].x = rstd * (df[t][i].x - s1 - f[i].x * s2);
          df[t][i].y = rstd * (df[t][i].y - s1 - f[i].y * s2);
          df[t][i].z = rstd * (df[t][i].z - s1 - f[i].z * s2);
          df[t][i].w = rstd * (df[t][i].w - s1 - f[i].w * s2);
        }
      }
      // gate backward: dw_k = <dfused, E_k>; softmax' ; dq += dl_k Wg_k
      float dw[MIX_MAXN], dot = 0.f;
#pragma unroll
      for (int k = 0; k < MIX_MAXN; ++k) {
        dw[k] = 0.f;
        if (k < n) {
          float acc = 0.f;
#pragma unroll
          for (int i = 0; i < NV; ++i) {
            const int c = (i * 32 + lane) * 4;
            if (c < d) { const float4 v = ld4(evb + k * d + c); acc += df[t][i].x * v.x + df[t][i].y * v.y + df[t][i].z * v.z + df[t][i].w * v.w; }
          }
          dw[k] = warp_sum(acc);
          dot += wt[t][k] * dw[k];
        }
      }
#pragma unroll
      for (int k = 0; k < MIX_MAXN; ++k) {
        if (k < n) {
          const float dlk = wt[t][k] * (dw[k] - dot);
          if (lane == 0) a.dl[((int64_t)t * a.B + b) * n + k] = dlk;
#pragma unroll
          for (int i = 0; i < NV; ++i) {
            const int c = (i * 32 + lane) * 4;
            if (c < d) { const float4 w = ld4(a.wg[t] + k * d + c); dq[i].x += dlk * w.x; dq[i].y += dlk * w.y; dq[i].z += dlk * w.z; dq[i].w += dlk * w.w; }
          }
        }
      }
    }
    // d E_k = w_good[k] dfused_good + w_best[k] dfused_best + dq / n
    float* dev = a.d_ev + b * (int64_t)n * d;
#pragma unroll
    for (int k = 0; k < MIX_MAXN; ++k) {
      if (k < n) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          const int c = (i * 32 + lane) * 4;
          if (c < d) {
            float4 v;
            v.x = wt[0][k] * df[0][i].x + wt[1][k] * df[1][i].x + inv_n * dq[i].x;
            v.y = wt[0][k] * df[0][i].y + wt[1][k] * df[1][i].y + inv_n * dq[i].y;
            v.z = wt[0][k] * df[0][i].z + wt[1][k] * df[1][i].z + inv_n * dq[i].z;
            v.w = wt[0][k] * df[0][i].w + wt[1][k] * df[1][i].w + inv_n * dq[i].w;
            st4(dev + k * d + c, v);
          }
        }
      }
    }
  }
  // tower LayerNorm weight / bias gradients: block reduction, one atomic per column per block
#pragma unroll
  for (int t = 0; t < 2; ++t) {
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c = (i * 32 + lane) * 4;
        if (c < d) *reinterpret_cast<float4*>(&red[wib][c]) = pass == 0 ? acc_g[t][i] : acc_b[t][i];
      }
      __syncthreads();
      float* dst = pass == 0 ? a.dgamma[t] : a.dbeta[t];
      for (int c = threadIdx.x; c < d; c += 128) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < WARPS; ++w) s += red[w][c];
        atomicAdd(dst + c, s);
      }
      __syncthreads();
    }
  }
}

// ------------------------------------------------------------------------------------------
// Register-resident variants for the common geometry (n_expert = NE, expert_dim <= NV*128): a warp first issues ALL of
// its sample's expert-vector loads (NE*NV independent 16-byte loads per lane = 18 KB in flight per warp), then does
// every pass from registers.  The v1 kernels above re-read expert_vecs through L1 for each pass and were latency-bound
// (forward 2.2 TB/s, backward 1.35 TB/s of algorithmic traffic at B = 65536).
// ------------------------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ void st4_as(T* p, float4 v);
template <> __device__ __forceinline__ void st4_as<float>(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
template <> __device__ __forceinline__ void st4_as<__nv_bfloat16>(__nv_bfloat16* p, float4 v) {
  const __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
  uint2 u; u.x = *reinterpret_cast<const uint32_t*>(&lo); u.y = *reinterpret_cast<const uint32_t*>(&hi);
  *reinterpret_cast<uint2*>(p) = u;
}
template <> __device__ __forceinline__ void st4_as<__half>(__half* p, float4 v) {
  const __half2 lo = __floats2half2_rn(v.x, v.y), hi = __floats2half2_rn(v.z, v.w);
  uint2 u; u.x = *reinterpret_cast<const uint32_t*>(&lo); u.y = *reinterpret_cast<const uint32_t*>(&hi);
  *reinterpret_cast<uint2*>(p) = u;
}
template <typename T> __device__ __forceinline__ float4 ld4_as(const T* p);
template <> __device__ __forceinline__ float4 ld4_as<float>(const float* p) { return *reinterpret_cast<const float4*>(p); }
template <> __device__ __forceinline__ float4 ld4_as<__nv_bfloat16>(const __nv_bfloat16* p) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  return make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xFFFF0000u), __uint_as_float(u.y << 16), __uint_as_float(u.y & 0xFFFF0000u));
}
template <> __device__ __forceinline__ float4 ld4_as<__half>(const __half* p) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
  return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ float dot4(float4 a, float4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
__device__ __forceinline__ void fma4(float4& acc, float s, float4 v) { acc.x = fmaf(s, v.x, acc.x); acc.y = fmaf(s, v.y, acc.y); acc.z = fmaf(s, v.z, acc.z); acc.w = fmaf(s, v.w, acc.w); }

template <typename T, int NE, int NV>
__global__ void __launch_bounds__(128, 2) head_fwd_reg_kernel(const HeadFusedDev a) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int d = a.d;
  const float inv_d = 1.f / (float)d, inv_n = 1.f / (float)NE;
  for (int64_t b = warp0; b < a.B; b += nw) {
    const float* evb = a.ev + b * (int64_t)NE * d;
    // the warp's NEXT sample (18 KB) goes to L2 through one bulk request while this one is computed from registers
    if (b + nw < a.B && lane == 0) prefetch_l2_bulk(a.ev + (b + nw) * (int64_t)NE * d, (uint32_t)(NE * d * 4));
    float4 E[NE][NV];
#pragma unroll
    for (int k = 0; k < NE; ++k)
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c = (i * 32 + lane) * 4;
        E[k][i] = c < d ? __ldcs(reinterpret_cast<const float4*>(evb + k * d + c)) : make_float4(0, 0, 0, 0);
      }
    float4 q[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      q[i] = E[0][i];
#pragma unroll
      for (int k = 1; k < NE; ++k) { q[i].x += E[k][i].x; q[i].y += E[k][i].y; q[i].z += E[k][i].z; q[i].w += E[k][i].w; }
      q[i].x *= inv_n; q[i].y *= inv_n; q[i].z *= inv_n; q[i].w *= inv_n;
      const int c = (i * 32 + lane) * 4;
      if (c < d) st4_as<T>((T*)a.query_t + b * d + c, q[i]);
    }
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      float wk[NE];
#pragma unroll
      for (int k = 0; k < NE; ++k) {
        float acc = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          const int c = (i * 32 + lane) * 4;
          if (c < d) acc += dot4(q[i], __ldg(reinterpret_cast<const float4*>(a.wg[t] + k * d + c)));
        }
        wk[k] = acc;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int k = 0; k < NE; ++k) wk[k] += __shfl_xor_sync(0xffffffffu, wk[k], o);
      float m = -INFINITY;
#pragma unroll
      for (int k = 0; k < NE; ++k) { wk[k] += a.bg[t][k]; m = fmaxf(m, wk[k]); }
      float ssum = 0.f;
#pragma unroll
      for (int k = 0; k < NE; ++k) { wk[k] = expf(wk[k] - m); ssum += wk[k]; }
      const float inv = 1.f / ssum;
#pragma unroll
      for (int k = 0; k < NE; ++k) wk[k] *= inv;
      if (lane < NE) {
        float mine = wk[0];
#pragma unroll
        for (int k = 1; k < NE; ++k) mine = lane == k ? wk[k] : mine;
        a.w[((int64_t)t * a.B + b) * NE + lane] = mine;
      }
      float4 f[NV];
      float sum = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        f[i] = make_float4(0, 0, 0, 0);
#pragma unroll
        for (int k = 0; k < NE; ++k) fma4(f[i], wk[k], E[k][i]);
        sum += f[i].x + f[i].y + f[i].z + f[i].w;
      }
      const float mean = warp_sum(sum) * inv_d;
      float sq = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c = (i * 32 + lane) * 4;
        if (c < d) { const float x0 = f[i].x - mean, x1 = f[i].y - mean, x2 = f[i].z - mean, x3 = f[i].w - mean; sq += x0 * x0 + x1 * x1 + x2 * x2 + x3 * x3; }
      }
      const float rstd = rsqrtf(warp_sum(sq) * inv_d + 1e-5f);
      if (lane == 0) { a.stats[((int64_t)t * a.B + b) * 2] = mean; a.stats[((int64_t)t * a.B + b) * 2 + 1] = rstd; }
      T* xo = (T*)a.xn + ((int64_t)t * a.B + b) * d;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c = (i * 32 + lane) * 4;
        if (c < d) {
          const float4 g = __ldg(reinterpret_cast<const float4*>(a.ln_w[t] + c)), be = __ldg(reinterpret_cast<const float4*>(a.ln_b[t] + c));
          st4_as<T>(xo + c, make_float4((f[i].x - mean) * rstd * g.x + be.x, (f[i].y - mean) * rstd * g.y + be.y,
                                        (f[i].z - mean) * rstd * g.z + be.z, (f[i].w - mean) * rstd * g.w + be.w));
        }
      }
    }
  }
}

// dynamic smem: per warp [2 tasks][dgamma | dbeta][NV*128] fp32 partial sums (warp-private columns: no atomics)
template <typename T, int NE, int NV>
__global__ void __launch_bounds__(128, 2) head_bwd_reg_kernel(const HeadFusedDev a) {
  constexpr int WARPS = 4, COLS = NV * 128;
  extern __shared__ float4 acc_sm4[];
  float* acc_sm = reinterpret_cast<float*>(acc_sm4);
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  float* mine = acc_sm + wib * (4 * COLS);
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int i = 0; i < NV; ++i) *reinterpret_cast<float4*>(mine + r * COLS + (i * 32 + lane) * 4) = make_float4(0, 0, 0, 0);
  const int64_t warp0 = (int64_t)blockIdx.x * WARPS + wib, nw = (int64_t)gridDim.x * WARPS;
  const int d = a.d;
  const float inv_d = 1.f / (float)d, inv_n = 1.f / (float)NE;
  for (int64_t b = warp0; b < a.B; b += nw) {
    const float* evb = a.ev + b * (int64_t)NE * d;
    if (b + nw < a.B) {
      if (lane == 0) prefetch_l2_bulk(a.ev + (b + nw) * (int64_t)NE * d, (uint32_t)(NE * d * 4));
      prefetch_l2_rows((const T*)a.dxn + (b + nw) * d, d * (int)sizeof(T), lane);
      prefetch_l2_rows((const T*)a.dxn + ((int64_t)a.B + b + nw) * d, d * (int)sizeof(T), lane);
    }
    float4 E[NE][NV];
#pragma unroll
    for (int k = 0; k < NE; ++k)
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c = (i * 32 + lane) * 4;
        E[k][i] = c < d ? __ldcs(reinterpret_cast<const float4*>(evb + k * d + c)) : make_float4(0, 0, 0, 0);
      }
    float wt[2][NE];
    float4 df[2][NV];
#pragma unroll
    for (int t = 0; t < 2; ++t) {
#pragma unroll
      for (int k = 0; k < NE; ++k) wt[t][k] = a.w[((int64_t)t * a.B + b) * NE + k];
      const float mean = a.stats[((int64_t)t * a.B + b) * 2], rstd = a.stats[((int64_t)t * a.B + b) * 2 + 1];
      const T* dy = (const T*)a.dxn + ((int64_t)t * a.B + b) * d;
      float4 dg[NV];
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c = (i * 32 + lane) * 4;
        float4 f = make_float4(0, 0, 0, 0);
#pragma unroll
        for (int k = 0; k < NE; ++k) fma4(f, wt[t][k], E[k][i]);
        dg[i] = make_float4(0, 0, 0, 0);
        df[t][i] = make_float4(0, 0, 0, 0);
        if (c < d) {
          const float4 y = ld4_as<T>(dy + c);
          const float4 g = __ldg(reinterpret_cast<const float4*>(a.ln_w[t] + c));
          const float4 xh = make_float4((f.x - mean) * rstd, (f.y - mean) * rstd, (f.z - mean) * rstd, (f.w - mean) * rstd);
          float4* pg = reinterpret_cast<float4*>(mine + (2 * t) * COLS + c);
          float4* pb = reinterpret_cast<float4*>(mine + (2 * t + 1) * COLS + c);
          float4 ag = *pg, ab = *pb;
          ag.x = fmaf(y.x, xh.x, ag.x); ag.y = fmaf(y.y, xh.y, ag.y); ag.z = fmaf(y.z, xh.z, ag.z); ag.w = fmaf(y.w, xh.w, ag.w);
          ab.x += y.x; ab.y += y.y; ab.z += y.z; ab.w += y.w;
          *pg = ag; *pb = ab;
          dg[i] = make_float4(y.x * g.x, y.y * g.y, y.z * g.z, y.w * g.w);
          s1 += dg[i].x + dg[i].y + dg[i].z + dg[i].w;
          s2 += dot4(dg[i], xh);
          df[t][i] = xh;
        }
      }
      s1 = warp_sum(s1) * inv_d; s2 = warp_sum(s2) * inv_d;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c = (i * 32 + lane) * 4;
        if (c < d) {
          df[t][i].x = rstd * (dg[i].x - s1 - df[t][i].x * s2);
          df[t][i].y = rstd * (dg[i].y - s1 - df[t][i].y * s2);
          df[t][i].z = rstd * (dg[i].z - s1 - df[t][i].z * s2);
          df[t][i].w = rstd * (dg[i].w - s1 - df[t][i].w * s2);
        }
      }
    }
    // gate backward: dw_t[k] = <dfused_t, E_k>  (12 dot products reduced together)
    float dw[2][NE];
#pragma unroll
    for (int t = 0; t < 2; ++t)
#pragma unroll
      for (int k = 0; k < NE; ++k) {
        float acc = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) acc += dot4(df[t][i], E[k][i]);
        dw[t][k] = acc;
      }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int t = 0; t < 2; ++t)
#pragma unroll
        for (int k = 0; k < NE; ++k) dw[t][k] += __shfl_xor_sync(0xffffffffu, dw[t][k], o);
    // expert_vecs is dead from here on
    float4 dq[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) dq[i] = make_float4(0, 0, 0, 0);
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      float dot = 0.f;
#pragma unroll
      for (int k = 0; k < NE; ++k) dot += wt[t][k] * dw[t][k];
      float sel = 0.f;
#pragma unroll
      for (int k = 0; k < NE; ++k) {
        const float dlk = wt[t][k] * (dw[t][k] - dot);
        sel = lane == k ? dlk : sel;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          const int c = (i * 32 + lane) * 4;
          if (c < d) fma4(dq[i], dlk, __ldg(reinterpret_cast<const float4*>(a.wg[t] + k * d + c)));
        }
      }
      if (lane < NE) a.dl[((int64_t)t * a.B + b) * NE + lane] = sel;
    }
    float* dev = a.d_ev + b * (int64_t)NE * d;
#pragma unroll
    for (int k = 0; k < NE; ++k)
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c = (i * 32 + lane) * 4;
        if (c < d) {
          float4 v;
          v.x = wt[0][k] * df[0][i].x + wt[1][k] * df[1][i].x + inv_n * dq[i].x;
          v.y = wt[0][k] * df[0][i].y + wt[1][k] * df[1][i].y + inv_n * dq[i].y;
          v.z = wt[0][k] * df[0][i].z + wt[1][k] * df[1][i].z + inv_n * dq[i].z;
          v.w = wt[0][k] * df[0][i].w + wt[1][k] * df[1][i].w + inv_n * dq[i].w;
          __stcs(reinterpret_cast<float4*>(dev + k * d + c), v);
        }
      }
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < 4 * COLS; idx += 128) {
    const int r = idx / COLS, c = idx - r * COLS;
    if (c < d) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < WARPS; ++w) s += acc_sm[w * (4 * COLS) + idx];
      atomicAdd(((r & 1) ? a.dbeta[r >> 1] : a.dgamma[r >> 1]) + c, s);
    }
  }
}

struct HeadSaved { void* query; float *w, *st; void *xn, *z1, *a1, *z2, *a2; };
static HeadSaved head_layout(Arena& A, const mmoe_head_cfg& cfg, int B, int dtype) {
  HeadSaved s{};
  const size_t es = dtype_size(dtype), Bz = (size_t)B; const int d = cfg.d, h1 = cfg.hidden, h2 = cfg.hidden / 2;
  s.query = A.take(Bz * d * es); s.w = (float*)A.take(2 * Bz * cfg.n_expert * 4);
  s.st = (float*)A.take(2 * Bz * 2 * 4); s.xn = A.take(2 * Bz * d * es);
  s.z1 = A.take(2 * Bz * h1 * es); s.a1 = A.take(2 * Bz * h1 * es); s.z2 = A.take(2 * Bz * h2 * es); s.a2 = A.take(2 * Bz * h2 * es);
  return s;
}
struct HeadScratch { void *dz2, *dz1, *dxn; float* dl; };
static HeadScratch head_scratch_layout(Arena& A, const mmoe_head_cfg& cfg, int B, int dtype) {
  HeadScratch t{};
  const size_t es = dtype_size(dtype), Bz = (size_t)B; const int d = cfg.d, h1 = cfg.hidden, h2 = cfg.hidden / 2;
  t.dz2 = A.take(2 * Bz * h2 * es); t.dz1 = A.take(2 * Bz * h1 * es); t.dxn = A.take(2 * Bz * d * es);
  t.dl = (float*)A.take(2 * Bz * cfg.n_expert * 4);
  return t;
}
static int check_head(const mmoe_head_cfg* cfg) {
  MMOE_CHECK(cfg->d % 8 == 0 && cfg->d <= 1024, "head: unsupported expert_dim %d", cfg->d);
  MMOE_CHECK(cfg->n_expert >= 1 && cfg->n_expert <= MIX_MAXN, "head: n_expert must be in [1,%d]", MIX_MAXN);
  MMOE_CHECK(cfg->hidden % 16 == 0 && cfg->hidden <= 1024, "head: unsupported tower_hidden %d", cfg->hidden);
  return 0;
}

template <typename T>
static HeadFusedDev head_dev(const mmoe_call* c, const mmoe_head_cfg* cfg, const float* ev, const HeadSaved& s, const HeadIdx& ix) {
  HeadFusedDev a{};
  const void* const* P = c->params;
  a.ev = ev;
  for (int t = 0; t < 2; ++t) {
    a.wg[t] = (const float*)P[ix.gw[t]]; a.bg[t] = (const float*)P[ix.gb[t]];
    a.ln_w[t] = (const float*)P[ix.ln_w[t]]; a.ln_b[t] = (const float*)P[ix.ln_b[t]];
  }
  a.query_t = s.query; a.w = s.w; a.stats = s.st; a.xn = s.xn;
  a.B = c->B; a.d = cfg->d; a.n = cfg->n_expert;
  return a;
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
    HeadFusedDev a = head_dev<T>(c, cfg, ev, s, ix);
    if (n == 6 && d <= 768) {
      int64_t blocks = ((int64_t)B + 3) / 4;
      if (blocks > (int64_t)sm_count() * 2) blocks = (int64_t)sm_count() * 2;
      head_fwd_reg_kernel<T, 6, 6><<<(int)blocks, 128, 0, st>>>(a);
    } else if (d <= 768) head_mix_ln_fwd_kernel<T, 6><<<rows_grid(B, 8), 256, 0, st>>>(a);
    else head_mix_ln_fwd_kernel<T, 8><<<rows_grid(B, 8), 256, 0, st>>>(a);
    MMOE_LAUNCH_OK("head forward mix/LN kernel");
    if (gate_w != nullptr) MMOE_CUDA(cudaMemcpyAsync(gate_w, s.w, (size_t)2 * B * n * 4, cudaMemcpyDeviceToDevice, st));
  }
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
    gemv_bwd_kernel<T><<<rows_grid(B, 4), 256, 0, st>>>(dlogits + (size_t)k * B, (const T*)s.a2 + (size_t)k * B * h2,
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
  {
    HeadFusedDev a = head_dev<T>(c, cfg, ev, s, ix);
    a.dxn = t.dxn; a.dl = t.dl; a.d_ev = d_ev;
    for (int k = 0; k < 2; ++k) { a.dgamma[k] = (float*)G[ix.ln_w[k]]; a.dbeta[k] = (float*)G[ix.ln_b[k]]; }
    int64_t blocks = ((int64_t)B + 3) / 4;
    const int64_t cap = (int64_t)sm_count() * 4;
    if (blocks > cap) blocks = cap;
    if (n == 6 && d <= 768) {
      constexpr int SM_BYTES = 4 * 4 * 6 * 128 * 4;      // 48 KB: 4 warps x [2 tasks][dgamma|dbeta][768]
      static bool attr = false;
      if (!attr) {
        MMOE_CUDA(cudaFuncSetAttribute(head_bwd_reg_kernel<T, 6, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_BYTES));
        attr = true;
      }
      if (blocks > (int64_t)sm_count() * 2) blocks = (int64_t)sm_count() * 2;
      head_bwd_reg_kernel<T, 6, 6><<<(int)blocks, 128, SM_BYTES, st>>>(a);
    } else if (d <= 768) head_ln_mix_bwd_kernel<T, 6><<<(int)blocks, 128, 0, st>>>(a);
    else head_ln_mix_bwd_kernel<T, 8><<<(int)blocks, 128, 0, st>>>(a);
    MMOE_LAUNCH_OK("head backward LN/mix kernel");
  }
  for (int k = 0; k < 2; ++k)
    MMOE_TRY(small_wgrad(t.dl + (size_t)k * B * n, s.query, dtype, (float*)G[ix.gw[k]], (float*)G[ix.gb[k]], B, d, n, st));
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
  return small_wgrad(dl, x, MMOE_F32, dwg, dbg, B, d, n, (cudaStream_t)stream);
}
