// Kernels shared by the two multi-task heads (TwoTaskMMoE, HOME_MMoE_Complete):
// dense-gate mix, the n-output gate weight gradient, and the final 1-output tower layer.
#pragma once
#include "encoder.cuh"

namespace mmoe {

constexpr int MIX_NV = 8;       // float4 per lane -> d <= 1024
constexpr int MIX_MAXN = 8;     // experts per gate

// ---------------------------------------------------------------- gate + mix forward
// One warp per sample.  query q[b] = (query_in ? query_in[b] : mean_n ev[b,n]);  for each task t:
// w_t = softmax(q Wg_t^T + bg_t);  fused_t[b] = sum_n w_t[n] * E_t[b,n]  where expert n of task t lives at
// experts + sel[t][n]*expert_stride + b*row_stride.   (model.py:564-572; model_HoME.py:628-632)
struct MixDev {
  const float* experts; int64_t expert_stride, row_stride;
  int sel[2][MIX_MAXN];
  const float* query_in;        // [B,d] or null (then the mean over the n experts of task 0 is used)
  const float* wg[2]; const float* bg[2];
  float* fused;                 // [2][B][d]
  float* query_out;             // [B][d] (mean query) or null
  float* w;                     // [2][B][n]
  // backward
  const float* dfused;          // [2][B][d]
  float* dl;                    // [2][B][n]   d(gate logits)
  float* dexperts; int accumulate_dexperts;   // same addressing as experts (written once per distinct expert)
  float* dquery;                // [B][d]: gradient flowing into the query (HoME) or null (v1: folded into dexperts)
  int64_t B; int d, n;
};

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

static __global__ void __launch_bounds__(256) mix_fwd_kernel(const MixDev a) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int d = a.d, n = a.n;
  for (int64_t b = warp0; b < a.B; b += nw) {
    float4 q[MIX_NV];
    if (a.query_in != nullptr) {
#pragma unroll
      for (int i = 0; i < MIX_NV; ++i) { const int c = (i * 32 + lane) * 4; q[i] = c < d ? ld4(a.query_in + b * d + c) : make_float4(0, 0, 0, 0); }
    } else {
#pragma unroll
      for (int i = 0; i < MIX_NV; ++i) q[i] = make_float4(0, 0, 0, 0);
      for (int k = 0; k < n; ++k) {
        const float* e = a.experts + a.sel[0][k] * a.expert_stride + b * a.row_stride;
#pragma unroll
        for (int i = 0; i < MIX_NV; ++i) {
          const int c = (i * 32 + lane) * 4;
          if (c < d) { const float4 v = ld4(e + c); q[i].x += v.x; q[i].y += v.y; q[i].z += v.z; q[i].w += v.w; }
        }
      }
      const float inv = 1.f / (float)n;
#pragma unroll
      for (int i = 0; i < MIX_NV; ++i) {
        q[i].x *= inv; q[i].y *= inv; q[i].z *= inv; q[i].w *= inv;
        const int c = (i * 32 + lane) * 4;
        if (a.query_out != nullptr && c < d) st4(a.query_out + b * d + c, q[i]);
      }
    }
    for (int t = 0; t < 2; ++t) {
      float logit[MIX_MAXN];
#pragma unroll
      for (int k = 0; k < MIX_MAXN; ++k) {
        logit[k] = -INFINITY;
        if (k < n) {
          float acc = 0.f;
#pragma unroll
          for (int i = 0; i < MIX_NV; ++i) {
            const int c = (i * 32 + lane) * 4;
            if (c < d) { const float4 w = ld4(a.wg[t] + k * d + c); acc += q[i].x * w.x + q[i].y * w.y + q[i].z * w.z + q[i].w * w.w; }
          }
          logit[k] = warp_sum(acc) + a.bg[t][k];
        }
      }
      float m = -INFINITY;
#pragma unroll
      for (int k = 0; k < MIX_MAXN; ++k) m = fmaxf(m, logit[k]);
      float ssum = 0.f;
#pragma unroll
      for (int k = 0; k < MIX_MAXN; ++k) { logit[k] = k < n ? expf(logit[k] - m) : 0.f; ssum += logit[k]; }
      const float inv = 1.f / ssum;
      float4 f[MIX_NV];
#pragma unroll
      for (int i = 0; i < MIX_NV; ++i) f[i] = make_float4(0, 0, 0, 0);
#pragma unroll
      for (int k = 0; k < MIX_MAXN; ++k) {
        if (k < n) {
          const float wk = logit[k] * inv;
          if (lane == 0) a.w[((int64_t)t * a.B + b) * n + k] = wk;
          const float* e = a.experts + a.sel[t][k] * a.expert_stride + b * a.row_stride;
#pragma unroll
          for (int i = 0; i < MIX_NV; ++i) {
            const int c = (i * 32 + lane) * 4;
            if (c < d) { const float4 v = ld4(e + c); f[i].x += wk * v.x; f[i].y += wk * v.y; f[i].z += wk * v.z; f[i].w += wk * v.w; }
          }
        }
      }
#pragma unroll
      for (int i = 0; i < MIX_NV; ++i) {
        const int c = (i * 32 + lane) * 4;
        if (c < d) st4(a.fused + ((int64_t)t * a.B + b) * d + c, f[i]);
      }
    }
  }
}

// ---------------------------------------------------------------- gate + mix backward
// per sample: dw_t[n] = <dfused_t, E_t[n]>; softmax backward -> dl_t; dq = sum_t dl_t Wg_t;
// dE[e] = sum_{t,n: sel[t][n]==e} w_t[n] dfused_t  (+ dq/n for the v1 mean query).
static __global__ void __launch_bounds__(256) mix_bwd_kernel(const MixDev a, int n_distinct) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int d = a.d, n = a.n;
  for (int64_t b = warp0; b < a.B; b += nw) {
    float4 df[2][MIX_NV];
    float wt[2][MIX_MAXN], dl[2][MIX_MAXN];
#pragma unroll
    for (int t = 0; t < 2; ++t) {
#pragma unroll
      for (int i = 0; i < MIX_NV; ++i) {
        const int c = (i * 32 + lane) * 4;
        df[t][i] = c < d ? ld4(a.dfused + ((int64_t)t * a.B + b) * d + c) : make_float4(0, 0, 0, 0);
      }
      float dot = 0.f;
#pragma unroll
      for (int k = 0; k < MIX_MAXN; ++k) {
        wt[t][k] = 0.f; dl[t][k] = 0.f;
        if (k < n) {
          const float* e = a.experts + a.sel[t][k] * a.expert_stride + b * a.row_stride;
          float acc = 0.f;
#pragma unroll
          for (int i = 0; i < MIX_NV; ++i) {
            const int c = (i * 32 + lane) * 4;
            if (c < d) { const float4 v = ld4(e + c); acc += df[t][i].x * v.x + df[t][i].y * v.y + df[t][i].z * v.z + df[t][i].w * v.w; }
          }
          acc = warp_sum(acc);
          wt[t][k] = a.w[((int64_t)t * a.B + b) * n + k];
          dl[t][k] = acc;
          dot += wt[t][k] * acc;
        }
      }
#pragma unroll
      for (int k = 0; k < MIX_MAXN; ++k) {
        if (k < n) {
          dl[t][k] = wt[t][k] * (dl[t][k] - dot);
          if (lane == 0) a.dl[((int64_t)t * a.B + b) * n + k] = dl[t][k];
        }
      }
    }
    // dq = sum_t sum_k dl_t[k] * Wg_t[k]
    float4 dq[MIX_NV];
#pragma unroll
    for (int i = 0; i < MIX_NV; ++i) {
      dq[i] = make_float4(0, 0, 0, 0);
      const int c = (i * 32 + lane) * 4;
      if (c < d) {
#pragma unroll
        for (int t = 0; t < 2; ++t)
#pragma unroll
          for (int k = 0; k < MIX_MAXN; ++k)
            if (k < n) {
              const float4 w = ld4(a.wg[t] + k * d + c);
              dq[i].x += dl[t][k] * w.x; dq[i].y += dl[t][k] * w.y; dq[i].z += dl[t][k] * w.z; dq[i].w += dl[t][k] * w.w;
            }
        if (a.dquery != nullptr) st4(a.dquery + b * d + c, dq[i]);
      }
    }
    const float qshare = a.dquery == nullptr ? 1.f / (float)n : 0.f;
    for (int e = 0; e < n_distinct; ++e) {
      float coef[2] = {0.f, 0.f};
      bool in_query = false;
#pragma unroll
      for (int t = 0; t < 2; ++t)
#pragma unroll
        for (int k = 0; k < MIX_MAXN; ++k)
          if (k < n && a.sel[t][k] == e) { coef[t] += wt[t][k]; if (t == 0) in_query = true; }
      float* o = a.dexperts + e * a.expert_stride + b * a.row_stride;
#pragma unroll
      for (int i = 0; i < MIX_NV; ++i) {
        const int c = (i * 32 + lane) * 4;
        if (c < d) {
          float4 v;
          const float qs = in_query ? qshare : 0.f;
          v.x = coef[0] * df[0][i].x + coef[1] * df[1][i].x + qs * dq[i].x;
          v.y = coef[0] * df[0][i].y + coef[1] * df[1][i].y + qs * dq[i].y;
          v.z = coef[0] * df[0][i].z + coef[1] * df[1][i].z + qs * dq[i].z;
          v.w = coef[0] * df[0][i].w + coef[1] * df[1][i].w + qs * dq[i].w;
          if (a.accumulate_dexperts) { const float4 p = ld4(o + c); v.x += p.x; v.y += p.y; v.z += p.z; v.w += p.w; }
          st4(o + c, v);
        }
      }
    }
  }
}

// ---------------------------------------------------------------- dW[k][c] += sum_b coef[b][k] * X[b][c]; db[k] += sum_b coef[b][k]
// (gate weight gradients: k <= 8 outputs).  grid = (ceil(d/64), row chunks), 256 threads.
static __global__ void __launch_bounds__(256) small_wgrad_kernel(const float* __restrict__ coef, const void* __restrict__ X, int x_dtype,
                                                          float* __restrict__ dW, float* __restrict__ db, int64_t B, int d, int n,
                                                          int64_t rows_per_block) {
  __shared__ float red[4][MIX_MAXN][64];
  const int cl = threadIdx.x & 63, sub = threadIdx.x >> 6;
  const int c = blockIdx.x * 64 + cl;
  const int64_t b0 = (int64_t)blockIdx.y * rows_per_block, b1 = min(B, b0 + rows_per_block);
  float acc[MIX_MAXN], accb[MIX_MAXN];
#pragma unroll
  for (int k = 0; k < MIX_MAXN; ++k) acc[k] = accb[k] = 0.f;
  // 4 rows per thread per trip: all of their loads are issued before the first FMA (the loop is latency-bound otherwise)
  for (int64_t bb = b0 + sub; bb < b1; bb += 16) {
    float x[4], cf[4][MIX_MAXN];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t b = bb + 4 * u;
      const bool ok = b < b1;
      x[u] = (ok && c < d) ? load_as_f(X, b * d + c, x_dtype) : 0.f;
#pragma unroll
      for (int k = 0; k < MIX_MAXN; ++k) cf[u][k] = (ok && k < n) ? __ldg(coef + b * n + k) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int k = 0; k < MIX_MAXN; ++k) { acc[k] = fmaf(cf[u][k], x[u], acc[k]); accb[k] += cf[u][k]; }
  }
#pragma unroll
  for (int k = 0; k < MIX_MAXN; ++k) red[sub][k][cl] = acc[k];
  __syncthreads();
  if (sub == 0 && c < d) {
#pragma unroll
    for (int k = 0; k < MIX_MAXN; ++k)
      if (k < n) atomicAdd(dW + k * d + c, red[0][k][cl] + red[1][k][cl] + red[2][k][cl] + red[3][k][cl]);
  }
  if (blockIdx.x == 0 && cl == 0) {
#pragma unroll
    for (int k = 0; k < MIX_MAXN; ++k)
      if (k < n) atomicAdd(db + k, accb[k]);
  }
}

inline int small_wgrad(const float* coef, const void* X, int x_dtype, float* dW, float* db, int64_t B, int d, int n, cudaStream_t s) {
  int64_t chunks = (B + 255) / 256;
  if (chunks > 128) chunks = 128;
  if (chunks < 1) chunks = 1;
  const int64_t rpb = (B + chunks - 1) / chunks;
  dim3 grid((d + 63) / 64, (unsigned)chunks);
  small_wgrad_kernel<<<grid, 256, 0, s>>>(coef, X, x_dtype, dW, db, B, d, n, rpb);
  MMOE_LAUNCH_OK("small_wgrad_kernel");
  return 0;
}

// ---------------------------------------------------------------- final tower layer (k -> 1)
// logit[b] = <a[b,:], w> + bias       one warp per row
template <typename T>
__global__ void __launch_bounds__(256) gemv_fwd_kernel(const T* __restrict__ a, const float* __restrict__ w,
                                                       const float* __restrict__ bias, float* __restrict__ out, int64_t B, int k) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t b = warp0; b < B; b += nw) {
    float acc = 0.f;
    for (int j = lane; j < k; j += 32) acc = fmaf(to_f<T>(a[b * k + j]), w[j], acc);
    acc = warp_sum(acc);
    if (lane == 0) out[b] = acc + bias[0];
  }
}
// backward through  logit = <drop(gelu(z)), w> + bias :
//   dz[b,j] = T(dlogit[b] * w[j] * dropmask * gelu'(z[b,j]));  db_prev[j] += colsum(dz);  dw[j] += dlogit[b]*a[b,j];  dbias += dlogit
template <typename T>
__global__ void __launch_bounds__(256) gemv_bwd_kernel(const float* __restrict__ dlogit, const T* __restrict__ a, const T* __restrict__ z,
                                                       const float* __restrict__ w, T* __restrict__ dz, float* __restrict__ dw,
                                                       float* __restrict__ dbias, float* __restrict__ db_prev, int64_t B, int k,
                                                       uint32_t thresh, float scale, uint32_t k0, uint32_t k1) {
  constexpr int MAXJ = 16;   // k <= 512
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  float acc_w[MAXJ], acc_b[MAXJ];
#pragma unroll
  for (int i = 0; i < MAXJ; ++i) acc_w[i] = acc_b[i] = 0.f;
  float acc_bias = 0.f;
  for (int64_t b = warp0; b < B; b += nw) {
    const float dl = dlogit[b];
    acc_bias += dl;
#pragma unroll
    for (int i = 0; i < MAXJ; ++i) {
      const int j = lane + 32 * i;
      if (j < k) {
        float v = dl * w[j] * gelu_grad_f(to_f<T>(z[b * k + j]));
        if (thresh != 0) v = drop_keep(k0, k1, (uint64_t)b * k + j, thresh) ? v * scale : 0.f;
        const T o = from_f<T>(v);
        dz[b * k + j] = o;
        acc_b[i] += to_f<T>(o);
        acc_w[i] = fmaf(dl, to_f<T>(a[b * k + j]), acc_w[i]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < MAXJ; ++i) {
    const int j = lane + 32 * i;
    if (j < k) { atomicAdd(dw + j, acc_w[i]); atomicAdd(db_prev + j, acc_b[i]); }
  }
  if (lane == 0) atomicAdd(dbias, acc_bias);
}

inline int rows_grid(int64_t rows, int mult) {
  int64_t blocks = (rows + 7) / 8;
  const int64_t cap = (int64_t)sm_count() * mult;
  if (blocks > cap) blocks = cap;
  return blocks < 1 ? 1 : (int)blocks;
}

}  // namespace mmoe
