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
// (gate weight gradients: k <= 8 outputs).  A block owns a chunk of rows and ALL columns: its coefficients are staged in
// shared memory once, a thread owns 4 adjacent columns (8- or 16-byte row loads, 8 rows in flight) and keeps the k x 4
// partial sums in registers; one atomic per (k, column) per block at the end.
constexpr int SWG_ROWS = 128;          // rows of coefficients staged per trip
template <typename XT>
static __global__ void __launch_bounds__(256) small_wgrad_kernel(const float* __restrict__ coef, const XT* __restrict__ X,
                                                          float* __restrict__ dW, float* __restrict__ db, int64_t B, int d, int n,
                                                          int64_t rows_per_block) {
  __shared__ float cs[SWG_ROWS][MIX_MAXN];
  const int64_t b0 = (int64_t)blockIdx.x * rows_per_block, b1 = min(B, b0 + rows_per_block);
  const int c = threadIdx.x * 4;
  const bool col_ok = c < d;
  float4 acc[MIX_MAXN];
  float accb = 0.f;                    // thread k < n sums coefficient column k
#pragma unroll
  for (int k = 0; k < MIX_MAXN; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int64_t r0 = b0; r0 < b1; r0 += SWG_ROWS) {
    const int rows = (int)min((int64_t)SWG_ROWS, b1 - r0);
    __syncthreads();
    for (int e = threadIdx.x; e < rows * MIX_MAXN; e += 256) {
      const int r = e / MIX_MAXN, k = e - r * MIX_MAXN;
      cs[r][k] = k < n ? coef[(r0 + r) * n + k] : 0.f;
    }
    __syncthreads();
    if (threadIdx.x < n)
      for (int r = 0; r < rows; ++r) accb += cs[r][threadIdx.x];
    if (col_ok) {
#pragma unroll 8
      for (int r = 0; r < rows; ++r) {
        float4 x;
        if (sizeof(XT) == 4) {
          x = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(X) + (r0 + r) * d + c);
        } else {
          const uint2 u = *reinterpret_cast<const uint2*>(reinterpret_cast<const uint16_t*>(X) + (r0 + r) * d + c);
          x = make_float4(to_f<XT>(reinterpret_cast<const XT*>(&u.x)[0]), to_f<XT>(reinterpret_cast<const XT*>(&u.x)[1]),
                          to_f<XT>(reinterpret_cast<const XT*>(&u.y)[0]), to_f<XT>(reinterpret_cast<const XT*>(&u.y)[1]));
        }
#pragma unroll
        for (int k = 0; k < MIX_MAXN; ++k) {
          const float cf = cs[r][k];
          acc[k].x = fmaf(cf, x.x, acc[k].x); acc[k].y = fmaf(cf, x.y, acc[k].y);
          acc[k].z = fmaf(cf, x.z, acc[k].z); acc[k].w = fmaf(cf, x.w, acc[k].w);
        }
      }
    }
  }
  if (col_ok) {
#pragma unroll
    for (int k = 0; k < MIX_MAXN; ++k)
      if (k < n) {
        atomicAdd(dW + k * d + c + 0, acc[k].x); atomicAdd(dW + k * d + c + 1, acc[k].y);
        atomicAdd(dW + k * d + c + 2, acc[k].z); atomicAdd(dW + k * d + c + 3, acc[k].w);
      }
  }
  if (threadIdx.x < n) atomicAdd(db + threadIdx.x, accb);
}

inline int small_wgrad(const float* coef, const void* X, int x_dtype, float* dW, float* db, int64_t B, int d, int n, cudaStream_t s) {
  MMOE_CHECK(d % 4 == 0 && d <= 1024 && n <= MIX_MAXN, "small_wgrad: unsupported d=%d n=%d", d, n);
  int64_t blocks = (B + 63) / 64;                       // at least 64 rows per block
  const int64_t cap = (int64_t)sm_count() * 4;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  const int64_t rpb = (B + blocks - 1) / blocks;
  blocks = (B + rpb - 1) / rpb;
  if (x_dtype == MMOE_F32) small_wgrad_kernel<float><<<(int)blocks, 256, 0, s>>>(coef, (const float*)X, dW, db, B, d, n, rpb);
  else if (x_dtype == MMOE_BF16) small_wgrad_kernel<__nv_bfloat16><<<(int)blocks, 256, 0, s>>>(coef, (const __nv_bfloat16*)X, dW, db, B, d, n, rpb);
  else small_wgrad_kernel<__half><<<(int)blocks, 256, 0, s>>>(coef, (const __half*)X, dW, db, B, d, n, rpb);
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
// A lane owns 4 adjacent columns (8-byte row accesses; k % 4 == 0, k <= 512); the per-column partial sums are reduced
// across the block's warps in shared memory, so a block issues ONE atomic per column (a warp-level atomic per column cost
// ~100 us of same-address contention at B = 65536).
template <typename T> __device__ __forceinline__ float4 gv_ld4(const T* p);
template <> __device__ __forceinline__ float4 gv_ld4<float>(const float* p) { return *reinterpret_cast<const float4*>(p); }
template <> __device__ __forceinline__ float4 gv_ld4<__nv_bfloat16>(const __nv_bfloat16* p) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  return make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xFFFF0000u), __uint_as_float(u.y << 16), __uint_as_float(u.y & 0xFFFF0000u));
}
template <> __device__ __forceinline__ float4 gv_ld4<__half>(const __half* p) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
  return make_float4(a.x, a.y, b.x, b.y);
}
template <typename T> __device__ __forceinline__ float4 gv_st4(T* p, float4 v);      // returns what was stored
template <> __device__ __forceinline__ float4 gv_st4<float>(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; return v; }
template <> __device__ __forceinline__ float4 gv_st4<__nv_bfloat16>(__nv_bfloat16* p, float4 v) {
  const __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
  uint2 u; u.x = *reinterpret_cast<const uint32_t*>(&lo); u.y = *reinterpret_cast<const uint32_t*>(&hi);
  *reinterpret_cast<uint2*>(p) = u;
  return make_float4(__low2float(lo), __high2float(lo), __low2float(hi), __high2float(hi));
}
template <> __device__ __forceinline__ float4 gv_st4<__half>(__half* p, float4 v) {
  const __half2 lo = __floats2half2_rn(v.x, v.y), hi = __floats2half2_rn(v.z, v.w);
  uint2 u; u.x = *reinterpret_cast<const uint32_t*>(&lo); u.y = *reinterpret_cast<const uint32_t*>(&hi);
  *reinterpret_cast<uint2*>(p) = u;
  return make_float4(__low2float(lo), __high2float(lo), __low2float(hi), __high2float(hi));
}
template <typename T>
__global__ void __launch_bounds__(256) gemv_bwd_kernel(const float* __restrict__ dlogit, const T* __restrict__ a, const T* __restrict__ z,
                                                       const float* __restrict__ w, T* __restrict__ dz, float* __restrict__ dw,
                                                       float* __restrict__ dbias, float* __restrict__ db_prev, int64_t B, int k,
                                                       uint32_t thresh, float scale, uint32_t k0, uint32_t k1) {
  constexpr int MAXI = 4;    // k <= 512
  __shared__ float red[8][2][512];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  float4 acc_w[MAXI], acc_b[MAXI], wv[MAXI];
#pragma unroll
  for (int i = 0; i < MAXI; ++i) {
    acc_w[i] = acc_b[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    const int c = lane * 4 + 128 * i;
    wv[i] = c < k ? *reinterpret_cast<const float4*>(w + c) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  float acc_bias = 0.f;
#pragma unroll 2
  for (int64_t b = warp0; b < B; b += nw) {
    const float dl = dlogit[b];
    acc_bias += dl;
#pragma unroll
    for (int i = 0; i < MAXI; ++i) {
      const int c = lane * 4 + 128 * i;
      if (c < k) {
        const float4 zv = gv_ld4<T>(z + b * k + c), av = gv_ld4<T>(a + b * k + c);
        float4 v = make_float4(dl * wv[i].x * gelu_grad_f(zv.x), dl * wv[i].y * gelu_grad_f(zv.y),
                               dl * wv[i].z * gelu_grad_f(zv.z), dl * wv[i].w * gelu_grad_f(zv.w));
        if (thresh != 0) {
          const uint64_t idx = (uint64_t)b * k + c;               // multiple of 4: two hash pairs
          const uint32_t h0 = drop_hash(k0, k1, idx >> 1), h1 = drop_hash(k0, k1, (idx >> 1) + 1);
          v.x = ((h0 & 0xFFFFu) >= thresh) ? v.x * scale : 0.f;
          v.y = ((h0 >> 16) >= thresh) ? v.y * scale : 0.f;
          v.z = ((h1 & 0xFFFFu) >= thresh) ? v.z * scale : 0.f;
          v.w = ((h1 >> 16) >= thresh) ? v.w * scale : 0.f;
        }
        const float4 o = gv_st4<T>(dz + b * k + c, v);
        acc_b[i].x += o.x; acc_b[i].y += o.y; acc_b[i].z += o.z; acc_b[i].w += o.w;
        acc_w[i].x = fmaf(dl, av.x, acc_w[i].x); acc_w[i].y = fmaf(dl, av.y, acc_w[i].y);
        acc_w[i].z = fmaf(dl, av.z, acc_w[i].z); acc_w[i].w = fmaf(dl, av.w, acc_w[i].w);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < MAXI; ++i) {
    const int c = lane * 4 + 128 * i;
    if (c < k) {
      *reinterpret_cast<float4*>(&red[wib][0][c]) = acc_w[i];
      *reinterpret_cast<float4*>(&red[wib][1][c]) = acc_b[i];
    }
  }
  __syncthreads();                       // (every lane of a warp holds the same acc_bias: lane 0 reports it)
  for (int e = threadIdx.x; e < 2 * k; e += 256) {
    const int r = e / k, c = e - r * k;
    float t = 0.f;
#pragma unroll
    for (int wv_ = 0; wv_ < 8; ++wv_) t += red[wv_][r][c];
    atomicAdd((r == 0 ? dw : db_prev) + c, t);
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
