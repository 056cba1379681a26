// Bandwidth-bound helper kernels: casts, dropout masks, LayerNorm forward/backward.
// All rows are d <= 1024 wide (d = 768 on this path), d % 4 == 0; one warp owns a row and
// moves it with 128-bit (fp32) / 64-bit (16-bit types) coalesced accesses.
#include <atomic>

#include "kernels.cuh"

namespace mmoe {

thread_local char g_error[512] = {0};
// process-wide: forward runs on the Python thread, backward on the autograd worker thread
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
bool pdl_enabled() {
  // opt-in: measured on the benchmark step 9.85 (on) / 9.85 (off) / 9.76 (on) ms — no gain; the step runs power-capped
  // (sw_power_cap, ~1.87 GHz), so closing the 2-3 us launch gaps only moves the clock
  static const bool on = [] { const char* e = getenv("MMOE_PDL"); return e != nullptr && e[0] == '1'; }();
  return on;
}

// ---------------------------------------------------------------- vector load/store helpers
template <typename T> __device__ __forceinline__ float4 load4(const T* p);
template <> __device__ __forceinline__ float4 load4<float>(const float* p) { return *reinterpret_cast<const float4*>(p); }
template <> __device__ __forceinline__ float4 load4<__nv_bfloat16>(const __nv_bfloat16* p) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&u.x);
  const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&u.y);
  return make_float4(__low2float(a), __high2float(a), __low2float(b), __high2float(b));
}
template <> __device__ __forceinline__ float4 load4<__half>(const __half* p) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  const __half2 a = *reinterpret_cast<const __half2*>(&u.x);
  const __half2 b = *reinterpret_cast<const __half2*>(&u.y);
  return make_float4(__low2float(a), __high2float(a), __low2float(b), __high2float(b));
}
// what a value becomes when stored as T and read back
template <typename T> __device__ __forceinline__ float4 round4(float4 v) {
  return make_float4(to_f<T>(from_f<T>(v.x)), to_f<T>(from_f<T>(v.y)), to_f<T>(from_f<T>(v.z)), to_f<T>(from_f<T>(v.w)));
}
template <typename T> __device__ __forceinline__ void store4(T* p, float4 v);
template <> __device__ __forceinline__ void store4<float>(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
template <> __device__ __forceinline__ void store4<__nv_bfloat16>(__nv_bfloat16* p, float4 v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
  uint2 u;
  u.x = *reinterpret_cast<uint32_t*>(&a);
  u.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = u;
}
template <> __device__ __forceinline__ void store4<__half>(__half* p, float4 v) {
  __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
  uint2 u;
  u.x = *reinterpret_cast<uint32_t*>(&a);
  u.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = u;
}

// ---------------------------------------------------------------- cast
template <typename T>
__global__ void cast_kernel(const float* __restrict__ x, T* __restrict__ y, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x * 4;
  for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
    if (i + 3 < n) {
      store4<T>(y + i, load4<float>(x + i));
    } else {
      for (int64_t j = i; j < n; ++j) y[j] = from_f<T>(x[j]);
    }
  }
}

int cast_f32(const float* x, void* y, int64_t n, int dtype, cudaStream_t s) {
  if (n == 0) return 0;
  MMOE_CHECK((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(y) & 7) == 0, "cast: unaligned pointer");
  const int threads = 256;
  int64_t blocks = (n / 4 + threads - 1) / threads;
  const int64_t cap = (int64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  if (dtype == MMOE_BF16) cast_kernel<__nv_bfloat16><<<(int)blocks, threads, 0, s>>>(x, (__nv_bfloat16*)y, n);
  else if (dtype == MMOE_F16) cast_kernel<__half><<<(int)blocks, threads, 0, s>>>(x, (__half*)y, n);
  else cast_kernel<float><<<(int)blocks, threads, 0, s>>>(x, (float*)y, n);
  MMOE_LAUNCH_OK("cast_kernel");
  return 0;
}

// ---------------------------------------------------------------- cast + dropout + column sums
// block = 256 threads, owns a panel of `rpb` rows x 1024 columns (blockIdx.y); thread t owns 4 adjacent columns
template <typename T>
__global__ void cast_drop_colsum_kernel(const float* __restrict__ x, const T* __restrict__ x_t, T* __restrict__ g, float* __restrict__ colsum,
                                        int64_t rows, int cols, uint32_t thresh, float scale, uint32_t k0, uint32_t k1, int rpb) {
  pdl_entry();
  const int64_t r0 = (int64_t)blockIdx.x * rpb;
  const int64_t r1 = min(rows, r0 + rpb);
  for (int c = blockIdx.y * 1024 + threadIdx.x * 4; c < min(cols, (int)(blockIdx.y + 1) * 1024); c += 1024) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    // four rows per trip: their loads are issued together (the one-row loop ran at 20 % of DRAM peak on [32768, 768])
    for (int64_t r = r0; r < r1; r += 4) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r + u < r1) {
          if (x != nullptr) v[u] = load4<float>(x + (r + u) * cols + c);
          if (x_t != nullptr) { const float4 e = load4<T>(x_t + (r + u) * cols + c); v[u].x += e.x; v[u].y += e.y; v[u].z += e.z; v[u].w += e.w; }
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (r + u >= r1) break;
        float4 w = v[u];
        if (thresh != 0) {
          const uint64_t idx = (uint64_t)(r + u) * cols + c;
          w.x = drop_keep(k0, k1, idx + 0, thresh) ? w.x * scale : 0.f;
          w.y = drop_keep(k0, k1, idx + 1, thresh) ? w.y * scale : 0.f;
          w.z = drop_keep(k0, k1, idx + 2, thresh) ? w.z * scale : 0.f;
          w.w = drop_keep(k0, k1, idx + 3, thresh) ? w.w * scale : 0.f;
        }
        if (g != nullptr) {
          store4<T>(g + (r + u) * cols + c, w);
          w = round4<T>(w);               // column sums are taken over the values as stored (rounded to T)
        }
        acc.x += w.x; acc.y += w.y; acc.z += w.z; acc.w += w.w;
      }
    }
    if (colsum != nullptr) {
      atomicAdd(colsum + c + 0, acc.x); atomicAdd(colsum + c + 1, acc.y);
      atomicAdd(colsum + c + 2, acc.z); atomicAdd(colsum + c + 3, acc.w);
    }
  }
}

// rows per block of the column-owning panel kernels: tall panels amortise the column-sum atomics, but the grid must
// still cover the chip a few times over (small-batch modules would otherwise run on a handful of SMs)
static int panel_rows(int64_t rows, int cols) {
  const int64_t col_chunks = (cols + 1023) / 1024;
  int64_t rpb = rows * col_chunks / ((int64_t)sm_count() * 4);
  if (rpb > 64) rpb = 64;
  if (rpb < 2) rpb = 2;
  return (int)rpb;
}

int cast_drop_colsum(const float* x, void* g, float* colsum, int64_t rows, int cols, float drop_p, uint32_t k0, uint32_t k1,
                     int dtype, cudaStream_t s, const void* x_t) {
  if (rows == 0) return 0;
  MMOE_CHECK(cols % 4 == 0, "cast_drop_colsum: cols must be a multiple of 4");
  MMOE_CHECK(x != nullptr || x_t != nullptr, "cast_drop_colsum: no input");
  const uint32_t thresh = drop_p > 0.f ? drop_threshold(drop_p) : 0u;
  const float scale = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  const int rpb = panel_rows(rows, cols);
  const dim3 grid((unsigned)((rows + rpb - 1) / rpb), (unsigned)((cols + 1023) / 1024));
  if (dtype == MMOE_BF16) MMOE_CUDA(launch_pdl(cast_drop_colsum_kernel<__nv_bfloat16>, grid, 256, 0, s, x, (const __nv_bfloat16*)x_t, (__nv_bfloat16*)g, colsum, rows, cols, thresh, scale, k0, k1, rpb));
  else if (dtype == MMOE_F16) MMOE_CUDA(launch_pdl(cast_drop_colsum_kernel<__half>, grid, 256, 0, s, x, (const __half*)x_t, (__half*)g, colsum, rows, cols, thresh, scale, k0, k1, rpb));
  else MMOE_CUDA(launch_pdl(cast_drop_colsum_kernel<float>, grid, 256, 0, s, x, (const float*)x_t, (float*)g, colsum, rows, cols, thresh, scale, k0, k1, rpb));
  MMOE_LAUNCH_OK("cast_drop_colsum_kernel");
  return 0;
}

// ---------------------------------------------------------------- dropout mask (debug/test)
__global__ void dropout_mask_kernel(uint32_t k0, uint32_t k1, uint32_t thresh, int64_t n, uint8_t* out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    out[i] = drop_keep(k0, k1, (uint64_t)i, thresh) ? 1 : 0;
}
int dropout_mask(uint32_t k0, uint32_t k1, float p, int64_t n, uint8_t* out, cudaStream_t s) {
  if (n == 0) return 0;
  int64_t blocks = (n + 255) / 256;
  if (blocks > 4096) blocks = 4096;
  dropout_mask_kernel<<<(int)blocks, 256, 0, s>>>(k0, k1, p > 0.f ? drop_threshold(p) : 0u, n, out);
  MMOE_LAUNCH_OK("dropout_mask_kernel");
  return 0;
}

// ---------------------------------------------------------------- LayerNorm forward
constexpr int LN_NV = 8;   // float4 per lane -> d <= 1024

template <typename XT, typename T>
__global__ void __launch_bounds__(256) ln_fwd_kernel(const XT* __restrict__ x, const float* __restrict__ gamma,
                                                     const float* __restrict__ beta, T* __restrict__ y_t,
                                                     float* __restrict__ y_f, float* __restrict__ stats, int64_t rows, int d,
                                                     const T* __restrict__ delta = nullptr, float* __restrict__ x_sum = nullptr) {
  pdl_entry();
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const float inv_d = 1.0f / (float)d;
  for (int64_t row = warp0; row < rows; row += nwarps) {
    const XT* xr = x + row * d;
    if (row + nwarps < rows) {                 // this warp's next row: into L2 while this one is worked on
      prefetch_l2_rows(x + (row + nwarps) * d, d * (int)sizeof(XT), lane);
      if (delta != nullptr) prefetch_l2_rows(delta + (row + nwarps) * d, d * (int)sizeof(T), lane);
    }
    float4 v[LN_NV];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < LN_NV; ++i) {
      const int c = (i * 32 + lane) * 4;
      if (c < d) {
        v[i] = load4<XT>(xr + c);
        if (delta != nullptr) {
          const float4 dl = load4<T>(delta + row * d + c);
          v[i].x += dl.x; v[i].y += dl.y; v[i].z += dl.z; v[i].w += dl.w;
          store4<float>(x_sum + row * d + c, v[i]);
        }
        sum += v[i].x + v[i].y + v[i].z + v[i].w;
      } else v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const float mean = warp_sum(sum) * inv_d;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < LN_NV; ++i) {
      const int c = (i * 32 + lane) * 4;
      if (c < d) {
        const float a = v[i].x - mean, b = v[i].y - mean, e = v[i].z - mean, f = v[i].w - mean;
        sq += a * a + b * b + e * e + f * f;
      }
    }
    const float rstd = rsqrtf(warp_sum(sq) * inv_d + 1e-5f);
    if (stats != nullptr && lane == 0) { stats[row * 2] = mean; stats[row * 2 + 1] = rstd; }
#pragma unroll
    for (int i = 0; i < LN_NV; ++i) {
      const int c = (i * 32 + lane) * 4;
      if (c < d) {
        const float4 g = load4<float>(gamma + c), b = load4<float>(beta + c);
        float4 o;
        o.x = (v[i].x - mean) * rstd * g.x + b.x;
        o.y = (v[i].y - mean) * rstd * g.y + b.y;
        o.z = (v[i].z - mean) * rstd * g.z + b.z;
        o.w = (v[i].w - mean) * rstd * g.w + b.w;
        if (y_t != nullptr) store4<T>(y_t + row * d + c, o);
        if (y_f != nullptr) store4<float>(y_f + row * d + c, o);
      }
    }
  }
}

template <typename T>
static int ln_fwd_dispatch(const void* x, bool x_f32, const float* gamma, const float* beta, void* y_t, float* y_f, float* stats,
                           int64_t rows, int d, cudaStream_t s) {
  int64_t blocks = (rows + 7) / 8;
  const int64_t cap = (int64_t)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  if (x_f32) MMOE_CUDA(launch_pdl(ln_fwd_kernel<float, T>, (int)blocks, 256, 0, s, (const float*)x, gamma, beta, (T*)y_t, y_f, stats, rows, d, nullptr, nullptr));
  else       MMOE_CUDA(launch_pdl(ln_fwd_kernel<T, T>, (int)blocks, 256, 0, s, (const T*)x, gamma, beta, (T*)y_t, y_f, stats, rows, d, nullptr, nullptr));
  MMOE_LAUNCH_OK("ln_fwd_kernel");
  return 0;
}

int layernorm_fwd(const void* x, int x_dtype, const float* gamma, const float* beta, void* y_t, float* y_f32, float* stats,
                  int64_t rows, int d, int dtype, cudaStream_t s) {
  if (rows == 0) return 0;
  MMOE_CHECK(d % 4 == 0 && d <= LN_NV * 128, "layernorm: d must be a multiple of 4 and <= %d (got %d)", LN_NV * 128, d);
  MMOE_CHECK(x_dtype == MMOE_F32 || x_dtype == dtype, "layernorm: x must be fp32 or the operand dtype");
  const bool xf = x_dtype == MMOE_F32;
  if (dtype == MMOE_BF16) return ln_fwd_dispatch<__nv_bfloat16>(x, xf, gamma, beta, y_t, y_f32, stats, rows, d, s);
  if (dtype == MMOE_F16) return ln_fwd_dispatch<__half>(x, xf, gamma, beta, y_t, y_f32, stats, rows, d, s);
  return ln_fwd_dispatch<float>(x, true, gamma, beta, y_t, y_f32, stats, rows, d, s);
}

template <typename T>
static int ln_fwd_add_dispatch(const float* x, const void* delta, float* x_sum, const float* gamma, const float* beta, void* y_t,
                               float* stats, int64_t rows, int d, cudaStream_t s) {
  int64_t blocks = (rows + 7) / 8;
  const int64_t cap = (int64_t)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  MMOE_CUDA(launch_pdl(ln_fwd_kernel<float, T>, (int)blocks, 256, 0, s, x, gamma, beta, (T*)y_t, nullptr, stats, rows, d, (const T*)delta, x_sum));
  MMOE_LAUNCH_OK("ln_fwd_kernel(add)");
  return 0;
}
int layernorm_fwd_add(const float* x, const void* delta, float* x_sum, const float* gamma, const float* beta, void* y_t,
                      float* stats, int64_t rows, int d, int dtype, cudaStream_t s) {
  if (rows == 0) return 0;
  MMOE_CHECK(d % 4 == 0 && d <= LN_NV * 128, "layernorm: d must be a multiple of 4 and <= %d (got %d)", LN_NV * 128, d);
  MMOE_CHECK(delta != nullptr && x_sum != nullptr, "layernorm_fwd_add: delta and x_sum are required");
  if (dtype == MMOE_BF16) return ln_fwd_add_dispatch<__nv_bfloat16>(x, delta, x_sum, gamma, beta, y_t, stats, rows, d, s);
  if (dtype == MMOE_F16) return ln_fwd_add_dispatch<__half>(x, delta, x_sum, gamma, beta, y_t, stats, rows, d, s);
  return ln_fwd_add_dispatch<float>(x, delta, x_sum, gamma, beta, y_t, stats, rows, d, s);
}

// ---------------------------------------------------------------- out_f = a + b ; out_t = T(out_f)
template <typename T>
__global__ void add_cast_kernel(const float* __restrict__ a, const T* __restrict__ b, float* __restrict__ out_f, T* __restrict__ out_t, int64_t n) {
  pdl_entry();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x * 4;
  for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
    float4 v = load4<float>(a + i);
    if (b != nullptr) { const float4 w = load4<T>(b + i); v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w; }
    if (out_f != nullptr) store4<float>(out_f + i, v);
    if (out_t != nullptr) store4<T>(out_t + i, v);
  }
}
int add_cast(const float* a, const void* b, float* out_f, void* out_t, int64_t n, int dtype, cudaStream_t s) {
  if (n == 0) return 0;
  MMOE_CHECK(n % 4 == 0, "add_cast: n must be a multiple of 4");
  int64_t blocks = (n / 4 + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  if (dtype == MMOE_BF16) MMOE_CUDA(launch_pdl(add_cast_kernel<__nv_bfloat16>, (int)blocks, 256, 0, s, a, (const __nv_bfloat16*)b, out_f, (__nv_bfloat16*)out_t, n));
  else if (dtype == MMOE_F16) MMOE_CUDA(launch_pdl(add_cast_kernel<__half>, (int)blocks, 256, 0, s, a, (const __half*)b, out_f, (__half*)out_t, n));
  else MMOE_CUDA(launch_pdl(add_cast_kernel<float>, (int)blocks, 256, 0, s, a, (const float*)b, out_f, (float*)out_t, n));
  MMOE_LAUNCH_OK("add_cast_kernel");
  return 0;
}

// ---------------------------------------------------------------- ReLU(+dropout) backward mask + bias-gradient column sums
// block = 256 threads, panel of `rpb` rows x 1024 columns (blockIdx.y); a thread owns 4 adjacent columns
template <typename T>
__global__ void __launch_bounds__(256) relu_mask_colsum_kernel(T* __restrict__ g, const T* __restrict__ h, float* __restrict__ colsum,
                                                               int64_t rows, int cols, float scale, int rpb) {
  const int64_t r0 = (int64_t)blockIdx.x * rpb, r1 = min(rows, r0 + rpb);
  for (int c = blockIdx.y * 1024 + threadIdx.x * 4; c < min(cols, (int)(blockIdx.y + 1) * 1024); c += 1024) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
    for (int64_t r = r0; r < r1; ++r) {
      const float4 hv = load4<T>(h + r * cols + c);
      float4 gv = load4<T>(g + r * cols + c);
      gv.x = hv.x != 0.f ? gv.x * scale : 0.f;
      gv.y = hv.y != 0.f ? gv.y * scale : 0.f;
      gv.z = hv.z != 0.f ? gv.z * scale : 0.f;
      gv.w = hv.w != 0.f ? gv.w * scale : 0.f;
      store4<T>(g + r * cols + c, gv);
      const float4 w = load4<T>(g + r * cols + c);   // as stored (rounded)
      acc.x += w.x; acc.y += w.y; acc.z += w.z; acc.w += w.w;
    }
    if (colsum != nullptr) {
      atomicAdd(colsum + c + 0, acc.x); atomicAdd(colsum + c + 1, acc.y);
      atomicAdd(colsum + c + 2, acc.z); atomicAdd(colsum + c + 3, acc.w);
    }
  }
}
int relu_mask_colsum(void* g, const void* h, float* colsum, int64_t rows, int cols, float scale, int dtype, cudaStream_t s) {
  if (rows == 0) return 0;
  MMOE_CHECK(cols % 4 == 0, "relu_mask_colsum: cols must be a multiple of 4");
  const int rpb = panel_rows(rows, cols);
  const dim3 grid((unsigned)((rows + rpb - 1) / rpb), (unsigned)((cols + 1023) / 1024));
  if (dtype == MMOE_BF16) relu_mask_colsum_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>((__nv_bfloat16*)g, (const __nv_bfloat16*)h, colsum, rows, cols, scale, rpb);
  else if (dtype == MMOE_F16) relu_mask_colsum_kernel<__half><<<grid, 256, 0, s>>>((__half*)g, (const __half*)h, colsum, rows, cols, scale, rpb);
  else relu_mask_colsum_kernel<float><<<grid, 256, 0, s>>>((float*)g, (const float*)h, colsum, rows, cols, scale, rpb);
  MMOE_LAUNCH_OK("relu_mask_colsum_kernel");
  return 0;
}

// ---------------------------------------------------------------- LayerNorm backward
struct LnBwdDev {
  const void* dy; const void* x; const float* stats; const float* gamma; const float* dres; const void* dres_t;
  float* dx; float* dgamma; float* dbeta; void* g_out; float* g_colsum;
  uint32_t thresh, k0, k1; float scale;
  int64_t rows; int d;
};

// NV = float4 per lane (6 covers d <= 768).  The kernel is a stream of long-latency row loads, so bytes in flight per SM set
// its speed: every load of a row (x, dy and the residual gradient) is issued before the first use, and the per-column
// partial sums (d gamma, d beta, bias gradient) live in warp-private shared memory instead of 72 registers, which lets
// 3 blocks of 4 warps share an SM without spilling.
template <typename DT_, typename XT, typename T, int NV>
__global__ void __launch_bounds__(128, 3) ln_bwd_kernel(const LnBwdDev a) {
  constexpr int LN_NV = NV;
  constexpr int WARPS = 4, COLS = NV * 128;
  __shared__ __align__(16) float accs[WARPS][3][COLS];           // [warp][dgamma | dbeta | colsum][column]
  pdl_entry();
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t warp0 = (int64_t)blockIdx.x * WARPS + wib;
  const int64_t nwarps = (int64_t)gridDim.x * WARPS;
  const int d = a.d;
  const float inv_d = 1.0f / (float)d;
  float* mine = &accs[wib][0][0];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int i = 0; i < LN_NV; ++i) *reinterpret_cast<float4*>(mine + r * COLS + (i * 32 + lane) * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
  const DT_* dy = reinterpret_cast<const DT_*>(a.dy);
  const XT* x = reinterpret_cast<const XT*>(a.x);
  T* gout = reinterpret_cast<T*>(a.g_out);
  const bool want_wb = a.dgamma != nullptr, want_g = gout != nullptr || a.g_colsum != nullptr;
  for (int64_t row = warp0; row < a.rows; row += nwarps) {
    if (row + nwarps < a.rows) {               // this warp's next row: into L2 while this one is worked on
      const int64_t nx = (row + nwarps) * d;
      prefetch_l2_rows(x + nx, d * (int)sizeof(XT), lane);
      prefetch_l2_rows(dy + nx, d * (int)sizeof(DT_), lane);
      if (a.dres != nullptr) prefetch_l2_rows(a.dres + nx, d * 4, lane);
      if (a.dres_t != nullptr) prefetch_l2_rows(reinterpret_cast<const T*>(a.dres_t) + nx, d * (int)sizeof(T), lane);
    }
    float4 xh[LN_NV], dg[LN_NV], rs[LN_NV];
#pragma unroll
    for (int i = 0; i < LN_NV; ++i) {
      const int c = (i * 32 + lane) * 4;
      xh[i] = dg[i] = rs[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c < d) {
        xh[i] = load4<XT>(x + row * d + c);
        dg[i] = load4<DT_>(dy + row * d + c);
        if (a.dres != nullptr) rs[i] = load4<float>(a.dres + row * d + c);
        if (a.dres_t != nullptr) {
          const float4 e = load4<T>(reinterpret_cast<const T*>(a.dres_t) + row * d + c);
          rs[i].x += e.x; rs[i].y += e.y; rs[i].z += e.z; rs[i].w += e.w;
        }
      }
    }
    const float mean = a.stats[row * 2], rstd = a.stats[row * 2 + 1];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < LN_NV; ++i) {
      const int c = (i * 32 + lane) * 4;
      if (c < d) {
        const float4 dv = dg[i];
        const float4 g = load4<float>(a.gamma + c);
        xh[i] = make_float4((xh[i].x - mean) * rstd, (xh[i].y - mean) * rstd, (xh[i].z - mean) * rstd, (xh[i].w - mean) * rstd);
        dg[i] = make_float4(dv.x * g.x, dv.y * g.y, dv.z * g.z, dv.w * g.w);
        s1 += dg[i].x + dg[i].y + dg[i].z + dg[i].w;
        s2 += dg[i].x * xh[i].x + dg[i].y * xh[i].y + dg[i].z * xh[i].z + dg[i].w * xh[i].w;
        if (want_wb) {
          float4* pg = reinterpret_cast<float4*>(mine + c);
          float4* pb = reinterpret_cast<float4*>(mine + COLS + c);
          float4 ag = *pg, ab = *pb;
          ag.x = fmaf(dv.x, xh[i].x, ag.x); ag.y = fmaf(dv.y, xh[i].y, ag.y); ag.z = fmaf(dv.z, xh[i].z, ag.z); ag.w = fmaf(dv.w, xh[i].w, ag.w);
          ab.x += dv.x; ab.y += dv.y; ab.z += dv.z; ab.w += dv.w;
          *pg = ag; *pb = ab;
        }
      }
    }
    s1 = warp_sum(s1) * inv_d;
    s2 = warp_sum(s2) * inv_d;
#pragma unroll
    for (int i = 0; i < LN_NV; ++i) {
      const int c = (i * 32 + lane) * 4;
      if (c < d) {
        float4 o;
        o.x = rstd * (dg[i].x - s1 - xh[i].x * s2) + rs[i].x;
        o.y = rstd * (dg[i].y - s1 - xh[i].y * s2) + rs[i].y;
        o.z = rstd * (dg[i].z - s1 - xh[i].z * s2) + rs[i].z;
        o.w = rstd * (dg[i].w - s1 - xh[i].w * s2) + rs[i].w;
        if (a.dx != nullptr) store4<float>(a.dx + row * d + c, o);
        if (want_g) {
          if (a.thresh != 0) {
            const uint64_t idx = (uint64_t)row * d + c;                   // multiple of 4: two hash pairs
            const uint32_t h0 = drop_hash(a.k0, a.k1, idx >> 1), h1 = drop_hash(a.k0, a.k1, (idx >> 1) + 1);
            o.x = ((h0 & 0xFFFFu) >= a.thresh) ? o.x * a.scale : 0.f;
            o.y = ((h0 >> 16) >= a.thresh) ? o.y * a.scale : 0.f;
            o.z = ((h1 & 0xFFFFu) >= a.thresh) ? o.z * a.scale : 0.f;
            o.w = ((h1 >> 16) >= a.thresh) ? o.w * a.scale : 0.f;
          }
          if (gout != nullptr) {
            store4<T>(gout + row * d + c, o);
            o = round4<T>(o);                                             // the bias gradient sums what was stored
          }
          float4* pc = reinterpret_cast<float4*>(mine + 2 * COLS + c);
          float4 ac = *pc;
          ac.x += o.x; ac.y += o.y; ac.z += o.z; ac.w += o.w;
          *pc = ac;
        }
      }
    }
  }
  __syncthreads();
  // block reduction of the warp-private column partials, then one atomic per column per block
  for (int idx = threadIdx.x; idx < 3 * COLS; idx += 128) {
    const int r = idx / COLS, c = idx - r * COLS;
    float* dst = r == 0 ? a.dgamma : (r == 1 ? a.dbeta : a.g_colsum);
    if (c < d && dst != nullptr && (r == 2 ? want_g : want_wb)) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < WARPS; ++w) t += accs[w][r][c];
      atomicAdd(dst + c, t);
    }
  }
}

template <typename T, int NV>
static int ln_bwd_dispatch_nv(const LnBwdDev& dev, bool dy_f32, bool x_f32, int blocks, cudaStream_t s) {
  if (dy_f32 && x_f32)       MMOE_CUDA(launch_pdl(ln_bwd_kernel<float, float, T, NV>, blocks, 128, 0, s, dev));
  else if (dy_f32 && !x_f32) MMOE_CUDA(launch_pdl(ln_bwd_kernel<float, T, T, NV>, blocks, 128, 0, s, dev));
  else if (!dy_f32 && x_f32) MMOE_CUDA(launch_pdl(ln_bwd_kernel<T, float, T, NV>, blocks, 128, 0, s, dev));
  else                       MMOE_CUDA(launch_pdl(ln_bwd_kernel<T, T, T, NV>, blocks, 128, 0, s, dev));
  MMOE_LAUNCH_OK("ln_bwd_kernel");
  return 0;
}
template <typename T>
static int ln_bwd_dispatch(const LnBwdDev& dev, bool dy_f32, bool x_f32, int blocks, cudaStream_t s) {
  if (dev.d <= 768) return ln_bwd_dispatch_nv<T, 6>(dev, dy_f32, x_f32, blocks, s);
  return ln_bwd_dispatch_nv<T, 8>(dev, dy_f32, x_f32, blocks, s);
}

int layernorm_bwd(const LnBwdArgs& a, cudaStream_t s) {
  if (a.rows == 0) return 0;
  MMOE_CHECK(a.d % 4 == 0 && a.d <= LN_NV * 128, "layernorm_bwd: d must be a multiple of 4 and <= %d", LN_NV * 128);
  MMOE_CHECK((a.dgamma == nullptr) == (a.dbeta == nullptr), "layernorm_bwd: dgamma and dbeta go together");
  LnBwdDev dev;
  dev.dy = a.dy; dev.x = a.x; dev.stats = a.stats; dev.gamma = a.gamma; dev.dres = a.dres; dev.dres_t = a.dres_t;
  dev.dx = a.dx; dev.dgamma = a.dgamma; dev.dbeta = a.dbeta; dev.g_out = a.g_out; dev.g_colsum = a.g_colsum;
  dev.thresh = a.drop_p > 0.f ? drop_threshold(a.drop_p) : 0u;
  dev.scale = a.drop_p > 0.f ? 1.f / (1.f - a.drop_p) : 1.f;
  dev.k0 = a.k0; dev.k1 = a.k1; dev.rows = a.rows; dev.d = a.d;
  int64_t blocks = (a.rows + 3) / 4;
  const int64_t cap = (int64_t)sm_count() * 3;      // resident blocks: 3 per SM
  if (blocks > cap) blocks = cap;
  MMOE_CHECK((a.dy_dtype == MMOE_F32 || a.dy_dtype == a.dtype) && (a.x_dtype == MMOE_F32 || a.x_dtype == a.dtype),
             "layernorm_bwd: dy and x must be fp32 or the operand dtype");
  const bool df = a.dy_dtype == MMOE_F32, xf = a.x_dtype == MMOE_F32;
  if (a.dtype == MMOE_BF16) return ln_bwd_dispatch<__nv_bfloat16>(dev, df, xf, (int)blocks, s);
  if (a.dtype == MMOE_F16) return ln_bwd_dispatch<__half>(dev, df, xf, (int)blocks, s);
  return ln_bwd_dispatch<float>(dev, true, true, (int)blocks, s);
}

}  // namespace mmoe

// ---------------------------------------------------------------- C ABI (elementary ops)
using namespace mmoe;

extern "C" int mmoe_abi_version(void) { return MMOE_ABI_VERSION; }
extern "C" size_t mmoe_abi_sizeof(int which) {
  switch (which) {
    case 0: return sizeof(mmoe_epilogue);
    case 1: return sizeof(mmoe_gemm_problem);
    case 2: return sizeof(mmoe_call);
    case 3: return sizeof(mmoe_head_cfg);
    case 4: return sizeof(mmoe_cross_cfg);
    case 5: return sizeof(mmoe_fuse_cfg);
    case 6: return sizeof(mmoe_home_cfg);
    default: return 0;
  }
}
extern "C" const char* mmoe_last_error(void) { return mmoe::g_error; }
extern "C" int64_t mmoe_launch_count(int reset) {
  const int64_t n = mmoe::g_launches.load();
  if (reset) mmoe::g_launches.store(0);
  return n;
}
extern "C" int mmoe_init(void) {
  int dev = 0;
  MMOE_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  MMOE_CUDA(cudaGetDeviceProperties(&prop, dev));
  MMOE_CHECK(prop.major == 10, "this library is built for sm_100a only; device %d is sm_%d%d (%s)", dev, prop.major, prop.minor, prop.name);
  return 0;
}
extern "C" int mmoe_cast_f32(const float* x, void* y, int64_t n, int dtype, void* stream) {
  return cast_f32(x, y, n, dtype, (cudaStream_t)stream);
}
extern "C" int mmoe_dropout_mask(uint32_t key0, uint32_t key1, float p, int64_t n, uint8_t* out, void* stream) {
  return dropout_mask(key0, key1, p, n, out, (cudaStream_t)stream);
}
extern "C" int mmoe_site_keys(uint64_t seed, uint32_t site, uint32_t* key0, uint32_t* key1) {
  MMOE_CHECK(key0 != nullptr && key1 != nullptr, "site_keys: null output");
  mmoe::site_keys(seed, site, key0, key1);
  return 0;
}
extern "C" int mmoe_layernorm_fwd(const void* x, int x_dtype, const float* gamma, const float* beta, void* y_t, float* y_f32,
                                  float* stats, int64_t rows, int32_t d, int dtype, void* stream) {
  return layernorm_fwd(x, x_dtype, gamma, beta, y_t, y_f32, stats, rows, d, dtype, (cudaStream_t)stream);
}
