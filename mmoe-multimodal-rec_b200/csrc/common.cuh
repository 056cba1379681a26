// Shared device/host helpers for the sm_100a MMoE kernels.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>

#include "../../include/mmoe_b200.h"

namespace mmoe {

// ---------------------------------------------------------------- errors / launch count
void set_error(const char* fmt, ...);
void count_launch(int n = 1);
cudaError_t last_launch_status(const char* what);

#define MMOE_CHECK(cond, ...)            \
  do {                                   \
    if (!(cond)) {                       \
      ::mmoe::set_error(__VA_ARGS__);    \
      return -1;                         \
    }                                    \
  } while (0)

#define MMOE_CUDA(expr)                                                              \
  do {                                                                               \
    cudaError_t _e = (expr);                                                         \
    if (_e != cudaSuccess) {                                                         \
      ::mmoe::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return -2;                                                                     \
    }                                                                                \
  } while (0)

#define MMOE_LAUNCH_OK(what)                                                         \
  do {                                                                               \
    ::mmoe::count_launch();                                                          \
    cudaError_t _e = cudaGetLastError();                                             \
    if (_e != cudaSuccess) {                                                         \
      ::mmoe::set_error("launch of %s failed: %s", what, cudaGetErrorString(_e));    \
      return -3;                                                                     \
    }                                                                                \
  } while (0)

#define MMOE_TRY(expr)        \
  do {                        \
    int _r = (expr);          \
    if (_r != 0) return _r;   \
  } while (0)

// ---------------------------------------------------------------- programmatic dependent launch (PDL)
// A kernel launched through launch_pdl() may be scheduled while the kernel before it in the stream is still running: its
// CTAs start as SM resources come free, run their on-chip prologue (barrier init, TMEM allocation, descriptor prefetch) and
// block in pdl_wait() until the WHOLE preceding grid has completed and its memory is visible.  Rule for every such kernel:
// nothing before pdl_wait() reads or writes global memory.  pdl_launch_dependents() — issued right AFTER the wait, so that
// at most one dependent grid is ever pre-launched — lets the next PDL kernel do the same behind this one.  The step has
// ~140 back-to-back launches on its critical stream with 2-3 us between them (profiles/r02_exchange_trace.md).  Opt-in
// (MMOE_PDL=1 sets the launch attribute; without it the two instructions are no-ops): see pdl_enabled() for the measurement.
bool pdl_enabled();
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_entry() { pdl_wait(); pdl_launch_dependents(); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

// ---------------------------------------------------------------- dtype helpers
template <typename T> struct DT;
template <> struct DT<float> { static constexpr int id = MMOE_F32; };
template <> struct DT<__nv_bfloat16> { static constexpr int id = MMOE_BF16; };
template <> struct DT<__half> { static constexpr int id = MMOE_F16; };

inline size_t dtype_size(int dtype) { return dtype == MMOE_F32 ? 4 : 2; }

template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float to_f<__half>(__half v) { return __half2float(v); }

template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ __half from_f<__half>(float v) { return __float2half_rn(v); }

// load/store element i of a buffer whose dtype is only known at run time
__device__ __forceinline__ float load_as_f(const void* p, int64_t i, int dtype) {
  if (dtype == MMOE_F32) return reinterpret_cast<const float*>(p)[i];
  if (dtype == MMOE_BF16) return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]);
  return __half2float(reinterpret_cast<const __half*>(p)[i]);
}
__device__ __forceinline__ void store_from_f(void* p, int64_t i, int dtype, float v) {
  if (dtype == MMOE_F32) reinterpret_cast<float*>(p)[i] = v;
  else if (dtype == MMOE_BF16) reinterpret_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16_rn(v);
  else reinterpret_cast<__half*>(p)[i] = __float2half_rn(v);
}

// ---------------------------------------------------------------- math
__device__ __forceinline__ float gelu_f(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
// d/dx gelu(x) = Phi(x) + x * phi(x)
__device__ __forceinline__ float gelu_grad_f(float x) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
  const float pdf = 0.39894228040143267794f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}
__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + __expf(-x)); }

// Branch-free GELU for the 16-bit tensor-core epilogues (8 epilogue warps per SM cannot hide erff's two divergent
// polynomial branches: the exact form made GELU tiles 4x slower than ReLU tiles).  erfc(z) = poly(t) exp(-z^2),
// t = 1/(1 + p z)  (Abramowitz & Stegun 7.1.26, |abs err| <= 1.5e-7; ~6e-7 in fp32 arithmetic) — three orders of
// magnitude below the rounding of the 16-bit value the result is stored as.  fp32 kernels keep erff.
__device__ __forceinline__ void gelu_parts_fast(float x, float& cdf, float& ez) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  const float t = __frcp_rn(fmaf(0.3275911f, z, 1.0f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  ez = __expf(-z * z);
  const float half_c = 0.5f * p * t * ez;
  cdf = x < 0.f ? half_c : 1.0f - half_c;
}
__device__ __forceinline__ float gelu_fast_f(float x) { float cdf, ez; gelu_parts_fast(x, cdf, ez); return x * cdf; }
__device__ __forceinline__ float gelu_grad_fast_f(float x) {
  float cdf, ez; gelu_parts_fast(x, cdf, ez);
  return fmaf(x * 0.39894228040143267794f, ez, cdf);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---------------------------------------------------------------- L2 prefetch
// The row / sample streaming kernels hold one unit of work per warp in registers, so a warp alternates between waiting for
// its loads and computing; with 8-12 resident warps per SM that left HBM at 40-58 % (ncu: long-scoreboard stalls, DRAM
// throughput 40 %).  Asking L2 for the warp's NEXT unit while it works on the current one keeps DRAM streaming without any
// extra registers or shared memory: the demand loads then hit L2.
//   prefetch_l2_rows : every lane touches its own 128-byte lines of [ptr, ptr + bytes)
//   prefetch_l2_bulk : one bulk request (TMA engine), bytes a multiple of 16, ptr 16-byte aligned
__device__ __forceinline__ void prefetch_l2_rows(const void* ptr, int bytes, int lane) {
  const char* p = reinterpret_cast<const char*>(ptr);
  for (int off = lane * 128; off < bytes; off += 32 * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + off));
}
__device__ __forceinline__ void prefetch_l2_bulk(const void* ptr, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(ptr), "r"(bytes) : "memory");
}

// ---------------------------------------------------------------- dropout keep-mask
// A counter-based hash of the flat element index, keyed per dropout site.  Any kernel can regenerate the mask in
// backward from (key0, key1, index); no mask tensor is ever stored.  One 32-bit hash serves the element PAIR
// (2i, 2i+1): 16 bits each, keep iff bits >= p * 2^16  (p = 0.1 -> 6553/65536 = 0.09999).
__host__ __device__ __forceinline__ uint32_t drop_hash(uint32_t key0, uint32_t key1, uint64_t pair_idx) {
  uint32_t x = ((uint32_t)pair_idx ^ ((uint32_t)(pair_idx >> 32) * 0x632BE5ABu) ^ key0) * 0x9E3779B1u;
  x ^= x >> 15; x *= 0x85EBCA77u;
  x ^= x >> 13; x = (x ^ key1) * 0xC2B2AE3Du;
  x ^= x >> 16;
  return x;
}
__host__ __device__ __forceinline__ uint32_t drop_threshold(float p) { return (uint32_t)(p * 65536.0f); }
__host__ __device__ __forceinline__ bool drop_keep(uint32_t key0, uint32_t key1, uint64_t idx, uint32_t thresh) {
  const uint32_t h = drop_hash(key0, key1, idx >> 1);
  return ((h >> ((uint32_t)(idx & 1) * 16)) & 0xFFFFu) >= thresh;
}
// per-site keys derived on the host from the call seed
inline void site_keys(uint64_t seed, uint32_t site, uint32_t* k0, uint32_t* k1) {
  uint64_t z = seed + 0x9E3779B97F4A7C15ull * (uint64_t)(site + 1);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  *k0 = (uint32_t)z;
  *k1 = (uint32_t)(z >> 32);
}

inline int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// ---------------------------------------------------------------- GEMM engine entry (gemm.cu)
int gemm_grouped(const mmoe_gemm_problem* problems, int n_problems, int dtype, int engine, cudaStream_t stream);
bool gemm_bitmask_supported(int dtype, int64_t M, int N);

// convenience builders used by the orchestrators
inline mmoe_epilogue epi_none() {
  mmoe_epilogue e{};
  e.alpha = 1.0f;
  return e;
}
inline mmoe_gemm_problem gemm_problem(const void* a, int64_t lda, int a_major, const void* b, int64_t ldb, int b_major,
                                      int M, int N, int K, const mmoe_epilogue& e, int k_splits = 1) {
  mmoe_gemm_problem p{};
  p.a = a; p.lda = lda; p.a_major = a_major;
  p.b = b; p.ldb = ldb; p.b_major = b_major;
  p.M = M; p.N = N; p.K = K; p.k_splits = k_splits; p.epi = e;
  return p;
}

}  // namespace mmoe
