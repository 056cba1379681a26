// Loss and metric tail of the training / evaluation scripts:
//
//  * two-task BCE-with-logits with pos_weight, mean reduction — train.py:189-192, 253-254 (nn.BCEWithLogitsLoss x 2):
//    forward value and d loss / d logits in ONE launch over the [2, B] logits the head kernel wrote.
//  * InfoNCE contrastive loss — train_HoME.py:43-51 (calculate_contrastive_loss) for the three (anchor, positive) pairs
//    of train_HoME.py:362-364: row L2-normalisation, [B,d] x [d,B] similarity GEMMs on the tcgen05 engine (all pairs in
//    one grouped launch), fused row log-sum-exp / cross-entropy that also writes d sim, two grouped GEMM launches and one
//    normalisation-backward kernel for the gradients.
//  * ROC-AUC on the device — inference_and_auc.py:150-178 (sklearn.metrics.roc_auc_score on the concatenated sigmoid
//    scores): bitonic sort of (order-preserving score key | label) words + rank-sum (Mann-Whitney U with average ranks for
//    ties, which is what the trapezoidal ROC area equals), so a scoring sweep needs no per-batch .cpu().numpy() sync.
#include "kernels.cuh"

namespace mmoe {

// ------------------------------------------------------------------------------------------ BCE
// PyTorch's stable form (binary_cross_entropy_with_logits): l = (1 - y) x + (1 + (pw - 1) y) * (log1p(exp(-|x|)) + max(-x, 0))
__global__ void __launch_bounds__(256) bce2_kernel(const float* __restrict__ logits, const float* __restrict__ y0,
                                                   const float* __restrict__ y1, float pw0, float pw1, int B, float* loss,
                                                   float* dlogits, float gscale) {
  __shared__ float red[8];
  float acc = 0.f;
  const float inv_b = 1.f / (float)B;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 2 * B; i += gridDim.x * blockDim.x) {
    const int t = i >= B;
    const float x = logits[i], y = t ? y1[i - B] : y0[i], pw = t ? pw1 : pw0;
    const float lw = 1.f + (pw - 1.f) * y;
    acc += ((1.f - y) * x + lw * (log1pf(expf(-fabsf(x))) + fmaxf(-x, 0.f))) * inv_b;
    // d/dx = (1 - y) - lw * sigmoid(-x)
    if (dlogits != nullptr) dlogits[i] = ((1.f - y) - lw / (1.f + expf(x))) * inv_b * gscale;
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int w = 0; w < 8; ++w) s += red[w];
    atomicAdd(loss, s);
  }
}

// ------------------------------------------------------------------------------------------ InfoNCE pieces
// xn = x / max(||x||, 1e-12) (F.normalize), one warp per row; writes T copy (GEMM operand) and the inverse norm
template <typename T>
__global__ void __launch_bounds__(256) l2norm_fwd_kernel(const float* __restrict__ x, T* __restrict__ xn, float* __restrict__ inv_norm,
                                                         int rows, int d) {
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (r >= rows) return;
  const float* xr = x + (int64_t)r * d;
  float s = 0.f;
  for (int c = lane; c < d; c += 32) s = fmaf(xr[c], xr[c], s);
  s = warp_sum(s);
  const float inv = 1.f / fmaxf(sqrtf(s), 1e-12f);
  if (lane == 0) inv_norm[r] = inv;
  for (int c = lane; c < d; c += 32) xn[(int64_t)r * d + c] = from_f<T>(xr[c] * inv);
}
// dx = inv * (dn - n <n, dn>),  n = x * inv;   dx accumulated into the caller's buffer (an input may feed several pairs)
__global__ void __launch_bounds__(256) l2norm_bwd_kernel(const float* __restrict__ x, const float* __restrict__ inv_norm,
                                                         const float* __restrict__ dn, float* __restrict__ dx, int rows, int d, float gscale) {
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (r >= rows) return;
  const float inv = inv_norm[r];
  const float* xr = x + (int64_t)r * d; const float* gr = dn + (int64_t)r * d;
  float dot = 0.f;
  for (int c = lane; c < d; c += 32) dot = fmaf(xr[c] * inv, gr[c], dot);
  dot = warp_sum(dot);
  for (int c = lane; c < d; c += 32) dx[(int64_t)r * d + c] += gscale * inv * (gr[c] - xr[c] * inv * dot);
}
// row i of sim [B,B] (already divided by the temperature): loss += (lse_i - sim_ii) / B;
// dsim_ij = (softmax_ij - [i == j]) / (B * temperature)   — the gradient w.r.t. the UNSCALED similarity, as a T GEMM operand
template <typename T>
__global__ void __launch_bounds__(256) ce_diag_kernel(const float* __restrict__ sim, T* __restrict__ dsim, float* loss, int B, float inv_temp) {
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (r >= B) return;
  const float* s = sim + (int64_t)r * B;
  float m = -INFINITY;
  for (int c = lane; c < B; c += 32) m = fmaxf(m, s[c]);
  m = warp_max(m);
  float z = 0.f;
  for (int c = lane; c < B; c += 32) z += expf(s[c] - m);
  z = warp_sum(z);
  const float lse = m + logf(z), inv_b = 1.f / (float)B;
  if (lane == 0) atomicAdd(loss, (lse - s[r]) * inv_b);
  for (int c = lane; c < B; c += 32)
    dsim[(int64_t)r * B + c] = from_f<T>((expf(s[c] - lse) - (c == r ? 1.f : 0.f)) * inv_b * inv_temp);
}

constexpr int NCE_MAX = 4;
struct NceLayout { void* an[NCE_MAX]; void* pn[NCE_MAX]; float* ia[NCE_MAX]; float* ip[NCE_MAX]; void* dsim[NCE_MAX]; float* sim; float* dan; float* dpn; };
static NceLayout nce_layout(char* saved, char* work, int n, int B, int d, size_t es) {
  NceLayout L{};
  size_t off = 0;
  auto take = [&](char* base, size_t& o, size_t bytes) { o = (o + 255) & ~(size_t)255; char* p = base + o; o += bytes; return (void*)p; };
  for (int i = 0; i < n; ++i) {
    L.an[i] = take(saved, off, (size_t)B * d * es); L.pn[i] = take(saved, off, (size_t)B * d * es);
    L.ia[i] = (float*)take(saved, off, (size_t)B * 4); L.ip[i] = (float*)take(saved, off, (size_t)B * 4);
    L.dsim[i] = take(saved, off, (size_t)B * B * es);
  }
  size_t w = 0;
  L.sim = (float*)take(work, w, (size_t)n * B * B * 4);
  L.dan = (float*)take(work, w, (size_t)B * d * 4);
  L.dpn = (float*)take(work, w, (size_t)B * d * 4);
  return L;
}

// ------------------------------------------------------------------------------------------ AUC
// key = order-preserving map of the float score to uint32, word = key << 1 | label  (so equal scores sort negatives first;
// tie groups are recovered from the key alone)
__device__ __forceinline__ uint32_t float_key(float f) {
  uint32_t u = __float_as_uint(f);
  if ((u << 1) == 0u) u = 0u;                     // -0.0 and +0.0 are one value (a tie), not neighbours
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__global__ void auc_pack_kernel(const float* __restrict__ score, const float* __restrict__ label, uint64_t* __restrict__ w, int n, int n_pad) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pad) return;
  w[i] = i < n ? (((uint64_t)float_key(score[i]) << 1) | (label[i] > 0.5f ? 1ull : 0ull)) : ~0ull;   // padding sorts last
}
// one bitonic compare-exchange step (k = size of the bitonic sequences being merged, j = partner distance)
__global__ void bitonic_global_kernel(uint64_t* w, int n_pad, int k, int j) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int p = i ^ j;
  if (i >= n_pad || p <= i) return;
  const uint64_t a = w[i], b = w[p];
  const bool up = (i & k) == 0;
  if ((a > b) == up) { w[i] = b; w[p] = a; }
}
// all steps with j < 1024 of the merges k0 .. k1 inside a 2048-element shared-memory tile
__global__ void __launch_bounds__(1024) bitonic_tile_kernel(uint64_t* w, int k_first, int k_last) {
  __shared__ uint64_t t[2048];
  const int base = blockIdx.x * 2048;
  t[threadIdx.x] = w[base + threadIdx.x];
  t[threadIdx.x + 1024] = w[base + threadIdx.x + 1024];
  __syncthreads();
  for (int k = k_first; k <= k_last; k <<= 1) {
    for (int j = min(k >> 1, 1024); j > 0; j >>= 1) {
      // thread handles the pair (i, i ^ j) with i the element whose bit j is clear
      const int i = ((threadIdx.x & ~(j - 1)) << 1) | (threadIdx.x & (j - 1));
      const int p = i | j;
      const uint64_t a = t[i], b = t[p];
      const bool up = ((base + i) & k) == 0;
      if ((a > b) == up) { t[i] = b; t[p] = a; }
      __syncthreads();
    }
  }
  w[base + threadIdx.x] = t[threadIdx.x];
  w[base + threadIdx.x + 1024] = t[threadIdx.x + 1024];
}
// rank sum of the positives with average ranks over tie groups; acc[0] += sum of ranks, acc[1] += #positives
__global__ void __launch_bounds__(256) auc_ranksum_kernel(const uint64_t* __restrict__ w, int n, double* acc) {
  __shared__ double red[2][8];
  double rs = 0.0, np = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint64_t v = w[i];
    if (!(v & 1ull)) continue;
    const uint64_t key = v >> 1;
    int lo = 0, hi = i;                       // first index whose key == key
    while (lo < hi) { const int mid = (lo + hi) >> 1; if ((w[mid] >> 1) < key) lo = mid + 1; else hi = mid; }
    const int first = lo;
    lo = i; hi = n;                           // one past the last index whose key == key
    while (lo < hi) { const int mid = (lo + hi) >> 1; if ((w[mid] >> 1) <= key) lo = mid + 1; else hi = mid; }
    rs += 0.5 * ((double)first + 1.0 + (double)lo);      // average of the 1-based ranks first+1 .. lo
    np += 1.0;
  }
  for (int o = 16; o > 0; o >>= 1) { rs += __shfl_xor_sync(0xffffffffu, rs, o); np += __shfl_xor_sync(0xffffffffu, np, o); }
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = rs; red[1][threadIdx.x >> 5] = np; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
    for (int q = 0; q < 8; ++q) { a += red[0][q]; b += red[1][q]; }
    atomicAdd(acc, a); atomicAdd(acc + 1, b);
  }
}
__global__ void auc_final_kernel(const double* acc, int n, double* out) {
  const double np = acc[1], nn = (double)n - np;
  out[0] = (np > 0.0 && nn > 0.0) ? (acc[0] - np * (np + 1.0) * 0.5) / (np * nn) : nan("");
}

static int next_pow2(int n) { int p = 2048; while (p < n) p <<= 1; return p; }

}  // namespace mmoe

using namespace mmoe;

// loss (1 float, device) is ACCUMULATED into: zero it first.  dlogits may be NULL (evaluation).  gscale multiplies the
// gradient (e.g. 1/grad_accum * GradScaler scale when the caller folds them in; 1 otherwise).
extern "C" int mmoe_bce2_fwd_bwd(const float* logits, const float* y_good, const float* y_best, float pos_weight_good,
                                 float pos_weight_best, int32_t B, float* loss, float* dlogits, float gscale, void* stream) {
  MMOE_CHECK(B >= 1 && logits && y_good && y_best && loss, "bce2: bad arguments");
  int blocks = (2 * B + 255) / 256;
  if (blocks > 2 * sm_count()) blocks = 2 * sm_count();
  bce2_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(logits, y_good, y_best, pos_weight_good, pos_weight_best, B, loss, dlogits, gscale);
  MMOE_LAUNCH_OK("bce2_kernel");
  return 0;
}

extern "C" size_t mmoe_info_nce_saved_bytes(int32_t n_pairs, int32_t B, int32_t d, int dtype) {
  const size_t es = dtype_size(dtype);
  return (size_t)n_pairs * (2 * ((size_t)B * d * es + 256) + 2 * ((size_t)B * 4 + 256) + (size_t)B * B * es + 256) + 256;
}
extern "C" size_t mmoe_info_nce_workspace_bytes(int32_t n_pairs, int32_t B, int32_t d, int dtype) {
  (void)dtype;
  return (size_t)n_pairs * B * B * 4 + 2 * ((size_t)B * d * 4 + 256) + 512;
}

template <typename T>
static int nce_fwd_t(const mmoe_call* c, int n, int d, const float* const* anchor, const float* const* positive, float temperature, float* loss) {
  const int B = c->B, dtype = c->dtype;
  cudaStream_t st = (cudaStream_t)c->stream;
  NceLayout L = nce_layout((char*)c->saved, (char*)c->workspace, n, B, d, sizeof(T));
  const int rb = (B + 7) / 8;
  mmoe_gemm_problem p[NCE_MAX];
  for (int i = 0; i < n; ++i) {
    l2norm_fwd_kernel<T><<<rb, 256, 0, st>>>(anchor[i], (T*)L.an[i], L.ia[i], B, d);
    MMOE_LAUNCH_OK("l2norm_fwd_kernel");
    l2norm_fwd_kernel<T><<<rb, 256, 0, st>>>(positive[i], (T*)L.pn[i], L.ip[i], B, d);
    MMOE_LAUNCH_OK("l2norm_fwd_kernel");
    mmoe_epilogue e = epi_none();
    e.out = L.sim + (size_t)i * B * B; e.out_dtype = MMOE_F32; e.ldo = B; e.alpha = 1.f / temperature;
    p[i] = gemm_problem(L.an[i], d, 0, L.pn[i], d, 0, B, B, d, e);          // sim = an pn^T / temperature
  }
  MMOE_TRY(gemm_grouped(p, n, dtype, 0, st));
  for (int i = 0; i < n; ++i) {
    ce_diag_kernel<T><<<rb, 256, 0, st>>>(L.sim + (size_t)i * B * B, (T*)L.dsim[i], loss + i, B, 1.f / temperature);
    MMOE_LAUNCH_OK("ce_diag_kernel");
  }
  return 0;
}
template <typename T>
static int nce_bwd_t(const mmoe_call* c, int n, int d, const float* const* anchor, const float* const* positive, const float* dloss,
                     float* const* d_anchor, float* const* d_positive) {
  const int B = c->B, dtype = c->dtype;
  cudaStream_t st = (cudaStream_t)c->stream;
  NceLayout L = nce_layout((char*)c->saved, (char*)c->workspace, n, B, d, sizeof(T));
  const int rb = (B + 7) / 8;
  float h_dloss[NCE_MAX];
  for (int i = 0; i < n; ++i) h_dloss[i] = dloss[i];
  for (int i = 0; i < n; ++i) {
    // d an = dsim pn ;  d pn = dsim^T an     (one grouped launch per pair: both read dsim)
    mmoe_epilogue ea = epi_none(), ep = epi_none();
    ea.out = L.dan; ea.out_dtype = MMOE_F32; ea.ldo = d;
    ep.out = L.dpn; ep.out_dtype = MMOE_F32; ep.ldo = d;
    mmoe_gemm_problem p[2] = {gemm_problem(L.dsim[i], B, 0, L.pn[i], d, 1, B, d, B, ea),
                              gemm_problem(L.dsim[i], B, 1, L.an[i], d, 1, B, d, B, ep)};
    MMOE_TRY(gemm_grouped(p, 2, dtype, 0, st));
    if (d_anchor[i] != nullptr) {
      l2norm_bwd_kernel<<<rb, 256, 0, st>>>(anchor[i], L.ia[i], L.dan, d_anchor[i], B, d, h_dloss[i]);
      MMOE_LAUNCH_OK("l2norm_bwd_kernel");
    }
    if (d_positive[i] != nullptr) {
      l2norm_bwd_kernel<<<rb, 256, 0, st>>>(positive[i], L.ip[i], L.dpn, d_positive[i], B, d, h_dloss[i]);
      MMOE_LAUNCH_OK("l2norm_bwd_kernel");
    }
  }
  return 0;
}

static int nce_check(const mmoe_call* c, int n, int d) {
  MMOE_CHECK(n >= 1 && n <= NCE_MAX, "info_nce: 1..%d pairs", NCE_MAX);
  MMOE_CHECK(c->B >= 1 && d >= 8 && d % 8 == 0 && c->B % 8 == 0, "info_nce: B and d must be multiples of 8 (B=%d d=%d)", c->B, d);
  MMOE_CHECK(c->saved && c->saved_bytes >= mmoe_info_nce_saved_bytes(n, c->B, d, c->dtype), "info_nce: saved blob too small");
  MMOE_CHECK(c->workspace && c->workspace_bytes >= mmoe_info_nce_workspace_bytes(n, c->B, d, c->dtype), "info_nce: workspace too small");
  return 0;
}
// loss: n_pairs floats (device), ACCUMULATED into (zero first).  anchor / positive: host arrays of n_pairs device pointers to
// fp32 [B, d].  Backward: dloss = host array of the n_pairs upstream scalars; d_anchor / d_positive: host arrays of fp32
// [B, d] device buffers that are ACCUMULATED into (an input that appears in several pairs passes the same buffer), NULL = skip.
extern "C" int mmoe_info_nce_fwd(const mmoe_call* c, int32_t n_pairs, int32_t d, const float* const* anchor, const float* const* positive,
                                 float temperature, float* loss) {
  MMOE_TRY(nce_check(c, n_pairs, d));
  if (c->dtype == MMOE_BF16) return nce_fwd_t<__nv_bfloat16>(c, n_pairs, d, anchor, positive, temperature, loss);
  if (c->dtype == MMOE_F16) return nce_fwd_t<__half>(c, n_pairs, d, anchor, positive, temperature, loss);
  return nce_fwd_t<float>(c, n_pairs, d, anchor, positive, temperature, loss);
}
extern "C" int mmoe_info_nce_bwd(const mmoe_call* c, int32_t n_pairs, int32_t d, const float* const* anchor, const float* const* positive,
                                 const float* dloss, float* const* d_anchor, float* const* d_positive) {
  MMOE_TRY(nce_check(c, n_pairs, d));
  if (c->dtype == MMOE_BF16) return nce_bwd_t<__nv_bfloat16>(c, n_pairs, d, anchor, positive, dloss, d_anchor, d_positive);
  if (c->dtype == MMOE_F16) return nce_bwd_t<__half>(c, n_pairs, d, anchor, positive, dloss, d_anchor, d_positive);
  return nce_bwd_t<float>(c, n_pairs, d, anchor, positive, dloss, d_anchor, d_positive);
}

extern "C" size_t mmoe_auc_workspace_bytes(int64_t n) {
  return (size_t)next_pow2((int)n) * 8 + 64;
}
// scores, labels: fp32 [n] on the device (label > 0.5 = positive); out: 1 double on the device (NaN if a class is empty).
extern "C" int mmoe_auc(const float* scores, const float* labels, int64_t n, void* workspace, size_t workspace_bytes, double* out, void* stream) {
  MMOE_CHECK(n >= 1 && n <= (1 << 28), "auc: n out of range");
  MMOE_CHECK(workspace != nullptr && workspace_bytes >= mmoe_auc_workspace_bytes(n), "auc: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const int n_pad = next_pow2((int)n);
  uint64_t* w = (uint64_t*)workspace;
  double* acc = (double*)(w + n_pad);
  MMOE_CUDA(cudaMemsetAsync(acc, 0, 16, st));
  auc_pack_kernel<<<(n_pad + 255) / 256, 256, 0, st>>>(scores, labels, w, (int)n, n_pad);
  MMOE_LAUNCH_OK("auc_pack_kernel");
  bitonic_tile_kernel<<<n_pad / 2048, 1024, 0, st>>>(w, 2, 2048);          // sorted runs of 2048, alternating direction
  MMOE_LAUNCH_OK("bitonic_tile_kernel");
  for (int k = 4096; k <= n_pad; k <<= 1) {
    for (int j = k >> 1; j >= 2048; j >>= 1) {
      bitonic_global_kernel<<<(n_pad + 255) / 256, 256, 0, st>>>(w, n_pad, k, j);
      MMOE_LAUNCH_OK("bitonic_global_kernel");
    }
    bitonic_tile_kernel<<<n_pad / 2048, 1024, 0, st>>>(w, k, k);
    MMOE_LAUNCH_OK("bitonic_tile_kernel");
  }
  int blocks = (int)((n + 255) / 256);
  if (blocks > 4 * sm_count()) blocks = 4 * sm_count();
  auc_ranksum_kernel<<<blocks, 256, 0, st>>>(w, (int)n, acc);
  MMOE_LAUNCH_OK("auc_ranksum_kernel");
  auc_final_kernel<<<1, 1, 0, st>>>(acc, (int)n, out);
  MMOE_LAUNCH_OK("auc_final_kernel");
  return 0;
}
