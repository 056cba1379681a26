// HomeExpertWrapper x n + torch.stack — train_HoME.py:100-116 (class), :350-356 (the six calls and the stack that feeds
// HOME_MMoE_Complete).  SURVEY.md §8f row 1.
//
// Reference, per expert e:  y_e = Dropout(SiLU(BatchNorm1d(x_e)))  on [B, d] with PER-RANK batch statistics in training
// (biased variance for the normalisation, unbiased for the running estimate, momentum 0.1, eps 1e-5), then
// expert_vecs = stack([y_0 .. y_{n-1}], dim=1)  -> [B, n, d].  That is 6 x (batch_norm, silu, dropout) launches plus a
// stack copy: ~30 launches and 3 passes over [B, 6, 768].  Here: ONE launch forward, ONE backward.  The data (9.4 MB at
// B = 512) is latency-, not bandwidth-bound, so the layout aims at filling the machine with one wave: a CTA per
// (expert, 32-column slab) = n x d/32 CTAs (144 for 6 x 768 on 148 SMs); a warp reads one 128-byte row segment per
// step (coalesced), 8 warps stride the rows; column statistics are reduced through shared memory.  Statistics use the
// two-pass form (mean, then centred second moment): the re-reads hit L2.
#include "kernels.cuh"

namespace mmoe {

constexpr int WRAP_MAX = 8;
struct WrapDev {
  const float* x[WRAP_MAX];
  const float* gamma[WRAP_MAX]; const float* beta[WRAP_MAX];
  float* run_mean[WRAP_MAX]; float* run_var[WRAP_MAX];
  float* dgamma[WRAP_MAX]; float* dbeta[WRAP_MAX]; float* dx[WRAP_MAX];
  float* out;              // [B, n, d]
  const float* dout;       // [B, n, d]
  float* save_mean; float* save_rstd;   // [n, d]
  int B, n, d, training;
  float eps, momentum, drop_scale;
  uint32_t thresh, k0, k1;
};

__device__ __forceinline__ float block_col_sum(float v, float (*red)[32], int rg, int cl) {
  red[rg][cl] = v;
  __syncthreads();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += red[i][cl];
  __syncthreads();
  return s;
}

__global__ void __launch_bounds__(256) bn_silu_stack_fwd_kernel(const WrapDev a) {
  __shared__ float red[8][32];
  const int slabs = a.d / 32;
  const int e = blockIdx.x / slabs, c = (blockIdx.x % slabs) * 32 + (threadIdx.x & 31);
  const int rg = threadIdx.x >> 5, cl = threadIdx.x & 31;
  const float* __restrict__ x = a.x[e];
  const int B = a.B, d = a.d;
  float mean, rstd;
  if (a.training) {
    float s = 0.f;
    for (int b = rg; b < B; b += 8) s += x[(int64_t)b * d + c];
    mean = block_col_sum(s, red, rg, cl) / (float)B;
    float q = 0.f;
    for (int b = rg; b < B; b += 8) { const float t = x[(int64_t)b * d + c] - mean; q = fmaf(t, t, q); }
    const float var = block_col_sum(q, red, rg, cl) / (float)B;           // biased: what normalises the batch
    rstd = rsqrtf(var + a.eps);
    if (rg == 0) {
      // running estimates (nn.BatchNorm1d): unbiased variance, momentum 0.1
      const float unb = B > 1 ? var * (float)B / (float)(B - 1) : var;
      a.run_mean[e][c] = (1.f - a.momentum) * a.run_mean[e][c] + a.momentum * mean;
      a.run_var[e][c] = (1.f - a.momentum) * a.run_var[e][c] + a.momentum * unb;
    }
  } else {
    mean = a.run_mean[e][c];
    rstd = rsqrtf(a.run_var[e][c] + a.eps);
  }
  if (rg == 0) { a.save_mean[e * d + c] = mean; a.save_rstd[e * d + c] = rstd; }
  const float g = a.gamma[e][c] * rstd, sh = a.beta[e][c] - mean * g;
  for (int b = rg; b < B; b += 8) {
    const float z = fmaf(x[(int64_t)b * d + c], g, sh);
    float y = z * sigmoid_f(z);
    const int64_t o = ((int64_t)b * a.n + e) * d + c;
    if (a.thresh != 0) y = drop_keep(a.k0, a.k1, (uint64_t)o, a.thresh) ? y * a.drop_scale : 0.f;
    a.out[o] = y;
  }
}

__global__ void __launch_bounds__(256) bn_silu_stack_bwd_kernel(const WrapDev a) {
  __shared__ float red[8][32];
  const int slabs = a.d / 32;
  const int e = blockIdx.x / slabs, c = (blockIdx.x % slabs) * 32 + (threadIdx.x & 31);
  const int rg = threadIdx.x >> 5, cl = threadIdx.x & 31;
  const float* __restrict__ x = a.x[e];
  const int B = a.B, d = a.d;
  const float mean = a.save_mean[e * d + c], rstd = a.save_rstd[e * d + c];
  const float gam = a.gamma[e][c], bet = a.beta[e][c];
  // dz = dout * keep/(1-p) * silu'(z),  silu'(z) = s (1 + z (1 - s))
  auto dz_of = [&](int b, float& xhat) {
    xhat = (x[(int64_t)b * d + c] - mean) * rstd;
    const float z = fmaf(xhat, gam, bet);
    const int64_t o = ((int64_t)b * a.n + e) * d + c;
    float g = a.dout[o];
    if (a.thresh != 0) g = drop_keep(a.k0, a.k1, (uint64_t)o, a.thresh) ? g * a.drop_scale : 0.f;
    const float s = sigmoid_f(z);
    return g * s * fmaf(z, 1.f - s, 1.f);
  };
  float s1 = 0.f, s2 = 0.f;
  for (int b = rg; b < B; b += 8) { float xh; const float dz = dz_of(b, xh); s1 += dz; s2 = fmaf(dz, xh, s2); }
  const float dbeta = block_col_sum(s1, red, rg, cl);
  const float dgamma = block_col_sum(s2, red, rg, cl);
  if (rg == 0) {
    if (a.dgamma[e] != nullptr) atomicAdd(a.dgamma[e] + c, dgamma);
    if (a.dbeta[e] != nullptr) atomicAdd(a.dbeta[e] + c, dbeta);
  }
  if (a.dx[e] == nullptr) return;
  const float inv_b = 1.f / (float)B;
  for (int b = rg; b < B; b += 8) {
    float xh; const float dz = dz_of(b, xh);
    float v;
    if (a.training) v = gam * rstd * (dz - inv_b * (dbeta + xh * dgamma));     // batch statistics depend on x
    else v = gam * rstd * dz;                                                   // running statistics are constants
    a.dx[e][(int64_t)b * d + c] = v;
  }
}

static int fill(WrapDev& a, const mmoe_call* c, int n, int d, const float* const* x, float* const* run_mean, float* const* run_var,
                float* save_mean, float* save_rstd, float momentum, float eps) {
  MMOE_CHECK(n >= 1 && n <= WRAP_MAX, "bn_silu_stack: 1..%d experts", WRAP_MAX);
  MMOE_CHECK(d % 32 == 0 && d >= 32, "bn_silu_stack: d must be a multiple of 32 (got %d)", d);
  MMOE_CHECK(c->B >= 0 && c->params != nullptr, "bn_silu_stack: bad call");
  MMOE_CHECK(!(c->training && c->B < 2), "bn_silu_stack: BatchNorm1d needs more than 1 value per channel in training (got B=%d)", c->B);
  a.B = c->B; a.n = n; a.d = d; a.training = c->training; a.eps = eps; a.momentum = momentum;
  const float p = c->training ? c->drop_p : 0.f;
  a.thresh = p > 0.f ? drop_threshold(p) : 0u; a.drop_scale = p > 0.f ? 1.f / (1.f - p) : 1.f;
  site_keys(c->seed, 0, &a.k0, &a.k1);
  for (int e = 0; e < n; ++e) {
    a.x[e] = x[e];
    a.gamma[e] = (const float*)c->params[2 * e]; a.beta[e] = (const float*)c->params[2 * e + 1];
    a.run_mean[e] = run_mean[e]; a.run_var[e] = run_var[e];
    MMOE_CHECK(a.x[e] && a.gamma[e] && a.beta[e] && a.run_mean[e] && a.run_var[e], "bn_silu_stack: null pointer for expert %d", e);
  }
  a.save_mean = save_mean; a.save_rstd = save_rstd;
  return 0;
}

}  // namespace mmoe

using namespace mmoe;

extern "C" int mmoe_bn_silu_stack_fwd(const mmoe_call* c, int32_t n, int32_t d, const float* const* x, float* out, float* save_mean,
                                      float* save_rstd, float* const* running_mean, float* const* running_var, float momentum, float eps) {
  WrapDev a{};
  MMOE_TRY(fill(a, c, n, d, x, running_mean, running_var, save_mean, save_rstd, momentum, eps));
  if (c->B == 0) return 0;
  a.out = out;
  bn_silu_stack_fwd_kernel<<<n * (d / 32), 256, 0, (cudaStream_t)c->stream>>>(a);
  MMOE_LAUNCH_OK("bn_silu_stack_fwd_kernel");
  return 0;
}

extern "C" int mmoe_bn_silu_stack_bwd(const mmoe_call* c, int32_t n, int32_t d, const float* const* x, const float* dout,
                                      const float* save_mean, const float* save_rstd, float* const* running_mean,
                                      float* const* running_var, float* const* dx, float eps) {
  WrapDev a{};
  MMOE_TRY(fill(a, c, n, d, x, running_mean, running_var, const_cast<float*>(save_mean), const_cast<float*>(save_rstd), 0.f, eps));
  if (c->B == 0) return 0;
  MMOE_CHECK(c->grads != nullptr, "bn_silu_stack_bwd: grads missing");
  a.dout = dout;
  for (int e = 0; e < n; ++e) {
    a.dgamma[e] = (float*)c->grads[2 * e]; a.dbeta[e] = (float*)c->grads[2 * e + 1];
    a.dx[e] = dx != nullptr ? dx[e] : nullptr;
  }
  bn_silu_stack_bwd_kernel<<<n * (d / 32), 256, 0, (cudaStream_t)c->stream>>>(a);
  MMOE_LAUNCH_OK("bn_silu_stack_bwd_kernel");
  return 0;
}
