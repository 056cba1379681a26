// The data formats either side of the fusion path (SURVEY.md §8f rows 2 and 3, K10 of §2.5):
//
//  * patch projection — the Conv2d(3, 768, 16, stride 16) inside HF ViTPatchEmbeddings that ItemImageExpert's backbone call
//    reaches (model.py:373-376).  Non-overlapping 16x16 patches make the convolution a plain GEMM
//    [B*196, 768] x [768, 768]^T on rows laid out (channel, py, px) — which is EXACTLY the byte layout newpatch.py:102-104
//    writes to disk ([196, 3*16*16] uint8) and data4model.py:254-258 ships as patch.bin.  The reference un-patchifies those
//    bytes on the CPU into a float [3,224,224] image (model.py:160-178), normalises, uploads 4 bytes per value, and the
//    conv re-patchifies.  Here the raw bytes are uploaded (1 byte per value), one kernel turns them into the 16-bit A operand
//    (0..255 are exact in bf16/fp16; the /255, mean and std of model.py:172-174 are folded into the weights and bias by
//    the caller, see ingest.py), and the tcgen05 GEMM engine does the projection.  A second entry point gathers the same A
//    operand from an already normalised float image, for callers that hand over [B,3,224,224] like the unchanged scripts.
//  * TextExpert's post-encoder step — model.py:286-338: gather the <SENT> hidden states, bucket per sample, pad to 64
//    sentence slots, data-derived padding mask, masked mean, LayerNorm, dropout.  The reference loops over samples in
//    Python (torch.cat / F.pad per sample, ~4B tiny launches); here one launch, a CTA per sample, a warp per sentence row.
#include "kernels.cuh"

namespace mmoe {

// ------------------------------------------------------------------------------------------ patches
// u8 [rows, 768] -> T [rows, 768], 16 bytes in / 32 bytes out per thread step
template <typename T>
__global__ void __launch_bounds__(256) patch_u8_to_t_kernel(const uint8_t* __restrict__ in, T* __restrict__ out, int64_t n16) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (int64_t)gridDim.x * blockDim.x) {
    const uint4 v = reinterpret_cast<const uint4*>(in)[i];
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    T o[16];
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int b = 0; b < 4; ++b) o[4 * q + b] = from_f<T>((float)((w[q] >> (8 * b)) & 0xFFu));
    uint4* dst = reinterpret_cast<uint4*>(out + i * 16);
    if (sizeof(T) == 2) { dst[0] = reinterpret_cast<const uint4*>(o)[0]; dst[1] = reinterpret_cast<const uint4*>(o)[1]; }
    else { for (int q = 0; q < 4; ++q) dst[q] = reinterpret_cast<const uint4*>(o)[q]; }
  }
}
// float image [B, C, H, W] -> patch-major T rows [B * (H/p)*(W/p), C*p*p] (the im2col of a stride-p, kernel-p convolution: a
// pure permutation).  One thread per 4 consecutive px of one (patch, channel, py) row segment: 16-byte reads, 8-byte writes.
template <typename T>
__global__ void __launch_bounds__(256) patchify_kernel(const float* __restrict__ img, T* __restrict__ out, int B, int Cc, int H, int W, int p) {
  const int gw = W / p, gh = H / p, quads = p / 4;
  const int64_t total = (int64_t)B * gh * gw * Cc * p * quads;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i;
    const int q = (int)(r % quads); r /= quads;
    const int py = (int)(r % p); r /= p;
    const int c = (int)(r % Cc); r /= Cc;
    const int px_blk = (int)(r % gw); r /= gw;
    const int py_blk = (int)(r % gh); r /= gh;
    const int b = (int)r;
    const float4 v = *reinterpret_cast<const float4*>(img + (((int64_t)b * Cc + c) * H + (py_blk * p + py)) * W + px_blk * p + q * 4);
    T* dst = out + (((int64_t)b * gh + py_blk) * gw + px_blk) * ((int64_t)Cc * p * p) + ((int64_t)c * p + py) * p + q * 4;
    dst[0] = from_f<T>(v.x); dst[1] = from_f<T>(v.y); dst[2] = from_f<T>(v.z); dst[3] = from_f<T>(v.w);
  }
}

// ------------------------------------------------------------------------------------------ sentence gather
struct GatherDev {
  const void* h; int h_dtype; int64_t ld_h;     // encoder hidden states, rows of d values; row index from src
  const int32_t* src;                           // [B, S]: row of h feeding sentence slot (b, s), or -1 for an empty slot
  const float* gamma; const float* beta;        // TextExpert.norm (null = HoME variant without the final LayerNorm/dropout)
  float* sent; uint8_t* mask; float* doc;       // [B,S,d], [B,S], [B,d]
  float* pre_doc; float* stats;                 // saved for backward: un-normalised doc [B,d]; (mean, rstd) per row [B*(S+1), 2]
  // backward
  const float* d_sent; const float* d_doc; float* dh; float* dgamma; float* dbeta;
  int B, S, d;
  uint32_t thresh, k0, k1, k2, k3; float drop_scale;
};

// One CTA per sample; warp w handles sentence slots w, w+8, ...  d <= 1024 (values kept in registers, 32 per lane max).
template <int VPL>
__global__ void __launch_bounds__(256) sent_gather_fwd_kernel(const GatherDev a) {
  __shared__ float dsum[8][VPL * 32];
  __shared__ int cnt[8];
  const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, d = a.d, S = a.S;
  float acc[VPL];
#pragma unroll
  for (int v = 0; v < VPL; ++v) acc[v] = 0.f;
  int valid = 0;
  for (int s = warp; s < S; s += 8) {
    const int64_t row = a.src[(int64_t)b * S + s];
    float x[VPL];
    float asum = 0.f;
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
      const int c = lane + 32 * v;
      x[v] = (row >= 0 && c < d) ? load_as_f(a.h, row * a.ld_h + c, a.h_dtype) : 0.f;
      asum += fabsf(x[v]);
    }
    asum = warp_sum(asum);
    const bool pad = asum == 0.f;                 // model.py:328: a slot is padding iff its values sum (abs) to exactly 0
    if (lane == 0) a.mask[(int64_t)b * S + s] = pad ? 1 : 0;
    if (!pad) ++valid;
#pragma unroll
    for (int v = 0; v < VPL; ++v) acc[v] += x[v];
    // sentence row: LayerNorm + dropout (v1), or the raw row (HoME: model_HoME.py:366-367 commented out)
    float* out = a.sent + ((int64_t)b * S + s) * d;
    if (a.gamma != nullptr) {
      float m = 0.f;
#pragma unroll
      for (int v = 0; v < VPL; ++v) m += x[v];
      m = warp_sum(m) / (float)d;
      float q = 0.f;
#pragma unroll
      for (int v = 0; v < VPL; ++v) { const int c = lane + 32 * v; const float t = c < d ? x[v] - m : 0.f; q = fmaf(t, t, q); }
      const float rstd = rsqrtf(warp_sum(q) / (float)d + 1e-5f);
      if (lane == 0) { a.stats[((int64_t)b * (S + 1) + s) * 2] = m; a.stats[((int64_t)b * (S + 1) + s) * 2 + 1] = rstd; }
#pragma unroll
      for (int v = 0; v < VPL; ++v) {
        const int c = lane + 32 * v;
        if (c >= d) continue;
        float y = (x[v] - m) * rstd * a.gamma[c] + a.beta[c];
        const uint64_t idx = ((uint64_t)b * S + s) * d + c;
        if (a.thresh != 0) y = drop_keep(a.k0, a.k1, idx, a.thresh) ? y * a.drop_scale : 0.f;
        out[c] = y;
      }
    } else {
#pragma unroll
      for (int v = 0; v < VPL; ++v) { const int c = lane + 32 * v; if (c < d) out[c] = x[v]; }
    }
  }
#pragma unroll
  for (int v = 0; v < VPL; ++v) dsum[warp][lane + 32 * v] = acc[v];
  if (lane == 0) cnt[warp] = valid;
  __syncthreads();
  // doc = sum over slots / max(#non-pad, 1)  (model.py:331-332), then LayerNorm + dropout (v1)
  if (warp == 0) {
    int n = 0;
    for (int w = 0; w < 8; ++w) n += cnt[w];
    const float inv = 1.f / (float)max(n, 1);
    float x[VPL];
    float m = 0.f;
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
      float s = 0.f;
      for (int w = 0; w < 8; ++w) s += dsum[w][lane + 32 * v];
      x[v] = s * inv;
      const int c = lane + 32 * v;
      if (c < d) { a.pre_doc[(int64_t)b * d + c] = x[v]; m += x[v]; }
    }
    float* out = a.doc + (int64_t)b * d;
    if (a.gamma != nullptr) {
      m = warp_sum(m) / (float)d;
      float q = 0.f;
#pragma unroll
      for (int v = 0; v < VPL; ++v) { const int c = lane + 32 * v; const float t = c < d ? x[v] - m : 0.f; q = fmaf(t, t, q); }
      const float rstd = rsqrtf(warp_sum(q) / (float)d + 1e-5f);
      if (lane == 0) { a.stats[((int64_t)b * (S + 1) + S) * 2] = m; a.stats[((int64_t)b * (S + 1) + S) * 2 + 1] = rstd; }
#pragma unroll
      for (int v = 0; v < VPL; ++v) {
        const int c = lane + 32 * v;
        if (c >= d) continue;
        float y = (x[v] - m) * rstd * a.gamma[c] + a.beta[c];
        if (a.thresh != 0) y = drop_keep(a.k2, a.k3, (uint64_t)b * d + c, a.thresh) ? y * a.drop_scale : 0.f;
        out[c] = y;
      }
    } else {
#pragma unroll
      for (int v = 0; v < VPL; ++v) { const int c = lane + 32 * v; if (c < d) out[c] = x[v]; }
    }
  }
}

// Backward: d h[row] += LN'(drop'(d_sent[b,s])) + (slot non-pad ? LN'(drop'(d_doc[b])) / n_b : 0); d gamma / d beta.
// (The padding mask is data-derived and piecewise constant: no gradient through it.  A padded slot's x is all zeros, so it
// adds nothing to the doc sum; its own LayerNorm path still yields a gradient into h when it maps to a real row whose values
// happen to be all zero — kept, as in the reference.)
template <int VPL>
__global__ void __launch_bounds__(256) sent_gather_bwd_kernel(const GatherDev a) {
  __shared__ float ddoc[VPL * 32];
  __shared__ float part[8][VPL * 32];        // per-warp partial column sums (gamma first, then beta: 48 KB static limit)
  __shared__ int cnt_s;
  const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, d = a.d, S = a.S;
  if (threadIdx.x == 0) {
    int n = 0;
    for (int s = 0; s < S; ++s) n += a.mask[(int64_t)b * S + s] ? 0 : 1;
    cnt_s = max(n, 1);
  }
  float ag[VPL], ab[VPL];
#pragma unroll
  for (int v = 0; v < VPL; ++v) { ag[v] = 0.f; ab[v] = 0.f; }
  // LayerNorm backward of one row held in registers: g = upstream (already dropout-masked), x = LN input
  auto ln_bwd_row = [&](const float (&x)[VPL], float (&g)[VPL], float m, float rstd) {
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
      const int c = lane + 32 * v;
      if (c >= d) { g[v] = 0.f; continue; }
      const float xh = (x[v] - m) * rstd;
      ag[v] = fmaf(g[v], xh, ag[v]); ab[v] += g[v];
      g[v] *= a.gamma[c];
      s1 += g[v]; s2 = fmaf(g[v], xh, s2);
    }
    s1 = warp_sum(s1) / (float)d; s2 = warp_sum(s2) / (float)d;
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
      const int c = lane + 32 * v;
      if (c < d) g[v] = rstd * (g[v] - s1 - (x[v] - m) * rstd * s2);
    }
  };
  // the doc path first (warp 0), its input gradient is shared by every non-pad slot
  if (warp == 0) {
    float x[VPL], g[VPL];
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
      const int c = lane + 32 * v;
      x[v] = c < d ? a.pre_doc[(int64_t)b * d + c] : 0.f;
      g[v] = (c < d && a.d_doc != nullptr) ? a.d_doc[(int64_t)b * d + c] : 0.f;
      if (a.gamma != nullptr && a.thresh != 0 && c < d)
        g[v] = drop_keep(a.k2, a.k3, (uint64_t)b * d + c, a.thresh) ? g[v] * a.drop_scale : 0.f;
    }
    if (a.gamma != nullptr) ln_bwd_row(x, g, a.stats[((int64_t)b * (S + 1) + S) * 2], a.stats[((int64_t)b * (S + 1) + S) * 2 + 1]);
#pragma unroll
    for (int v = 0; v < VPL; ++v) ddoc[lane + 32 * v] = g[v];
  }
  __syncthreads();
  const float inv_n = 1.f / (float)cnt_s;
  for (int s = warp; s < S; s += 8) {
    const int64_t row = a.src[(int64_t)b * S + s];
    if (row < 0) {
      // an empty slot: its LayerNorm input is the zero row — only beta (and nothing of h) sees its gradient
      if (a.gamma != nullptr && a.d_sent != nullptr) {
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
          const int c = lane + 32 * v;
          if (c >= d) continue;
          float g = a.d_sent[((int64_t)b * S + s) * d + c];
          if (a.thresh != 0) g = drop_keep(a.k0, a.k1, ((uint64_t)b * S + s) * d + c, a.thresh) ? g * a.drop_scale : 0.f;
          ab[v] += g;                       // xhat = 0 for a constant row: no gamma gradient
        }
      }
      continue;
    }
    float x[VPL], g[VPL];
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
      const int c = lane + 32 * v;
      x[v] = c < d ? load_as_f(a.h, row * a.ld_h + c, a.h_dtype) : 0.f;
      g[v] = (c < d && a.d_sent != nullptr) ? a.d_sent[((int64_t)b * S + s) * d + c] : 0.f;
      if (a.gamma != nullptr && a.thresh != 0 && c < d)
        g[v] = drop_keep(a.k0, a.k1, ((uint64_t)b * S + s) * d + c, a.thresh) ? g[v] * a.drop_scale : 0.f;
    }
    if (a.gamma != nullptr) ln_bwd_row(x, g, a.stats[((int64_t)b * (S + 1) + s) * 2], a.stats[((int64_t)b * (S + 1) + s) * 2 + 1]);
    const bool pad = a.mask[(int64_t)b * S + s] != 0;
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
      const int c = lane + 32 * v;
      if (c >= d) continue;
      // every slot's x enters the doc SUM (padding rows are zero rows but still summands); the divisor is the non-pad count
      const float t = g[v] + ddoc[c] * inv_n;
      (void)pad;
      atomicAdd(a.dh + row * (int64_t)d + c, t);      // clamped positions can repeat a row (model.py:293-295)
    }
  }
  if (a.gamma != nullptr && a.dgamma != nullptr) {
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
#pragma unroll
      for (int v = 0; v < VPL; ++v) part[warp][lane + 32 * v] = pass == 0 ? ag[v] : ab[v];
      __syncthreads();
      for (int c = threadIdx.x; c < d; c += 256) {
        float sum = 0.f;
        for (int w = 0; w < 8; ++w) sum += part[w][c];
        atomicAdd((pass == 0 ? a.dgamma : a.dbeta) + c, sum);
      }
      __syncthreads();
    }
  }
}

}  // namespace mmoe

using namespace mmoe;

// patch bytes [rows, k] uint8 (rows = B * patches per image, k = C*p*p, k % 16 == 0) -> T [rows, k] holding the byte values.
extern "C" int mmoe_patch_u8_to_operand(const uint8_t* patches, void* out, int64_t rows, int32_t k, int dtype, void* stream) {
  MMOE_CHECK(rows >= 0 && k > 0 && k % 16 == 0, "patch_u8_to_operand: k must be a multiple of 16");
  MMOE_CHECK((reinterpret_cast<uintptr_t>(patches) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0, "patch_u8_to_operand: 16-byte alignment");
  if (rows == 0) return 0;
  const int64_t n16 = rows * k / 16;
  int64_t blocks = (n16 + 255) / 256;
  if (blocks > (int64_t)sm_count() * 8) blocks = (int64_t)sm_count() * 8;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == MMOE_BF16) patch_u8_to_t_kernel<__nv_bfloat16><<<(int)blocks, 256, 0, st>>>(patches, (__nv_bfloat16*)out, n16);
  else if (dtype == MMOE_F16) patch_u8_to_t_kernel<__half><<<(int)blocks, 256, 0, st>>>(patches, (__half*)out, n16);
  else patch_u8_to_t_kernel<float><<<(int)blocks, 256, 0, st>>>(patches, (float*)out, n16);
  MMOE_LAUNCH_OK("patch_u8_to_t_kernel");
  return 0;
}
// float image [B, C, H, W] -> T [B * (H/p) * (W/p), C*p*p]  (rows in the HF ViTPatchEmbeddings order: row-major over the patch grid)
extern "C" int mmoe_patchify(const float* images, void* out, int32_t B, int32_t C, int32_t H, int32_t W, int32_t p, int dtype, void* stream) {
  MMOE_CHECK(B >= 0 && C >= 1 && p >= 4 && p % 4 == 0 && H % p == 0 && W % p == 0, "patchify: bad geometry %dx%dx%d patch %d", C, H, W, p);
  MMOE_CHECK((reinterpret_cast<uintptr_t>(images) & 15) == 0 && W % 4 == 0, "patchify: 16-byte alignment");
  if (B == 0) return 0;
  const int64_t total = (int64_t)B * C * H * W / 4;
  int64_t blocks = (total + 255) / 256;
  if (blocks > (int64_t)sm_count() * 8) blocks = (int64_t)sm_count() * 8;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == MMOE_BF16) patchify_kernel<__nv_bfloat16><<<(int)blocks, 256, 0, st>>>(images, (__nv_bfloat16*)out, B, C, H, W, p);
  else if (dtype == MMOE_F16) patchify_kernel<__half><<<(int)blocks, 256, 0, st>>>(images, (__half*)out, B, C, H, W, p);
  else patchify_kernel<float><<<(int)blocks, 256, 0, st>>>(images, (float*)out, B, C, H, W, p);
  MMOE_LAUNCH_OK("patchify_kernel");
  return 0;
}
// out[rows, N] = operand[rows, K] W[N, K]^T + bias   on the GEMM engine (tcgen05 in 16-bit modes); out_dtype F32 or T
extern "C" int mmoe_patch_project(const void* operand, const void* weight, const float* bias, void* out, int out_dtype, int64_t rows,
                                  int32_t N, int32_t K, int dtype, void* stream) {
  MMOE_CHECK(rows >= 0 && rows < (1ll << 31), "patch_project: rows out of range");
  if (rows == 0) return 0;
  mmoe_epilogue e = epi_none();
  e.out = out; e.out_dtype = out_dtype; e.ldo = N; e.bias = bias;
  mmoe_gemm_problem p = gemm_problem(operand, K, 0, weight, K, 0, (int)rows, N, K, e);
  return gemm_grouped(&p, 1, dtype, 0, (cudaStream_t)stream);
}

static int gather_fill(GatherDev& a, const mmoe_call* c, int S, int d, const void* h, int h_dtype, const int32_t* src) {
  MMOE_CHECK(c->B >= 0 && S >= 1 && d >= 32 && d <= 1024 && d % 32 == 0, "sent_gather: d must be a multiple of 32 in [32,1024], S >= 1");
  a.h = h; a.h_dtype = h_dtype; a.ld_h = d; a.src = src; a.B = c->B; a.S = S; a.d = d;
  a.gamma = c->params != nullptr ? (const float*)c->params[0] : nullptr;
  a.beta = c->params != nullptr ? (const float*)c->params[1] : nullptr;
  MMOE_CHECK((a.gamma == nullptr) == (a.beta == nullptr), "sent_gather: norm.weight and norm.bias go together");
  const float p = (c->training && a.gamma != nullptr) ? c->drop_p : 0.f;
  a.thresh = p > 0.f ? drop_threshold(p) : 0u; a.drop_scale = p > 0.f ? 1.f / (1.f - p) : 1.f;
  site_keys(c->seed, 0, &a.k0, &a.k1);
  site_keys(c->seed, 1, &a.k2, &a.k3);
  return 0;
}
// TextExpert.forward after the encoder — model.py:286-338 (HoME: model_HoME.py:328-369, params = NULL: no final LayerNorm /
// dropout).  h: encoder hidden states viewed as rows of d values (fp32 or T); src int32 [B, S]: the row of h that feeds sentence
// slot (b, s) — chunk * seq_len + clamp(sent_pos) for the slot's chunk/position, -1 where the slot is empty (built on the
// host from chunk2sample / sent_pos, see ingest.py).  params = {norm.weight, norm.bias} or NULL.  Dropout sites: 0 =
// sentence rows [B,S,d], 1 = doc vectors [B,d].  Outputs: sent fp32 [B,S,d], mask uint8 [B,S] (1 = padding), doc fp32 [B,d];
// saved: pre_doc fp32 [B,d], stats fp32 [B*(S+1), 2].
extern "C" int mmoe_sent_gather_fwd(const mmoe_call* c, int32_t S, int32_t d, const void* h, int h_dtype, const int32_t* src,
                                    float* sent, uint8_t* mask, float* doc, float* pre_doc, float* stats) {
  GatherDev a{};
  MMOE_TRY(gather_fill(a, c, S, d, h, h_dtype, src));
  if (c->B == 0) return 0;
  a.sent = sent; a.mask = mask; a.doc = doc; a.pre_doc = pre_doc; a.stats = stats;
  cudaStream_t st = (cudaStream_t)c->stream;
  const int vpl = (d + 31) / 32;
  if (vpl <= 8) sent_gather_fwd_kernel<8><<<c->B, 256, 0, st>>>(a);
  else if (vpl <= 24) sent_gather_fwd_kernel<24><<<c->B, 256, 0, st>>>(a);
  else sent_gather_fwd_kernel<32><<<c->B, 256, 0, st>>>(a);
  MMOE_LAUNCH_OK("sent_gather_fwd_kernel");
  return 0;
}
// dh fp32 [rows of h, d], ACCUMULATED into (zero it first); grads = {d norm.weight, d norm.bias} accumulated, or NULL.
extern "C" int mmoe_sent_gather_bwd(const mmoe_call* c, int32_t S, int32_t d, const void* h, int h_dtype, const int32_t* src,
                                    const uint8_t* mask, const float* pre_doc, const float* stats, const float* d_sent,
                                    const float* d_doc, float* dh) {
  GatherDev a{};
  MMOE_TRY(gather_fill(a, c, S, d, h, h_dtype, src));
  if (c->B == 0) return 0;
  a.mask = const_cast<uint8_t*>(mask); a.pre_doc = const_cast<float*>(pre_doc); a.stats = const_cast<float*>(stats);
  a.d_sent = d_sent; a.d_doc = d_doc; a.dh = dh;
  a.dgamma = c->grads != nullptr ? (float*)c->grads[0] : nullptr;
  a.dbeta = c->grads != nullptr ? (float*)c->grads[1] : nullptr;
  cudaStream_t st = (cudaStream_t)c->stream;
  const int vpl = (d + 31) / 32;
  if (vpl <= 8) sent_gather_bwd_kernel<8><<<c->B, 256, 0, st>>>(a);
  else if (vpl <= 24) sent_gather_bwd_kernel<24><<<c->B, 256, 0, st>>>(a);
  else sent_gather_bwd_kernel<32><<<c->B, 256, 0, st>>>(a);
  MMOE_LAUNCH_OK("sent_gather_bwd_kernel");
  return 0;
}
