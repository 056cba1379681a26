// Internal C++ interface of the non-GEMM kernels (kernels.cu, attention.cu).
#pragma once
#include "common.cuh"

namespace mmoe {

// y = T(x), n elements
int cast_f32(const float* x, void* y, int64_t n, int dtype, cudaStream_t s);

// g[r,c] = T(dropmask(x[r,c] + x_t[r,c]));  colsum[c] += sum_r g  (either output optional; x fp32 and x_t (T) are both
// optional summands, at least one must be given)
int cast_drop_colsum(const float* x, void* g, float* colsum, int64_t rows, int cols, float drop_p, uint32_t k0,
                     uint32_t k1, int dtype, cudaStream_t s, const void* x_t = nullptr);

int dropout_mask(uint32_t k0, uint32_t k1, float p, int64_t n, uint8_t* out, cudaStream_t s);

// LayerNorm forward: x fp32 or T; outputs T and/or fp32; stats [rows,2]
int layernorm_fwd(const void* x, int x_dtype, const float* gamma, const float* beta, void* y_t, float* y_f32, float* stats,
                  int64_t rows, int d, int dtype, cudaStream_t s);
// Residual add fused into LayerNorm:  x_sum = x (fp32) + delta (T);  x_sum is written (fp32) and normalised into y_t.
int layernorm_fwd_add(const float* x, const void* delta, float* x_sum, const float* gamma, const float* beta, void* y_t,
                      float* stats, int64_t rows, int d, int dtype, cudaStream_t s);
// out_f = a (fp32) + b (T);  out_t = T(out_f)   (either output optional)
int add_cast(const float* a, const void* b, float* out_f, void* out_t, int64_t n, int dtype, cudaStream_t s);
// ReLU(+dropout) backward on a stored gradient, in place:  g = (h != 0) ? g * scale : 0;  colsum += column sums of g
int relu_mask_colsum(void* g, const void* h, float* colsum, int64_t rows, int cols, float scale, int dtype, cudaStream_t s);

// LayerNorm backward fused with the residual add and the next dropout/cast:
//   dx[r,:]  = (dres ? dres[r,:] : 0) + LN'(dy[r,:])          (fp32 out, may alias dres)
//   dgamma  += sum_r dy*xhat ; dbeta += sum_r dy
//   g_out    = T(dropmask(dx))  and  g_colsum += colsum(g_out)   (both optional)
struct LnBwdArgs {
  const void* dy; int dy_dtype;          // [rows,d]
  const void* x; int x_dtype;            // LN input
  const float* stats; const float* gamma;
  const float* dres;                     // optional
  const void* dres_t;                    // optional second residual gradient, operand dtype T (added to dres)
  float* dx;                             // optional
  float* dgamma; float* dbeta;           // optional (both or none)
  void* g_out; float* g_colsum;          // optional
  float drop_p; uint32_t k0, k1;
  int64_t rows; int d; int dtype;
};
int layernorm_bwd(const LnBwdArgs& a, cudaStream_t s);

// attention core (see mmoe_attention_fwd/bwd in the header)
struct AttnArgs {
  const void *q, *k, *v; int64_t ldq, ldk, ldv;
  const uint8_t* mask;                   // [B,Sk] or null
  void* ctx; int64_t ldc;                // fwd out / bwd: dctx in
  void *dq, *dk, *dv;                    // bwd outs (strides ldq/ldk/ldv)
  float *bgq, *bgk, *bgv;                // bwd: bias-grad column sums (optional)
  int B, Sq, Sk, H, hd;
  float drop_p; uint32_t k0, k1;
  int dtype;
};
// tensor-core (mma.sync) variant for the 16-bit modes, attention_tc.cu
bool attention_tc_supported(const AttnArgs& a);
int attention_tc(const AttnArgs& a, bool bwd, cudaStream_t s);
int attention_fwd(const AttnArgs& a, cudaStream_t s);
int attention_bwd(const AttnArgs& a, cudaStream_t s);

}  // namespace mmoe
