// HOME_MMoE_Complete — model_HoME.py:530-638.
//
// forward : cast -> [input_projection GEMM] -> fused [LN + GELU + mean residual] -> {3 FeatureGate GEMMs} (grouped)
//           -> fused [x * 2 sigmoid(.)] -> {8 expert GEMMs +GELU +dropout} (ONE grouped launch) -> {8 expert GEMMs}
//           (one launch) -> {3 SelfGate GEMMs +sigmoid} (grouped) -> fused [shared + s*y] -> gate/mix kernel ->
//           LN x2 -> {2 tower GEMMs +GELU +dropout} -> final 512->1 layer.
// The SelfGate of a group is evaluated once (the reference recomputes it per expert).
#include "head_kernels.cuh"

namespace mmoe {

constexpr int HOME_MAXE = 8;

// shared = gelu(LN(z0)) + mean_n ev                                (model_HoME.py:597-602)   one warp per row
template <typename T>
__global__ void __launch_bounds__(256) home_shared_fwd_kernel(const float* __restrict__ z0, const float* __restrict__ gamma,
                                                              const float* __restrict__ beta, const float* __restrict__ ev,
                                                              float* __restrict__ stats, float* __restrict__ shared_f,
                                                              T* __restrict__ shared_t, int64_t B, int d, int n_in) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const float inv_d = 1.f / (float)d;
  for (int64_t b = warp0; b < B; b += nw) {
    float sum = 0.f;
    for (int c = lane; c < d; c += 32) sum += z0[b * d + c];
    const float mean = warp_sum(sum) * inv_d;
    float sq = 0.f;
    for (int c = lane; c < d; c += 32) { const float t = z0[b * d + c] - mean; sq += t * t; }
    const float rstd = rsqrtf(warp_sum(sq) * inv_d + 1e-5f);
    if (lane == 0) { stats[b * 2] = mean; stats[b * 2 + 1] = rstd; }
    for (int c = lane; c < d; c += 32) {
      const float ln = (z0[b * d + c] - mean) * rstd * gamma[c] + beta[c];
      float m = 0.f;
      for (int k = 0; k < n_in; ++k) m += ev[(b * n_in + k) * d + c];
      const float s = gelu_f(ln) + m / (float)n_in;
      shared_f[b * d + c] = s;
      shared_t[b * d + c] = from_f<T>(s);
    }
  }
}

// dln = dshared * gelu'(LN(z0));  d_ev[b,n,:] = dshared / n_in
__global__ void __launch_bounds__(256) home_shared_bwd_kernel(const float* __restrict__ dshared, const float* __restrict__ z0,
                                                              const float* __restrict__ stats, const float* __restrict__ gamma,
                                                              const float* __restrict__ beta, float* __restrict__ dln,
                                                              float* __restrict__ d_ev, int64_t B, int d, int n_in) {
  const int64_t n = B * d;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / d; const int c = (int)(i - b * d);
    const float ln = (z0[i] - stats[b * 2]) * stats[b * 2 + 1] * gamma[c] + beta[c];
    const float ds = dshared[i];
    dln[i] = ds * gelu_grad_f(ln);
    const float share = ds / (float)n_in;
    for (int k = 0; k < n_in; ++k) d_ev[(b * n_in + k) * d + c] = share;
  }
}

// xin[b, e*d + j] = shared[b,j] * 2 sigmoid(z[b, e*d + j])        (FeatureGate, model_HoME.py:232-234)
template <typename T>
__global__ void fg_apply_kernel(const T* __restrict__ z, const float* __restrict__ shared_f, T* __restrict__ xin, int64_t B, int d, int ne) {
  const int64_t n = B * ne * d;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / ((int64_t)ne * d); const int j = (int)(i % d);
    xin[i] = from_f<T>(shared_f[b * d + j] * 2.f * sigmoid_f(to_f<T>(z[i])));
  }
}
// dz = dxin * shared * 2 s (1-s)  (+colsum -> d bias);  dshared[b,j] += sum_e dxin * 2 s
// block = 32-row panel, thread owns columns j (all experts)
template <typename T>
__global__ void __launch_bounds__(256) fg_apply_bwd_kernel(const T* __restrict__ dxin, const T* __restrict__ z, const float* __restrict__ shared_f,
                                                           T* __restrict__ dz, float* const* __restrict__ dbias /*[ne] host-built device table*/,
                                                           float* __restrict__ dshared, int64_t B, int d, int ne,
                                                           float* db0, float* db1, float* db2, int n0, int n1) {
  const int64_t r0 = (int64_t)blockIdx.x * 32, r1 = min(B, r0 + 32);
  for (int j = threadIdx.x; j < d; j += blockDim.x) {
    float cs[HOME_MAXE];
#pragma unroll
    for (int e = 0; e < HOME_MAXE; ++e) cs[e] = 0.f;
    for (int64_t b = r0; b < r1; ++b) {
      const float sh = shared_f[b * d + j];
      float acc = 0.f;
#pragma unroll
      for (int e = 0; e < HOME_MAXE; ++e) {
        if (e < ne) {
          const int64_t i = (b * ne + e) * d + j;
          const float s = sigmoid_f(to_f<T>(z[i]));
          const float g = to_f<T>(dxin[i]);
          acc += g * 2.f * s;
          const T o = from_f<T>(g * sh * 2.f * s * (1.f - s));
          dz[i] = o;
          cs[e] += to_f<T>(o);
        }
      }
      dshared[b * d + j] += acc;
    }
#pragma unroll
    for (int e = 0; e < HOME_MAXE; ++e) {
      if (e < ne) {
        // bias layout follows the three FeatureGate modules: meta [n0*d], good [n1*d], best [n1*d]
        float* dst = e < n0 ? db0 + (int64_t)e * d : (e < n0 + n1 ? db1 + (int64_t)(e - n0) * d : db2 + (int64_t)(e - n0 - n1) * d);
        atomicAdd(dst + j, cs[e]);
      }
    }
  }
}

// enh[e,b,j] = shared[b,j] + sg[g(e),b,j] * y[e,b,j]              (SelfGate, model_HoME.py:242-243)
template <typename T>
__global__ void sg_apply_kernel(const T* __restrict__ y, const T* __restrict__ sg, const float* __restrict__ shared_f,
                                float* __restrict__ enh, int64_t B, int d, int ne, int n0, int n1) {
  const int64_t n = (int64_t)ne * B * d;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int e = (int)(i / (B * d)); const int64_t r = i - (int64_t)e * B * d;
    const int g = e < n0 ? 0 : (e < n0 + n1 ? 1 : 2);
    enh[i] = shared_f[r] + to_f<T>(sg[(int64_t)g * B * d + r]) * to_f<T>(y[i]);
  }
}
// dshared[b,j] += sum_e denh;  dy[e] = T(denh * s) (+colsum -> d b2_e);  dzs[g] = T(sum_{e in g} denh*y * s(1-s)) (+colsum -> d bsg_g)
template <typename T>
__global__ void __launch_bounds__(256) sg_apply_bwd_kernel(const float* __restrict__ denh, const T* __restrict__ y, const T* __restrict__ sg,
                                                           T* __restrict__ dy, T* __restrict__ dzs, float* __restrict__ dshared,
                                                           float* const* __restrict__ unused, int64_t B, int d, int ne, int n0, int n1,
                                                           float** db2_tbl /*device [ne]*/, float* dbs0, float* dbs1, float* dbs2) {
  const int64_t r0 = (int64_t)blockIdx.x * 32, r1 = min(B, r0 + 32);
  for (int j = threadIdx.x; j < d; j += blockDim.x) {
    float cs_y[HOME_MAXE], cs_s[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int e = 0; e < HOME_MAXE; ++e) cs_y[e] = 0.f;
    for (int64_t b = r0; b < r1; ++b) {
      const int64_t r = b * d + j;
      float s[3], dsg[3] = {0.f, 0.f, 0.f};
#pragma unroll
      for (int g = 0; g < 3; ++g) s[g] = to_f<T>(sg[(int64_t)g * B * d + r]);
      float acc = 0.f;
#pragma unroll
      for (int e = 0; e < HOME_MAXE; ++e) {
        if (e < ne) {
          const int g = e < n0 ? 0 : (e < n0 + n1 ? 1 : 2);
          const int64_t i = (int64_t)e * B * d + r;
          const float de = denh[i];
          acc += de;
          const T o = from_f<T>(de * s[g]);
          dy[i] = o;
          cs_y[e] += to_f<T>(o);
          dsg[g] += de * to_f<T>(y[i]);
        }
      }
      dshared[r] += acc;
#pragma unroll
      for (int g = 0; g < 3; ++g) {
        const T o = from_f<T>(dsg[g] * s[g] * (1.f - s[g]));
        dzs[(int64_t)g * B * d + r] = o;
        cs_s[g] += to_f<T>(o);
      }
    }
#pragma unroll
    for (int e = 0; e < HOME_MAXE; ++e)
      if (e < ne) atomicAdd(db2_tbl[e] + j, cs_y[e]);
    atomicAdd(dbs0 + j, cs_s[0]); atomicAdd(dbs1 + j, cs_s[1]); atomicAdd(dbs2 + j, cs_s[2]);
  }
}

struct HomeIdx {
  int p_w, p_b, pln_w, pln_b, exp0, fg0, sg0, gate0, tower0, count, ne;
  int e_w1(int e) const { return exp0 + 4 * e; }
  int e_b1(int e) const { return exp0 + 4 * e + 1; }
  int e_w2(int e) const { return exp0 + 4 * e + 2; }
  int e_b2(int e) const { return exp0 + 4 * e + 3; }
  int fg_w(int g) const { return fg0 + 2 * g; }
  int fg_b(int g) const { return fg0 + 2 * g + 1; }
  int sg_w(int g) const { return sg0 + 2 * g; }
  int sg_b(int g) const { return sg0 + 2 * g + 1; }
  int gate_w(int t) const { return gate0 + 2 * t; }
  int gate_b(int t) const { return gate0 + 2 * t + 1; }
  int t_lnw(int t) const { return tower0 + 6 * t; }
  int t_lnb(int t) const { return tower0 + 6 * t + 1; }
  int t_w1(int t) const { return tower0 + 6 * t + 2; }
  int t_b1(int t) const { return tower0 + 6 * t + 3; }
  int t_w2(int t) const { return tower0 + 6 * t + 4; }
  int t_b2(int t) const { return tower0 + 6 * t + 5; }
};
static HomeIdx home_idx(const mmoe_home_cfg& cfg) {
  HomeIdx i;
  i.ne = cfg.n_shared + 2 * cfg.n_task;
  i.p_w = 0; i.p_b = 1; i.pln_w = 2; i.pln_b = 3; i.exp0 = 4;
  i.fg0 = i.exp0 + 4 * i.ne; i.sg0 = i.fg0 + 6; i.gate0 = i.sg0 + 6; i.tower0 = i.gate0 + 4; i.count = i.tower0 + 12;
  return i;
}

struct HomeSaved {
  void* ct; float* z0; float* st0; float* shared_f; void* shared_t; void* zfg; void* xin; void* z1; void* h; void* y; void* sg;
  float* enh; float* fused; float* w; float* st_t; void* xn; void* zt; void* at;
};
static HomeSaved home_layout(Arena& A, const mmoe_home_cfg& cfg, int B, int dtype) {
  HomeSaved s{};
  const size_t es = dtype_size(dtype), Bz = (size_t)B; const int d = cfg.d, ne = cfg.n_shared + 2 * cfg.n_task;
  const int n = cfg.n_shared + cfg.n_task;
  s.ct = A.take(Bz * cfg.n_in * d * es); s.z0 = (float*)A.take(Bz * d * 4); s.st0 = (float*)A.take(Bz * 2 * 4);
  s.shared_f = (float*)A.take(Bz * d * 4); s.shared_t = A.take(Bz * d * es);
  s.zfg = A.take(Bz * ne * d * es); s.xin = A.take(Bz * ne * d * es);
  s.z1 = A.take(Bz * ne * cfg.expert_hidden * es); s.h = A.take(Bz * ne * cfg.expert_hidden * es);
  s.y = A.take(Bz * ne * d * es); s.sg = A.take(Bz * 3 * d * es);
  s.enh = (float*)A.take(Bz * ne * d * 4); s.fused = (float*)A.take(Bz * 2 * d * 4); s.w = (float*)A.take(Bz * 2 * n * 4);
  s.st_t = (float*)A.take(Bz * 2 * 2 * 4); s.xn = A.take(Bz * 2 * d * es);
  s.zt = A.take(Bz * 2 * cfg.tower_hidden * es); s.at = A.take(Bz * 2 * cfg.tower_hidden * es);
  return s;
}
struct HomeScratch {
  void* dzt; void* dxn; float* dfused; float* dl; float* denh; float* dshared; void* dy; void* dzs; void* dh; void* dxin; void* dzfg;
  float* dln; void* g0; float** tbl;
};
static HomeScratch home_scratch_layout(Arena& A, const mmoe_home_cfg& cfg, int B, int dtype) {
  HomeScratch t{};
  const size_t es = dtype_size(dtype), Bz = (size_t)B; const int d = cfg.d, ne = cfg.n_shared + 2 * cfg.n_task;
  const int n = cfg.n_shared + cfg.n_task;
  t.dzt = A.take(Bz * 2 * cfg.tower_hidden * es); t.dxn = A.take(Bz * 2 * d * es); t.dfused = (float*)A.take(Bz * 2 * d * 4);
  t.dl = (float*)A.take(Bz * 2 * n * 4); t.denh = (float*)A.take(Bz * ne * d * 4); t.dshared = (float*)A.take(Bz * d * 4);
  t.dy = A.take(Bz * ne * d * es); t.dzs = A.take(Bz * 3 * d * es); t.dh = A.take(Bz * ne * cfg.expert_hidden * es);
  t.dxin = A.take(Bz * ne * d * es); t.dzfg = A.take(Bz * ne * d * es); t.dln = (float*)A.take(Bz * d * 4); t.g0 = A.take(Bz * d * es);
  t.tbl = (float**)A.take(sizeof(float*) * 16);
  return t;
}
static int check_home(const mmoe_home_cfg* cfg) {
  const int ne = cfg->n_shared + 2 * cfg->n_task;
  MMOE_CHECK(cfg->d % 8 == 0 && cfg->d <= 1024, "HoME head: unsupported expert_dim %d", cfg->d);
  MMOE_CHECK(ne >= 1 && ne <= HOME_MAXE && cfg->n_shared >= 0 && cfg->n_task >= 0, "HoME head: at most %d experts in total", HOME_MAXE);
  MMOE_CHECK(cfg->n_shared + cfg->n_task <= MIX_MAXN && cfg->n_shared + cfg->n_task >= 1, "HoME head: too many experts per gate");
  MMOE_CHECK(cfg->tower_hidden % 16 == 0 && cfg->tower_hidden <= 512, "HoME head: tower_hidden must be a multiple of 16, <= 512");
  MMOE_CHECK(cfg->expert_hidden % 16 == 0, "HoME head: expert_hidden must be a multiple of 16");
  MMOE_CHECK(cfg->n_in >= 1 && cfg->n_in <= 16, "HoME head: unsupported num_input_experts");
  return 0;
}
static inline int ew_grid(int64_t n) {
  int64_t b = (n + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 16;
  if (b > cap) b = cap;
  return b < 1 ? 1 : (int)b;
}

template <typename T>
static int home_fwd_t(const mmoe_call* c, const mmoe_home_cfg* cfg, const float* ev, float* logits, float* gate_w) {
  const int B = c->B, d = cfg->d, dtype = c->dtype, n0 = cfg->n_shared, n1 = cfg->n_task, ne = n0 + 2 * n1, n = n0 + n1;
  const int eh = cfg->expert_hidden, th = cfg->tower_hidden, nin = cfg->n_in;
  const size_t es = dtype_size(dtype);
  cudaStream_t st = (cudaStream_t)c->stream;
  Arena A(c->saved);
  HomeSaved s = home_layout(A, *cfg, B, dtype);
  const HomeIdx ix = home_idx(*cfg);
  const void* const* P = c->params;
  const float drop_p = c->training ? c->drop_p : 0.f;
  uint32_t k0, k1;
  MMOE_TRY(cast_f32(ev, s.ct, (int64_t)B * nin * d, dtype, st));
  {
    mmoe_epilogue e = epi_none();
    e.out = s.z0; e.out_dtype = MMOE_F32; e.ldo = d; e.bias = (const float*)P[ix.p_b];
    mmoe_gemm_problem p = linear_fwd(s.ct, (int64_t)nin * d, P[ix.p_w], B, d, nin * d, e);
    MMOE_TRY(gemm_grouped(&p, 1, dtype, 0, st));
  }
  home_shared_fwd_kernel<T><<<rows_grid(B, 8), 256, 0, st>>>(s.z0, (const float*)P[ix.pln_w], (const float*)P[ix.pln_b], ev, s.st0,
                                                             s.shared_f, (T*)s.shared_t, B, d, nin);
  MMOE_LAUNCH_OK("home_shared_fwd_kernel");
  {
    // FeatureGate pre-activations for the three groups, written side by side into zfg [B, ne*d]
    mmoe_gemm_problem p[3];
    const int col0[3] = {0, n0 * d, (n0 + n1) * d}, width[3] = {n0 * d, n1 * d, n1 * d};
    int np = 0;
    for (int g = 0; g < 3; ++g) {
      if (width[g] == 0) continue;
      mmoe_epilogue e = epi_none();
      e.out = (char*)s.zfg + (size_t)col0[g] * es; e.out_dtype = dtype; e.ldo = (int64_t)ne * d; e.bias = (const float*)P[ix.fg_b(g)];
      p[np++] = linear_fwd(s.shared_t, d, P[ix.fg_w(g)], B, width[g], d, e);
    }
    MMOE_TRY(gemm_grouped(p, np, dtype, 0, st));
  }
  fg_apply_kernel<T><<<ew_grid((int64_t)B * ne * d), 256, 0, st>>>((const T*)s.zfg, s.shared_f, (T*)s.xin, B, d, ne);
  MMOE_LAUNCH_OK("fg_apply_kernel");
  {
    mmoe_gemm_problem p[HOME_MAXE];
    for (int e2 = 0; e2 < ne; ++e2) {
      mmoe_epilogue e = epi_none();
      e.out = (char*)s.h + (size_t)e2 * B * eh * es; e.preact = (char*)s.z1 + (size_t)e2 * B * eh * es;
      e.out_dtype = dtype; e.ldo = eh; e.bias = (const float*)P[ix.e_b1(e2)]; e.act = 2;
      site_keys(c->seed, 10 + e2, &k0, &k1);
      e.drop_p = drop_p; e.drop_key0 = k0; e.drop_key1 = k1;
      p[e2] = linear_fwd((char*)s.xin + (size_t)e2 * d * es, (int64_t)ne * d, P[ix.e_w1(e2)], B, eh, d, e);
    }
    MMOE_TRY(gemm_grouped(p, ne, dtype, 0, st));
    for (int e2 = 0; e2 < ne; ++e2) {
      mmoe_epilogue e = epi_none();
      e.out = (char*)s.y + (size_t)e2 * B * d * es; e.out_dtype = dtype; e.ldo = d; e.bias = (const float*)P[ix.e_b2(e2)];
      p[e2] = linear_fwd((char*)s.h + (size_t)e2 * B * eh * es, eh, P[ix.e_w2(e2)], B, d, eh, e);
    }
    MMOE_TRY(gemm_grouped(p, ne, dtype, 0, st));
  }
  {
    mmoe_gemm_problem p[3];
    for (int g = 0; g < 3; ++g) {
      mmoe_epilogue e = epi_none();
      e.out = (char*)s.sg + (size_t)g * B * d * es; e.out_dtype = dtype; e.ldo = d; e.bias = (const float*)P[ix.sg_b(g)]; e.act = 3;
      p[g] = linear_fwd(s.shared_t, d, P[ix.sg_w(g)], B, d, d, e);
    }
    MMOE_TRY(gemm_grouped(p, 3, dtype, 0, st));
  }
  sg_apply_kernel<T><<<ew_grid((int64_t)ne * B * d), 256, 0, st>>>((const T*)s.y, (const T*)s.sg, s.shared_f, s.enh, B, d, ne, n0, n1);
  MMOE_LAUNCH_OK("sg_apply_kernel");
  {
    MixDev a{};
    a.experts = s.enh; a.expert_stride = (int64_t)B * d; a.row_stride = d;
    for (int k = 0; k < n; ++k) { a.sel[0][k] = k; a.sel[1][k] = k < n0 ? k : k + n1; }
    a.query_in = s.shared_f;
    a.wg[0] = (const float*)P[ix.gate_w(0)]; a.wg[1] = (const float*)P[ix.gate_w(1)];
    a.bg[0] = (const float*)P[ix.gate_b(0)]; a.bg[1] = (const float*)P[ix.gate_b(1)];
    a.fused = s.fused; a.w = s.w; a.B = B; a.d = d; a.n = n;
    mix_fwd_kernel<<<rows_grid(B, 8), 256, 0, st>>>(a);
    MMOE_LAUNCH_OK("mix_fwd_kernel");
    if (gate_w != nullptr) MMOE_CUDA(cudaMemcpyAsync(gate_w, s.w, (size_t)2 * B * n * 4, cudaMemcpyDeviceToDevice, st));
  }
  for (int t = 0; t < 2; ++t)
    MMOE_TRY(layernorm_fwd(s.fused + (size_t)t * B * d, MMOE_F32, (const float*)P[ix.t_lnw(t)], (const float*)P[ix.t_lnb(t)],
                           (char*)s.xn + (size_t)t * B * d * es, nullptr, s.st_t + (size_t)t * B * 2, B, d, dtype, st));
  {
    mmoe_gemm_problem p[2];
    for (int t = 0; t < 2; ++t) {
      mmoe_epilogue e = epi_none();
      e.out = (char*)s.at + (size_t)t * B * th * es; e.preact = (char*)s.zt + (size_t)t * B * th * es;
      e.out_dtype = dtype; e.ldo = th; e.bias = (const float*)P[ix.t_b1(t)]; e.act = 2;
      site_keys(c->seed, 30 + t, &k0, &k1);
      e.drop_p = drop_p; e.drop_key0 = k0; e.drop_key1 = k1;
      p[t] = linear_fwd((char*)s.xn + (size_t)t * B * d * es, d, P[ix.t_w1(t)], B, th, d, e);
    }
    MMOE_TRY(gemm_grouped(p, 2, dtype, 0, st));
  }
  for (int t = 0; t < 2; ++t) {
    gemv_fwd_kernel<T><<<rows_grid(B, 8), 256, 0, st>>>((const T*)s.at + (size_t)t * B * th, (const float*)P[ix.t_w2(t)],
                                                        (const float*)P[ix.t_b2(t)], logits + (size_t)t * B, B, th);
    MMOE_LAUNCH_OK("gemv_fwd_kernel");
  }
  return 0;
}

template <typename T>
static int home_bwd_t(const mmoe_call* c, const mmoe_home_cfg* cfg, const float* ev, const float* dlogits, float* d_ev) {
  const int B = c->B, d = cfg->d, dtype = c->dtype, n0 = cfg->n_shared, n1 = cfg->n_task, ne = n0 + 2 * n1, n = n0 + n1;
  const int eh = cfg->expert_hidden, th = cfg->tower_hidden, nin = cfg->n_in;
  const size_t es = dtype_size(dtype);
  cudaStream_t st = (cudaStream_t)c->stream;
  Arena A(c->saved);
  HomeSaved s = home_layout(A, *cfg, B, dtype);
  Arena W(c->workspace);
  HomeScratch t = home_scratch_layout(W, *cfg, B, dtype);
  const HomeIdx ix = home_idx(*cfg);
  const void* const* P = c->params;
  void* const* G = c->grads;
  const float drop_p = c->training ? c->drop_p : 0.f;
  const uint32_t thresh = drop_p > 0.f ? drop_threshold(drop_p) : 0u;
  const float scale = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  uint32_t k0, k1;
  // towers
  for (int k = 0; k < 2; ++k) {
    site_keys(c->seed, 30 + k, &k0, &k1);
    gemv_bwd_kernel<T><<<rows_grid(B, 4), 256, 0, st>>>(dlogits + (size_t)k * B, (const T*)s.at + (size_t)k * B * th,
                                                        (const T*)s.zt + (size_t)k * B * th, (const float*)P[ix.t_w2(k)],
                                                        (T*)t.dzt + (size_t)k * B * th, (float*)G[ix.t_w2(k)], (float*)G[ix.t_b2(k)],
                                                        (float*)G[ix.t_b1(k)], B, th, thresh, scale, k0, k1);
    MMOE_LAUNCH_OK("gemv_bwd_kernel");
  }
  {
    mmoe_gemm_problem p[4];
    for (int k = 0; k < 2; ++k) {
      mmoe_epilogue e = epi_none();
      e.out = (char*)t.dxn + (size_t)k * B * d * es; e.out_dtype = dtype; e.ldo = d;
      const char* dz = (const char*)t.dzt + (size_t)k * B * th * es;
      p[2 * k] = linear_dgrad(dz, th, P[ix.t_w1(k)], B, th, d, e);
      p[2 * k + 1] = linear_wgrad(dz, th, (char*)s.xn + (size_t)k * B * d * es, d, (float*)G[ix.t_w1(k)], B, th, d);
    }
    MMOE_TRY(gemm_grouped(p, 4, dtype, 0, st));
  }
  for (int k = 0; k < 2; ++k) {
    LnBwdArgs a{};
    a.dy = (char*)t.dxn + (size_t)k * B * d * es; a.dy_dtype = dtype; a.x = s.fused + (size_t)k * B * d; a.x_dtype = MMOE_F32;
    a.stats = s.st_t + (size_t)k * B * 2; a.gamma = (const float*)P[ix.t_lnw(k)]; a.dx = t.dfused + (size_t)k * B * d;
    a.dgamma = (float*)G[ix.t_lnw(k)]; a.dbeta = (float*)G[ix.t_lnb(k)]; a.rows = B; a.d = d; a.dtype = dtype;
    MMOE_TRY(layernorm_bwd(a, st));
  }
  // gates + mix: denh (every expert written exactly once), dshared initialised with the query gradient
  {
    MixDev a{};
    a.experts = s.enh; a.expert_stride = (int64_t)B * d; a.row_stride = d;
    for (int k = 0; k < n; ++k) { a.sel[0][k] = k; a.sel[1][k] = k < n0 ? k : k + n1; }
    a.wg[0] = (const float*)P[ix.gate_w(0)]; a.wg[1] = (const float*)P[ix.gate_w(1)];
    a.w = s.w; a.dfused = t.dfused; a.dl = t.dl; a.dexperts = t.denh; a.accumulate_dexperts = 0; a.dquery = t.dshared;
    a.B = B; a.d = d; a.n = n;
    mix_bwd_kernel<<<rows_grid(B, 8), 256, 0, st>>>(a, ne);
    MMOE_LAUNCH_OK("mix_bwd_kernel");
  }
  for (int k = 0; k < 2; ++k)
    MMOE_TRY(small_wgrad(t.dl + (size_t)k * B * n, s.shared_f, MMOE_F32, (float*)G[ix.gate_w(k)], (float*)G[ix.gate_b(k)], B, d, n, st));
  // SelfGate backward
  {
    float* host_tbl[16];
    for (int e2 = 0; e2 < ne; ++e2) host_tbl[e2] = (float*)G[ix.e_b2(e2)];
    MMOE_CUDA(cudaMemcpyAsync(t.tbl, host_tbl, sizeof(float*) * ne, cudaMemcpyHostToDevice, st));
    sg_apply_bwd_kernel<T><<<(B + 31) / 32, 256, 0, st>>>(t.denh, (const T*)s.y, (const T*)s.sg, (T*)t.dy, (T*)t.dzs, t.dshared, nullptr,
                                                          B, d, ne, n0, n1, t.tbl, (float*)G[ix.sg_b(0)], (float*)G[ix.sg_b(1)],
                                                          (float*)G[ix.sg_b(2)]);
    MMOE_LAUNCH_OK("sg_apply_bwd_kernel");
  }
  {
    mmoe_gemm_problem p[6];
    for (int g = 0; g < 3; ++g) {
      mmoe_epilogue e = epi_none();   // dshared += dzs_g Wsg_g   (three writers -> atomic accumulation)
      e.out = t.dshared; e.out_dtype = MMOE_F32; e.ldo = d; e.accumulate = 1;
      const char* dz = (const char*)t.dzs + (size_t)g * B * d * es;
      p[2 * g] = linear_dgrad(dz, d, P[ix.sg_w(g)], B, d, d, e);
      p[2 * g + 1] = linear_wgrad(dz, d, s.shared_t, d, (float*)G[ix.sg_w(g)], B, d, d);
    }
    MMOE_TRY(gemm_grouped(p, 6, dtype, 0, st));
  }
  // experts
  {
    mmoe_gemm_problem p[HOME_MAXE];
    for (int e2 = 0; e2 < ne; ++e2) {
      mmoe_epilogue e = epi_none();
      e.out = (char*)t.dh + (size_t)e2 * B * eh * es; e.out_dtype = dtype; e.ldo = eh; e.bwd_mode = 2;
      e.aux = (char*)s.z1 + (size_t)e2 * B * eh * es; e.ld_aux = eh; e.colsum = (float*)G[ix.e_b1(e2)];
      site_keys(c->seed, 10 + e2, &k0, &k1);
      e.drop_p = drop_p; e.drop_key0 = k0; e.drop_key1 = k1;
      p[e2] = linear_dgrad((char*)t.dy + (size_t)e2 * B * d * es, d, P[ix.e_w2(e2)], B, d, eh, e);
    }
    MMOE_TRY(gemm_grouped(p, ne, dtype, 0, st));
    for (int e2 = 0; e2 < ne; ++e2)
      p[e2] = linear_wgrad((char*)t.dy + (size_t)e2 * B * d * es, d, (char*)s.h + (size_t)e2 * B * eh * es, eh, (float*)G[ix.e_w2(e2)], B, d, eh);
    MMOE_TRY(gemm_grouped(p, ne, dtype, 0, st));
    for (int e2 = 0; e2 < ne; ++e2) {
      mmoe_epilogue e = epi_none();
      e.out = (char*)t.dxin + (size_t)e2 * d * es; e.out_dtype = dtype; e.ldo = (int64_t)ne * d;
      p[e2] = linear_dgrad((char*)t.dh + (size_t)e2 * B * eh * es, eh, P[ix.e_w1(e2)], B, eh, d, e);
    }
    MMOE_TRY(gemm_grouped(p, ne, dtype, 0, st));
    for (int e2 = 0; e2 < ne; ++e2)
      p[e2] = linear_wgrad((char*)t.dh + (size_t)e2 * B * eh * es, eh, (char*)s.xin + (size_t)e2 * d * es, (int64_t)ne * d,
                           (float*)G[ix.e_w1(e2)], B, eh, d);
    MMOE_TRY(gemm_grouped(p, ne, dtype, 0, st));
  }
  // FeatureGate backward
  fg_apply_bwd_kernel<T><<<(B + 31) / 32, 256, 0, st>>>((const T*)t.dxin, (const T*)s.zfg, s.shared_f, (T*)t.dzfg, nullptr, t.dshared, B, d, ne,
                                                        (float*)G[ix.fg_b(0)], (float*)G[ix.fg_b(1)], (float*)G[ix.fg_b(2)], n0, n1);
  MMOE_LAUNCH_OK("fg_apply_bwd_kernel");
  {
    mmoe_gemm_problem p[6];
    const int col0[3] = {0, n0 * d, (n0 + n1) * d}, width[3] = {n0 * d, n1 * d, n1 * d};
    int np = 0;
    for (int g = 0; g < 3; ++g) {
      if (width[g] == 0) continue;
      mmoe_epilogue e = epi_none();
      e.out = t.dshared; e.out_dtype = MMOE_F32; e.ldo = d; e.accumulate = 1;
      const char* dz = (const char*)t.dzfg + (size_t)col0[g] * es;
      p[np++] = linear_dgrad(dz, (int64_t)ne * d, P[ix.fg_w(g)], B, width[g], d, e);
      p[np++] = linear_wgrad(dz, (int64_t)ne * d, s.shared_t, d, (float*)G[ix.fg_w(g)], B, width[g], d);
    }
    MMOE_TRY(gemm_grouped(p, np, dtype, 0, st));
  }
  // shared = gelu(LN(z0)) + mean(ev);  z0 = ct Wp^T + bp
  home_shared_bwd_kernel<<<ew_grid((int64_t)B * d), 256, 0, st>>>(t.dshared, s.z0, s.st0, (const float*)P[ix.pln_w], (const float*)P[ix.pln_b],
                                                                  t.dln, d_ev, B, d, nin);
  MMOE_LAUNCH_OK("home_shared_bwd_kernel");
  {
    LnBwdArgs a{};
    a.dy = t.dln; a.dy_dtype = MMOE_F32; a.x = s.z0; a.x_dtype = MMOE_F32; a.stats = s.st0; a.gamma = (const float*)P[ix.pln_w];
    a.dgamma = (float*)G[ix.pln_w]; a.dbeta = (float*)G[ix.pln_b]; a.g_out = t.g0; a.g_colsum = (float*)G[ix.p_b];
    a.rows = B; a.d = d; a.dtype = dtype;
    MMOE_TRY(layernorm_bwd(a, st));
  }
  {
    mmoe_epilogue e = epi_none();   // d_ev += g0 Wp
    e.out = d_ev; e.out_dtype = MMOE_F32; e.ldo = (int64_t)nin * d; e.residual = d_ev; e.ld_res = (int64_t)nin * d;
    mmoe_gemm_problem p[2] = {linear_dgrad(t.g0, d, P[ix.p_w], B, d, nin * d, e),
                              linear_wgrad(t.g0, d, s.ct, (int64_t)nin * d, (float*)G[ix.p_w], B, d, nin * d)};
    MMOE_TRY(gemm_grouped(p, 2, dtype, 0, st));
  }
  return 0;
}

}  // namespace mmoe

using namespace mmoe;

extern "C" size_t mmoe_home_saved_bytes(const mmoe_home_cfg* cfg, int32_t B, int dtype) {
  Arena A(nullptr);
  home_layout(A, *cfg, B, dtype);
  return A.off + 256;
}
extern "C" size_t mmoe_home_workspace_bytes(const mmoe_home_cfg* cfg, int32_t B, int dtype) {
  Arena A(nullptr);
  home_scratch_layout(A, *cfg, B, dtype);
  return A.off + 256;
}
extern "C" int mmoe_home_fwd(const mmoe_call* c, const mmoe_home_cfg* cfg, const float* expert_vecs, float* logits, float* gate_w) {
  MMOE_TRY(check_home(cfg));
  if (c->B == 0) return 0;
  MMOE_CHECK(c->saved != nullptr && c->saved_bytes >= mmoe_home_saved_bytes(cfg, c->B, c->dtype), "home_fwd: saved blob too small");
  if (c->dtype == MMOE_BF16) return home_fwd_t<__nv_bfloat16>(c, cfg, expert_vecs, logits, gate_w);
  if (c->dtype == MMOE_F16) return home_fwd_t<__half>(c, cfg, expert_vecs, logits, gate_w);
  return home_fwd_t<float>(c, cfg, expert_vecs, logits, gate_w);
}
extern "C" int mmoe_home_bwd(const mmoe_call* c, const mmoe_home_cfg* cfg, const float* expert_vecs, const float* dlogits, float* d_expert_vecs) {
  MMOE_TRY(check_home(cfg));
  if (c->B == 0) return 0;
  MMOE_CHECK(c->saved != nullptr && c->saved_bytes >= mmoe_home_saved_bytes(cfg, c->B, c->dtype), "home_bwd: saved blob too small");
  MMOE_CHECK(c->workspace != nullptr && c->workspace_bytes >= mmoe_home_workspace_bytes(cfg, c->B, c->dtype), "home_bwd: workspace too small");
  if (c->dtype == MMOE_BF16) return home_bwd_t<__nv_bfloat16>(c, cfg, expert_vecs, dlogits, d_expert_vecs);
  if (c->dtype == MMOE_F16) return home_bwd_t<__half>(c, cfg, expert_vecs, dlogits, d_expert_vecs);
  return home_bwd_t<float>(c, cfg, expert_vecs, dlogits, d_expert_vecs);
}
