// Tensor-core variant of the fused masked-softmax attention (16-bit modes).
//
// One CTA per (batch, head), up to 4 warps; warp w owns query rows [16w, 16w+16).  Q, K, V (and dO in backward)
// tiles are staged once in shared memory with 128-bit loads; S = QK^T, P.V, dP = dO V^T, dQ = dS K, dV = P^T dO and
// dK = dS^T Q all run on mma.sync.m16n8k16 (fp32 accumulate) with the softmax / mask / dropout done on the
// accumulator fragments in registers (P never goes to shared or global memory in forward).  Transposed operands
// are read with ldmatrix.trans, so no transposed copy of V / P / dS is ever made.  S <= 64 keys fit one tile:
// no online-softmax rescaling is needed.
// The score tile is 64x64 per head, i.e. far too small to amortise a tcgen05/TMEM round trip (alloc + commit +
// tcgen05.ld per 2 x 0.8 MFLOP); the warp-level MMA keeps the whole attention in registers instead.
#include "kernels.cuh"

namespace mmoe {

template <typename T> struct MmaT;
template <> struct MmaT<__nv_bfloat16> {
  static __device__ __forceinline__ void mma(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  }
  static __device__ __forceinline__ uint32_t pack(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
  }
};
template <> struct MmaT<__half> {
  static __device__ __forceinline__ void mma(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  }
  static __device__ __forceinline__ uint32_t pack(float lo, float hi) {
    __half2 v = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
  }
};

// ---- fragment loaders (g = lane>>2, t = lane&3) ------------------------------------------------
// A (16 x 16) from memory [m][k], k contiguous
template <typename T>
__device__ __forceinline__ void frag_a_mk(uint32_t (&a)[4], const T* base, int pitch, int m0, int k0, int lane) {
  // four 8x8 blocks: a0 = (m 0-7, k 0-7)  a1 = (m 8-15, k 0-7)  a2 = (m 0-7, k 8-15)  a3 = (m 8-15, k 8-15)
  const int mat = lane >> 3, r = lane & 7;
  const uint32_t addr = (uint32_t)__cvta_generic_to_shared(base + (m0 + 8 * (mat & 1) + r) * pitch + k0 + 8 * (mat >> 1));
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]) : "r"(addr));
}
// B fragments of TWO adjacent 8-key tiles (n0..n0+7 and n0+8..n0+15) from memory [n][k], k contiguous
template <typename T>
__device__ __forceinline__ void frag_b_nk2(uint32_t (&b0)[2], uint32_t (&b1)[2], const T* base, int pitch, int n0, int k0, int lane) {
  const int mat = lane >> 3, r = lane & 7;
  const uint32_t addr = (uint32_t)__cvta_generic_to_shared(base + (n0 + 8 * (mat >> 1) + r) * pitch + k0 + 8 * (mat & 1));
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(b0[0]), "=r"(b0[1]), "=r"(b1[0]), "=r"(b1[1]) : "r"(addr));
}
// B (16 k x 8 n) from memory [k][n], n contiguous: two 8x8 blocks (k 0-7, k 8-15), transposed on load
template <typename T>
__device__ __forceinline__ void frag_b_kn(uint32_t (&b)[2], const T* base, int pitch, int k0, int n0, int lane) {
  const uint32_t addr = (uint32_t)__cvta_generic_to_shared(base + (k0 + (lane & 15)) * pitch + n0);
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(b[0]), "=r"(b[1]) : "r"(addr));
}
// A (16 m x 16 k) from memory [k][m], m contiguous: four 8x8 blocks transposed on load
//   a0 = (m 0-7, k 0-7)  a1 = (m 8-15, k 0-7)  a2 = (m 0-7, k 8-15)  a3 = (m 8-15, k 8-15)
template <typename T>
__device__ __forceinline__ void frag_a_km(uint32_t (&a)[4], const T* base, int pitch, int k0, int m0, int lane) {
  const int mat = lane >> 3, r = lane & 7;
  const uint32_t addr = (uint32_t)__cvta_generic_to_shared(base + (k0 + 8 * (mat >> 1) + r) * pitch + m0 + 8 * (mat & 1));
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]) : "r"(addr));
}

// global [rows][HD] (row stride ld) -> smem [64][PITCH]; rows >= valid are zero-filled.  Asynchronous 16-byte copies
// (cp.async, LDGSTS): every thread puts all of its chunks in flight before anyone waits, so a tile costs one memory
// latency instead of one per chunk.
// A half-warp owns a row (lanes 0..CH-1 of the half carry its 16-byte chunks), so addresses advance by a constant per
// trip: the previous element-indexed loop spent a fifth of the backward kernel's instructions on div/mod address math.
template <typename T, int HD>
__device__ __forceinline__ void stage_tile(T* dst, const T* src, int64_t ld, int valid, int rows_padded, int nthreads) {
  constexpr int PITCH = HD + 8, CH = HD / 8;
  static_assert(CH <= 16, "a row must fit a half-warp of 16-byte chunks");
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = nthreads >> 5;
  const int ch = lane & 15;
  if (ch >= CH) return;
  int r = 2 * warp + (lane >> 4);
  const T* g = src + (int64_t)r * ld + ch * 8;
  uint32_t d = (uint32_t)__cvta_generic_to_shared(dst + r * PITCH + ch * 8);
  const int64_t gstep = (int64_t)2 * nw * ld;
  const uint32_t dstep = (uint32_t)(2 * nw * PITCH * sizeof(T));
  for (; r < rows_padded; r += 2 * nw, g += gstep, d += dstep) {
    const bool ok = r < valid;
    const int bytes = ok ? 16 : 0;           // src-size 0 -> the 16 destination bytes are zero-filled
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(ok ? g : src), "r"(bytes) : "memory");
  }
}
__device__ __forceinline__ void stage_wait() {
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

struct AttnTcDev {
  const void *q, *k, *v; int64_t ldq, ldk, ldv;
  const uint8_t* mask;
  void* ctx; int64_t ldc;
  void *dq, *dk, *dv;
  float *bgq, *bgk, *bgv;
  int B, Sq, Sk, H;
  float qscale, drop_scale; uint32_t thresh, k0, k1;
};

// scores -> probabilities on the accumulator fragments of one warp (rows g and g+8 of its 16-row slab)
// s[nt][0..1]: row g, keys nt*8+2t,+1;  s[nt][2..3]: row g+8.
__device__ __forceinline__ void softmax_frag(float (&s)[8][4], int NT, int Sk, const uint8_t* mrow, float qscale, int lane) {
  const int t = lane & 3;
  float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    if (nt < NT) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int key = nt * 8 + 2 * t + e;
        const bool dead = key >= Sk || (mrow != nullptr && mrow[key]);
        s[nt][e] = dead ? -INFINITY : s[nt][e] * qscale;
        s[nt][2 + e] = dead ? -INFINITY : s[nt][2 + e] * qscale;
        mx0 = fmaxf(mx0, s[nt][e]);
        mx1 = fmaxf(mx1, s[nt][2 + e]);
      }
    }
  }
  mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
  mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
  float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    if (nt < NT) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        // a fully padded row gives exp(-inf - -inf) = NaN, as torch's softmax does
        s[nt][e] = __expf(s[nt][e] - mx0); sum0 += s[nt][e];          // ex2.approx path: 16-bit operands downstream
        s[nt][2 + e] = __expf(s[nt][2 + e] - mx1); sum1 += s[nt][2 + e];
      }
    }
  }
  sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1); sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
  sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1); sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
  const float i0 = 1.f / sum0, i1 = 1.f / sum1;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    if (nt < NT) { s[nt][0] *= i0; s[nt][1] *= i0; s[nt][2] *= i1; s[nt][3] *= i1; }
  }
}

template <typename T, int HD>
__global__ void __launch_bounds__(128) attn_tc_fwd_kernel(const AttnTcDev a) {
  constexpr int PITCH = HD + 8;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  pdl_entry();
  T* Qs = reinterpret_cast<T*>(smem_raw);
  T* Ks = Qs + 64 * PITCH;
  T* Vs = Ks + 64 * PITCH;
  const int b = blockIdx.x / a.H, h = blockIdx.x - b * a.H;
  const int Sq = a.Sq, Sk = a.Sk;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int KT = (Sk + 15) / 16, NT = (Sk + 7) / 8;
  const int qrows = ((Sq + 15) / 16) * 16, krows = KT * 16;
  stage_tile<T, HD>(Qs, (const T*)a.q + (int64_t)b * Sq * a.ldq + h * HD, a.ldq, Sq, qrows, blockDim.x);
  stage_tile<T, HD>(Ks, (const T*)a.k + (int64_t)b * Sk * a.ldk + h * HD, a.ldk, Sk, krows, blockDim.x);
  stage_tile<T, HD>(Vs, (const T*)a.v + (int64_t)b * Sk * a.ldv + h * HD, a.ldv, Sk, krows, blockDim.x);
  __shared__ uint8_t smask[64];
  if (a.mask != nullptr && threadIdx.x < Sk) smask[threadIdx.x] = a.mask[(int64_t)b * Sk + threadIdx.x];
  stage_wait();
  __syncthreads();
  const int row0 = warp * 16;
  if (row0 >= Sq) return;
  float s[8][4];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
#pragma unroll
  for (int kk = 0; kk < HD / 16; ++kk) {
    uint32_t af[4];
    frag_a_mk<T>(af, Qs, PITCH, row0, kk * 16, lane);
#pragma unroll
    for (int nt = 0; nt < 8; nt += 2) {
      if (nt < NT) {                   // an odd NT reads one zero-filled key tile too many (rows < krows are staged)
        uint32_t b0[2], b1[2];
        frag_b_nk2<T>(b0, b1, Ks, PITCH, nt * 8, kk * 16, lane);
        MmaT<T>::mma(s[nt], af, b0);
        MmaT<T>::mma(s[nt + 1], af, b1);
      }
    }
  }
  softmax_frag(s, NT, Sk, a.mask ? smask : nullptr, a.qscale, lane);
  if (a.thresh != 0) {
    // element (row, key) has flat index base + row*Sk + key; keys 2t, 2t+1 share one hash when Sk is even
    const uint64_t base = (uint64_t)blockIdx.x * Sq * Sk;
    const bool paired = (Sk & 1) == 0;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      if (nt < NT) {
#pragma unroll
        for (int hrow = 0; hrow < 2; ++hrow) {
          const int row = row0 + g + hrow * 8, key = nt * 8 + 2 * t;
          const uint64_t idx = base + (uint64_t)row * Sk + key;
          bool k0, k1;
          if (paired) {
            const uint32_t hsh = drop_hash(a.k0, a.k1, idx >> 1);
            k0 = (hsh & 0xFFFFu) >= a.thresh; k1 = (hsh >> 16) >= a.thresh;
          } else {
            k0 = drop_keep(a.k0, a.k1, idx, a.thresh); k1 = drop_keep(a.k0, a.k1, idx + 1, a.thresh);
          }
          s[nt][2 * hrow] = k0 ? s[nt][2 * hrow] * a.drop_scale : 0.f;
          s[nt][2 * hrow + 1] = k1 ? s[nt][2 * hrow + 1] * a.drop_scale : 0.f;
        }
      }
    }
  }
  float o[HD / 8][4];
#pragma unroll
  for (int nt = 0; nt < HD / 8; ++nt) o[nt][0] = o[nt][1] = o[nt][2] = o[nt][3] = 0.f;
#pragma unroll
  for (int kt = 0; kt < 4; ++kt) {
    if (kt < KT) {
      uint32_t af[4];
      af[0] = MmaT<T>::pack(s[2 * kt][0], s[2 * kt][1]);
      af[1] = MmaT<T>::pack(s[2 * kt][2], s[2 * kt][3]);
      af[2] = MmaT<T>::pack(s[2 * kt + 1][0], s[2 * kt + 1][1]);
      af[3] = MmaT<T>::pack(s[2 * kt + 1][2], s[2 * kt + 1][3]);
#pragma unroll
      for (int nt = 0; nt < HD / 8; ++nt) {
        uint32_t bf[2];
        frag_b_kn<T>(bf, Vs, PITCH, kt * 16, nt * 8, lane);
        MmaT<T>::mma(o[nt], af, bf);
      }
    }
  }
  T* out = (T*)a.ctx + (int64_t)b * Sq * a.ldc + h * HD;
#pragma unroll
  for (int nt = 0; nt < HD / 8; ++nt) {
    const int col = nt * 8 + 2 * t;
    if (row0 + g < Sq) *reinterpret_cast<uint32_t*>(out + (int64_t)(row0 + g) * a.ldc + col) = MmaT<T>::pack(o[nt][0], o[nt][1]);
    if (row0 + g + 8 < Sq) *reinterpret_cast<uint32_t*>(out + (int64_t)(row0 + g + 8) * a.ldc + col) = MmaT<T>::pack(o[nt][2], o[nt][3]);
  }
}

// column sums of a warp's 16 x (HD) accumulator slab -> this warp's private row of partials (plain stores: a float
// atomicAdd on shared memory is a CAS loop, and four warps contending on it cost a quarter of the kernel)
template <int HD>
__device__ __forceinline__ void colsum_to_smem(const float (&acc)[HD / 8][4], float* dst, int lane) {
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int nt = 0; nt < HD / 8; ++nt) {
    float c0 = acc[nt][0] + acc[nt][2], c1 = acc[nt][1] + acc[nt][3];
#pragma unroll
    for (int o = 4; o < 32; o <<= 1) { c0 += __shfl_xor_sync(0xffffffffu, c0, o); c1 += __shfl_xor_sync(0xffffffffu, c1, o); }
    if (g == 0) *reinterpret_cast<float2*>(dst + nt * 8 + 2 * t) = make_float2(c0, c1);
  }
}

template <typename T, int HD>
__global__ void __launch_bounds__(128) attn_tc_bwd_kernel(const AttnTcDev a) {
  constexpr int PITCH = HD + 8, PP = 72;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  pdl_entry();
  T* Qs = reinterpret_cast<T*>(smem_raw);
  T* Ks = Qs + 64 * PITCH;
  T* Vs = Ks + 64 * PITCH;
  T* dOs = Vs + 64 * PITCH;
  T* Ps = Vs;                      // probabilities after dropout, [query][key]: takes over V's tile once S and dP exist
  T* dSs = dOs + 64 * PITCH;       // qscale * dS, [query][key]
  float* bg = reinterpret_cast<float*>(dSs + 64 * PP);   // [4 warps][3][HD] warp-private column-sum partials
  const int b = blockIdx.x / a.H, h = blockIdx.x - b * a.H;
  const int Sq = a.Sq, Sk = a.Sk;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int KT = (Sk + 15) / 16, NT = (Sk + 7) / 8, QT = (Sq + 15) / 16;
  const int qrows = QT * 16, krows = KT * 16;
  stage_tile<T, HD>(Qs, (const T*)a.q + (int64_t)b * Sq * a.ldq + h * HD, a.ldq, Sq, qrows, blockDim.x);
  stage_tile<T, HD>(Ks, (const T*)a.k + (int64_t)b * Sk * a.ldk + h * HD, a.ldk, Sk, krows, blockDim.x);
  stage_tile<T, HD>(Vs, (const T*)a.v + (int64_t)b * Sk * a.ldv + h * HD, a.ldv, Sk, krows, blockDim.x);
  stage_tile<T, HD>(dOs, (const T*)a.ctx + (int64_t)b * Sq * a.ldc + h * HD, a.ldc, Sq, qrows, blockDim.x);
  __shared__ uint8_t smask[64];
  if (a.mask != nullptr && threadIdx.x < Sk) smask[threadIdx.x] = a.mask[(int64_t)b * Sk + threadIdx.x];
  for (int e = threadIdx.x; e < 4 * 3 * HD; e += blockDim.x) bg[e] = 0.f;
  // keys between Sk rounded up to 8 and Sk rounded up to 16 are written by nobody: zero them for phase 2
  const bool zero_scores = NT * 8 != krows;
  if (zero_scores)
    for (int e = threadIdx.x; e < 64 * PP / 2; e += blockDim.x) reinterpret_cast<uint32_t*>(dSs)[e] = 0u;
  stage_wait();
  __syncthreads();
  const int row0 = warp * 16;
  float s[8][4], dp[8][4];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) { s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f; dp[nt][0] = dp[nt][1] = dp[nt][2] = dp[nt][3] = 0.f; }
  if (row0 < Sq) {
#pragma unroll
    for (int kk = 0; kk < HD / 16; ++kk) {
      uint32_t aq[4], ad[4];
      frag_a_mk<T>(aq, Qs, PITCH, row0, kk * 16, lane);
      frag_a_mk<T>(ad, dOs, PITCH, row0, kk * 16, lane);
#pragma unroll
      for (int nt = 0; nt < 8; nt += 2) {
        if (nt < NT) {               // an odd NT reads one zero-filled key tile too many (rows < krows are staged)
          uint32_t bk0[2], bk1[2], bv0[2], bv1[2];
          frag_b_nk2<T>(bk0, bk1, Ks, PITCH, nt * 8, kk * 16, lane);
          frag_b_nk2<T>(bv0, bv1, Vs, PITCH, nt * 8, kk * 16, lane);
          MmaT<T>::mma(s[nt], aq, bk0);      // S  = Q K^T
          MmaT<T>::mma(dp[nt], ad, bv0);     // dP = dO V^T
          MmaT<T>::mma(s[nt + 1], aq, bk1);
          MmaT<T>::mma(dp[nt + 1], ad, bv1);
        }
      }
    }
  }
  __syncthreads();                   // every warp is done with V: its tile becomes P
  if (zero_scores)
    for (int e = threadIdx.x; e < 64 * PP / 2; e += blockDim.x) reinterpret_cast<uint32_t*>(Ps)[e] = 0u;
  if (zero_scores) __syncthreads();
  if (row0 < Sq) {
    softmax_frag(s, NT, Sk, a.mask ? smask : nullptr, a.qscale, lane);
    // dropout on P and dP, delta = sum_j P * dP, dS = P * (dP - delta)
    const uint64_t base = (uint64_t)blockIdx.x * Sq * Sk;
    float dl0 = 0.f, dl1 = 0.f;
    float pd[8][4];
    const bool paired = (Sk & 1) == 0;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      bool keep[4] = {true, true, true, true};
      if (nt < NT && a.thresh != 0) {
#pragma unroll
        for (int hrow = 0; hrow < 2; ++hrow) {
          const int row = row0 + g + hrow * 8, key = nt * 8 + 2 * t;
          const uint64_t idx = base + (uint64_t)row * Sk + key;
          if (paired) {
            const uint32_t hsh = drop_hash(a.k0, a.k1, idx >> 1);
            keep[2 * hrow] = (hsh & 0xFFFFu) >= a.thresh; keep[2 * hrow + 1] = (hsh >> 16) >= a.thresh;
          } else {
            keep[2 * hrow] = drop_keep(a.k0, a.k1, idx, a.thresh); keep[2 * hrow + 1] = drop_keep(a.k0, a.k1, idx + 1, a.thresh);
          }
        }
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        pd[nt][e] = 0.f;
        if (nt < NT) {
          const float p = s[nt][e];
          float d = dp[nt][e], pdrop = p;
          if (a.thresh != 0) {
            d = keep[e] ? d * a.drop_scale : 0.f;
            pdrop = keep[e] ? p * a.drop_scale : 0.f;
          }
          pd[nt][e] = pdrop;
          dp[nt][e] = d;
          if (e < 2) dl0 += p * d; else dl1 += p * d;
        }
      }
    }
    dl0 += __shfl_xor_sync(0xffffffffu, dl0, 1); dl0 += __shfl_xor_sync(0xffffffffu, dl0, 2);
    dl1 += __shfl_xor_sync(0xffffffffu, dl1, 1); dl1 += __shfl_xor_sync(0xffffffffu, dl1, 2);
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      if (nt < NT) {
        dp[nt][0] = s[nt][0] * (dp[nt][0] - dl0) * a.qscale;
        dp[nt][1] = s[nt][1] * (dp[nt][1] - dl0) * a.qscale;
        dp[nt][2] = s[nt][2] * (dp[nt][2] - dl1) * a.qscale;
        dp[nt][3] = s[nt][3] * (dp[nt][3] - dl1) * a.qscale;
        const int key = nt * 8 + 2 * t;
        *reinterpret_cast<uint32_t*>(Ps + (row0 + g) * PP + key) = MmaT<T>::pack(pd[nt][0], pd[nt][1]);
        *reinterpret_cast<uint32_t*>(Ps + (row0 + g + 8) * PP + key) = MmaT<T>::pack(pd[nt][2], pd[nt][3]);
        *reinterpret_cast<uint32_t*>(dSs + (row0 + g) * PP + key) = MmaT<T>::pack(dp[nt][0], dp[nt][1]);
        *reinterpret_cast<uint32_t*>(dSs + (row0 + g + 8) * PP + key) = MmaT<T>::pack(dp[nt][2], dp[nt][3]);
      }
    }
    // dQ = (qscale dS) K
    float dq[HD / 8][4];
#pragma unroll
    for (int nt = 0; nt < HD / 8; ++nt) dq[nt][0] = dq[nt][1] = dq[nt][2] = dq[nt][3] = 0.f;
#pragma unroll
    for (int kt = 0; kt < 4; ++kt) {
      if (kt < KT) {
        uint32_t af[4];
        af[0] = MmaT<T>::pack(dp[2 * kt][0], dp[2 * kt][1]);
        af[1] = MmaT<T>::pack(dp[2 * kt][2], dp[2 * kt][3]);
        af[2] = MmaT<T>::pack(dp[2 * kt + 1][0], dp[2 * kt + 1][1]);
        af[3] = MmaT<T>::pack(dp[2 * kt + 1][2], dp[2 * kt + 1][3]);
#pragma unroll
        for (int nt = 0; nt < HD / 8; ++nt) {
          uint32_t bf[2];
          frag_b_kn<T>(bf, Ks, PITCH, kt * 16, nt * 8, lane);
          MmaT<T>::mma(dq[nt], af, bf);
        }
      }
    }
    {
      const bool ok0 = row0 + g < Sq, ok1 = row0 + g + 8 < Sq;
      T* r0p = (T*)a.dq + ((int64_t)b * Sq + row0 + g) * a.ldq + h * HD + 2 * t;
      T* r1p = r0p + 8 * a.ldq;
#pragma unroll
      for (int nt = 0; nt < HD / 8; ++nt) {
        if (ok0) *reinterpret_cast<uint32_t*>(r0p + nt * 8) = MmaT<T>::pack(dq[nt][0], dq[nt][1]);
        else { dq[nt][0] = dq[nt][1] = 0.f; }
        if (ok1) *reinterpret_cast<uint32_t*>(r1p + nt * 8) = MmaT<T>::pack(dq[nt][2], dq[nt][3]);
        else { dq[nt][2] = dq[nt][3] = 0.f; }
      }
    }
    if (a.bgq != nullptr) colsum_to_smem<HD>(dq, bg + warp * 3 * HD, lane);
  }
  __syncthreads();
  // phase 2: warp w owns keys [16w, 16w+16):  dV = Pd^T dO,  dK = (qscale dS)^T Q
  const int key0 = warp * 16;
  if (key0 < Sk) {
    float dv[HD / 8][4], dk[HD / 8][4];
#pragma unroll
    for (int nt = 0; nt < HD / 8; ++nt) { dv[nt][0] = dv[nt][1] = dv[nt][2] = dv[nt][3] = 0.f; dk[nt][0] = dk[nt][1] = dk[nt][2] = dk[nt][3] = 0.f; }
#pragma unroll
    for (int qt = 0; qt < 4; ++qt) {
      if (qt < QT) {
        uint32_t ap[4], as_[4];
        frag_a_km<T>(ap, Ps, PP, qt * 16, key0, lane);
        frag_a_km<T>(as_, dSs, PP, qt * 16, key0, lane);
#pragma unroll
        for (int nt = 0; nt < HD / 8; ++nt) {
          uint32_t bo[2], bq[2];
          frag_b_kn<T>(bo, dOs, PITCH, qt * 16, nt * 8, lane);
          frag_b_kn<T>(bq, Qs, PITCH, qt * 16, nt * 8, lane);
          MmaT<T>::mma(dv[nt], ap, bo);
          MmaT<T>::mma(dk[nt], as_, bq);
        }
      }
    }
    {
      const bool ok0 = key0 + g < Sk, ok1 = key0 + g + 8 < Sk;
      T* k0p = (T*)a.dk + ((int64_t)b * Sk + key0 + g) * a.ldk + h * HD + 2 * t;
      T* k1p = k0p + 8 * a.ldk;
      T* v0p = (T*)a.dv + ((int64_t)b * Sk + key0 + g) * a.ldv + h * HD + 2 * t;
      T* v1p = v0p + 8 * a.ldv;
#pragma unroll
      for (int nt = 0; nt < HD / 8; ++nt) {
        if (ok0) {
          *reinterpret_cast<uint32_t*>(k0p + nt * 8) = MmaT<T>::pack(dk[nt][0], dk[nt][1]);
          *reinterpret_cast<uint32_t*>(v0p + nt * 8) = MmaT<T>::pack(dv[nt][0], dv[nt][1]);
        } else { dk[nt][0] = dk[nt][1] = dv[nt][0] = dv[nt][1] = 0.f; }
        if (ok1) {
          *reinterpret_cast<uint32_t*>(k1p + nt * 8) = MmaT<T>::pack(dk[nt][2], dk[nt][3]);
          *reinterpret_cast<uint32_t*>(v1p + nt * 8) = MmaT<T>::pack(dv[nt][2], dv[nt][3]);
        } else { dk[nt][2] = dk[nt][3] = dv[nt][2] = dv[nt][3] = 0.f; }
      }
    }
    if (a.bgk != nullptr) colsum_to_smem<HD>(dk, bg + warp * 3 * HD + HD, lane);
    if (a.bgv != nullptr) colsum_to_smem<HD>(dv, bg + warp * 3 * HD + 2 * HD, lane);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 3 * HD; c += blockDim.x) {
    float* dst = c < HD ? a.bgq : (c < 2 * HD ? a.bgk : a.bgv);
    if (dst != nullptr) {
      const float v = bg[c] + bg[3 * HD + c] + bg[6 * HD + c] + bg[9 * HD + c];
      atomicAdd(dst + h * HD + (c % HD), v);
    }
  }
}

static void fill_tc(AttnTcDev* d, const AttnArgs& a) {
  d->q = a.q; d->k = a.k; d->v = a.v; d->ldq = a.ldq; d->ldk = a.ldk; d->ldv = a.ldv;
  d->mask = a.mask; d->ctx = a.ctx; d->ldc = a.ldc; d->dq = a.dq; d->dk = a.dk; d->dv = a.dv;
  d->bgq = a.bgq; d->bgk = a.bgk; d->bgv = a.bgv;
  d->B = a.B; d->Sq = a.Sq; d->Sk = a.Sk; d->H = a.H;
  d->qscale = 1.0f / sqrtf((float)a.hd);
  d->thresh = a.drop_p > 0.f ? drop_threshold(a.drop_p) : 0u;
  d->drop_scale = a.drop_p > 0.f ? 1.f / (1.f - a.drop_p) : 1.f;
  d->k0 = a.k0; d->k1 = a.k1;
}

bool attention_tc_supported(const AttnArgs& a) {
  if (a.dtype == MMOE_F32) return false;
  if (!(a.hd == 64 || a.hd == 96 || a.hd == 128)) return false;
  if (a.Sq > 64 || a.Sk > 64) return false;
  // 128-bit staging and 32-bit packed stores need 16-byte aligned rows
  auto ok = [](const void* p, int64_t ld) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0 && (ld % 8) == 0; };
  return ok(a.q, a.ldq) && ok(a.k, a.ldk) && ok(a.v, a.ldv) && ok(a.ctx, a.ldc) &&
         (a.dq == nullptr || (ok(a.dq, a.ldq) && ok(a.dk, a.ldk) && ok(a.dv, a.ldv)));
}

template <typename T, int HD>
static int launch_tc(const AttnArgs& a, bool bwd, cudaStream_t s) {
  AttnTcDev d;
  fill_tc(&d, a);
  const int warps = bwd ? ((a.Sq > a.Sk ? a.Sq : a.Sk) + 15) / 16 : (a.Sq + 15) / 16;
  const int threads = warps * 32;
  if (!bwd) {
    const size_t smem = (size_t)3 * 64 * (HD + 8) * sizeof(T);
    static bool cfg = false;
    if (!cfg) { MMOE_CUDA(cudaFuncSetAttribute(attn_tc_fwd_kernel<T, HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); cfg = true; }
    MMOE_CUDA(launch_pdl(attn_tc_fwd_kernel<T, HD>, a.B * a.H, threads, smem, s, d));
    MMOE_LAUNCH_OK("attn_tc_fwd_kernel");
  } else {
    const size_t smem = (size_t)4 * 64 * (HD + 8) * sizeof(T) + (size_t)64 * 72 * sizeof(T) + 4 * 3 * HD * sizeof(float);
    static bool cfg = false;
    if (!cfg) { MMOE_CUDA(cudaFuncSetAttribute(attn_tc_bwd_kernel<T, HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); cfg = true; }
    MMOE_CUDA(launch_pdl(attn_tc_bwd_kernel<T, HD>, a.B * a.H, threads, smem, s, d));
    MMOE_LAUNCH_OK("attn_tc_bwd_kernel");
  }
  return 0;
}

template <typename T>
static int dispatch_hd(const AttnArgs& a, bool bwd, cudaStream_t s) {
  if (a.hd == 64) return launch_tc<T, 64>(a, bwd, s);
  if (a.hd == 96) return launch_tc<T, 96>(a, bwd, s);
  return launch_tc<T, 128>(a, bwd, s);
}

int attention_tc(const AttnArgs& a, bool bwd, cudaStream_t s) {
  if (a.dtype == MMOE_BF16) return dispatch_hd<__nv_bfloat16>(a, bwd, s);
  return dispatch_hd<__half>(a, bwd, s);
}

}  // namespace mmoe
