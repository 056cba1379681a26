/*
 * mmoe_b200.h — C ABI of the B200-native (sm_100a) MMoE / HoME fusion-and-head path.
 *
 * The reference (JingxiangQU/mmoe-multimodal-rec) is 100 % Python and has no FFI of its
 * own: its "operator interface" for this path is the forward() of the
 * torch modules in model.py / model_HoME.py.  This header is the boundary a maintainer
 * binds instead of those forward() bodies; every entry point names the reference lines
 * it replaces.  The Python drop-ins (model.py, model_HoME.py at the repo root) bind it
 * with ctypes; INTEGRATION.md shows the stub.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless it says host.
 *   - the caller owns all memory (inputs, outputs, saved-for-backward blobs, workspaces);
 *     the library never allocates device memory.  *_saved_bytes / *_workspace_bytes give
 *     the sizes to allocate.
 *   - every call is asynchronous on the CUDA stream passed as `stream` (a cudaStream_t).
 *   - re-entrant and thread-safe: forward runs on the Python thread, backward on the
 *     autograd worker thread.  No global mutable state besides a thread-local error string.
 *   - return value: 0 on success, negative on error; mmoe_last_error() returns the message
 *     for the calling thread.
 *   - there is NO CPU fallback: a call without a usable sm_100 device fails.
 *
 * dtype
 *   MMOE_F32: fp32 activations/weights everywhere (exact path, SIMT FFMA GEMMs).
 *   MMOE_BF16 / MMOE_F16: 16-bit GEMM operands on tcgen05 tensor cores with fp32 TMEM
 *   accumulation; the residual stream, LayerNorm/softmax statistics, biases, LayerNorm
 *   gains, outputs and all gradients of parameters stay fp32.
 *   "T" below means the 16-bit type in 16-bit modes and float in MMOE_F32 mode.
 */
#ifndef MMOE_B200_H_
#define MMOE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MMOE_ABI_VERSION 2

typedef enum { MMOE_F32 = 0, MMOE_BF16 = 1, MMOE_F16 = 2 } mmoe_dtype;

/* ---- library ------------------------------------------------------------------------- */
int         mmoe_abi_version(void);
/* sizeof() of the ABI structs as compiled (0 epilogue, 1 gemm_problem, 2 call, 3 head_cfg, 4 cross_cfg,
 * 5 fuse_cfg, 6 home_cfg) so a foreign-language binding can verify its mirror of the layout. */
size_t      mmoe_abi_sizeof(int which);
const char* mmoe_last_error(void);
/* Checks that the current device is sm_100 and resolves the driver entry points. */
int         mmoe_init(void);
/* Number of kernels launched by this library on the calling thread since the last reset
 * (bench.py's "gpu_launches"). */
int64_t     mmoe_launch_count(int reset);

/* ---- elementary operators (used by the module entry points; exported for unit tests) -- */

/* y = cast(x): fp32 -> T (n elements).  Weight down-cast once per optimizer step. */
int mmoe_cast_f32(const float* x, void* y, int64_t n, int dtype, void* stream);

/* GEMM  D[M,N] = epilogue( sum_k A(m,k) * B(n,k) ).
 *   a_major / b_major: 0 = K-major  (memory [M or N][K], K contiguous),
 *                      1 = MN-major (memory [K][M or N], M/N contiguous).
 *   nn.Linear forward  x[M,K] W[N,K]^T            : a_major 0, b_major 0
 *   dgrad  dY[M,N'] W[N',K'] -> dX[M,K']           : a_major 0, b_major 1
 *   wgrad  dY[M',N]^T X[M',K] -> dW[N,K]           : a_major 1, b_major 1
 * Epilogue, in order: v = alpha*acc (+ bias[n]); preact store; activation; backward
 * multiplier; dropout; (+ residual[m,n]); column-sum; store / atomic-accumulate. */
typedef struct {
  void*        out;         /* [M, ldo]; NULL = no store (colsum only)                     */
  int32_t      out_dtype;   /* MMOE_F32 or the operand dtype                                */
  int32_t      accumulate;  /* 1: fp32 atomic add into out (needed when k_splits > 1)       */
  int64_t      ldo;
  const float* bias;        /* [N] or NULL                                                  */
  void*        preact;      /* T [M, ldo] or NULL: v after bias, before the activation      */
  int32_t      act;         /* 0 none, 1 ReLU, 2 GELU(erf), 3 sigmoid                       */
  int32_t      bwd_mode;    /* 0 none; 1: v *= (aux != 0) (ReLU[+dropout] backward, aux =    */
                            /*    saved post-dropout output; scale by 1/(1-p) if drop_p>0);  */
                            /* 2: v *= gelu'(aux); 3: v *= s(1-s), s = sigmoid(aux);         */
                            /* 4: as 1, but aux is the BIT mask written through mask_out     */
                            /*    (uint64 [M, N/64]); 16-bit tensor-core path only           */
  const void*  aux;         /* T [M, ld_aux]  (mode 4: uint64 [M, N/64], ld_aux ignored)    */
  int64_t      ld_aux;
  const float* residual;    /* fp32 [M, ld_res] or NULL, added last                         */
  int64_t      ld_res;
  float*       colsum;      /* fp32 [N] or NULL: += sum_m of the stored value               */
  float        alpha;
  float        drop_p;      /* dropout after the activation / backward multiplier           */
  uint32_t     drop_key0, drop_key1;   /* mask(m,n) = f(key, m*N + n), see mmoe_dropout_mask */
  void*        mask_out;    /* uint64 [M, N/64] or NULL: one flag per element = (stored 16-bit value != 0), for the columns
                               64w .. 64w+63 of row m in word [m, w].  Flag of column 64w + c:  bit  32*(c/32) + (c odd ?
                               31 : 15) - (c%32)/2  (the order the epilogue's packed 16-bit pairs produce with 3 integer
                               operations per pair).  Lets the ReLU(+dropout) backward read 1 bit per element instead of
                               the saved activation; consumed by bwd_mode 4.  16-bit tensor-core path only (N % 64 == 0,
                               act == ReLU); fails loudly otherwise. */
} mmoe_epilogue;

typedef struct {
  const void* a; int64_t lda; int32_t a_major;
  const void* b; int64_t ldb; int32_t b_major;
  int32_t M, N, K;
  int32_t k_splits;         /* >1 splits K over CTAs; requires epi.accumulate; 0 = let the
                               launcher choose (1 unless epi.accumulate)                    */
  mmoe_epilogue epi;
} mmoe_gemm_problem;

/* Runs n_problems (<= 8) GEMMs of one dtype in ONE launch (grouped persistent kernel).
 * engine: 0 = default for the dtype (tcgen05 for 16-bit, SIMT for fp32),
 *         1 = force the SIMT kernel (cross-check in tests). */
int mmoe_gemm_grouped(const mmoe_gemm_problem* problems /*host*/, int n_problems, int dtype,
                      int engine, void* stream);

/* Live timing of the GEMM launches (bench.py's roofline): while enabled, CUDA events are recorded on the launching
 * stream around every mmoe_gemm_grouped launch (also those issued by the module entry points).  _read sums the
 * durations / algorithmic FLOPs (2*M*N*K) recorded since the last read — call it after a device synchronize.
 * enable > 1 also pre-creates event pairs for that many launches, so no event is created inside a timed region. */
int mmoe_gemm_timing(int enable);
int mmoe_gemm_timing_read(double* total_ms, double* total_flops, int64_t* launches, int tc_only);
/* Per-launch records since the last _read, not cleared (call after a device synchronize and before _read): up to max_rows
 * rows of 10 doubles {ms, 2*M*N*K summed over the launch, tensor-core launch?, tile width, CTAs per tile, problems in the
 * launch, M, N, K of the first problem, its a_major | b_major << 1}.  Returns the number of records available. */
int mmoe_gemm_timing_dump(double* rows, int max_rows);

/* Launch trace of the tensor-core GEMM launches (tests: which kernel variant a shape resolved to).  enable != 0 clears the
 * record and starts recording, 0 stops.  _read copies up to max_entries records of 4 ints {tile width, CTAs per tile (2 =
 * cta_group::2 pair kernel), rich epilogue, tiles} into out (may be NULL) and returns the number recorded so far. */
int mmoe_launch_trace(int enable);
int mmoe_launch_trace_read(int32_t* out, int max_entries);

/* Number of SMs the persistent GEMM leaves free (its CTAs own a whole SM each and cannot co-reside with a running
 * NCCL kernel).  Set it to the number of NCCL CTAs when gradient all-reduce overlaps backward; default 0 or the
 * environment variable MMOE_SM_RESERVE. */
int mmoe_set_sm_reserve(int n_sms);

/* keep-mask a kernel would use for flat element index i in [0,n): out[i] = 1/0 (uint8). */
int mmoe_dropout_mask(uint32_t key0, uint32_t key1, float p, int64_t n, uint8_t* out, void* stream);
/* Keys of dropout site `site` of a module call made with mmoe_call.seed = seed (host function).  Site numbers, with the
 * tensor whose row-major flat index the mask is taken over (the reference's nn.Dropout call sites in brackets):
 *   encoder layer l (model.py:207-212 / :460-465; stack base s0 = 16*l, +8 for the item stack of the cross expert):
 *     s0+0 attention probabilities [B,H,Sq,Sk] (MHA dropout), s0+1 dropout1 [B*S,d], s0+2 FFN dropout [B*S,4d],
 *     s0+3 dropout2 [B*S,d];
 *   RobustTextCrossExpert: 100 cross-attention probabilities [B,H,S,S] (model.py:407-408), 101 AttnPool1D weights [B,S]
 *     (:197,204), 102 mlp[2] [B,4d], 103 mlp[4] [B,d] (:421-423);
 *   EnhancedCrossFuse: 100 proj[3] [B,d] (model.py:488);
 *   TwoTaskMMoE: 10+t tower_t[3] [B,hidden], 20+t tower_t[6] [B,hidden/2] (model.py:545-556; t = 0 good, 1 best);
 *   HOME_MMoE_Complete: 10+e ExpertMLP e dropout [B,1024] (model_HoME.py:33; e: meta 0..n_shared-1, then good, then best),
 *     30+t tower_t[3] [B,tower_hidden] (:586);
 *   ItemImageExpert: 0 dropout [B,d] (model.py:384). */
int mmoe_site_keys(uint64_t seed, uint32_t site, uint32_t* key0, uint32_t* key1);

/* LayerNorm over the last dim (eps 1e-5, biased variance — nn.LayerNorm).
 * x: fp32 or T ([rows, d]); y: T and/or fp32 copies (either may be NULL); stats: fp32 [rows,2]
 * (mean, rstd) or NULL. */
int mmoe_layernorm_fwd(const void* x, int x_dtype, const float* gamma, const float* beta,
                       void* y_t, float* y_f32, float* stats, int64_t rows, int32_t d,
                       int dtype, void* stream);

/* Multi-head attention core on packed projections (nn.MultiheadAttention semantics,
 * SURVEY.md Appendix A): q [B,Sq,*], k,v [B,Sk,*] with row strides ldq/ldk/ldv (elements),
 * head h at columns h*hd..; key_padding_mask uint8 [B,Sk] (1 = padded) or NULL;
 * ctx [B,Sq,n_head*hd] T.  Dropout acts on the probabilities. */
int mmoe_attention_fwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                       const uint8_t* key_padding_mask, void* ctx, int64_t ldc,
                       int32_t B, int32_t Sq, int32_t Sk, int32_t n_head, int32_t hd,
                       float drop_p, uint32_t key0, uint32_t key1, int dtype, void* stream);
/* Gradients dq,dk,dv (T, same strides as q,k,v); bias_grad_{q,k,v}: fp32 [n_head*hd] column
 * sums accumulated atomically, or NULL. */
int mmoe_attention_bwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                       const uint8_t* key_padding_mask, const void* dctx, int64_t ldc,
                       void* dq, void* dk, void* dv,
                       float* bias_grad_q, float* bias_grad_k, float* bias_grad_v,
                       int32_t B, int32_t Sq, int32_t Sk, int32_t n_head, int32_t hd,
                       float drop_p, uint32_t key0, uint32_t key1, int dtype, void* stream);

/* ---- module entry points ---------------------------------------------------------------
 * `params`: host array of device pointers in the order of the module's state_dict()
 * (SURVEY.md §8b lists the keys).  2-D weights are in T (the caller casts its fp32 masters
 * with mmoe_cast_f32, once per optimizer step); 1-D tensors (biases, LayerNorm gains), the
 * pooling query and the scalar gate are fp32.
 * `grads`: host array, same order, fp32 device buffers the backward ACCUMULATES into
 * (zero them first); an entry may be NULL for parameters that take no part (HoME variants).
 * `training` != 0 enables dropout with probability `drop_p`, keyed by `seed`.           */

typedef struct {
  int32_t dtype, B, training, home;   /* home: the model_HoME.py variant of the module     */
  float   drop_p;
  uint64_t seed;
  const void* const* params;          /* host array                                        */
  void* const* grads;                 /* host array (backward only)                        */
  void* saved;      size_t saved_bytes;
  void* workspace;  size_t workspace_bytes;
  void* stream;
} mmoe_call;

/* TwoTaskMMoE.forward — model.py:562-577 (ctor 532-559): mean query, two DenseGate softmaxes
 * (model.py:522-524), gate-weighted expert sums, two towers.  expert_vecs fp32 [B,n_expert,d];
 * logits fp32 [2,B] (good, best); gate_w fp32 [2,B,n_expert] or NULL. */
typedef struct { int32_t d, n_expert, hidden; float tower_drop_p; } mmoe_head_cfg;
size_t mmoe_head_saved_bytes(const mmoe_head_cfg* cfg, int32_t B, int dtype);
size_t mmoe_head_workspace_bytes(const mmoe_head_cfg* cfg, int32_t B, int dtype);
int mmoe_head_fwd(const mmoe_call* c, const mmoe_head_cfg* cfg, const float* expert_vecs,
                  float* logits, float* gate_w);
int mmoe_head_bwd(const mmoe_call* c, const mmoe_head_cfg* cfg, const float* expert_vecs,
                  const float* dlogits /*[2,B]*/, float* d_expert_vecs);

/* DenseGate.forward called on its own — model.py:522-524 / model_HoME.py:251-252:
 * w[B,n] = softmax(x[B,d] Wg[n,d]^T + bg), all fp32.  Backward: dl is a [B,n] fp32 scratch; dwg/dbg are
 * accumulated into (zero them first). */
int mmoe_dense_gate_fwd(const float* x, const float* wg, const float* bg, float* out, int64_t B, int32_t d, int32_t n, void* stream);
int mmoe_dense_gate_bwd(const float* x, const float* wg, const float* w, const float* dw, float* dl, float* dx,
                        float* dwg, float* dbg, int64_t B, int32_t d, int32_t n, void* stream);

/* RobustTextCrossExpert.forward — model.py:426-451 (RobustTransformerLayer 207-212,
 * AttnPool1D 199-206); home=1: model_HoME.py:441-466 (returns the pooled vector, AttnPool1D
 * finite-row guard 205-215).  user,item fp32 [B,S,d]; masks uint8 [B,S] (1 = padded);
 * out fp32 [B,d]. */
typedef struct { int32_t d, S, n_head, n_layer; } mmoe_cross_cfg;
size_t mmoe_cross_saved_bytes(const mmoe_cross_cfg* cfg, int32_t B, int dtype);
size_t mmoe_cross_workspace_bytes(const mmoe_cross_cfg* cfg, int32_t B, int dtype);
int mmoe_cross_fwd(const mmoe_call* c, const mmoe_cross_cfg* cfg, const float* user, const uint8_t* user_mask,
                   const float* item, const uint8_t* item_mask, float* out);
/* user,item: the same tensors that were given to the forward call. */
int mmoe_cross_bwd(const mmoe_call* c, const mmoe_cross_cfg* cfg, const float* user, const uint8_t* user_mask,
                   const float* item, const uint8_t* item_mask, const float* dout, float* d_user, float* d_item);

/* The same backward in stages, so that finished parameter gradients can be handed to DDP's bucketed all-reduce while
 * later stages still run.  stage: -1 = everything; 0 = tail (MLP, LN, pooling, cross attention; leaves the stream
 * gradients in the workspace); 100 + l = encoder layer l of the user stack; 200 + l = layer l of the item stack.
 * Order: tail first, then each stack from its last layer down.  All stages of one backward share the workspace and the
 * (zeroed) grads; d_user / d_item are written by the layer-0 stages. */
int mmoe_cross_bwd_stage(const mmoe_call* c, const mmoe_cross_cfg* cfg, int stage, const float* user, const uint8_t* user_mask,
                         const float* item, const uint8_t* item_mask, const float* dout, float* d_user, float* d_item);

/* Layout query for tests/debugging: byte range inside the saved blob of an encoder layer's activation.
 * stream_id 0 = user stack, 1 = item stack; which 0 = FFN activation h [B*S,4d] T, 1 = layer output [B*S,d] fp32. */
int mmoe_cross_saved_offset(const mmoe_cross_cfg* cfg, int32_t B, int dtype, int home, int stream_id, int layer, int which,
                            size_t* offset, size_t* bytes);

/* EnhancedCrossFuse.forward — model.py:491-507; home=1: model_HoME.py:506-522 (returns
 * fused + identity).  v_cls,t_cls fp32 [B,d]; out fp32 [B,d]. */
typedef struct { int32_t d, n_head, depth; } mmoe_fuse_cfg;
size_t mmoe_fuse_saved_bytes(const mmoe_fuse_cfg* cfg, int32_t B, int dtype);
size_t mmoe_fuse_workspace_bytes(const mmoe_fuse_cfg* cfg, int32_t B, int dtype);
int mmoe_fuse_fwd(const mmoe_call* c, const mmoe_fuse_cfg* cfg, const float* v_cls, const float* t_cls, float* out);
/* d_cat fp32 [B,2,d]: d_cat[:,0,:] is the gradient of v_cls, d_cat[:,1,:] of t_cls. */
int mmoe_fuse_bwd(const mmoe_call* c, const mmoe_fuse_cfg* cfg, const float* dout, float* d_cat);

int mmoe_fuse_saved_offset(const mmoe_fuse_cfg* cfg, int32_t B, int dtype, int home, int layer, int which,
                           size_t* offset, size_t* bytes);

/* HOME_MMoE_Complete.forward — model_HoME.py:590-638 (ExpertMLP 28-35, FeatureGate 232-234,
 * SelfGate 242-243, DenseGate 251-252, tower 581-588).  expert_vecs fp32 [B,n_in,d];
 * logits fp32 [2,B]; gate_w fp32 [2,B,n_shared+n_task] or NULL. */
typedef struct { int32_t d, n_in, n_shared, n_task, tower_hidden, expert_hidden; } mmoe_home_cfg;
size_t mmoe_home_saved_bytes(const mmoe_home_cfg* cfg, int32_t B, int dtype);
size_t mmoe_home_workspace_bytes(const mmoe_home_cfg* cfg, int32_t B, int dtype);
int mmoe_home_fwd(const mmoe_call* c, const mmoe_home_cfg* cfg, const float* expert_vecs, float* logits, float* gate_w);
int mmoe_home_bwd(const mmoe_call* c, const mmoe_home_cfg* cfg, const float* expert_vecs,
                  const float* dlogits, float* d_expert_vecs);

/* ItemImageExpert.forward after the backbone — model.py:377-385: token mean (pool_cls=0) or
 * CLS token (pool_cls=1), LayerNorm, dropout.  tokens fp32 or T [B,n_tok,d]; params =
 * {norm.weight, norm.bias}; out fp32 [B,d].  Backward gives d_tokens (same dtype as tokens). */
int mmoe_img_pool_fwd(const mmoe_call* c, const void* tokens, int tok_dtype, int32_t n_tok, int32_t d,
                      int32_t pool_cls, float* out, float* stats /*[B,2] fp32, saved*/, float* pooled /*[B,d] saved*/);
int mmoe_img_pool_bwd(const mmoe_call* c, int32_t n_tok, int32_t d, int32_t pool_cls,
                      const float* stats, const float* pooled, const float* dout, void* d_tokens, int tok_dtype);

/* ImageExpertWithProjection.projection_head — model_HoME.py:383-387,397:
 * Linear(d,2d) GELU Linear(2d,proj).  img_vec fp32 [B,d]; out fp32 [B,proj]. */
size_t mmoe_img_proj_saved_bytes(int32_t B, int32_t d, int32_t proj, int dtype);
size_t mmoe_img_proj_workspace_bytes(int32_t B, int32_t d, int32_t proj, int dtype);
int mmoe_img_proj_fwd(const mmoe_call* c, int32_t d, int32_t proj, const float* img_vec, float* out);
int mmoe_img_proj_bwd(const mmoe_call* c, int32_t d, int32_t proj, const float* dout, float* d_img_vec);

/* ---- the callers and data formats either side of the path ------------------------------------------ */

/* HomeExpertWrapper x n + torch.stack — train_HoME.py:100-116 (Dropout(SiLU(BatchNorm1d(x)))) and :350-356 (six wrappers
 * and the stack feeding HOME_MMoE_Complete), in one launch.  x: host array of n device pointers to fp32 [B, d];
 * params = {norm_0.weight, norm_0.bias, ..., norm_{n-1}.weight, norm_{n-1}.bias}; running_mean / running_var: host arrays of n
 * device pointers to the BatchNorm buffers (fp32 [d]), UPDATED in training (momentum, unbiased variance) and used as the
 * statistics in eval; out fp32 [B, n, d]; save_mean / save_rstd fp32 [n, d] for the backward.  Dropout site 0 over the
 * stacked [B, n, d] index.  Backward: dout fp32 [B, n, d]; dx host array of n fp32 [B, d] buffers (NULL entries / NULL = skip);
 * grads as params, accumulated. */
int mmoe_bn_silu_stack_fwd(const mmoe_call* c, int32_t n, int32_t d, const float* const* x, float* out, float* save_mean,
                           float* save_rstd, float* const* running_mean, float* const* running_var, float momentum, float eps);
int mmoe_bn_silu_stack_bwd(const mmoe_call* c, int32_t n, int32_t d, const float* const* x, const float* dout,
                           const float* save_mean, const float* save_rstd, float* const* running_mean,
                           float* const* running_var, float* const* dx, float eps);

/* Two-task nn.BCEWithLogitsLoss(pos_weight) with mean reduction — train.py:189-192, 253-254: value and gradient in one
 * launch over the head's logits [2, B] (good, best).  loss: 1 float on the device, ACCUMULATED into (zero it first);
 * dlogits fp32 [2, B] or NULL; gscale multiplies the gradient. */
int mmoe_bce2_fwd_bwd(const float* logits, const float* y_good, const float* y_best, float pos_weight_good,
                      float pos_weight_best, int32_t B, float* loss, float* dlogits, float gscale, void* stream);

/* calculate_contrastive_loss — train_HoME.py:43-51 (InfoNCE with in-batch negatives, temperature 0.07) for up to 4
 * (anchor, positive) pairs at once (train_HoME.py:362-364 uses three): F.normalize rows, similarity GEMMs on the GEMM engine,
 * cross-entropy against the diagonal.  anchor / positive: host arrays of n_pairs device pointers to fp32 [B, d]
 * (B % 8 == 0, d % 8 == 0); loss: n_pairs floats on the device, ACCUMULATED into.  Backward: dloss = host array of the
 * n_pairs upstream scalars; d_anchor / d_positive: host arrays of fp32 [B, d] device buffers ACCUMULATED into (an input
 * used by several pairs passes the same buffer), NULL entries are skipped. */
size_t mmoe_info_nce_saved_bytes(int32_t n_pairs, int32_t B, int32_t d, int dtype);
size_t mmoe_info_nce_workspace_bytes(int32_t n_pairs, int32_t B, int32_t d, int dtype);
int mmoe_info_nce_fwd(const mmoe_call* c, int32_t n_pairs, int32_t d, const float* const* anchor, const float* const* positive,
                      float temperature, float* loss);
int mmoe_info_nce_bwd(const mmoe_call* c, int32_t n_pairs, int32_t d, const float* const* anchor, const float* const* positive,
                      const float* dloss, float* const* d_anchor, float* const* d_positive);

/* ROC-AUC of n scores on the device — what inference_and_auc.py:150-178 computes with sklearn.metrics.roc_auc_score after
 * copying every batch to the host: sort + rank-sum with average ranks for ties.  scores, labels fp32 [n] (label > 0.5 =
 * positive); out: 1 double on the device (NaN when a class is empty). */
size_t mmoe_auc_workspace_bytes(int64_t n);
int mmoe_auc(const float* scores, const float* labels, int64_t n, void* workspace, size_t workspace_bytes, double* out, void* stream);

/* Patch projection — the Conv2d(3,768,16,stride 16) of HF ViTPatchEmbeddings reached by ItemImageExpert's backbone call
 * (model.py:373-376) as a GEMM over patch-major rows [B*196, C*16*16], the on-disk layout of newpatch.py:102-104 /
 * data4model.py:254-258 (patch.bin).  _u8_to_operand: raw patch bytes -> T rows holding the byte values 0..255 (exact in 16
 * bits; the caller folds /255, mean and std of model.py:172-174 into weight and bias).  _patchify: normalised float images
 * [B,C,H,W] -> the same row layout.  _project: out[rows,N] = operand[rows,K] weight[N,K]^T + bias on the GEMM engine. */
int mmoe_patch_u8_to_operand(const uint8_t* patches, void* out, int64_t rows, int32_t k, int dtype, void* stream);
int mmoe_patchify(const float* images, void* out, int32_t B, int32_t C, int32_t H, int32_t W, int32_t p, int dtype, void* stream);
int mmoe_patch_project(const void* operand, const void* weight, const float* bias, void* out, int out_dtype, int64_t rows,
                       int32_t N, int32_t K, int dtype, void* stream);

/* TextExpert.forward after the encoder — model.py:286-338 (gather the <SENT> hidden states, bucket per sample, pad to S
 * slots, data-derived padding mask, masked mean, LayerNorm, dropout); HoME variant model_HoME.py:328-369 (params = NULL: no
 * final LayerNorm / dropout).  h: encoder hidden states as rows of d values (fp32 or T); src int32 [B, S]: row of h feeding
 * slot (b, s), -1 = empty slot (built on the host from chunk2sample / sent_pos); params = {norm.weight, norm.bias} or NULL.
 * Dropout sites: 0 sentence rows [B,S,d], 1 doc vectors [B,d].  Outputs: sent fp32 [B,S,d], mask uint8 [B,S] (1 = padding),
 * doc fp32 [B,d]; saved for backward: pre_doc fp32 [B,d], stats fp32 [B*(S+1), 2].  Backward: dh fp32 [rows of h, d]
 * ACCUMULATED into (zero first); grads = {d norm.weight, d norm.bias} or NULL. */
int mmoe_sent_gather_fwd(const mmoe_call* c, int32_t S, int32_t d, const void* h, int h_dtype, const int32_t* src,
                         float* sent, uint8_t* mask, float* doc, float* pre_doc, float* stats);
int mmoe_sent_gather_bwd(const mmoe_call* c, int32_t S, int32_t d, const void* h, int h_dtype, const int32_t* src,
                         const uint8_t* mask, const float* pre_doc, const float* stats, const float* d_sent,
                         const float* d_doc, float* dh);

#ifdef __cplusplus
}
#endif
#endif /* MMOE_B200_H_ */
