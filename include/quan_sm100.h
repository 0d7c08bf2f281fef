/*
 * quan_sm100.h — C ABI of libquan_sm100.so: the B200 (sm_100a) implementation of QUAN's
 * quaternion layer stack (QConv2D separable Hamilton convolution, IQBN, QUpsample, Poincare map).
 *
 * This is the drop-in boundary for the hot path named in BASELINE.json.north_star / SURVEY.md §8.
 * Every entry point cites the reference interface it replaces (paths relative to the reference
 * repository root, bryceag11/QUAN_ultralytics).
 *
 * Conventions
 *   - Plain C: raw device pointers, sizes, enums.  No torch / ATen types.
 *   - The CALLER allocates every output and the workspace (SURVEY §8(b) "Ownership"); kernels never
 *     allocate device memory.  All launches go to the `stream` argument (a cudaStream_t passed as void*).
 *   - Return value: 0 = ok, <0 = argument/shape error (QUAN_E_*), >0 = a cudaError_t raised by the launch.
 *     `quan_last_error()` returns a thread-local human-readable message for the last non-zero return.
 *   - Logical activation shape is always the reference's BCHWQ = [B, C, H, W, 4] with C counted in
 *     QUATERNION channels (ultralytics/nn/modules/conv.py:118-126).  Two physical layouts are accepted:
 *       QUAN_LAYOUT_BCHWQ (0): contiguous [B][C][H][W][4]       — the reference's layout (conv.py:441).
 *       QUAN_LAYOUT_BHWQC (1): contiguous [B][H][W][4][C]       — torch.channels_last_3d of the same logical
 *                              tensor; the tensor-core (TMA/tcgen05) kernels require this one.
 *   - dtype: QUAN_F32 (fp32 storage, TF32 tensor-core math, fp32 accumulate) or QUAN_BF16 (bf16 storage,
 *     fp32 accumulate).  Parameters/statistics (gamma, beta, mean, var, bias, master weights, weight grads)
 *     are always fp32.
 *   - mix[16]: the 4x4 mixing matrix M, row-major, out_p = sum_q mix[4*p+q] * S_q.  The reference has two:
 *       M_A  ultralytics/nn/modules/conv.py:493-496  {+1,-1,-1,-1, -1,+1,+1,-1, -1,-1,+1,+1, -1,+1,-1,+1}
 *       M_B  classification/quaternion/qconv.py:606-609 and ultralytics/nn/cuda/quaternion_ops.cu:152-155
 *            {+1,+1,+1,+1, +1,-1,-1,+1, +1,+1,-1,-1, +1,-1,+1,-1}
 */
#ifndef QUAN_SM100_H
#define QUAN_SM100_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QUAN_ABI_VERSION 1

enum quan_dtype  { QUAN_F32 = 0, QUAN_BF16 = 1 };
enum quan_layout { QUAN_LAYOUT_BCHWQ = 0, QUAN_LAYOUT_BHWQC = 1 };
enum quan_act    { QUAN_ACT_NONE = 0, QUAN_ACT_SILU = 1 };

enum quan_status {
  QUAN_OK = 0,
  QUAN_E_ARG = -1,          /* null pointer / bad enum / non-positive size */
  QUAN_E_SHAPE = -2,        /* inconsistent dims (e.g. Ci % groups != 0) */
  QUAN_E_UNSUPPORTED = -3,  /* valid request this build cannot serve (message says why) */
  QUAN_E_WORKSPACE = -4,    /* workspace too small; see quan_*_workspace_bytes */
  QUAN_E_DRIVER = -5        /* CUDA driver entry point (TMA descriptor encode) unavailable */
};

/* Convolution geometry.  Ci/Co are quaternion channels per component (conv.py:118-131);
 * H/W are the INPUT spatial sizes; the output size follows torch's conv2d rule. */
typedef struct quan_conv_dims {
  int32_t B, Ci, Co, H, W;
  int32_t kH, kW, sH, sW, pH, pW, dH, dW, groups;
} quan_conv_dims;

/* Which engine a conv call should use.  AUTO picks tcgen05 when the shape qualifies. */
enum quan_conv_algo { QUAN_ALGO_AUTO = 0, QUAN_ALGO_DIRECT = 1, QUAN_ALGO_TCGEN05 = 2, QUAN_ALGO_DEPTHWISE = 3,
                      QUAN_ALGO_SMALLC = 4 /* 1..8 quaternion channels: one thread per pixel */ };

/* ---- library ------------------------------------------------------------------------------- */
int         quan_version(void);                 /* QUAN_ABI_VERSION */
const char* quan_last_error(void);              /* thread-local message of the last failure */
const char* quan_build_info(void);              /* "sm_100a nvcc 12.9 ..." */
uint64_t    quan_launch_count(void);            /* kernels launched by this library so far (process-wide) */

/* ---- Poincare RGB -> quaternion ------------------------------------------------------------
 * Replaces QConv2D._rgb_to_quaternion (poincare branch), ultralytics/nn/modules/conv.py:378-408
 * (map at :388-397) == classification/quaternion/qconv.py:514-544.
 * rgb: fp32 [B,3,H,W] contiguous.  out: [B,1,H,W,4] (both layouts coincide for C=1), dtype out_dtype.
 * bwd: grad_rgb (fp32 [B,3,H,W]) from grad_out (out_dtype) and rgb. */
int quan_poincare_fwd(const float* rgb, void* out, int32_t B, int32_t H, int32_t W,
                      int out_dtype, void* stream);
int quan_poincare_bwd(const float* rgb, const void* grad_out, float* grad_rgb,
                      int32_t B, int32_t H, int32_t W, int out_dtype, void* stream);

/* ---- IQBN -----------------------------------------------------------------------------------
 * Training branch: IQBN.forward, ultralytics/nn/modules/conv.py:553-571 ==
 * classification/quaternion/qconv.py:378-396.  Eval branch: conv.py:546-552 and the extension entry
 * iqbn_forward, ultralytics/nn/cuda/quaternion_ops_py.cpp:113-128 -> quaternion_ops.cu:8-39,683-731.
 *
 * quan_iqbn_train_stats: one pass over x; per (c,q): mean, biased var (+1e-8 as the reference adds,
 *   conv.py:557), rstd = 1/sqrt(var+eps); updates running_mean/var in place with `momentum`
 *   (conv.py:561-562) when running_mean != NULL.  Writes the [20C] stats table described below.
 * quan_iqbn_apply_fwd: y = act(gamma*(x-mean)*rstd + beta)  (SiLU fused: conv.py:789,809).
 * quan_iqbn_eval_fwd: y = act(gamma*(x-running_mean)/sqrt(running_var+eps)+beta) (no +1e-8).
 * quan_iqbn_bwd_reduce: sums[0..4C) = sum dz, sums[4C..8C) = sum dz*xhat, dz = dy*act'(z).
 * quan_iqbn_bwd_apply: dx = gamma*rstd*(dz - sums_dz/n - xhat*sums_dzxhat/n); if mix_t != NULL the result
 *   is additionally multiplied per quaternion by mix_t (used to emit G = M^T dY for the preceding QConv2D).
 * Workspace: quan_iqbn_workspace_bytes(C) bytes, ZERO-FILLED ONCE by the caller before its first use and private to one stream:
 *   per-row-split fp64 partial sums (plain stores, folded by a second launch: deterministic), followed by [8C] fp64 accumulators
 *   and a ticket that the single-launch reduction of small tensors (<= 40 MB streamed) adds to with L2 atomics; its last block
 *   finishes the statistics and leaves accumulators and ticket zero again (summation order across blocks is then not fixed:
 *   fp64, ~1e-16 relative).  QUAN_IQBN_SMALL_MB=0 keeps every reduction on the two-launch deterministic path. */
size_t quan_iqbn_workspace_bytes(int32_t C);
/* stats: [20*C] floats = mean | var(+1e-8) | rstd (index c*4+q) | scaleT | shiftT (index q*C+c, the BHWQC column order:
 * scale = gamma*rstd, shift = beta - mean*scale) — the table lets the streaming kernels fetch coefficients with 16-byte loads. */
int quan_iqbn_train_stats(const void* x, int32_t B, int32_t C, int32_t H, int32_t W, int dtype, int layout,
                          const float* gamma, const float* beta, float eps, float momentum, float* running_mean,
                          float* running_var, float* stats /* [20*C] */, void* workspace, size_t ws_bytes, void* stream);
/* synced-IQBN building blocks: raw per-(c,q) sums in fp64 {sum, sumsq}[c*4+q], then finalize after the caller
 * all-reduced them across ranks (new work, no reference call site). */
int quan_iqbn_partial_sums(const void* x, int32_t B, int32_t C, int32_t H, int32_t W, int dtype, int layout,
                           double* sums /* [8*C]: sum[4C], sumsq[4C] */, void* workspace, size_t ws_bytes,
                           void* stream);
int quan_iqbn_finalize_stats(const double* sums, double count, int32_t C, const float* gamma, const float* beta, float eps,
                             float momentum, float* running_mean, float* running_var, float* stats /* [20*C] */,
                             void* stream);
/* eval mode: the same [20*C] table from the running statistics (no +1e-8, conv.py:550); then quan_iqbn_apply_fwd. */
int quan_iqbn_eval_stats(const float* gamma, const float* beta, const float* running_mean, const float* running_var,
                         float eps, int32_t C, float* stats /* [20*C] */, void* stream);
int quan_iqbn_apply_fwd(const void* x, void* y, int32_t B, int32_t C, int32_t H, int32_t W, int dtype,
                        int layout, const float* stats, const float* gamma, const float* beta, int act,
                        void* stream);
int quan_iqbn_eval_fwd(const void* x, void* y, int32_t B, int32_t C, int32_t H, int32_t W, int dtype,
                       int layout, const float* gamma, const float* beta, const float* running_mean,
                       const float* running_var, float eps, int act, void* stream);
/* sums: [14*C] doubles = {sum dz, sum dz*xhat}[c*4+q] (8C) followed by the backward coefficient table k1T|k2T|k3T
 * (12C floats, index q*C+c; dx = k1*dz + k2*x + k3).  count > 0: the table is written by the same launch;
 * count <= 0: sums only — all-reduce sums[0..8C) across ranks, then quan_iqbn_bwd_coef with the global count. */
int quan_iqbn_bwd_reduce(const void* dy, const void* x, int32_t B, int32_t C, int32_t H, int32_t W,
                         int dtype, int layout, const float* stats, const float* gamma, const float* beta,
                         int act, double count, double* sums /* [14*C] */, void* workspace, size_t ws_bytes,
                         void* stream);
int quan_iqbn_bwd_coef(double* sums /* [14*C] */, double count, int32_t C, const float* stats, const float* gamma,
                       void* stream);
int quan_iqbn_bwd_apply(const void* dy, const void* x, void* dx, int32_t B, int32_t C, int32_t H, int32_t W,
                        int dtype, int layout, const float* stats, const float* gamma, const float* beta,
                        int act, const double* sums /* [14*C] */, double count, float* dgamma, float* dbeta,
                        const float* mix_t /* NULL or [16] */, void* stream);
/* backward of the eval-mode affine (running stats are constants): dx = dz*gamma*rstd_run. */
int quan_iqbn_eval_bwd(const void* dy, const void* x, void* dx, int32_t B, int32_t C, int32_t H, int32_t W,
                       int dtype, int layout, const float* gamma, const float* beta,
                       const float* running_mean, const float* running_var, float eps, int act,
                       void* stream);

/* ---- QUpsample (nearest) --------------------------------------------------------------------
 * Replaces QUpsample.forward, ultralytics/nn/modules/conv.py:1229-1246 (F.interpolate nearest on each
 * component).  x [B,C,H,W,4] -> y [B,C,H*s,W*s,4]; bwd sums the s*s taps. */
int quan_qupsample_nearest_fwd(const void* x, void* y, int32_t B, int32_t C, int32_t H, int32_t W,
                               int32_t scale, int dtype, int layout, void* stream);
int quan_qupsample_nearest_bwd(const void* dy, void* dx, int32_t B, int32_t C, int32_t H, int32_t W,
                               int32_t scale, int dtype, int layout, void* stream);

/* ---- QuaternionMaxPool ---------------------------------------------------------------------
 * Replaces QuaternionMaxPool.forward, ultralytics/nn/modules/block.py:85-109 (== classification/models/blocks/
 * quaternion_blocks.py:236-260): nn.MaxPool2d(kernel, stride, padding) on each of the four components, stacked back
 * (users: QSPPF block.py:270-302, the Q-ResNet stems quaternion_models.py:193,236).  x [B,C,H,W,4] ->
 * y [B,C,Ho,Wo,4], Ho = (H + 2p - k)/s + 1; padding counts as -inf, the first maximum of the row-major window scan wins
 * ties, NaN propagates (nn.MaxPool2d).  idx: one byte per output element holding the winning tap kh*kW + kw (what the
 * backward routes the gradient by); NULL for inference.  bwd: dx = gather of dy over the windows whose tap points at
 * the element (autograd of the reference; deterministic).  kH*kW <= 255, 2*pad <= kernel. */
int quan_qmaxpool_fwd(const void* x, void* y, uint8_t* idx, int32_t B, int32_t C, int32_t H, int32_t W, int32_t kH,
                      int32_t kW, int32_t sH, int32_t sW, int32_t pH, int32_t pW, int dtype, int layout, void* stream);
int quan_qmaxpool_bwd(const void* dy, const uint8_t* idx, void* dx, int32_t B, int32_t C, int32_t H, int32_t W,
                      int32_t kH, int32_t kW, int32_t sH, int32_t sW, int32_t pH, int32_t pW, int dtype, int layout,
                      void* stream);

/* ---- helpers ---------------------------------------------------------------------------------
 * quan_mix: out_p = sum_q mix[4p+q] in_q per quaternion (qmix_forward/backward kernels,
 *   ultralytics/nn/cuda/quaternion_ops_head.cu:8-95).  n_quat = B*C*H*W.
 * quan_layout_convert: BCHWQ <-> BHWQC (what `.contiguous()` at conv.py:441 does for the reference). */
int quan_mix(const void* in, void* out, int32_t B, int32_t C, int32_t H, int32_t W, int dtype, int layout,
             const float* mix, void* stream);
int quan_layout_convert(const void* src, int src_layout, void* dst, int dst_layout, int32_t B, int32_t C,
                        int32_t H, int32_t W, int dtype, void* stream);

/* ---- QConv2D ---------------------------------------------------------------------------------
 * fwd replaces QConv2D.forward's compute (conv.py:472-499, classification/quaternion/qconv.py:592-612)
 *   and the extension entry qconv_forward (quaternion_ops_py.cpp:48-86 -> quaternion_ops.cu:43-181,735-799):
 *   S_q = conv2d(x_q, w_q) (+ bias_r on q=0), y_p = sum_q mix[4p+q] S_q.
 * bwd replaces QConvFunction.backward (quaternion_autograd_cuda.py:42-67) and qconv_backward
 *   (quaternion_ops_py.cpp:89-111 -> quaternion_ops.cu:532-679): G = M^T dY,
 *   dX_q = conv_transpose(G_q, w_q), dW_q = corr(G_q, x_q), db_r = sum G_r (the autograd-correct bias grad;
 *   the reference extension's `sum dY_r`, quaternion_ops.cu:500, is a documented defect — SURVEY §8(c)).
 * w[4]: fp32 master weights, each [Co, Ci/groups, kH, kW] contiguous (conv.py:133-141); bias_r fp32 [Co] or NULL.
 * dw[4]: fp32 grads, same shape, OVERWRITTEN.  Any of dx / dw / dbias may be NULL to skip that product.
 * Workspace sized by quan_qconv2d_workspace_bytes (0 allowed for the direct engine's fwd); layout
 *   [ G = M^T dY | packed weights (fwd / dgrad) | wgrad split-K partials ].
 * Engines (quan_conv_algo; AUTO picks the first that serves the shape): TCGEN05 — tcgen05/TMEM/TMA implicit GEMM, layout
 *   BHWQC, groups = 1, in its separable form (one GEMM per component, 4x4 mix in the epilogue) for wide layers and its dense
 *   Hamilton form (one GEMM over 4*C channels, mix folded into the packed weights) for narrow ones; DEPTHWISE (groups = C)
 *   and SMALLC (1..8 channels) — streaming kernels that apply the mix in registers; DIRECT — every other shape and the
 *   BCHWQ layout.  The backward of a narrow layer runs dgrad and wgrad concurrently on an internal side stream and joins
 *   before returning (stream order towards the caller is unchanged; works under CUDA graph capture). */
size_t quan_qconv2d_workspace_bytes(const quan_conv_dims* d, int dtype, int layout, int algo);
int quan_qconv2d_fwd(const void* x, const float* const w[4], const float* bias_r, void* y,
                     const quan_conv_dims* d, int dtype, int layout, const float* mix, int algo,
                     void* workspace, size_t ws_bytes, void* stream);
int quan_qconv2d_bwd(const void* dy, const void* x, const float* const w[4], void* dx, float* const dw[4],
                     float* dbias_r, const quan_conv_dims* d, int dtype, int layout, const float* mix,
                     int algo, void* workspace, size_t ws_bytes, void* stream);
/* Forward conv that also emits the IQBN batch statistics of its OUTPUT as per-CTA partial sums (the Conv block's
 * conv -> IQBN order, conv.py:805-809): the tensor-core epilogue adds every output value and its square into shared-memory
 * accumulators and each CTA writes its slot of the IQBN partials buffer (`iqbn_workspace`, quan_iqbn_workspace_bytes(Co)
 * bytes).  *nparts = slots written; 0 means this shape/engine produced none and the caller runs quan_iqbn_train_stats on y.
 * quan_iqbn_finalize_partials folds the slots exactly like the second launch of quan_iqbn_train_stats (mean, var + 1e-8,
 * rstd, coefficient tables, running statistics); count = B*Ho*Wo. */
int quan_qconv2d_fwd_stats(const void* x, const float* const w[4], const float* bias_r, void* y, const quan_conv_dims* d,
                           int dtype, int layout, const float* mix, int algo, void* workspace, size_t ws_bytes,
                           void* iqbn_workspace, size_t iqbn_ws_bytes, int* nparts, void* stream);
int quan_iqbn_finalize_partials(const void* iqbn_workspace, int32_t nparts, double count, int32_t C, const float* gamma,
                                const float* beta, float eps, float momentum, float* running_mean, float* running_var,
                                float* stats /* [20*C] */, void* stream);

/* The Conv block's backward (conv -> IQBN, ultralytics/nn/modules/conv.py:805-809) can skip the G = M^T dY pass:
 * quan_iqbn_bwd_apply(mix_t = M^T) emits G directly and quan_qconv2d_bwd_premixed consumes it.  Only the passes that
 * read G qualify (separable tensor-core form, direct engine, bias gradient); the dense Hamilton form and the depthwise
 * kernels take dY itself — quan_qconv2d_bwd_wants_mixed() returns 1 when every requested pass reads G, else 0. */
int quan_qconv2d_bwd_premixed(const void* g, const void* x, const float* const w[4], void* dx, float* const dw[4],
                              float* dbias_r, const quan_conv_dims* d, int dtype, int layout, const float* mix, int algo,
                              void* workspace, size_t ws_bytes, void* stream);
int quan_qconv2d_bwd_wants_mixed(const quan_conv_dims* d, int dtype, int layout, int algo, int need_dx, int need_dw);
/* reports which engine AUTO would pick for this shape: QUAN_ALGO_DIRECT, QUAN_ALGO_TCGEN05, QUAN_ALGO_DEPTHWISE or QUAN_ALGO_SMALLC */
int quan_qconv2d_pick_algo(const quan_conv_dims* d, int dtype, int layout, int pass /*0 fwd,1 dgrad,2 wgrad*/);

/* ---- the reference's Conv block in one call per direction ---------------------------------------------------------------
 * QConv2D (bias-free) -> IQBN with batch statistics -> act (ultralytics/nn/modules/conv.py:805-809 `Conv.forward`, and its
 * autograd).  Sequences the entry points above (same kernels, same results); exists because the eager step of a narrow layer
 * is bound by host time per call.  y: conv output (kept for the backward), out: activated output, stats: [20*Co];
 * g: scratch tensor shaped like y; sums: [14*Co] doubles; dx / dw may be NULL; workspaces as for the separate calls. */
int quan_conv_block_fwd(const void* x, const float* const w[4], const float* gamma, const float* beta, float* running_mean,
                        float* running_var, void* y, void* out, float* stats, const quan_conv_dims* d, int dtype, int layout,
                        const float* mix, int algo, float eps, float momentum, int act, int epilogue_stats, void* conv_ws,
                        size_t conv_ws_bytes, void* iqbn_ws, size_t iqbn_ws_bytes, void* stream);
int quan_conv_block_bwd(const void* dout, const void* x, const void* y, const float* const w[4], const float* stats,
                        const float* gamma, const float* beta, void* g, void* dx, float* const dw[4], float* dgamma,
                        float* dbeta, double* sums, const quan_conv_dims* d, int dtype, int layout, const float* mix, int algo,
                        int act, void* conv_ws, size_t conv_ws_bytes, void* iqbn_ws, size_t iqbn_ws_bytes, void* stream);

/* Inference form of the block: IQBN uses the running statistics (conv.py:546-552) and, on the tensor-core engine, runs with
 * the activation inside the conv epilogue (one kernel).  stats: [20*Co] scratch for the coefficient table; y_scratch: tensor
 * shaped like the output, only used by the engines without an epilogue hook (may be NULL when quan_qconv2d_pick_algo says
 * QUAN_ALGO_TCGEN05). */
int quan_conv_block_eval_fwd(const void* x, const float* const w[4], const float* gamma, const float* beta,
                             const float* running_mean, const float* running_var, void* out, float* stats, void* y_scratch,
                             const quan_conv_dims* d, int dtype, int layout, const float* mix, int algo, float eps, int act,
                             void* conv_ws, size_t conv_ws_bytes, void* stream);

/* Channel concatenation in BHWQC — `torch.cat(xs, 1)` of the reference's blocks (C2f / C3k2 block.py:350-352, C3 / C3k, QSPPF, QC2PSA,
 * `Concat` conv.py) on tensors whose memory is one row per (pixel, component): source i contributes row_bytes of every destination row,
 * read from rows ld_bytes apart (dense tensors: ld_bytes == row_bytes; channel chunks of a wider tensor: its row pitch).  Sources fill
 * the destination rows left to right; one launch, 16-byte vectors when every pointer / pitch / width allows.  srcs is a HOST array. */
#define QUAN_CAT_MAX 8
typedef struct quan_cat_src {
  const void* ptr;
  int64_t ld_bytes;
  int64_t row_bytes;
} quan_cat_src;
int quan_rows_cat(const quan_cat_src* srcs, int32_t nsrc, void* dst, int64_t dst_ld_bytes, int64_t nrows, void* stream);
/* The same geometry backwards (the gradient of the concatenation, `CatBackward` at those call sites): part i RECEIVES its row_bytes
 * columns of every row of the wide tensor `src` (rows src_ld_bytes apart) — dense gradient slices in one launch instead of strided
 * views that every consumer gathers (and adds through the generic strided kernel) for itself.  parts[i].ptr is written. */
int quan_rows_split(const quan_cat_src* parts, int32_t nparts, const void* src, int64_t src_ld_bytes, int64_t nrows, void* stream);

/* Weight-gradient chains off the critical path.  The backward of a narrow layer already runs its wgrad next to its dgrad (fork / join
 * on a library-owned stream inside quan_qconv2d_bwd / quan_conv_block_bwd).  With a LENT stream the join is deferred: after
 *   quan_bwd_side_stream_set(side, ws, ws_bytes)   side: caller-owned stream; ws: workspace used by nothing else (>= the largest
 *                                                  quan_qconv2d_workspace_bytes of the layers; smaller -> that layer joins as before)
 * every tensor-core wgrad of a narrow layer (and its fold) is forked onto `side` and the call returns without waiting for it — dW is
 * then only valid after
 *   quan_bwd_side_stream_join(stream)              `stream` waits for everything forked so far (call it before the optimizer / the
 *                                                  gradient all-reduce reads dW; once per step is enough).
 * The caller keeps dY (or G), x and dW alive until the join (torch: record_stream on `side`).  quan_bwd_side_stream_set(NULL, NULL, 0)
 * switches back.  Per device (process-wide); works under stream capture (the fork and the join become graph edges). */
int quan_bwd_side_stream_set(void* side_stream, void* wgrad_workspace, size_t ws_bytes);
int quan_bwd_side_stream_join(void* stream);

/* ---- pack plan: all packed weights of a training step in one launch ------------------------------------------------------------
 * The tensor-core engine reads weights in a packed operand layout (per component or, for narrow layers, the dense Hamilton matrix with
 * the mixing matrix folded in), rebuilt from the fp32 masters by a small kernel in front of every forward and every dgrad.  Weights
 * change once per optimizer step, so a step driver can hoist all of them into one launch:
 *   quan_pack_plan_record(1)  start noting every pack the convolutions perform (shape, form, dtype, mix, weight pointers)
 *   ... run one training step ...
 *   quan_pack_plan_record(0)  stop; returns the number of distinct packs
 *   quan_pack_plan_bytes      arena bytes for all of them (*table_bytes: bytes of the device job table)
 *   quan_pack_plan_commit     bind caller-owned device memory (arena 1024-byte aligned) and ACTIVATE: from now on a convolution whose
 *                             pack is in the plan reads its slot of the arena and launches no pack kernel — the caller must run
 *   quan_pack_plan_run        (one launch: every slot rebuilt from the current masters) after each weight update and before the
 *                             first convolution of the step, on a stream ordered before them (inside a captured graph: first node)
 *   quan_pack_plan_release    deactivate (kernels already captured keep their arena pointers; the arena must outlive them).
 * One plan per process; the reference has no counterpart (its extension re-reads the masters in every kernel). */
int quan_pack_plan_record(int on);
size_t quan_pack_plan_bytes(size_t* table_bytes);
int quan_pack_plan_commit(void* arena, size_t arena_bytes, void* table, size_t table_bytes, void* stream);
int quan_pack_plan_run(void* stream);
int quan_pack_plan_release(void);

/* Channel-slice gather: the dense copy of a channel `chunk` / `split` of a BHWQC tensor (what `.contiguous()` at conv.py:441 does for the
 * strided halves C2f / C3k2 / QC2PSA pass on, block.py:350-352): dst row r (row_bytes bytes) = src + r * src_ld_bytes. */
int quan_rows_gather(const void* src, void* dst, int64_t nrows, int32_t row_bytes, int64_t src_ld_bytes, void* stream);

/* ---- QER: quaternion -> real extraction of the detection heads ------------------------------------------
 * Replaces `QER.forward` ultralytics/nn/modules/head.py:40-47 — `x.permute(0, 1, 4, 2, 3).contiguous().view(B, 4C, H, W)` followed by
 * the real 1x1 `nn.Conv2d(4C, N)` (`output_proj`, head.py:36; input channel index c*4 + q) — and its autograd, reading the
 * activation in QUAN_LAYOUT_BHWQC (a pixel is a row of 4C contiguous values, index q*C + c) without the permute copy.
 *   x      [npix][4][C]  activation rows (npix = B*H*W), dense
 *   weight [N][4C]       fp32, the reference's `output_proj.weight` as stored (index c*4 + q);  bias [N] fp32 or NULL
 *   out    npix rows of N values, `out_ld` elements apart (>= N; any alignment): `out_ld` = the channel count of the concatenated
 *          head tensor lets the box / class extractions write straight into `torch.cat((cv2, cv3), 1)` of head.py:143
 *   dy     npix rows of N values, `dy_ld` apart (the gradient arrives as a channel slice of the concatenated tensor's gradient)
 *   dx     like x (NULL: skip);  dweight [N][4C] / dbias [N] fp32 (dweight NULL: skip both; dbias NULL: skip it)
 *   out_writable / dy_readable: columns of every out / dy row (counted from the row pointer; N <= value <= pitch, smaller values mean
 *          N) that belong to this call — row padding.  A ragged width (15 classes in a 16-column slot) then moves as whole 16-byte
 *          vectors: the forward writes zeros into the padding, the backward reads it and ignores what it holds.
 * 4C <= 256 and N <= 64; bf16 needs C % 4 == 0.  bf16 contracts with fp32 accumulation (operands as stored, weight rounded to bf16);
 * fp32 is exact fp32 FMA.  workspace: quan_qer_workspace_bytes (per-CTA partial weight gradients, folded deterministically). */
size_t quan_qer_workspace_bytes(int64_t npix, int32_t C, int32_t N, int dtype);
int quan_qer_fwd(const void* x, const float* weight, const float* bias, void* out, int64_t npix, int32_t C, int32_t N, int64_t out_ld,
                 int32_t out_writable, int dtype, void* stream);
int quan_qer_bwd(const void* dy, int64_t dy_ld, int32_t dy_readable, const void* x, const float* weight, void* dx, float* dweight, float* dbias,
                 int64_t npix, int32_t C, int32_t N, int dtype, void* workspace, size_t ws_bytes, void* stream);

/* ---- QAttention core ---------------------------------------------------------------------------------
 * Replaces the attention arithmetic of `QAttention.forward` ultralytics/nn/modules/block.py:1520-1540 (split of the qkv QConv2D output,
 * per-component `matmul(q, k) * scale` -> `softmax` -> `matmul(attn, v)`) and its autograd, fused: the N x N score matrix never
 * reaches HBM.  qkv: [B, heads*(2*key_dim + head_dim), H, W, 4] in QUAN_LAYOUT_BHWQC with the reference's channel order
 * (q | k | v, each head-major); o: [B, heads*head_dim, H, W, 4]; lse: [B*4*heads*H*W] floats (log2-domain log-sum-exp per
 * query, saved for the backward); N = H*W tokens; scale = key_dim^-0.5.  Built for (key_dim, head_dim) in {(1,2), (2,4), (4,8),
 * (8,16)} — the QUAN yamls give (2,4) at every model scale; anything else returns QUAN_E_UNSUPPORTED. */
int quan_qattention_fwd(const void* qkv, void* o, float* lse, int32_t B, int32_t H, int32_t W, int32_t heads, int32_t key_dim,
                        int32_t head_dim, float scale, int dtype, int layout, void* stream);
int quan_qattention_bwd(const void* qkv, const void* o, const void* d_o, const float* lse, void* dqkv, int32_t B, int32_t H, int32_t W,
                        int32_t heads, int32_t key_dim, int32_t head_dim, float scale, int dtype, int layout, void* stream);

/* ---- rotated task-aligned assigner (the non-differentiable half of v8OBBLoss) -------------------------
 * Replaces `RotatedTaskAlignedAssigner.forward` ultralytics/utils/tal.py:298-330 (+ base class :40-296) as called by
 * `v8OBBLoss.__call__` ultralytics/utils/loss.py:985-993: static shapes, no host synchronisation (graph-capturable).
 * All tensors fp32, dense:  pd_scores [B,A,nc] (sigmoid of the class logits), pd_bboxes [B,A,5] (xywhr, pixels), anc_points [A,2]
 * (pixels), gt_labels [B,n], gt_bboxes [B,n,5] (xywhr, pixels), mask_gt [B,n] (0/1: padding rows are 0).
 * Outputs: target_bboxes [B,A,5] (the assigned box; box 0 of the image for background anchors, as tal.py:212-216 yields),
 * target_scores [B,A,nc], fg_mask [B,A] (0/1 bytes), target_gt_idx [B,A] (int64).  Top-k ties resolve to the lowest anchor index
 * (torch.topk leaves tie order unspecified).  workspace: quan_rotated_tal_workspace_bytes(B, A, n). */
size_t quan_rotated_tal_workspace_bytes(int32_t B, int32_t A, int32_t n);
int quan_rotated_tal_assign(const float* pd_scores, const float* pd_bboxes, const float* anc_points, const float* gt_labels,
                            const float* gt_bboxes, const float* mask_gt, int32_t B, int32_t A, int32_t n, int32_t nc, int32_t topk,
                            float alpha, float beta, float eps, float* target_bboxes, float* target_scores, uint8_t* fg_mask,
                            int64_t* target_gt_idx, void* workspace, size_t ws_bytes, void* stream);

/* ---- OBB loss, differentiable half ----------------------------------------------------------------------
 * Replaces ultralytics/utils/loss.py:941-1033 around the assigner.  The OBB head's training outputs are read where they lie:
 * feats[l] = the level-l output [B, no, H_l, W_l] in channels-last memory ([B][H_l][W_l][no], no = 4*reg_max + nc: the DFL logits of
 * the four sides, then the class logits), pred_angle = [B][1][A] (A = sum H_l*W_l, levels concatenated in order), both of `dtype`.
 * hw = {H_0, W_0, H_1, W_1, H_2, W_2}; strides = the three level strides.
 *   quan_obb_decode       -> pd_scores [B][A][nc] = sigmoid(class logits), pd_bboxes [B][A][5] = decoded xywhr in pixels (fp32):
 *                            the assigner's inputs (loss.py:978-993; decode :1035-1050, tal.py:366-385).
 *   quan_obb_loss_fwd_bwd -> items[4] = (box, cls, dfl, angle) with the gains applied, total = sum(items) * B (loss.py:1027-1033),
 *                            AND d(total)/d(feats[l]), d(total)/d(pred_angle) in the layouts / dtype of the inputs, from the assigner's
 *                            target_bboxes [B][A][5] (pixels), target_scores [B][A][nc], fg_mask [B][A] (bytes).  Terms: BCE with logits
 *                            (:998), (1 - ProbIoU) * weight (:366-368, metrics.py:198-233), DFL (:306-329), quaternion geodesic angle
 *                            (:870-903); all divided by max(sum target_scores, 1).  scratch: 5 doubles.
 * feat_ld: [3] row pitches (elements) of feats[l] AND d_feats[l], or NULL for dense rows of `no`: the fused head (quan_qer_fwd with
 * out_ld) pads the concatenated rows to a multiple of 8 elements so that every row starts on a 16-byte boundary. */
int quan_obb_decode(const void* const feats[3], const void* pred_angle, const int32_t* hw, const float* strides, int32_t B, int32_t nc,
                    int32_t reg_max, const int32_t* feat_ld, float* pd_scores, float* pd_bboxes, int dtype, void* stream);
int quan_obb_loss_fwd_bwd(const void* const feats[3], const void* pred_angle, const int32_t* hw, const float* strides, int32_t B, int32_t nc,
                          int32_t reg_max, const float* target_bboxes, const float* target_scores, const uint8_t* fg_mask, float box_gain,
                          float cls_gain, float dfl_gain, float angle_gain, const int32_t* feat_ld, void* const d_feats[3], void* d_angle,
                          double* scratch, float* items, float* total, int dtype, void* stream);

/* ---- optimizer step -------------------------------------------------------------------------------
 * Replaces `BaseTrainer.optimizer_step` ultralytics/engine/trainer.py:586-594 — torch.nn.utils.clip_grad_norm_(max_norm) followed by
 * torch.optim.SGD(momentum, nesterov, per-group lr / weight_decay).step() and zero_grad(), built by trainer.py:766-806 — and
 * classification/utils/training.py:78-79, for ALL parameter tensors in two launches.  The caller describes the tensors with a device
 * table of chunks (<= 8192 consecutive fp32 elements of one tensor each; one thread block per chunk):
 *   p / g     device pointers to the chunk's parameters and gradients (fp32, dense)
 *   buf_off   element offset of the chunk in the caller-owned momentum buffer (zero-initialised: the first step then equals torch's)
 *   group     index of the parameter group (selects lr / weight_decay)
 * hyper (DEVICE, fp32): [lr_0, wd_0, ..., lr_{G-1}, wd_{G-1}, momentum, max_norm (<= 0: no clipping), nesterov (0/1), dampening] —
 * device-resident so that a captured CUDA graph follows schedules.  partial: [nchunks] doubles of scratch.  total_norm_out (device,
 * may be NULL) receives the pre-clip global gradient norm (what clip_grad_norm_ returns).  zero_grad != 0 clears the gradients
 * (optimizer.zero_grad()); otherwise they are left scaled by the clip coefficient, as clip_grad_norm_ leaves them. */
typedef struct quan_opt_chunk {
  void* p;
  void* g;
  int64_t buf_off;
  int32_t n;
  int32_t group;
} quan_opt_chunk;
int quan_sgd_clip_step(const void* chunk_table, int nchunks, float* momentum_buf, const float* hyper, int ngroups, double* partial,
                       float* total_norm_out, int zero_grad, void* stream);
/* ModelEMA.update ultralytics/utils/torch_utils.py:514-525: ema = d * ema + (1 - d) * value over every chunk (p = the live value,
 * buf_off = offset in ema_buf; g / group unused); decay: one DEVICE float (the reference ramps it with the update count). */
int quan_ema_update(const void* chunk_table, int nchunks, float* ema_buf, const float* decay, void* stream);

/* Optional per-kernel device timing for benchmarks (no reference counterpart): while enabled, every kernel the library
 * launches outside stream capture is bracketed by a CUDA-event pair on its own stream.  `enable(1)` clears earlier
 * records.  `report` synchronises the recorded events and writes one "kernel_name launches total_ms" line per kernel
 * into buf (NUL-terminated, truncated to cap); returns the bytes a complete report needs. */
int quan_kernel_timing_enable(int on);
size_t quan_kernel_timing_report(char* buf, size_t cap);

#ifdef __cplusplus
}
#endif
#endif /* QUAN_SM100_H */
