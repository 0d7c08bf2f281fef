// block_api.cu — the reference's `Conv` block (QConv2D -> IQBN with batch statistics -> act, ultralytics/nn/modules/
// conv.py:805-809) as ONE C call per direction.  Same kernels as the separate entry points (this file only sequences
// them); what it removes is host time: the eager step of a narrow layer spends more time in Python / ctypes glue
// (~360 us per block step, tools/host_profile.py) than on the GPU, and four calls per direction become one.
#include "qconv_internal.cuh"

using namespace quan;

extern "C" {

int quan_conv_block_fwd(const void* x, const float* const w[4], const float* gamma, const float* beta, float* running_mean,
                        float* running_var, void* y, void* out, float* stats, const quan_conv_dims* d, int dtype, int layout,
                        const float* mix, int algo, float eps, float momentum, int act, int epilogue_stats, void* conv_ws,
                        size_t conv_ws_bytes, void* iqbn_ws, size_t iqbn_ws_bytes, void* stream) {
  QUAN_REQUIRE(d != nullptr && y != nullptr && out != nullptr && stats != nullptr, QUAN_E_ARG, "conv_block_fwd: null pointer");
  QUAN_REQUIRE(iqbn_ws != nullptr && iqbn_ws_bytes >= quan_iqbn_workspace_bytes(d->Co), QUAN_E_WORKSPACE,
               "conv_block_fwd: IQBN workspace needs %zu bytes", quan_iqbn_workspace_bytes(d->Co));
  int nparts = -1;       // opt in: the narrow-layer epilogue may ADD its statistics to the workspace's accumulators (answers nparts < 0)
  int rc = epilogue_stats
               ? quan_qconv2d_fwd_stats(x, w, nullptr, y, d, dtype, layout, mix, algo, conv_ws, conv_ws_bytes, iqbn_ws,
                                        iqbn_ws_bytes, &nparts, stream)
               : quan_qconv2d_fwd(x, w, nullptr, y, d, dtype, layout, mix, algo, conv_ws, conv_ws_bytes, stream);
  if (rc) return rc;
  const int Ho = conv_out(d->H, d->kH, d->sH, d->pH, d->dH), Wo = conv_out(d->W, d->kW, d->sW, d->pW, d->dW);
  if (nparts < 0)        // raw sums are in L2: statistics + table + running buffers + normalisation + act in one launch
    return iqbn_apply_fwd_from_acc(y, out, d->B, d->Co, Ho, Wo, dtype, layout, iqbn_ws, (double)d->B * Ho * Wo, gamma, beta, eps, momentum,
                                   running_mean, running_var, stats, act, stream);
  if (nparts > 0)
    rc = quan_iqbn_finalize_partials(iqbn_ws, nparts, (double)d->B * Ho * Wo, d->Co, gamma, beta, eps, momentum, running_mean,
                                     running_var, stats, stream);
  else
    rc = quan_iqbn_train_stats(y, d->B, d->Co, Ho, Wo, dtype, layout, gamma, beta, eps, momentum, running_mean, running_var,
                               stats, iqbn_ws, iqbn_ws_bytes, stream);
  if (rc) return rc;
  return quan_iqbn_apply_fwd(y, out, d->B, d->Co, Ho, Wo, dtype, layout, stats, gamma, beta, act, stream);
}

// g: scratch tensor shaped like y (receives the IQBN input gradient, or G = M^T dY when the conv backward reads G)
int quan_conv_block_bwd(const void* dout, const void* x, const void* y, const float* const w[4], const float* stats,
                        const float* gamma, const float* beta, void* g, void* dx, float* const dw[4], float* dgamma,
                        float* dbeta, double* sums, const quan_conv_dims* d, int dtype, int layout, const float* mix, int algo,
                        int act, void* conv_ws, size_t conv_ws_bytes, void* iqbn_ws, size_t iqbn_ws_bytes, void* stream) {
  QUAN_REQUIRE(d != nullptr && dout != nullptr && y != nullptr && g != nullptr && sums != nullptr && mix != nullptr, QUAN_E_ARG,
               "conv_block_bwd: null pointer");
  const int Ho = conv_out(d->H, d->kH, d->sH, d->pH, d->dH), Wo = conv_out(d->W, d->kW, d->sW, d->pW, d->dW);
  const double count = (double)d->B * Ho * Wo;
  int rc = quan_iqbn_bwd_reduce(dout, y, d->B, d->Co, Ho, Wo, dtype, layout, stats, gamma, beta, act, count, sums, iqbn_ws,
                                iqbn_ws_bytes, stream);
  if (rc) return rc;
  const bool any = dx != nullptr || dw != nullptr;
  const bool mixed = any && quan_qconv2d_bwd_wants_mixed(d, dtype, layout, algo, dx != nullptr, dw != nullptr) == 1;
  float mix_t[16];
  for (int p = 0; p < 4; ++p)
    for (int q = 0; q < 4; ++q) mix_t[q * 4 + p] = mix[p * 4 + q];
  rc = quan_iqbn_bwd_apply(dout, y, g, d->B, d->Co, Ho, Wo, dtype, layout, stats, gamma, beta, act, sums, count, dgamma, dbeta,
                           mixed ? mix_t : nullptr, stream);
  if (rc || !any) return rc;
  return mixed ? quan_qconv2d_bwd_premixed(g, x, w, dx, dw, nullptr, d, dtype, layout, mix, algo, conv_ws, conv_ws_bytes, stream)
               : quan_qconv2d_bwd(g, x, w, dx, dw, nullptr, d, dtype, layout, mix, algo, conv_ws, conv_ws_bytes, stream);
}

// Inference: QConv2D -> IQBN with running statistics -> act (`Conv.forward` in eval mode, conv.py:546-552 + :805-809).  On
// the tensor-core engine the normalisation and the activation run in the conv epilogue (one kernel after the tiny
// coefficient-table kernel and the weight packing); other engines write the conv output to `y_scratch` and apply.
int quan_conv_block_eval_fwd(const void* x, const float* const w[4], const float* gamma, const float* beta,
                             const float* running_mean, const float* running_var, void* out, float* stats, void* y_scratch,
                             const quan_conv_dims* d, int dtype, int layout, const float* mix, int algo, float eps, int act,
                             void* conv_ws, size_t conv_ws_bytes, void* stream) {
  QUAN_REQUIRE(d != nullptr && x != nullptr && out != nullptr && stats != nullptr && mix != nullptr && w != nullptr, QUAN_E_ARG,
               "conv_block_eval_fwd: null pointer");
  int rc = quan_iqbn_eval_stats(gamma, beta, running_mean, running_var, eps, d->Co, stats, stream);
  if (rc) return rc;
  const int Ho = conv_out(d->H, d->kH, d->sH, d->pH, d->dH), Wo = conv_out(d->W, d->kW, d->sW, d->pW, d->dW);
  const int picked = algo == QUAN_ALGO_AUTO ? quan_qconv2d_pick_algo(d, dtype, layout, PASS_FWD) : algo;
  if (picked == QUAN_ALGO_TCGEN05 && qconv_tc_supported(*d, dtype, layout, PASS_FWD) && Ho > 0 && Wo > 0) {
    const size_t need = qconv_tc_workspace_bytes(*d, dtype, layout, PASS_FWD);
    QUAN_REQUIRE(conv_ws != nullptr && conv_ws_bytes >= need, QUAN_E_WORKSPACE,
                 "conv_block_eval_fwd: tcgen05 engine needs %zu workspace bytes, got %zu", need, conv_ws_bytes);
    return qconv_tc_fwd(x, w, nullptr, out, *d, dtype, qconv_tc_mode(*d, dtype, layout, PASS_FWD), mix, conv_ws, conv_ws_bytes,
                        (cudaStream_t)stream, nullptr, nullptr, stats + 12 * (size_t)d->Co, stats + 16 * (size_t)d->Co, act);
  }
  QUAN_REQUIRE(y_scratch != nullptr, QUAN_E_ARG, "conv_block_eval_fwd: this shape needs the y scratch tensor");
  rc = quan_qconv2d_fwd(x, w, nullptr, y_scratch, d, dtype, layout, mix, algo, conv_ws, conv_ws_bytes, stream);
  if (rc) return rc;
  return quan_iqbn_apply_fwd(y_scratch, out, d->B, d->Co, Ho, Wo, dtype, layout, stats, gamma, beta, act, stream);
}

}  // extern "C"
