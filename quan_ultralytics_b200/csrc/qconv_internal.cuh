// qconv_internal.cuh — engine entry points shared between qconv_api.cu, qconv_direct.cu and qconv_tc.cu.
#pragma once
#include "common.cuh"

namespace quan {

// ---- direct (CUDA-core) engine, qconv_direct.cu -------------------------------------------------
int qconv_fwd_direct_launch(const void* x, const float* const w[4], const float* bias_r, void* y,
                            const quan_conv_dims& d, int dtype, int layout, const float* mix, cudaStream_t st);
int qconv_dgrad_direct_launch(const void* gq, const float* const w[4], void* dx, const quan_conv_dims& d, int dtype,
                              int layout, cudaStream_t st);
int qconv_wgrad_direct_launch(const void* gq, const void* x, float* const dw[4], const quan_conv_dims& d, int dtype,
                              int layout, cudaStream_t st);
int qconv_bias_grad_launch(const void* gq, float* db, const quan_conv_dims& d, int dtype, int layout, cudaStream_t st);

// ---- depthwise streaming kernels (BHWQC, groups == C), qconv_dw.cu: take dY itself (M^T applied on load) ----------
bool qconv_dw_supported(const quan_conv_dims& d, int dtype, int layout, int pass);
int qconv_dw_fwd(const void* x, const float* const w[4], const float* bias_r, void* y, const quan_conv_dims& d, int dtype,
                 const float* mix, cudaStream_t st);
int qconv_dw_dgrad(const void* dy, const float* const w[4], void* dx, const quan_conv_dims& d, int dtype, const float* mix,
                   cudaStream_t st);
int qconv_dw_wgrad(const void* dy, const void* x, float* const dw[4], const quan_conv_dims& d, int dtype, const float* mix,
                   cudaStream_t st);

// ---- small-channel streaming kernels (BHWQC, 1..8 quaternion channels), qconv_small.cu: take dY itself -----------
bool qconv_small_supported(const quan_conv_dims& d, int dtype, int layout, int pass);
int qconv_small_fwd(const void* x, const float* const w[4], const float* bias_r, void* y, const quan_conv_dims& d, int dtype,
                    const float* mix, cudaStream_t st);
int qconv_small_dgrad(const void* dy, const float* const w[4], void* dx, const quan_conv_dims& d, int dtype, const float* mix,
                      cudaStream_t st);
int qconv_small_wgrad(const void* dy, const void* x, float* const dw[4], const quan_conv_dims& d, int dtype, const float* mix,
                      cudaStream_t st);

// ---- tcgen05 (tensor-core) engine, qconv_tc.cu ----------------------------------------------------
enum { PASS_FWD = 0, PASS_DGRAD = 1, PASS_WGRAD = 2 };
// how the implicit-GEMM kernels serve a shape: not at all, one GEMM per quaternion component (mix in the epilogue),
// or the dense Hamilton form (one GEMM over all 4*C_q channels, mix folded into the weights) used for narrow layers
enum { TC_NONE = 0, TC_SEPARABLE = 1, TC_DENSE = 2 };
int qconv_tc_mode(const quan_conv_dims& d, int dtype, int layout, int pass);
bool qconv_tc_supported(const quan_conv_dims& d, int dtype, int layout, int pass);
// bytes of workspace the tc engine needs for a pass (packed weights, split-K partials)
size_t qconv_tc_workspace_bytes(const quan_conv_dims& d, int dtype, int layout, int pass);
// `mode` is qconv_tc_mode()'s answer for the pass.  dgrad / wgrad take G = M^T dY in the separable form and dY itself
// in the dense form (`mix` is always the forward mixing matrix).
// stat_part (or NULL): the IQBN partials buffer — every CTA writes the partial sums / sums of squares of the outputs it
// produced to its own slot [2][4*C_o]; *stat_nparts = slots written (0: this launch produced none, run the stats kernel)
int qconv_tc_fwd(const void* x, const float* const w[4], const float* bias_r, void* y, const quan_conv_dims& d,
                 int dtype, int mode, const float* mix, void* ws, size_t ws_bytes, cudaStream_t st,
                 double* stat_part = nullptr, int* stat_nparts = nullptr,
                 // eval-mode IQBN + activation in the epilogue: y = act(y * scale + shift), tables [4][C_o] (component, channel)
                 const float* post_scale = nullptr, const float* post_shift = nullptr, int post_act = 0);
int qconv_tc_dgrad(const void* g, const float* const w[4], void* dx, const quan_conv_dims& d, int dtype, int mode,
                   const float* mix, void* ws, size_t ws_bytes, cudaStream_t st);
int qconv_tc_wgrad(const void* g, const void* x, float* const dw[4], const quan_conv_dims& d, int dtype, int mode,
                   const float* mix, void* ws, size_t ws_bytes, cudaStream_t st);

// iqbn.cu: IQBN apply from the raw sums the conv epilogue accumulated (see ApplyArgs.fsums)
int iqbn_apply_fwd_from_acc(const void* x, void* y, int B, int C, int H, int W, int dtype, int layout, void* iqbn_ws, double count,
                            const float* gamma, const float* beta, float eps, float momentum, float* running_mean, float* running_var,
                            float* stats, int act, void* stream);

}  // namespace quan
