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

// ---- tcgen05 (tensor-core) engine, qconv_tc.cu ----------------------------------------------------
enum { PASS_FWD = 0, PASS_DGRAD = 1, PASS_WGRAD = 2 };
// true when the implicit-GEMM kernels can serve this shape (layout BHWQC, channel multiples, ...)
bool qconv_tc_supported(const quan_conv_dims& d, int dtype, int layout, int pass);
// bytes of workspace the tc engine needs for a pass (packed weights, split-K partials)
size_t qconv_tc_workspace_bytes(const quan_conv_dims& d, int dtype, int pass);
int qconv_tc_fwd(const void* x, const float* const w[4], const float* bias_r, void* y, const quan_conv_dims& d,
                 int dtype, const float* mix, void* ws, size_t ws_bytes, cudaStream_t st);
int qconv_tc_dgrad(const void* gq, const float* const w[4], void* dx, const quan_conv_dims& d, int dtype, void* ws,
                   size_t ws_bytes, cudaStream_t st);
int qconv_tc_wgrad(const void* gq, const void* x, float* const dw[4], const quan_conv_dims& d, int dtype, void* ws,
                   size_t ws_bytes, cudaStream_t st);

}  // namespace quan
