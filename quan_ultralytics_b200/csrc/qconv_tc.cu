// qconv_tc.cu — tcgen05/TMEM/TMA implicit-GEMM engine for QConv2D (placeholder: engine reports "unsupported"
// until the kernels land; qconv_api.cu then routes everything to the direct engine).
#include "qconv_internal.cuh"
namespace quan {
bool qconv_tc_supported(const quan_conv_dims&, int, int, int) { return false; }
size_t qconv_tc_workspace_bytes(const quan_conv_dims&, int, int) { return 0; }
int qconv_tc_fwd(const void*, const float* const*, const float*, void*, const quan_conv_dims&, int, const float*, void*, size_t, cudaStream_t) { set_error("tcgen05 engine not built"); return QUAN_E_UNSUPPORTED; }
int qconv_tc_dgrad(const void*, const float* const*, void*, const quan_conv_dims&, int, void*, size_t, cudaStream_t) { set_error("tcgen05 engine not built"); return QUAN_E_UNSUPPORTED; }
int qconv_tc_wgrad(const void*, const void*, float* const*, const quan_conv_dims&, int, void*, size_t, cudaStream_t) { set_error("tcgen05 engine not built"); return QUAN_E_UNSUPPORTED; }
}
