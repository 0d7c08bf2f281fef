// optim.cu — the optimizer half of the training step: gradient-norm clipping + SGD(momentum, nesterov,
// weight decay) for ALL parameter tensors of a model in two launches.
//
// Reference: ultralytics/engine/trainer.py:586-594 `optimizer_step` = clip_grad_norm_(max_norm=10) -> optimizer.step() ->
// zero_grad(), with the optimizer of trainer.py:766-806 (torch.optim.SGD, nesterov, three parameter groups);
// classification/utils/training.py:78-79 (clip 1.0 + SGD).  torch runs this as a Python loop / foreach lists over ~540 tensors
// (YOLO11n: 2.6 ms of device time and ~5 ms of host time per step, tools/yolo_step_profile.py); here it is HBM-bound work over
// 0.7-6 M floats: a chunk table (<= 8192 elements of one tensor per thread block) makes every tensor addressable from one grid.
//
//   kernel 1  partial[c] = sum g^2 over chunk c                                  (double)
//   kernel 2  total = sqrt(sum_c partial[c]);  k = min(1, max_norm / (total + 1e-6))      (clip_grad_norm_ semantics)
//             g' = k g (+ wd p);  buf = mu buf + g';  p -= lr (nesterov ? g' + mu buf : buf);   g = 0 when zero_grad
//
// Hyper-parameters live in a small DEVICE array so that a captured CUDA graph follows learning-rate schedules.
#include "common.cuh"

namespace quan {

constexpr int OPT_THREADS = 256;

__device__ __forceinline__ double block_sum(double v, double* sh) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) sh[w] = v;
  __syncthreads();
  double r = 0.0;
  if (w == 0) {
    r = l < (OPT_THREADS >> 5) ? sh[l] : 0.0;
    for (int o = 4; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
    if (l == 0) sh[0] = r;
  }
  __syncthreads();
  r = sh[0];
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(OPT_THREADS) sgd_sqnorm_kernel(const quan_opt_chunk* __restrict__ tab, double* __restrict__ partial) {
  pdl_prologue();
  __shared__ double sh[OPT_THREADS / 32];
  const quan_opt_chunk c = tab[blockIdx.x];
  const float* g = reinterpret_cast<const float*>(c.g);
  double acc = 0.0;
  if ((reinterpret_cast<uintptr_t>(g) & 15) == 0) {
    const int n4 = c.n >> 2;
    const float4* g4 = reinterpret_cast<const float4*>(g);
    for (int i = threadIdx.x; i < n4; i += OPT_THREADS) {
      const float4 v = g4[i];
      acc += (double)(v.x * v.x + v.y * v.y) + (double)(v.z * v.z + v.w * v.w);
    }
    for (int i = (n4 << 2) + threadIdx.x; i < c.n; i += OPT_THREADS) acc += (double)g[i] * g[i];
  } else {
    for (int i = threadIdx.x; i < c.n; i += OPT_THREADS) acc += (double)g[i] * g[i];
  }
  acc = block_sum(acc, sh);
  if (threadIdx.x == 0) partial[blockIdx.x] = acc;
}

// hyper layout: [lr_0, wd_0, lr_1, wd_1, ..., lr_{G-1}, wd_{G-1}, momentum, max_norm, nesterov(0/1), dampening]
__global__ void __launch_bounds__(OPT_THREADS) sgd_update_kernel(const quan_opt_chunk* __restrict__ tab, int nchunks,
                                                                 float* __restrict__ mom, const float* __restrict__ hyper, int ngroups,
                                                                 const double* __restrict__ partial, float* __restrict__ total_norm_out,
                                                                 int zero_grad) {
  pdl_prologue();
  __shared__ double sh[OPT_THREADS / 32];
  double s = 0.0;
  for (int i = threadIdx.x; i < nchunks; i += OPT_THREADS) s += partial[i];
  s = block_sum(s, sh);
  const float total = (float)sqrt(s);
  const float mu = hyper[2 * ngroups], max_norm = hyper[2 * ngroups + 1], damp = hyper[2 * ngroups + 3];
  const bool nesterov = hyper[2 * ngroups + 2] != 0.f;
  float k = 1.f;
  if (max_norm > 0.f) {
    k = max_norm / (total + 1e-6f);
    k = k < 1.f ? k : 1.f;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && total_norm_out != nullptr) *total_norm_out = total;
  const quan_opt_chunk c = tab[blockIdx.x];
  const float lr = hyper[2 * c.group], wd = hyper[2 * c.group + 1];
  float* p = reinterpret_cast<float*>(c.p);
  float* g = reinterpret_cast<float*>(c.g);
  float* b = mom + c.buf_off;
  for (int i = threadIdx.x; i < c.n; i += OPT_THREADS) {
    float gi = g[i] * k;
    const float pi = p[i];
    if (wd != 0.f) gi = fmaf(wd, pi, gi);
    const float bi = fmaf(mu, b[i], (1.f - damp) * gi);
    b[i] = bi;
    const float step = nesterov ? fmaf(mu, bi, gi) : bi;
    p[i] = fmaf(-lr, step, pi);
    g[i] = zero_grad ? 0.f : g[i] * k;        // clip_grad_norm_ scales .grad in place
  }
}

// ---- EMA of the parameters / float buffers (ultralytics/utils/torch_utils.py:514-525 ModelEMA.update): e = d e + (1 - d) v
__global__ void __launch_bounds__(OPT_THREADS) ema_update_kernel(const quan_opt_chunk* __restrict__ tab, float* __restrict__ ema,
                                                                 const float* __restrict__ decay) {
  pdl_prologue();
  const quan_opt_chunk c = tab[blockIdx.x];
  const float d = *decay;
  const float* v = reinterpret_cast<const float*>(c.p);
  float* e = ema + c.buf_off;
  for (int i = threadIdx.x; i < c.n; i += OPT_THREADS) e[i] = fmaf(d, e[i], (1.f - d) * v[i]);
}

}  // namespace quan

extern "C" {

int quan_sgd_clip_step(const void* chunk_table, int nchunks, float* momentum_buf, const float* hyper, int ngroups, double* partial,
                       float* total_norm_out, int zero_grad, void* stream) {
  using namespace quan;
  QUAN_REQUIRE(nchunks >= 0 && ngroups >= 1 && ngroups <= 16, QUAN_E_ARG, "quan_sgd_clip_step: nchunks=%d ngroups=%d", nchunks, ngroups);
  if (nchunks == 0) return QUAN_OK;
  QUAN_REQUIRE(chunk_table && momentum_buf && hyper && partial, QUAN_E_ARG, "quan_sgd_clip_step: null pointer");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const quan_opt_chunk* tab = reinterpret_cast<const quan_opt_chunk*>(chunk_table);
  QUAN_TIMED(st);
  QUAN_LAUNCH((sgd_sqnorm_kernel), nchunks, OPT_THREADS, 0, st, tab, partial);
  QUAN_CHECK_LAUNCH("sgd_sqnorm");
  QUAN_TIMED(st);
  QUAN_LAUNCH((sgd_update_kernel), nchunks, OPT_THREADS, 0, st, tab, nchunks, momentum_buf, hyper, ngroups, partial, total_norm_out, zero_grad);
  QUAN_CHECK_LAUNCH("sgd_update");
  return QUAN_OK;
}

int quan_ema_update(const void* chunk_table, int nchunks, float* ema_buf, const float* decay, void* stream) {
  using namespace quan;
  QUAN_REQUIRE(nchunks >= 0, QUAN_E_ARG, "quan_ema_update: nchunks=%d", nchunks);
  if (nchunks == 0) return QUAN_OK;
  QUAN_REQUIRE(chunk_table && ema_buf && decay, QUAN_E_ARG, "quan_ema_update: null pointer");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  QUAN_TIMED(st);
  QUAN_LAUNCH((ema_update_kernel), nchunks, OPT_THREADS, 0, st, reinterpret_cast<const quan_opt_chunk*>(chunk_table), ema_buf, decay);
  QUAN_CHECK_LAUNCH("ema_update");
  return QUAN_OK;
}

}  // extern "C"
