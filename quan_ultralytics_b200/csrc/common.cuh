// common.cuh — shared helpers for libquan_sm100.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <stdlib.h>
#include "../../include/quan_sm100.h"

#define QUAN_NUM_SMS 148  // B200: 2 dies x 74 SMs; grids are sized in multiples of this.
#define QUAN_IQBN_MAX_PARTS (4 * QUAN_NUM_SMS)   // slots of the IQBN partials buffer; the [8C] fp64 accumulators follow them
#define QUAN_STR_(x) #x
#define QUAN_STR(x) QUAN_STR_(x)

namespace quan {

// ---- error plumbing ---------------------------------------------------------------------------
void set_error(const char* fmt, ...);   // defined in api.cu (thread-local buffer)

#define QUAN_REQUIRE(cond, code, ...)                    \
  do {                                                   \
    if (!(cond)) {                                       \
      ::quan::set_error(__VA_ARGS__);                    \
      return (code);                                     \
    }                                                    \
  } while (0)

// Always check the launch (the reference never does: quaternion_ops.cu:777-799).
void count_launch();                    // defined in api.cu (process-wide atomic counter)

// optional per-kernel timing (api.cu): QUAN_TIMED(st) before a launch, QUAN_CHECK_LAUNCH(name) after it
void timing_begin(cudaStream_t st);
void timing_end(const char* name);
void timing_work(const char* prefix, const char* skip, double bytes, double flops);   // algorithmic work of the next matching launch
#define QUAN_TIMED(st) ::quan::timing_begin(st)

#define QUAN_CHECK_LAUNCH(name)                                                     \
  do {                                                                              \
    ::quan::timing_end(name);                                                       \
    ::quan::count_launch();                                                         \
    cudaError_t e__ = cudaGetLastError();                                           \
    if (e__ != cudaSuccess) {                                                       \
      ::quan::set_error("%s: launch failed: %s", name, cudaGetErrorString(e__));    \
      return (int)e__;                                                              \
    }                                                                               \
  } while (0)

#define QUAN_CUDA(call)                                                             \
  do {                                                                              \
    cudaError_t e__ = (call);                                                       \
    if (e__ != cudaSuccess) {                                                       \
      ::quan::set_error("%s failed: %s", #call, cudaGetErrorString(e__));           \
      return (int)e__;                                                              \
    }                                                                               \
  } while (0)

// ---- dtype traits -----------------------------------------------------------------------------
template <typename T> struct VecTraits;
template <> struct VecTraits<float> {
  static constexpr int kMaxVec = 4;  // 16 B
};
template <> struct VecTraits<__nv_bfloat16> {
  static constexpr int kMaxVec = 8;  // 16 B
};

__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) {
  return __float2bfloat16_rn(v);
}

// A V-wide vector of T moved with one (<=16 B) load/store.
template <typename T, int V> struct alignas(sizeof(T) * V) Vec {
  T v[V];
};

template <typename T, int V>
__device__ __forceinline__ void load_vec(const T* __restrict__ p, float (&out)[V]) {
  Vec<T, V> t = *reinterpret_cast<const Vec<T, V>*>(p);
#pragma unroll
  for (int i = 0; i < V; ++i) out[i] = to_f32(t.v[i]);
}
template <typename T, int V>
__device__ __forceinline__ void store_vec(T* __restrict__ p, const float (&in)[V]) {
  Vec<T, V> t;
#pragma unroll
  for (int i = 0; i < V; ++i) t.v[i] = from_f32<T>(in[i]);
  *reinterpret_cast<Vec<T, V>*>(p) = t;
}

// ---- activation -------------------------------------------------------------------------------
__device__ __forceinline__ float sigmoid_f(float z) { return __fdividef(1.0f, 1.0f + __expf(-z)); }
// One MUFU op instead of two (ex2 + rcp): sigmoid(z) = 0.5 tanh(z/2) + 0.5 with tanh.approx.f32 (relative error 2^-11).
// Used by the bf16 kernels only — their outputs round to 2^-9 — where the SiLU kernels would otherwise sit at ~70% of
// the SM's 16 MUFU/clk while streaming at HBM rate.
__device__ __forceinline__ float sigmoid_fast(float z) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * z));
  return fmaf(0.5f, t, 0.5f);
}
template <int ACT, bool FAST = false> __device__ __forceinline__ float act_fwd(float z) {
  if constexpr (ACT == QUAN_ACT_SILU) return z * (FAST ? sigmoid_fast(z) : sigmoid_f(z));
  return z;
}
// d act(z) / dz
template <int ACT, bool FAST = false> __device__ __forceinline__ float act_grad(float z) {
  if constexpr (ACT == QUAN_ACT_SILU) {
    float s = FAST ? sigmoid_fast(z) : sigmoid_f(z);
    return s * (1.0f + z * (1.0f - s));
  }
  return 1.0f;
}

// ---- packed fp32x2 arithmetic (sm_100: one instruction for two fp32 lanes in a 64-bit register pair) -------------------
// The TMA-fed IQBN kernels are issue-bound (ncu: 62-68 % issue-active, memory stalls gone), so halving the FP32
// instruction count per element is what moves them.
__device__ __forceinline__ uint64_t f2_pack(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void f2_unpack(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t f2_add(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t f2_mul(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t f2_fma(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
// two consecutive elements of a vector as an fp32 pair (bf16: one 32-bit word -> shift / mask; fp32: the two floats)
__device__ __forceinline__ uint64_t f2_from(const __nv_bfloat16* p) {
  const uint32_t w = *reinterpret_cast<const uint32_t*>(p);
  return f2_pack(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
}
__device__ __forceinline__ uint64_t f2_from(const float* p) { return f2_pack(p[0], p[1]); }

// ---- small host helpers -----------------------------------------------------------------------
static inline int conv_out(int in, int k, int s, int p, int d) { return (in + 2 * p - d * (k - 1) - 1) / s + 1; }
static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int largest_pow2_divisor(int c, int cap) {
  int v = 1;
  while (v < cap && (c % (v * 2)) == 0) v *= 2;
  return v;
}
static inline int grid_for(int64_t work_items, int threads, int blocks_per_sm) {
  int64_t need = ceil_div64(work_items, threads);
  int64_t cap = (int64_t)QUAN_NUM_SMS * blocks_per_sm;
  if (need < 1) need = 1;
  return (int)(need < cap ? need : cap);
}

// Once-per-(host thread, device) latch.  cudaFuncSetAttribute and occupancy answers belong to the device that is current when they
// are made; a process driving several GPUs (or a model on cuda:1) must repeat them per device.  All GPUs of a box are the same B200,
// so cached VALUES (occupancy, SM count) are shared; only the latch is per device.
struct DeviceOnce {
  uint64_t mask = 0;
  bool first() {
    int d = 0;
    cudaGetDevice(&d);
    const uint64_t b = 1ull << (d & 63);
    if (mask & b) return false;
    mask |= b;
    return true;
  }
};

// ---- programmatic dependent launch (PDL) --------------------------------------------------------------------------------------
// The narrow QUAN layers are chains of ~1000 kernels of 3-40 us per training step; between two dependent kernels of a stream (or of
// a captured graph) the grid-launch latency is exposed.  Every kernel of the library starts with pdl_prologue(): it waits until the
// preceding grid has completed and its writes are visible (griddepcontrol.wait — a no-op when the kernel was launched without the
// attribute) and then lets ITS dependents be launched (griddepcontrol.launch_dependents), so the next kernel's CTAs are scheduled
// and parked at their own wait while this one runs.  No kernel touches global memory before the wait: ordering is exactly stream
// order.  Host side: QUAN_LAUNCH(...) = cudaLaunchKernelEx with programmaticStreamSerializationAllowed when QUAN_PDL=1.
// Measured on B200 (bench.py default workload, CUDA-graph replay): 17.32 ms/step with PDL, 17.17 ms without — a replayed graph
// already keeps the GPU 99 % busy (tools/graph_step_profile.py), the dependents only park on SMs earlier — so it is OFF by default.
__device__ __forceinline__ void pdl_prologue() {
#if defined(__CUDA_ARCH__)
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}

inline bool pdl_enabled() {
  static const bool on = [] { const char* e = getenv("QUAN_PDL"); return e != nullptr && atoi(e) != 0; }();
  return on;
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
#define QUAN_LAUNCH(kern, grid, block, smem, st, ...) (void)::quan::launch_pdl(kern, grid, block, smem, st, __VA_ARGS__)

struct Mix16 {
  float m[16];
};
static inline Mix16 make_mix(const float* mix) {
  Mix16 r;
  for (int i = 0; i < 16; ++i) r.m[i] = mix[i];
  return r;
}
static inline Mix16 make_mix_t(const float* mix) {  // transpose
  Mix16 r;
  for (int p = 0; p < 4; ++p)
    for (int q = 0; q < 4; ++q) r.m[q * 4 + p] = mix[p * 4 + q];
  return r;
}
__device__ __forceinline__ void apply_mix(const Mix16& M, const float (&in)[4], float (&out)[4]) {
#pragma unroll
  for (int p = 0; p < 4; ++p)
    out[p] = M.m[p * 4 + 0] * in[0] + M.m[p * 4 + 1] * in[1] + M.m[p * 4 + 2] * in[2] + M.m[p * 4 + 3] * in[3];
}

// element offset of (b,c,h,w,q) in the two physical layouts
template <int LAYOUT>
__device__ __forceinline__ int64_t elem_off(int b, int c, int h, int w, int q, int C, int H, int W) {
  if constexpr (LAYOUT == QUAN_LAYOUT_BCHWQ)
    return ((((int64_t)b * C + c) * H + h) * W + w) * 4 + q;
  else
    return ((((int64_t)b * H + h) * W + w) * 4 + q) * C + c;
}

}  // namespace quan
