// elementwise.cu — Poincare RGB->quaternion map, QUpsample (nearest), 4x4 mix, layout conversion.
// Bandwidth-bound streaming kernels: 16-byte vector accesses, grid-stride, grids capped at 148 x 8 blocks.
//
// Reference semantics: ultralytics/nn/modules/conv.py:388-397 (Poincare), :1229-1246 (QUpsample),
// ultralytics/nn/cuda/quaternion_ops_head.cu:8-95 (mix), conv.py:441 (`.contiguous()` layout copy).
#include "common.cuh"

namespace quan {

// ---- Poincare ---------------------------------------------------------------------------------
// n = R^2+G^2+B^2, q = [(1-n)/(1+n), 2R/(1+n), 2G/(1+n), 2B/(1+n)]
template <typename T>
__global__ void __launch_bounds__(256) poincare_fwd_kernel(const float* __restrict__ rgb, T* __restrict__ out,
                                                           int64_t HW, int64_t total) {
  pdl_prologue();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < total; p += stride) {
    const int64_t b = p / HW, i = p - b * HW;
    const float* src = rgb + b * 3 * HW + i;
    const float r = __ldg(src), g = __ldg(src + HW), bl = __ldg(src + 2 * HW);
    const float n = r * r + g * g + bl * bl;
    const float den = 1.0f + n;
    float o[4] = {(1.0f - n) / den, 2.0f * r / den, 2.0f * g / den, 2.0f * bl / den};
    store_vec<T, 4>(out + p * 4, o);
  }
}

// d/dx of the map, contracted with grad_out g = (g0, g1, g2, g3):
//   d q0/d x_a = -4 x_a / den^2 ;  d q_b/d x_a = 2 delta_ab/den - 4 x_a x_b / den^2
template <typename T>
__global__ void __launch_bounds__(256) poincare_bwd_kernel(const float* __restrict__ rgb,
                                                           const T* __restrict__ gout, float* __restrict__ grgb,
                                                           int64_t HW, int64_t total) {
  pdl_prologue();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < total; p += stride) {
    const int64_t b = p / HW, i = p - b * HW;
    const float* src = rgb + b * 3 * HW + i;
    const float x[3] = {__ldg(src), __ldg(src + HW), __ldg(src + 2 * HW)};
    float g[4];
    load_vec<T, 4>(gout + p * 4, g);
    const float n = x[0] * x[0] + x[1] * x[1] + x[2] * x[2];
    const float inv = 1.0f / (1.0f + n);
    const float dot = x[0] * g[1] + x[1] * g[2] + x[2] * g[3];
    const float common = -4.0f * inv * inv * (g[0] + dot);
    float* dst = grgb + b * 3 * HW + i;
#pragma unroll
    for (int a = 0; a < 3; ++a) dst[a * HW] = fmaf(common, x[a], 2.0f * inv * g[a + 1]);
  }
}

// ---- QUpsample nearest ------------------------------------------------------------------------
// Both layouts reduce to: "groups" of `inner` contiguous elements indexed by (outer, h, w); the group of input
// pixel (h,w) is replicated to the s*s output pixels (h*s+dy, w*s+dx).
//   BCHWQ: outer = b*C + c, inner = 4          BHWQC: outer = b, inner = 4C
template <typename T, int V, bool BWD>
__global__ void __launch_bounds__(256) upsample_kernel(const T* __restrict__ src, T* __restrict__ dst,
                                                       int64_t outer, int H, int W, int inner_vecs, int s) {
  pdl_prologue();
  // src/dst roles: FWD reads the small tensor and writes the big one; BWD reads big, writes small.
  const int64_t total = outer * H * W * inner_vecs;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int Ho = H * s, Wo = W * s;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
    const int iv = (int)(t % inner_vecs);
    int64_t pix = t / inner_vecs;
    const int w = (int)(pix % W);
    pix /= W;
    const int h = (int)(pix % H);
    const int64_t o = pix / H;
    const int64_t small_off = (((o * H + h) * W + w) * inner_vecs + iv) * V;
    float v[V];
    if constexpr (!BWD) {
      load_vec<T, V>(src + small_off, v);
    } else {
#pragma unroll
      for (int i = 0; i < V; ++i) v[i] = 0.f;
    }
    for (int dy = 0; dy < s; ++dy) {
      for (int dx = 0; dx < s; ++dx) {
        const int64_t big_off = (((o * Ho + (h * s + dy)) * Wo + (w * s + dx)) * inner_vecs + iv) * V;
        if constexpr (!BWD) {
          store_vec<T, V>(dst + big_off, v);
        } else {
          float t2[V];
          load_vec<T, V>(src + big_off, t2);
#pragma unroll
          for (int i = 0; i < V; ++i) v[i] += t2[i];
        }
      }
    }
    if constexpr (BWD) store_vec<T, V>(dst + small_off, v);
  }
}

// ---- 4x4 mix ----------------------------------------------------------------------------------
// BCHWQ: one quaternion = 4 contiguous elements.
template <typename T>
__global__ void __launch_bounds__(256) mix_a_kernel(const T* __restrict__ in, T* __restrict__ out, int64_t nquat,
                                                    Mix16 M) {
  pdl_prologue();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < nquat; t += stride) {
    float v[4], o[4];
    load_vec<T, 4>(in + t * 4, v);
    apply_mix(M, v, o);
    store_vec<T, 4>(out + t * 4, o);
  }
}
// BHWQC: a thread takes V channels of one pixel: 4 vector loads at stride C.
template <typename T, int V>
__global__ void __launch_bounds__(256) mix_b_kernel(const T* __restrict__ in, T* __restrict__ out, int64_t rows,
                                                    int C, Mix16 M) {
  pdl_prologue();
  const int chunks = C / V;
  const int64_t total = rows * chunks;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
    const int64_t r = t / chunks;
    const int j = (int)(t - r * chunks);
    const int64_t base = r * 4 * C + (int64_t)j * V;
    float v[4][V], o[4][V];
#pragma unroll
    for (int q = 0; q < 4; ++q) load_vec<T, V>(in + base + (int64_t)q * C, v[q]);
#pragma unroll
    for (int i = 0; i < V; ++i) {
      float a[4] = {v[0][i], v[1][i], v[2][i], v[3][i]}, b[4];
      apply_mix(M, a, b);
      o[0][i] = b[0]; o[1][i] = b[1]; o[2][i] = b[2]; o[3][i] = b[3];
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) store_vec<T, V>(out + base + (int64_t)q * C, o[q]);
  }
}

// ---- layout conversion: per image, transpose [C][HW] quaternions <-> [HW][4][C] ----------------------
// Tile = 32 channels x 32 pixels (x4 components), staged through padded shared memory so both sides are coalesced.
template <typename T, bool A2B>
__global__ void __launch_bounds__(256) layout_kernel(const T* __restrict__ src, T* __restrict__ dst, int C, int HW) {
  pdl_prologue();
  __shared__ float tile[32][4][33];  // [c][q][pix] (+1 pad)
  const int b = blockIdx.z;
  const int c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
  const int64_t imgA = (int64_t)b * C * HW * 4;  // BCHWQ image base (elements)
  const int64_t imgB = imgA;                     // same element count
  const int tid = threadIdx.x;
  if constexpr (A2B) {
    // read BCHWQ: for channel c the (pix,q) run is contiguous: 32 pix * 4 q = 128 elements per channel row
    for (int e = tid; e < 32 * 128; e += 256) {
      const int c = e / 128, r = e % 128, pix = r / 4, q = r % 4;
      float v = 0.f;
      if (c0 + c < C && p0 + pix < HW) v = to_f32(src[imgA + ((int64_t)(c0 + c) * HW + p0 + pix) * 4 + q]);
      tile[c][q][pix] = v;
    }
    __syncthreads();
    // write BHWQC: for pixel p, component q the c run is contiguous
    for (int e = tid; e < 32 * 128; e += 256) {
      const int c = e % 32, q = (e / 32) % 4, pix = e / 128;
      if (c0 + c < C && p0 + pix < HW)
        dst[imgB + ((int64_t)(p0 + pix) * 4 + q) * C + c0 + c] = from_f32<T>(tile[c][q][pix]);
    }
  } else {
    for (int e = tid; e < 32 * 128; e += 256) {
      const int c = e % 32, q = (e / 32) % 4, pix = e / 128;
      float v = 0.f;
      if (c0 + c < C && p0 + pix < HW) v = to_f32(src[imgB + ((int64_t)(p0 + pix) * 4 + q) * C + c0 + c]);
      tile[c][q][pix] = v;
    }
    __syncthreads();
    for (int e = tid; e < 32 * 128; e += 256) {
      const int c = e / 128, r = e % 128, pix = r / 4, q = r % 4;
      if (c0 + c < C && p0 + pix < HW)
        dst[imgA + ((int64_t)(c0 + c) * HW + p0 + pix) * 4 + q] = from_f32<T>(tile[c][q][pix]);
    }
  }
}

static int check_dims(const char* who, const void* a, const void* b, int B, int C, int H, int W, int dtype, int layout) {
  QUAN_REQUIRE(a != nullptr && b != nullptr, QUAN_E_ARG, "%s: null pointer", who);
  QUAN_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, QUAN_E_ARG, "%s: non-positive dims", who);
  QUAN_REQUIRE(dtype == QUAN_F32 || dtype == QUAN_BF16, QUAN_E_ARG, "%s: bad dtype %d", who, dtype);
  QUAN_REQUIRE(layout == QUAN_LAYOUT_BCHWQ || layout == QUAN_LAYOUT_BHWQC, QUAN_E_ARG, "%s: bad layout %d", who, layout);
  return QUAN_OK;
}

template <typename T, bool BWD>
static int launch_upsample(const void* src, void* dst, int B, int C, int H, int W, int s, int layout, cudaStream_t st) {
  const T* sp = reinterpret_cast<const T*>(src);
  T* dp = reinterpret_cast<T*>(dst);
  int64_t outer;
  int inner;
  if (layout == QUAN_LAYOUT_BCHWQ) { outer = (int64_t)B * C; inner = 4; }
  else { outer = B; inner = 4 * C; }
  int V = largest_pow2_divisor(inner, VecTraits<T>::kMaxVec);
  int inner_vecs = inner / V;
  int64_t total = outer * H * W * inner_vecs;
  int grid = grid_for(total, 256, 8);
  switch (V) {
    case 8:
      if constexpr (sizeof(T) == 2) QUAN_LAUNCH((upsample_kernel<T, 8, BWD>), grid, 256, 0, st, sp, dp, outer, H, W, inner_vecs, s);
      break;
    case 4: QUAN_LAUNCH((upsample_kernel<T, 4, BWD>), grid, 256, 0, st, sp, dp, outer, H, W, inner_vecs, s); break;
    default: set_error("upsample: unreachable vector width %d", V); return QUAN_E_UNSUPPORTED;
  }
  QUAN_CHECK_LAUNCH("upsample");
  return QUAN_OK;
}

}  // namespace quan

using namespace quan;

extern "C" {

int quan_poincare_fwd(const float* rgb, void* out, int32_t B, int32_t H, int32_t W, int out_dtype, void* stream) {
  QUAN_REQUIRE(rgb != nullptr && out != nullptr, QUAN_E_ARG, "poincare_fwd: null pointer");
  QUAN_REQUIRE(B > 0 && H > 0 && W > 0, QUAN_E_ARG, "poincare_fwd: non-positive dims");
  QUAN_REQUIRE(out_dtype == QUAN_F32 || out_dtype == QUAN_BF16, QUAN_E_ARG, "poincare_fwd: bad dtype");
  const int64_t HW = (int64_t)H * W, total = HW * B;
  int grid = grid_for(total, 256, 8);
  cudaStream_t st = (cudaStream_t)stream;
  QUAN_TIMED(st);
  if (out_dtype == QUAN_F32) QUAN_LAUNCH((poincare_fwd_kernel<float>), grid, 256, 0, st, rgb, (float*)out, HW, total);
  else QUAN_LAUNCH((poincare_fwd_kernel<__nv_bfloat16>), grid, 256, 0, st, rgb, (__nv_bfloat16*)out, HW, total);
  QUAN_CHECK_LAUNCH("poincare_fwd");
  return QUAN_OK;
}

int quan_poincare_bwd(const float* rgb, const void* grad_out, float* grad_rgb, int32_t B, int32_t H, int32_t W,
                      int out_dtype, void* stream) {
  QUAN_REQUIRE(rgb != nullptr && grad_out != nullptr && grad_rgb != nullptr, QUAN_E_ARG, "poincare_bwd: null pointer");
  QUAN_REQUIRE(B > 0 && H > 0 && W > 0, QUAN_E_ARG, "poincare_bwd: non-positive dims");
  QUAN_REQUIRE(out_dtype == QUAN_F32 || out_dtype == QUAN_BF16, QUAN_E_ARG, "poincare_bwd: bad dtype");
  const int64_t HW = (int64_t)H * W, total = HW * B;
  int grid = grid_for(total, 256, 8);
  cudaStream_t st = (cudaStream_t)stream;
  if (out_dtype == QUAN_F32)
    QUAN_LAUNCH((poincare_bwd_kernel<float>), grid, 256, 0, st, rgb, (const float*)grad_out, grad_rgb, HW, total);
  else
    QUAN_LAUNCH((poincare_bwd_kernel<__nv_bfloat16>), grid, 256, 0, st, rgb, (const __nv_bfloat16*)grad_out, grad_rgb, HW, total);
  QUAN_CHECK_LAUNCH("poincare_bwd");
  return QUAN_OK;
}

int quan_qupsample_nearest_fwd(const void* x, void* y, int32_t B, int32_t C, int32_t H, int32_t W, int32_t scale,
                               int dtype, int layout, void* stream) {
  int rc = check_dims("qupsample_fwd", x, y, B, C, H, W, dtype, layout);
  if (rc) return rc;
  QUAN_REQUIRE(scale >= 1 && scale <= 8, QUAN_E_ARG, "qupsample_fwd: scale %d outside [1,8]", scale);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == QUAN_F32) return launch_upsample<float, false>(x, y, B, C, H, W, scale, layout, st);
  return launch_upsample<__nv_bfloat16, false>(x, y, B, C, H, W, scale, layout, st);
}

int quan_qupsample_nearest_bwd(const void* dy, void* dx, int32_t B, int32_t C, int32_t H, int32_t W, int32_t scale,
                               int dtype, int layout, void* stream) {
  int rc = check_dims("qupsample_bwd", dy, dx, B, C, H, W, dtype, layout);
  if (rc) return rc;
  QUAN_REQUIRE(scale >= 1 && scale <= 8, QUAN_E_ARG, "qupsample_bwd: scale %d outside [1,8]", scale);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == QUAN_F32) return launch_upsample<float, true>(dy, dx, B, C, H, W, scale, layout, st);
  return launch_upsample<__nv_bfloat16, true>(dy, dx, B, C, H, W, scale, layout, st);
}

int quan_mix(const void* in, void* out, int32_t B, int32_t C, int32_t H, int32_t W, int dtype, int layout,
             const float* mix, void* stream) {
  int rc = check_dims("mix", in, out, B, C, H, W, dtype, layout);
  if (rc) return rc;
  QUAN_REQUIRE(mix != nullptr, QUAN_E_ARG, "mix: null matrix");
  Mix16 M = make_mix(mix);
  cudaStream_t st = (cudaStream_t)stream;
  QUAN_TIMED(st);
  const int64_t rows = (int64_t)B * H * W;
  if (layout == QUAN_LAYOUT_BCHWQ || C == 1) {
    const int64_t nquat = rows * C;
    int grid = grid_for(nquat, 256, 8);
    if (dtype == QUAN_F32) QUAN_LAUNCH((mix_a_kernel<float>), grid, 256, 0, st, (const float*)in, (float*)out, nquat, M);
    else QUAN_LAUNCH((mix_a_kernel<__nv_bfloat16>), grid, 256, 0, st, (const __nv_bfloat16*)in, (__nv_bfloat16*)out, nquat, M);
  } else {
    if (dtype == QUAN_F32) {
      int V = largest_pow2_divisor(C, 4);
      int grid = grid_for(rows * (C / V), 256, 8);
      if (V == 4) QUAN_LAUNCH((mix_b_kernel<float, 4>), grid, 256, 0, st, (const float*)in, (float*)out, rows, C, M);
      else if (V == 2) QUAN_LAUNCH((mix_b_kernel<float, 2>), grid, 256, 0, st, (const float*)in, (float*)out, rows, C, M);
      else QUAN_LAUNCH((mix_b_kernel<float, 1>), grid, 256, 0, st, (const float*)in, (float*)out, rows, C, M);
    } else {
      int V = largest_pow2_divisor(C, 8);
      int grid = grid_for(rows * (C / V), 256, 8);
      const __nv_bfloat16* ip = (const __nv_bfloat16*)in;
      __nv_bfloat16* op = (__nv_bfloat16*)out;
      if (V == 8) QUAN_LAUNCH((mix_b_kernel<__nv_bfloat16, 8>), grid, 256, 0, st, ip, op, rows, C, M);
      else if (V == 4) QUAN_LAUNCH((mix_b_kernel<__nv_bfloat16, 4>), grid, 256, 0, st, ip, op, rows, C, M);
      else if (V == 2) QUAN_LAUNCH((mix_b_kernel<__nv_bfloat16, 2>), grid, 256, 0, st, ip, op, rows, C, M);
      else QUAN_LAUNCH((mix_b_kernel<__nv_bfloat16, 1>), grid, 256, 0, st, ip, op, rows, C, M);
    }
  }
  QUAN_CHECK_LAUNCH("mix");
  return QUAN_OK;
}

int quan_layout_convert(const void* src, int src_layout, void* dst, int dst_layout, int32_t B, int32_t C, int32_t H,
                        int32_t W, int dtype, void* stream) {
  int rc = check_dims("layout_convert", src, dst, B, C, H, W, dtype, src_layout);
  if (rc) return rc;
  QUAN_REQUIRE(dst_layout == QUAN_LAYOUT_BCHWQ || dst_layout == QUAN_LAYOUT_BHWQC, QUAN_E_ARG, "layout_convert: bad dst layout");
  QUAN_REQUIRE(src != dst, QUAN_E_ARG, "layout_convert: in-place conversion is not supported");
  cudaStream_t st = (cudaStream_t)stream;
  QUAN_TIMED(st);
  const size_t esz = dtype == QUAN_F32 ? 4 : 2;
  const int HW = H * W;
  if (src_layout == dst_layout || C == 1) {
    QUAN_CUDA(cudaMemcpyAsync(dst, src, (size_t)B * C * HW * 4 * esz, cudaMemcpyDeviceToDevice, st));
    return QUAN_OK;
  }
  QUAN_REQUIRE(B <= 65535 && (C + 31) / 32 <= 65535, QUAN_E_UNSUPPORTED, "layout_convert: B or C too large for the grid");
  dim3 grid((HW + 31) / 32, (C + 31) / 32, B);
  const bool a2b = src_layout == QUAN_LAYOUT_BCHWQ;
  if (dtype == QUAN_F32) {
    if (a2b) QUAN_LAUNCH((layout_kernel<float, true>), grid, 256, 0, st, (const float*)src, (float*)dst, C, HW);
    else QUAN_LAUNCH((layout_kernel<float, false>), grid, 256, 0, st, (const float*)src, (float*)dst, C, HW);
  } else {
    if (a2b) QUAN_LAUNCH((layout_kernel<__nv_bfloat16, true>), grid, 256, 0, st, (const __nv_bfloat16*)src, (__nv_bfloat16*)dst, C, HW);
    else QUAN_LAUNCH((layout_kernel<__nv_bfloat16, false>), grid, 256, 0, st, (const __nv_bfloat16*)src, (__nv_bfloat16*)dst, C, HW);
  }
  QUAN_CHECK_LAUNCH("layout_convert");
  return QUAN_OK;
}

}  // extern "C"


// ---- channel-slice gather ---------------------------------------------------------------------------------------------------------
// A `chunk` / `split` along the channel axis of a BHWQC tensor (C2f / C3k2 / QC2PSA in block.py:337-365, :1548-1600 hand one half of a
// Conv output to the next layer) is a strided view: rows of C' contiguous elements, one per (pixel, component), `src_ld` elements
// apart.  Kernels of this library read dense tensors; torch's generic strided copy takes ~15 us for these (103 per YOLO11n step).
// This is the same copy with 16-byte vectors: row r of the dense destination = src[r * src_ld .. + row_elems).
namespace quan {
template <int VB>
__global__ void __launch_bounds__(256) rows_gather_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int64_t nrows, int row_bytes,
                                                          int64_t src_ld_bytes) {
  pdl_prologue();
  const int vpr = row_bytes / VB;
  const int64_t n = nrows * vpr;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / vpr;
    const int v = (int)(i - r * vpr);
    if constexpr (VB == 16) *reinterpret_cast<uint4*>(dst + r * row_bytes + v * 16) = *reinterpret_cast<const uint4*>(src + r * src_ld_bytes + v * 16);
    else if constexpr (VB == 8) *reinterpret_cast<uint2*>(dst + r * row_bytes + v * 8) = *reinterpret_cast<const uint2*>(src + r * src_ld_bytes + v * 8);
    else *reinterpret_cast<uint32_t*>(dst + r * row_bytes + v * 4) = *reinterpret_cast<const uint32_t*>(src + r * src_ld_bytes + v * 4);
  }
}
}  // namespace quan

extern "C" int quan_rows_gather(const void* src, void* dst, int64_t nrows, int32_t row_bytes, int64_t src_ld_bytes, void* stream) {
  using namespace quan;
  QUAN_REQUIRE(src != nullptr && dst != nullptr && nrows > 0 && row_bytes > 0 && src_ld_bytes >= row_bytes, QUAN_E_ARG, "rows_gather: bad argument");
  QUAN_REQUIRE(row_bytes % 4 == 0 && src_ld_bytes % 4 == 0, QUAN_E_UNSUPPORTED, "rows_gather: rows must be multiples of 4 bytes");
  cudaStream_t st = (cudaStream_t)stream;
  const uintptr_t al = reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst) | (uintptr_t)row_bytes | (uintptr_t)src_ld_bytes;
  QUAN_REQUIRE(al % 4 == 0, QUAN_E_UNSUPPORTED, "rows_gather: pointers must be 4-byte aligned");
  const int vb = (al % 16 == 0) ? 16 : (al % 8 == 0) ? 8 : 4;
  const int64_t n = nrows * (row_bytes / vb);
  const int grid = grid_for(n, 256, 8);
  QUAN_TIMED(st);
  if (vb == 16) QUAN_LAUNCH((rows_gather_kernel<16>), grid, 256, 0, st, (const uint8_t*)src, (uint8_t*)dst, nrows, row_bytes, src_ld_bytes);
  else if (vb == 8) QUAN_LAUNCH((rows_gather_kernel<8>), grid, 256, 0, st, (const uint8_t*)src, (uint8_t*)dst, nrows, row_bytes, src_ld_bytes);
  else QUAN_LAUNCH((rows_gather_kernel<4>), grid, 256, 0, st, (const uint8_t*)src, (uint8_t*)dst, nrows, row_bytes, src_ld_bytes);
  QUAN_CHECK_LAUNCH("rows_gather");
  return QUAN_OK;
}


// Channel concatenation in BHWQC (`torch.cat(xs, 1)` of C2f / C3k2 / C3k / QSPPF / QC2PSA / Concat, ultralytics/nn/modules/block.py:350-352,
// conv.py Concat): every source is rows of row_bytes (one (pixel, component) each) ld_bytes apart — dense tensors and channel chunks of
// other tensors alike — written side by side into the destination's rows.  torch copies each source with its generic strided-copy kernel
// (~15 us per source at 16 x 128^2, 65 launches per training step); here one launch moves all sources with 16-byte vectors.
namespace quan {
struct CatArgs {
  const uint8_t* src[QUAN_CAT_MAX];
  int64_t ld[QUAN_CAT_MAX];
  int vend[QUAN_CAT_MAX];      // running vector count per destination row: source s owns vectors [vend[s-1], vend[s])
  int nsrc;
};
// SPLIT = false: sources -> destination rows (concatenation); SPLIT = true: the same geometry backwards — every part receives its
// columns of the wide tensor's rows (the dense gradient slices of a concatenation's backward, one launch)
template <int VB, bool SPLIT = false>
__global__ void __launch_bounds__(256) rows_cat_kernel(const CatArgs a, uint8_t* __restrict__ dst, int64_t dst_ld, int64_t nrows) {
  pdl_prologue();
  const int vpr = a.vend[a.nsrc - 1];
  const int64_t n = nrows * vpr;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / vpr;
    const int v = (int)(i - r * vpr);
    int s = 0;
#pragma unroll
    for (int k = 0; k < QUAN_CAT_MAX - 1; ++k) s += (k < a.nsrc - 1 && v >= a.vend[k]) ? 1 : 0;
    const int v0 = s == 0 ? 0 : a.vend[s - 1];
    const uint8_t* part = a.src[s] + r * a.ld[s] + (int64_t)(v - v0) * VB;
    uint8_t* wide = dst + r * dst_ld + (int64_t)v * VB;
    const uint8_t* sp = SPLIT ? wide : part;
    uint8_t* dp = SPLIT ? const_cast<uint8_t*>(part) : wide;
    if constexpr (VB == 16) *reinterpret_cast<uint4*>(dp) = *reinterpret_cast<const uint4*>(sp);
    else if constexpr (VB == 8) *reinterpret_cast<uint2*>(dp) = *reinterpret_cast<const uint2*>(sp);
    else *reinterpret_cast<uint32_t*>(dp) = *reinterpret_cast<const uint32_t*>(sp);
  }
}
}  // namespace quan

static int rows_cat_impl(const quan_cat_src* srcs, int32_t nsrc, void* dst, int64_t dst_ld_bytes, int64_t nrows, void* stream, bool split);

extern "C" int quan_rows_cat(const quan_cat_src* srcs, int32_t nsrc, void* dst, int64_t dst_ld_bytes, int64_t nrows, void* stream) {
  return rows_cat_impl(srcs, nsrc, dst, dst_ld_bytes, nrows, stream, false);
}
extern "C" int quan_rows_split(const quan_cat_src* parts, int32_t nparts, const void* src, int64_t src_ld_bytes, int64_t nrows, void* stream) {
  return rows_cat_impl(parts, nparts, const_cast<void*>(src), src_ld_bytes, nrows, stream, true);
}

static int rows_cat_impl(const quan_cat_src* srcs, int32_t nsrc, void* dst, int64_t dst_ld_bytes, int64_t nrows, void* stream, bool split) {
  using namespace quan;
  QUAN_REQUIRE(srcs != nullptr && dst != nullptr && nsrc >= 1 && nsrc <= QUAN_CAT_MAX && nrows > 0, QUAN_E_ARG, "rows_cat: bad argument (1..%d sources)",
               QUAN_CAT_MAX);
  uintptr_t al = reinterpret_cast<uintptr_t>(dst) | (uintptr_t)dst_ld_bytes;
  int64_t total = 0;
  for (int i = 0; i < nsrc; ++i) {
    QUAN_REQUIRE(srcs[i].ptr != nullptr && srcs[i].row_bytes > 0 && srcs[i].ld_bytes >= srcs[i].row_bytes, QUAN_E_ARG, "rows_cat: source %d", i);
    al |= reinterpret_cast<uintptr_t>(srcs[i].ptr) | (uintptr_t)srcs[i].ld_bytes | (uintptr_t)srcs[i].row_bytes;
    total += srcs[i].row_bytes;
  }
  QUAN_REQUIRE(total <= dst_ld_bytes, QUAN_E_SHAPE, "rows_cat: sources are %lld bytes wide, destination rows %lld", (long long)total, (long long)dst_ld_bytes);
  QUAN_REQUIRE(al % 4 == 0, QUAN_E_UNSUPPORTED, "rows_cat: pointers, pitches and widths must be multiples of 4 bytes");
  const int vb = (al % 16 == 0) ? 16 : (al % 8 == 0) ? 8 : 4;
  CatArgs a = {};
  a.nsrc = nsrc;
  int vend = 0;
  for (int i = 0; i < nsrc; ++i) {
    a.src[i] = (const uint8_t*)srcs[i].ptr; a.ld[i] = srcs[i].ld_bytes;
    vend += srcs[i].row_bytes / vb;
    a.vend[i] = vend;
  }
  cudaStream_t st = (cudaStream_t)stream;
  timing_work("rows_cat", "", 2.0 * nrows * total, 0.0);
  const int grid = grid_for(nrows * vend, 256, 8);
  QUAN_TIMED(st);
  if (split) {
    if (vb == 16) QUAN_LAUNCH((rows_cat_kernel<16, true>), grid, 256, 0, st, a, (uint8_t*)dst, dst_ld_bytes, nrows);
    else if (vb == 8) QUAN_LAUNCH((rows_cat_kernel<8, true>), grid, 256, 0, st, a, (uint8_t*)dst, dst_ld_bytes, nrows);
    else QUAN_LAUNCH((rows_cat_kernel<4, true>), grid, 256, 0, st, a, (uint8_t*)dst, dst_ld_bytes, nrows);
  } else {
    if (vb == 16) QUAN_LAUNCH((rows_cat_kernel<16>), grid, 256, 0, st, a, (uint8_t*)dst, dst_ld_bytes, nrows);
    else if (vb == 8) QUAN_LAUNCH((rows_cat_kernel<8>), grid, 256, 0, st, a, (uint8_t*)dst, dst_ld_bytes, nrows);
    else QUAN_LAUNCH((rows_cat_kernel<4>), grid, 256, 0, st, a, (uint8_t*)dst, dst_ld_bytes, nrows);
  }
  QUAN_CHECK_LAUNCH(split ? "rows_split" : "rows_cat");
  return QUAN_OK;
}
