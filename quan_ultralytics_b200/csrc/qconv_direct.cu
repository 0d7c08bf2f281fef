// qconv_direct.cu — CUDA-core ("direct") engine for QConv2D: every shape the reference accepts (any groups,
// stride, dilation, kernel size, both layouts, fp32/bf16).  It serves the narrow / depthwise / first-layer
// convolutions of the real QUAN models (C_q 1..32, HBM- and latency-bound) and is the in-library
// cross-check for the tcgen05 engine (qconv_tc.cu) which takes the wide layers.
//
// Math (reference semantics, not code): ultralytics/nn/modules/conv.py:472-499,
// classification/quaternion/qconv.py:592-612, ultralytics/nn/cuda/quaternion_ops.cu:43-181 (fwd),
// :185-311 (dgrad), :314-470 (wgrad), :473-530 (bias).
//   S_q = conv2d(x_q, w_q) (+ b_r on q = r);  y_p = sum_q M[p][q] S_q
//   G = M^T dY;  dX_q = conv_transpose(G_q, w_q);  dW_q = corr(G_q, x_q);  db_r = sum G_r
#include "common.cuh"

namespace quan {

struct ConvGeom {
  int B, Ci, Co, H, W, Ho, Wo;
  int kH, kW, sH, sW, pH, pW, dH, dW, groups;
  int Cig, Cog;   // per-group channel counts
};

static inline ConvGeom make_geom(const quan_conv_dims& d) {
  ConvGeom g;
  g.B = d.B; g.Ci = d.Ci; g.Co = d.Co; g.H = d.H; g.W = d.W;
  g.kH = d.kH; g.kW = d.kW; g.sH = d.sH; g.sW = d.sW; g.pH = d.pH; g.pW = d.pW; g.dH = d.dH; g.dW = d.dW;
  g.groups = d.groups;
  g.Ho = conv_out(d.H, d.kH, d.sH, d.pH, d.dH);
  g.Wo = conv_out(d.W, d.kW, d.sW, d.pW, d.dW);
  g.Cig = d.Ci / d.groups;
  g.Cog = d.Co / d.groups;
  return g;
}

struct W4 {
  const float* w[4];
};

template <typename T, int LAYOUT>
__device__ __forceinline__ void load_quat(const T* __restrict__ p, int b, int c, int h, int w, int C, int H, int W,
                                          float (&v)[4]) {
  if constexpr (LAYOUT == QUAN_LAYOUT_BCHWQ) {
    load_vec<T, 4>(p + elem_off<LAYOUT>(b, c, h, w, 0, C, H, W), v);
  } else {
    const int64_t base = elem_off<LAYOUT>(b, c, h, w, 0, C, H, W);
#pragma unroll
    for (int q = 0; q < 4; ++q) v[q] = to_f32(p[base + (int64_t)q * C]);
  }
}
template <typename T, int LAYOUT>
__device__ __forceinline__ void store_quat(T* __restrict__ p, int b, int c, int h, int w, int C, int H, int W,
                                           const float (&v)[4]) {
  if constexpr (LAYOUT == QUAN_LAYOUT_BCHWQ) {
    store_vec<T, 4>(p + elem_off<LAYOUT>(b, c, h, w, 0, C, H, W), v);
  } else {
    const int64_t base = elem_off<LAYOUT>(b, c, h, w, 0, C, H, W);
#pragma unroll
    for (int q = 0; q < 4; ++q) p[base + (int64_t)q * C] = from_f32<T>(v[q]);
  }
}

// ------------------------------------------------------------------------------------------------
// forward: one thread = one output pixel x COT output channels (one group) x 4 components.
// Weights of the block's (group, co-chunk) are staged in shared memory as [ci][tap][co][q] so a single
// broadcast LDS.128 feeds 4 FMAs.  grid = (pixel blocks, groups * chunks_per_group).
// ------------------------------------------------------------------------------------------------
template <typename T, int LAYOUT, int COT>
__global__ void __launch_bounds__(128) qconv_fwd_direct(const T* __restrict__ x, W4 w, const float* __restrict__ bias_r,
                                                        T* __restrict__ y, ConvGeom g, Mix16 M, int cib) {
  pdl_prologue();
  extern __shared__ float4 wsm[];  // [cib][taps][COT]
  const int taps = g.kH * g.kW;
  const int chunks_per_group = (g.Cog + COT - 1) / COT;
  const int grp = blockIdx.y / chunks_per_group;
  const int co0 = grp * g.Cog + (blockIdx.y % chunks_per_group) * COT;   // absolute first out channel
  const int co_end = min(co0 + COT, (grp + 1) * g.Cog);
  const int nco = co_end - co0;

  const int64_t npix = (int64_t)g.B * g.Ho * g.Wo;
  const int64_t pix = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = pix < npix;
  int b = 0, ho = 0, wo = 0;
  if (active) {
    wo = (int)(pix % g.Wo);
    int64_t t = pix / g.Wo;
    ho = (int)(t % g.Ho);
    b = (int)(t / g.Ho);
  }
  float acc[COT][4];
#pragma unroll
  for (int c = 0; c < COT; ++c)
#pragma unroll
    for (int q = 0; q < 4; ++q) acc[c][q] = 0.f;

  for (int ci0 = 0; ci0 < g.Cig; ci0 += cib) {
    const int ncib = min(cib, g.Cig - ci0);
    __syncthreads();
    for (int e = threadIdx.x; e < ncib * taps * COT; e += blockDim.x) {
      const int c = e % COT, tap = (e / COT) % taps, ci = e / (COT * taps);
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c < nco) {
        const int64_t widx = ((int64_t)(co0 + c) * g.Cig + (ci0 + ci)) * taps + tap;
        v = make_float4(__ldg(w.w[0] + widx), __ldg(w.w[1] + widx), __ldg(w.w[2] + widx), __ldg(w.w[3] + widx));
      }
      wsm[e] = v;
    }
    __syncthreads();
    if (!active) continue;
    for (int ci = 0; ci < ncib; ++ci) {
      const int cin = grp * g.Cig + ci0 + ci;
      for (int kh = 0; kh < g.kH; ++kh) {
        const int hi = ho * g.sH - g.pH + kh * g.dH;
        if (hi < 0 || hi >= g.H) continue;
        for (int kw = 0; kw < g.kW; ++kw) {
          const int wi = wo * g.sW - g.pW + kw * g.dW;
          if (wi < 0 || wi >= g.W) continue;
          float xv[4];
          load_quat<T, LAYOUT>(x, b, cin, hi, wi, g.Ci, g.H, g.W, xv);
          const float4* wrow = wsm + ((size_t)ci * taps + kh * g.kW + kw) * COT;
#pragma unroll
          for (int c = 0; c < COT; ++c) {
            const float4 wv = wrow[c];
            acc[c][0] = fmaf(xv[0], wv.x, acc[c][0]);
            acc[c][1] = fmaf(xv[1], wv.y, acc[c][1]);
            acc[c][2] = fmaf(xv[2], wv.z, acc[c][2]);
            acc[c][3] = fmaf(xv[3], wv.w, acc[c][3]);
          }
        }
      }
    }
  }
  if (!active) return;
#pragma unroll
  for (int c = 0; c < COT; ++c) {
    if (c < nco) {
      float s[4] = {acc[c][0], acc[c][1], acc[c][2], acc[c][3]}, o[4];
      if (bias_r != nullptr) s[0] += __ldg(bias_r + co0 + c);   // conv.py:480: bias enters S_r before the mix
      apply_mix(M, s, o);
      store_quat<T, LAYOUT>(y, b, co0 + c, ho, wo, g.Co, g.Ho, g.Wo, o);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// dgrad: one thread = one INPUT pixel x CIT input channels (one group) x 4 components, reading G = M^T dY.
// Shared weights [co][tap][ci][q].  grid = (input-pixel blocks, groups * chunks_per_group).
// ------------------------------------------------------------------------------------------------
template <typename T, int LAYOUT, int CIT>
__global__ void __launch_bounds__(128) qconv_dgrad_direct(const T* __restrict__ gq, W4 w, T* __restrict__ dx,
                                                          ConvGeom g, int cob) {
  pdl_prologue();
  extern __shared__ float4 wsm[];  // [cob][taps][CIT]
  const int taps = g.kH * g.kW;
  const int chunks_per_group = (g.Cig + CIT - 1) / CIT;
  const int grp = blockIdx.y / chunks_per_group;
  const int cil0 = (blockIdx.y % chunks_per_group) * CIT;        // first in-channel, local to the group
  const int nci = min(CIT, g.Cig - cil0);

  const int64_t npix = (int64_t)g.B * g.H * g.W;
  const int64_t pix = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = pix < npix;
  int b = 0, h = 0, wq = 0;
  if (active) {
    wq = (int)(pix % g.W);
    int64_t t = pix / g.W;
    h = (int)(t % g.H);
    b = (int)(t / g.H);
  }
  float acc[CIT][4];
#pragma unroll
  for (int c = 0; c < CIT; ++c)
#pragma unroll
    for (int q = 0; q < 4; ++q) acc[c][q] = 0.f;

  for (int co0 = 0; co0 < g.Cog; co0 += cob) {
    const int ncob = min(cob, g.Cog - co0);
    __syncthreads();
    for (int e = threadIdx.x; e < ncob * taps * CIT; e += blockDim.x) {
      const int c = e % CIT, tap = (e / CIT) % taps, co = e / (CIT * taps);
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c < nci) {
        const int64_t widx = ((int64_t)(grp * g.Cog + co0 + co) * g.Cig + (cil0 + c)) * taps + tap;
        v = make_float4(__ldg(w.w[0] + widx), __ldg(w.w[1] + widx), __ldg(w.w[2] + widx), __ldg(w.w[3] + widx));
      }
      wsm[e] = v;
    }
    __syncthreads();
    if (!active) continue;
    for (int co = 0; co < ncob; ++co) {
      const int cout = grp * g.Cog + co0 + co;
      for (int kh = 0; kh < g.kH; ++kh) {
        const int th = h + g.pH - kh * g.dH;
        if (th < 0 || (th % g.sH) != 0) continue;
        const int ho = th / g.sH;
        if (ho >= g.Ho) continue;
        for (int kw = 0; kw < g.kW; ++kw) {
          const int tw = wq + g.pW - kw * g.dW;
          if (tw < 0 || (tw % g.sW) != 0) continue;
          const int wo = tw / g.sW;
          if (wo >= g.Wo) continue;
          float gv[4];
          load_quat<T, LAYOUT>(gq, b, cout, ho, wo, g.Co, g.Ho, g.Wo, gv);
          const float4* wrow = wsm + ((size_t)co * taps + kh * g.kW + kw) * CIT;
#pragma unroll
          for (int c = 0; c < CIT; ++c) {
            const float4 wv = wrow[c];
            acc[c][0] = fmaf(gv[0], wv.x, acc[c][0]);
            acc[c][1] = fmaf(gv[1], wv.y, acc[c][1]);
            acc[c][2] = fmaf(gv[2], wv.z, acc[c][2]);
            acc[c][3] = fmaf(gv[3], wv.w, acc[c][3]);
          }
        }
      }
    }
  }
  if (!active) return;
#pragma unroll
  for (int c = 0; c < CIT; ++c) {
    if (c < nci) {
      float o[4] = {acc[c][0], acc[c][1], acc[c][2], acc[c][3]};
      store_quat<T, LAYOUT>(dx, b, grp * g.Cig + cil0 + c, h, wq, g.Ci, g.H, g.W, o);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// wgrad (general): one thread = one weight element (co, ci_local, tap) x 4 components; blockIdx.y splits the
// output pixels; partial sums are combined with fp32 atomics into zero-initialised dW.
// ------------------------------------------------------------------------------------------------
template <typename T, int LAYOUT>
__global__ void __launch_bounds__(128) qconv_wgrad_direct(const T* __restrict__ gq, const T* __restrict__ x,
                                                          float* __restrict__ dw0, float* __restrict__ dw1,
                                                          float* __restrict__ dw2, float* __restrict__ dw3,
                                                          ConvGeom g, int64_t pix_per_split) {
  pdl_prologue();
  const int taps = g.kH * g.kW;
  const int64_t nelem = (int64_t)g.Co * g.Cig * taps;
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nelem) return;
  const int tap = (int)(e % taps);
  const int cil = (int)((e / taps) % g.Cig);
  const int co = (int)(e / ((int64_t)taps * g.Cig));
  const int kh = tap / g.kW, kw = tap % g.kW;
  const int cin = (co / g.Cog) * g.Cig + cil;

  const int64_t npix = (int64_t)g.B * g.Ho * g.Wo;
  const int64_t p0 = (int64_t)blockIdx.y * pix_per_split;
  const int64_t p1 = min(p0 + pix_per_split, npix);
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  int wo = (int)(p0 % g.Wo);
  int64_t t = p0 / g.Wo;
  int ho = (int)(t % g.Ho);
  int b = (int)(t / g.Ho);
  for (int64_t p = p0; p < p1; ++p) {
    const int hi = ho * g.sH - g.pH + kh * g.dH;
    const int wi = wo * g.sW - g.pW + kw * g.dW;
    if (hi >= 0 && hi < g.H && wi >= 0 && wi < g.W) {
      float gv[4], xv[4];
      load_quat<T, LAYOUT>(gq, b, co, ho, wo, g.Co, g.Ho, g.Wo, gv);
      load_quat<T, LAYOUT>(x, b, cin, hi, wi, g.Ci, g.H, g.W, xv);
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[q] = fmaf(gv[q], xv[q], acc[q]);
    }
    if (++wo == g.Wo) {
      wo = 0;
      if (++ho == g.Ho) { ho = 0; ++b; }
    }
  }
  atomicAdd(dw0 + e, acc[0]);
  atomicAdd(dw1 + e, acc[1]);
  atomicAdd(dw2 + e, acc[2]);
  atomicAdd(dw3 + e, acc[3]);
}

// bias grad: db_r[co] = sum over pixels of G_r.  grid = (Co, splits), fp32 atomics into zeroed db.
template <typename T, int LAYOUT>
__global__ void __launch_bounds__(256) qconv_bias_grad(const T* __restrict__ gq, float* __restrict__ db, ConvGeom g) {
  pdl_prologue();
  const int co = blockIdx.x;
  const int64_t npix = (int64_t)g.B * g.Ho * g.Wo;
  const int64_t hw = (int64_t)g.Ho * g.Wo;
  float acc = 0.f;
  for (int64_t p = (int64_t)blockIdx.y * blockDim.x + threadIdx.x; p < npix; p += (int64_t)gridDim.y * blockDim.x) {
    const int64_t b = p / hw, i = p - b * hw;
    int64_t off;
    if constexpr (LAYOUT == QUAN_LAYOUT_BCHWQ) off = ((b * g.Co + co) * hw + i) * 4;
    else off = ((b * hw + i) * 4) * g.Co + co;
    acc += to_f32(gq[off]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __shared__ float red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += red[i];
    atomicAdd(db + co, s);
  }
}

// bias grad in BHWQC: the r-component row of a pixel is C_o contiguous elements.  A thread owns one 16-byte channel vector
// and walks pixels; lanes fold through shared memory, one fp32 atomic per channel and block.  (The per-channel kernel above
// reads one element per thread at a 4*C_o stride — every block pulls the same sectors: 29 us per Q-ResNet-34 layer, 1.05 ms
// per step under ncu.)
template <typename T, int V>
__global__ void __launch_bounds__(256) qconv_bias_grad_rows(const T* __restrict__ gq, float* __restrict__ db, int64_t npix, int Co) {
  pdl_prologue();
  __shared__ float red[256][V + 1];
  const int cvs = Co / V;
  const int lanes = blockDim.x / cvs;
  const int cv = threadIdx.x % cvs, pl = threadIdx.x / cvs;
  float acc[V];
#pragma unroll
  for (int v = 0; v < V; ++v) acc[v] = 0.f;
  if (pl < lanes) {
    for (int64_t p = (int64_t)blockIdx.x * lanes + pl; p < npix; p += (int64_t)gridDim.x * lanes) {
      float t[V];
      load_vec<T, V>(gq + p * 4 * Co + cv * V, t);
#pragma unroll
      for (int v = 0; v < V; ++v) acc[v] += t[v];
    }
  }
#pragma unroll
  for (int v = 0; v < V; ++v) red[threadIdx.x][v] = pl < lanes ? acc[v] : 0.f;
  __syncthreads();
  if (pl == 0) {
#pragma unroll
    for (int v = 0; v < V; ++v) {
      float s = 0.f;
      for (int l = 0; l < lanes; ++l) s += red[l * cvs + cv][v];
      atomicAdd(db + cv * V + v, s);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// host launchers (called from qconv_api.cu)
// ------------------------------------------------------------------------------------------------
static int pick_tile(int c_per_group) { return c_per_group >= 8 ? 8 : (c_per_group >= 4 ? 4 : (c_per_group >= 2 ? 2 : 1)); }

template <typename T, int LAYOUT>
static int fwd_direct_t(const void* x, const float* const w[4], const float* bias_r, void* y, const ConvGeom& g,
                        const Mix16& M, cudaStream_t st) {
  const int taps = g.kH * g.kW;
  const int cot = pick_tile(g.Cog);
  int cib = 2048 / (taps * cot);   // <= 32 KB of float4 weights
  if (cib < 1) cib = 1;
  if (cib > g.Cig) cib = g.Cig;
  const size_t smem = (size_t)cib * taps * cot * sizeof(float4);
  QUAN_REQUIRE(smem <= 48 * 1024, QUAN_E_UNSUPPORTED, "qconv direct fwd: kernel %dx%d too large", g.kH, g.kW);
  const int64_t npix = (int64_t)g.B * g.Ho * g.Wo;
  const int chunks = (g.Cog + cot - 1) / cot;
  QUAN_REQUIRE((int64_t)g.groups * chunks <= 65535, QUAN_E_UNSUPPORTED, "qconv direct fwd: too many channel chunks");
  dim3 grid((unsigned)ceil_div64(npix, 128), (unsigned)(g.groups * chunks));
  W4 w4 = {{w[0], w[1], w[2], w[3]}};
  QUAN_TIMED(st);
  const T* xp = (const T*)x;
  T* yp = (T*)y;
  switch (cot) {
    case 8: QUAN_LAUNCH((qconv_fwd_direct<T, LAYOUT, 8>), grid, 128, smem, st, xp, w4, bias_r, yp, g, M, cib); break;
    case 4: QUAN_LAUNCH((qconv_fwd_direct<T, LAYOUT, 4>), grid, 128, smem, st, xp, w4, bias_r, yp, g, M, cib); break;
    case 2: QUAN_LAUNCH((qconv_fwd_direct<T, LAYOUT, 2>), grid, 128, smem, st, xp, w4, bias_r, yp, g, M, cib); break;
    default: QUAN_LAUNCH((qconv_fwd_direct<T, LAYOUT, 1>), grid, 128, smem, st, xp, w4, bias_r, yp, g, M, cib); break;
  }
  QUAN_CHECK_LAUNCH("qconv_fwd_direct");
  return QUAN_OK;
}

template <typename T, int LAYOUT>
static int dgrad_direct_t(const void* gq, const float* const w[4], void* dx, const ConvGeom& g, cudaStream_t st) {
  const int taps = g.kH * g.kW;
  const int cit = pick_tile(g.Cig);
  int cob = 2048 / (taps * cit);
  if (cob < 1) cob = 1;
  if (cob > g.Cog) cob = g.Cog;
  const size_t smem = (size_t)cob * taps * cit * sizeof(float4);
  QUAN_REQUIRE(smem <= 48 * 1024, QUAN_E_UNSUPPORTED, "qconv direct dgrad: kernel %dx%d too large", g.kH, g.kW);
  const int64_t npix = (int64_t)g.B * g.H * g.W;
  const int chunks = (g.Cig + cit - 1) / cit;
  QUAN_REQUIRE((int64_t)g.groups * chunks <= 65535, QUAN_E_UNSUPPORTED, "qconv direct dgrad: too many channel chunks");
  dim3 grid((unsigned)ceil_div64(npix, 128), (unsigned)(g.groups * chunks));
  W4 w4 = {{w[0], w[1], w[2], w[3]}};
  QUAN_TIMED(st);
  const T* gp = (const T*)gq;
  T* dp = (T*)dx;
  switch (cit) {
    case 8: QUAN_LAUNCH((qconv_dgrad_direct<T, LAYOUT, 8>), grid, 128, smem, st, gp, w4, dp, g, cob); break;
    case 4: QUAN_LAUNCH((qconv_dgrad_direct<T, LAYOUT, 4>), grid, 128, smem, st, gp, w4, dp, g, cob); break;
    case 2: QUAN_LAUNCH((qconv_dgrad_direct<T, LAYOUT, 2>), grid, 128, smem, st, gp, w4, dp, g, cob); break;
    default: QUAN_LAUNCH((qconv_dgrad_direct<T, LAYOUT, 1>), grid, 128, smem, st, gp, w4, dp, g, cob); break;
  }
  QUAN_CHECK_LAUNCH("qconv_dgrad_direct");
  return QUAN_OK;
}

template <typename T, int LAYOUT>
static int wgrad_direct_t(const void* gq, const void* x, float* const dw[4], const ConvGeom& g, cudaStream_t st) {
  const int taps = g.kH * g.kW;
  const int64_t nelem = (int64_t)g.Co * g.Cig * taps;
  for (int q = 0; q < 4; ++q) QUAN_CUDA(cudaMemsetAsync(dw[q], 0, nelem * sizeof(float), st));
  const int64_t npix = (int64_t)g.B * g.Ho * g.Wo;
  const int64_t eblocks = ceil_div64(nelem, 128);
  // enough pixel splits to fill the machine ~4 deep, but keep >= 64 pixels per split
  int64_t splits = ceil_div64((int64_t)QUAN_NUM_SMS * 16, eblocks);
  int64_t max_splits = ceil_div64(npix, 64);
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  if (splits > 65535) splits = 65535;
  const int64_t pps = ceil_div64(npix, splits);
  splits = ceil_div64(npix, pps);
  dim3 grid((unsigned)eblocks, (unsigned)splits);
  QUAN_TIMED(st);
  QUAN_LAUNCH((qconv_wgrad_direct<T, LAYOUT>), grid, 128, 0, st, (const T*)gq, (const T*)x, dw[0], dw[1], dw[2], dw[3], g, pps);
  QUAN_CHECK_LAUNCH("qconv_wgrad_direct");
  return QUAN_OK;
}

template <typename T, int LAYOUT>
static int bias_grad_t(const void* gq, float* db, const ConvGeom& g, cudaStream_t st) {
  QUAN_CUDA(cudaMemsetAsync(db, 0, (size_t)g.Co * sizeof(float), st));
  const int64_t npix = (int64_t)g.B * g.Ho * g.Wo;
  if constexpr (LAYOUT == QUAN_LAYOUT_BHWQC) {
    constexpr int VMAX = 16 / (int)sizeof(T);
    const int V = g.Co % VMAX == 0 ? VMAX : (g.Co % (VMAX / 2) == 0 ? VMAX / 2 : 0);
    if (V != 0 && g.Co / V <= 256) {
      const int lanes = 256 / (g.Co / V);
      int64_t blocks = ceil_div64(npix, (int64_t)lanes * 8);
      if (blocks > QUAN_NUM_SMS * 4) blocks = QUAN_NUM_SMS * 4;
      if (blocks < 1) blocks = 1;
      QUAN_TIMED(st);
      if (V == VMAX) QUAN_LAUNCH((qconv_bias_grad_rows<T, VMAX>), (unsigned)blocks, 256, 0, st, (const T*)gq, db, npix, g.Co);
      else QUAN_LAUNCH((qconv_bias_grad_rows<T, VMAX / 2>), (unsigned)blocks, 256, 0, st, (const T*)gq, db, npix, g.Co);
      QUAN_CHECK_LAUNCH("qconv_bias_grad");
      return QUAN_OK;
    }
  }
  int64_t splits = ceil_div64((int64_t)QUAN_NUM_SMS * 4, g.Co);
  int64_t max_splits = ceil_div64(npix, 256 * 4);
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  dim3 grid((unsigned)g.Co, (unsigned)splits);
  QUAN_LAUNCH((qconv_bias_grad<T, LAYOUT>), grid, 256, 0, st, (const T*)gq, db, g);
  QUAN_CHECK_LAUNCH("qconv_bias_grad");
  return QUAN_OK;
}

#define QUAN_DISPATCH_TL(dtype, layout, CALL)                                                    \
  do {                                                                                           \
    if ((dtype) == QUAN_F32) {                                                                   \
      if ((layout) == QUAN_LAYOUT_BCHWQ) { using T = float; constexpr int L = QUAN_LAYOUT_BCHWQ; return CALL; } \
      else { using T = float; constexpr int L = QUAN_LAYOUT_BHWQC; return CALL; }                \
    } else {                                                                                     \
      if ((layout) == QUAN_LAYOUT_BCHWQ) { using T = __nv_bfloat16; constexpr int L = QUAN_LAYOUT_BCHWQ; return CALL; } \
      else { using T = __nv_bfloat16; constexpr int L = QUAN_LAYOUT_BHWQC; return CALL; }        \
    }                                                                                            \
  } while (0)

int qconv_fwd_direct_launch(const void* x, const float* const w[4], const float* bias_r, void* y,
                            const quan_conv_dims& d, int dtype, int layout, const float* mix, cudaStream_t st) {
  ConvGeom g = make_geom(d);
  Mix16 M = make_mix(mix);
  QUAN_DISPATCH_TL(dtype, layout, (fwd_direct_t<T, L>(x, w, bias_r, y, g, M, st)));
}
int qconv_dgrad_direct_launch(const void* gq, const float* const w[4], void* dx, const quan_conv_dims& d, int dtype,
                              int layout, cudaStream_t st) {
  ConvGeom g = make_geom(d);
  QUAN_DISPATCH_TL(dtype, layout, (dgrad_direct_t<T, L>(gq, w, dx, g, st)));
}
int qconv_wgrad_direct_launch(const void* gq, const void* x, float* const dw[4], const quan_conv_dims& d, int dtype,
                              int layout, cudaStream_t st) {
  ConvGeom g = make_geom(d);
  QUAN_DISPATCH_TL(dtype, layout, (wgrad_direct_t<T, L>(gq, x, dw, g, st)));
}
int qconv_bias_grad_launch(const void* gq, float* db, const quan_conv_dims& d, int dtype, int layout, cudaStream_t st) {
  ConvGeom g = make_geom(d);
  QUAN_DISPATCH_TL(dtype, layout, (bias_grad_t<T, L>(gq, db, g, st)));
}

}  // namespace quan
