// qconv_dw.cu — depthwise QConv2D (groups == C_i == C_o, the DWConv blocks of QUAN-YOLO11: conv.py:918-923) in the
// BHWQC layout.  HBM-bound streaming kernels: a pixel is 4 rows of C contiguous channels, a thread owns one 16-byte
// channel vector (fwd / dgrad) or one channel pair (wgrad) of ALL FOUR components, so the mixing matrix is applied in
// registers on the way out (fwd: y = M S) or on the way in (dgrad / wgrad: G = M^T dY) and neither S nor G touches HBM.
//
// Math (reference semantics): ultralytics/nn/modules/conv.py:472-499 with groups = C; backward = autograd of it.
//   S_q[c] = sum_tap x_q[c](pix (+) tap) w_q[c][tap] (+ b_r[c] on q = r);   y_p = sum_q M[p][q] S_q
//   dX_q[c] = sum_tap G_q[c](pix (-) tap) w_q[c][tap];   dW_q[c][tap] = sum_pix G_q[c](pix) x_q[c](pix (+) tap)
#include "qconv_internal.cuh"

namespace quan {

struct DwGeom {
  int B, C, H, W, Ho, Wo;          // x [B][H][W][4][C], y / dY [B][Ho][Wo][4][C]
  int kH, kW, sH, sW, pH, pW, dH, dW;
};

static DwGeom make_dw_geom(const quan_conv_dims& d) {
  DwGeom g;
  g.B = d.B; g.C = d.Ci; g.H = d.H; g.W = d.W;
  g.Ho = conv_out(d.H, d.kH, d.sH, d.pH, d.dH);
  g.Wo = conv_out(d.W, d.kW, d.sW, d.pW, d.dW);
  g.kH = d.kH; g.kW = d.kW; g.sH = d.sH; g.sW = d.sW; g.pH = d.pH; g.pW = d.pW; g.dH = d.dH; g.dW = d.dW;
  return g;
}

struct W4p {
  const float* w[4];
};

// TRANSPOSED = false: forward (one thread = one output pixel x V channels x 4 components, final mix with M).
// TRANSPOSED = true : dgrad (one thread = one input pixel; dY is mixed with M^T as it is loaded, taps that do not
//                     land on the stride grid are skipped).
// Weights are staged once per block in shared memory as [tap][q][C] so a thread's V coefficients are contiguous.
template <typename T, int V, bool TRANSPOSED>
__global__ void __launch_bounds__(256) qconv_dw_kernel(const T* __restrict__ in, W4p w, const float* __restrict__ bias_r,
                                                       T* __restrict__ out, DwGeom g, Mix16 M) {
  pdl_prologue();
  extern __shared__ float wsm[];   // [taps][4][C]
  const int taps = g.kH * g.kW;
  for (int e = threadIdx.x; e < taps * 4 * g.C; e += blockDim.x) {
    const int c = e % g.C, q = (e / g.C) & 3, tap = e / (4 * g.C);
    wsm[e] = __ldg(w.w[q] + (int64_t)c * taps + tap);
  }
  __syncthreads();
  const int cvs = g.C / V;
  const int Hout = TRANSPOSED ? g.H : g.Ho, Wout = TRANSPOSED ? g.W : g.Wo;   // grid this kernel writes
  const int Hin = TRANSPOSED ? g.Ho : g.H, Win = TRANSPOSED ? g.Wo : g.W;     // grid it reads
  const int64_t items = (int64_t)g.B * Hout * Wout * cvs;
  for (int64_t it = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; it < items; it += (int64_t)gridDim.x * blockDim.x) {
    const int cv = (int)(it % cvs);
    int64_t pix = it / cvs;
    const int wo = (int)(pix % Wout);
    pix /= Wout;
    const int ho = (int)(pix % Hout);
    const int b = (int)(pix / Hout);
    float acc[4][V];
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int v = 0; v < V; ++v) acc[q][v] = 0.f;
    if constexpr (!TRANSPOSED) {
      if (bias_r != nullptr) {
#pragma unroll
        for (int v = 0; v < V; ++v) acc[0][v] = __ldg(bias_r + cv * V + v);   // conv.py:480: bias joins S_r before the mix
      }
    }
    for (int kh = 0; kh < g.kH; ++kh) {
      int hi;
      if constexpr (TRANSPOSED) {
        const int th = ho + g.pH - kh * g.dH;
        if (th < 0 || th % g.sH != 0) continue;
        hi = th / g.sH;
      } else {
        hi = ho * g.sH - g.pH + kh * g.dH;
      }
      if (hi < 0 || hi >= Hin) continue;
      for (int kw = 0; kw < g.kW; ++kw) {
        int wi;
        if constexpr (TRANSPOSED) {
          const int tw = wo + g.pW - kw * g.dW;
          if (tw < 0 || tw % g.sW != 0) continue;
          wi = tw / g.sW;
        } else {
          wi = wo * g.sW - g.pW + kw * g.dW;
        }
        if (wi < 0 || wi >= Win) continue;
        const T* src = in + ((((int64_t)b * Hin + hi) * Win + wi) * 4) * g.C + cv * V;
        float xv[4][V];
#pragma unroll
        for (int q = 0; q < 4; ++q) load_vec<T, V>(src + (int64_t)q * g.C, xv[q]);
        const float* wt = wsm + (size_t)(kh * g.kW + kw) * 4 * g.C + cv * V;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
#pragma unroll
          for (int v = 0; v < V; ++v) {
            float a;
            if constexpr (TRANSPOSED)   // G_q = sum_p M[p][q] dY_p
              a = M.m[0 * 4 + q] * xv[0][v] + M.m[1 * 4 + q] * xv[1][v] + M.m[2 * 4 + q] * xv[2][v] + M.m[3 * 4 + q] * xv[3][v];
            else
              a = xv[q][v];
            acc[q][v] = fmaf(a, wt[q * g.C + v], acc[q][v]);
          }
        }
      }
    }
    T* dst = out + ((((int64_t)b * Hout + ho) * Wout + wo) * 4) * g.C + cv * V;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      float o[V];
#pragma unroll
      for (int v = 0; v < V; ++v) {
        if constexpr (TRANSPOSED) o[v] = acc[p][v];
        else o[v] = M.m[p * 4 + 0] * acc[0][v] + M.m[p * 4 + 1] * acc[1][v] + M.m[p * 4 + 2] * acc[2][v] + M.m[p * 4 + 3] * acc[3][v];
      }
      store_vec<T, V>(dst + (int64_t)p * g.C, o);
    }
  }
}

// wgrad: one thread = one 16-byte channel vector of ONE component (q); a block covers blockDim / (4 * C/V) output pixels per
// step and strides over the image.  Per pixel the thread loads dY of all four components of its channels (G_q = sum_p
// M[p][q] dY_p) and the TAPS shifted x vectors of its own component: 4 + TAPS 16-byte loads (the first version used 4-byte
// channel pairs, 40 loads per pixel, and was load-issue bound).  Accumulators [TAPS][V] in registers, folded over the
// block's pixel lanes through shared memory, one fp32 atomic per (q, c, tap) and block into the zero-initialised dW
// (atomic order is not deterministic).
template <typename T, int V, int TAPS>
__global__ void __launch_bounds__(256) qconv_dw_wgrad_kernel(const T* __restrict__ dy, const T* __restrict__ x, float* dw0,
                                                             float* dw1, float* dw2, float* dw3, DwGeom g, Mix16 M) {
  pdl_prologue();
  __shared__ float red[256][V + 1];
  const int cvs = g.C / V;
  const int tpp = 4 * cvs;                       // threads per pixel: (q, channel vector)
  const int ppb = blockDim.x / tpp;              // pixel lanes per block
  const int tl = threadIdx.x % tpp, pl = threadIdx.x / tpp;
  const int q = tl / cvs, cv = tl - q * cvs;
  const int64_t npix = (int64_t)g.B * g.Ho * g.Wo;
  const int taps = g.kH * g.kW;
  float acc[TAPS][V];
#pragma unroll
  for (int t = 0; t < TAPS; ++t)
#pragma unroll
    for (int v = 0; v < V; ++v) acc[t][v] = 0.f;
  if (pl < ppb) {
    const float m0 = M.m[0 * 4 + q], m1 = M.m[1 * 4 + q], m2 = M.m[2 * 4 + q], m3 = M.m[3 * 4 + q];
    for (int64_t pix = (int64_t)blockIdx.x * ppb + pl; pix < npix; pix += (int64_t)gridDim.x * ppb) {
      const int wo = (int)(pix % g.Wo);
      const int64_t r = pix / g.Wo;
      const int ho = (int)(r % g.Ho);
      const int b = (int)(r / g.Ho);
      float gy[4][V], gq[V];
      const T* gsrc = dy + (pix * 4) * g.C + cv * V;
#pragma unroll
      for (int p = 0; p < 4; ++p) load_vec<T, V>(gsrc + (int64_t)p * g.C, gy[p]);
#pragma unroll
      for (int v = 0; v < V; ++v) gq[v] = m0 * gy[0][v] + m1 * gy[1][v] + m2 * gy[2][v] + m3 * gy[3][v];
#pragma unroll
      for (int t = 0; t < TAPS; ++t) {
        const int kh = t / g.kW, kw = t - kh * g.kW;
        const int hi = ho * g.sH - g.pH + kh * g.dH, wi = wo * g.sW - g.pW + kw * g.dW;
        if (t < taps && hi >= 0 && hi < g.H && wi >= 0 && wi < g.W) {
          float xv[V];
          load_vec<T, V>(x + ((((int64_t)b * g.H + hi) * g.W + wi) * 4 + q) * g.C + cv * V, xv);
#pragma unroll
          for (int v = 0; v < V; ++v) acc[t][v] = fmaf(gq[v], xv[v], acc[t][v]);
        }
      }
    }
  }
  float* dw = q == 0 ? dw0 : q == 1 ? dw1 : q == 2 ? dw2 : dw3;
#pragma unroll
  for (int t = 0; t < TAPS; ++t) {               // unrolled: acc[][] must stay in registers (no dynamic indexing)
    __syncthreads();
#pragma unroll
    for (int v = 0; v < V; ++v) red[threadIdx.x][v] = pl < ppb ? acc[t][v] : 0.f;
    __syncthreads();
    if (pl == 0 && t < taps) {
#pragma unroll
      for (int v = 0; v < V; ++v) {
        float s = 0.f;
        for (int l = 0; l < ppb; ++l) s += red[l * tpp + tl][v];
        atomicAdd(dw + (int64_t)(cv * V + v) * taps + t, s);
      }
    }
  }
}

// wgrad, second form (filters up to 3x3): a thread owns (q, channel vector, FILTER ROW kh) of its pixel lane — [KW][V]
// accumulators instead of [9][V] — so three times as many threads fit a pixel, two blocks fit an SM (the first form needs
// 150 registers: one 256-thread block per SM, 362 GB/s on the 128^2 DWConv of QUAN-YOLO11n, 185 us under ncu), U pixels
// are in flight per loop trip and the index arithmetic is 32-bit.  The dY vectors a pixel's 12 threads share come
// from L1.  Same fold: block lanes through shared memory, one fp32 atomic per (q, c, tap) and block.
template <typename T, int V, int KW, int U>
__global__ void __launch_bounds__(256, 2) qconv_dw_wgrad_rows_kernel(const T* __restrict__ dy, const T* __restrict__ x,
                                                                     float* dw0, float* dw1, float* dw2, float* dw3,
                                                                     DwGeom g, Mix16 M) {
  pdl_prologue();
  __shared__ float red[256][V + 1];
  const int cvs = g.C / V;
  const int tpp = 4 * cvs * g.kH;                // threads per pixel: (kh, q, channel vector)
  const int ppb = blockDim.x / tpp;              // pixel lanes per block
  const int tl = threadIdx.x % tpp, pl = threadIdx.x / tpp;
  const int cv = tl % cvs, q = (tl / cvs) & 3, kh = tl / (4 * cvs);
  const int npix = g.B * g.Ho * g.Wo;            // < 2^31 (qconv_dw_supported)
  const int stride = (int)gridDim.x * ppb;
  const int taps = g.kH * g.kW;
  float acc[KW][V];
#pragma unroll
  for (int t = 0; t < KW; ++t)
#pragma unroll
    for (int v = 0; v < V; ++v) acc[t][v] = 0.f;
  if (pl < ppb) {
    const float m0 = M.m[0 * 4 + q], m1 = M.m[1 * 4 + q], m2 = M.m[2 * 4 + q], m3 = M.m[3 * 4 + q];
    for (int pix0 = (int)blockIdx.x * ppb + pl; pix0 < npix; pix0 += U * stride) {
      Vec<T, V> gy[U][4], xv[U][KW];              // raw 16-byte registers: converted after all loads are issued
      bool row[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int pix = pix0 + u * stride;
        const int wo = pix % g.Wo, r = pix / g.Wo;
        const int ho = r % g.Ho, b = r / g.Ho;
        const int hi = ho * g.sH - g.pH + kh * g.dH;
        row[u] = pix < npix && hi >= 0 && hi < g.H;
        if (row[u]) {
          const T* gsrc = dy + ((int64_t)pix * 4) * g.C + cv * V;
#pragma unroll
          for (int p = 0; p < 4; ++p) gy[u][p] = *reinterpret_cast<const Vec<T, V>*>(gsrc + (int64_t)p * g.C);
          const T* xrow = x + ((((int64_t)b * g.H + hi) * g.W) * 4 + q) * g.C + cv * V;
#pragma unroll
          for (int kw = 0; kw < KW; ++kw) {
            const int wi = wo * g.sW - g.pW + kw * g.dW;
            if (kw < g.kW && wi >= 0 && wi < g.W) {
              xv[u][kw] = *reinterpret_cast<const Vec<T, V>*>(xrow + (int64_t)wi * 4 * g.C);
            } else {
#pragma unroll
              for (int v = 0; v < V; ++v) xv[u][kw].v[v] = from_f32<T>(0.f);
            }
          }
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (!row[u]) continue;
#pragma unroll
        for (int v = 0; v < V; ++v) {
          const float gq = m0 * to_f32(gy[u][0].v[v]) + m1 * to_f32(gy[u][1].v[v]) + m2 * to_f32(gy[u][2].v[v]) +
                           m3 * to_f32(gy[u][3].v[v]);
#pragma unroll
          for (int kw = 0; kw < KW; ++kw) acc[kw][v] = fmaf(gq, to_f32(xv[u][kw].v[v]), acc[kw][v]);
        }
      }
    }
  }
  float* dw = q == 0 ? dw0 : q == 1 ? dw1 : q == 2 ? dw2 : dw3;
#pragma unroll
  for (int kw = 0; kw < KW; ++kw) {              // unrolled: acc[][] must stay in registers (no dynamic indexing)
    __syncthreads();
#pragma unroll
    for (int v = 0; v < V; ++v) red[threadIdx.x][v] = pl < ppb ? acc[kw][v] : 0.f;
    __syncthreads();
    if (pl == 0 && kw < g.kW && threadIdx.x < tpp) {
#pragma unroll
      for (int v = 0; v < V; ++v) {
        float s = 0.f;
        for (int l = 0; l < ppb; ++l) s += red[l * tpp + tl][v];
        atomicAdd(dw + (int64_t)(cv * V + v) * taps + kh * g.kW + kw, s);
      }
    }
  }
}

// ---- host ------------------------------------------------------------------------------------------------------------
bool qconv_dw_supported(const quan_conv_dims& d, int dtype, int layout, int pass) {
  if (layout != QUAN_LAYOUT_BHWQC || d.groups != d.Ci || d.Ci != d.Co || d.groups < 2) return false;
  const int V = dtype == QUAN_BF16 ? 8 : 4;
  if (d.Ci % V != 0 && d.Ci % (V / 2) != 0) return false;
  if ((size_t)d.kH * d.kW * 4 * d.Ci * sizeof(float) > 40 * 1024) return false;
  if (pass == PASS_WGRAD) return d.kH * d.kW <= 9 && 4 * (d.Ci / (d.Ci % V == 0 ? V : V / 2)) <= 256;
  return true;
}

template <typename T, bool TRANSPOSED>
static int dw_launch_t(const void* in, const float* const w[4], const float* bias_r, void* out, const DwGeom& g, const Mix16& M,
                       cudaStream_t st) {
  constexpr int VMAX = 16 / (int)sizeof(T);
  const size_t smem = (size_t)g.kH * g.kW * 4 * g.C * sizeof(float);
  const int64_t opix = (int64_t)g.B * (TRANSPOSED ? g.H * g.W : g.Ho * g.Wo);
  W4p w4 = {{w[0], w[1], w[2], w[3]}};
  QUAN_TIMED(st);
  if (g.C % VMAX == 0) {
    const int grid = grid_for(opix * (g.C / VMAX), 256, 8);
    QUAN_LAUNCH((qconv_dw_kernel<T, VMAX, TRANSPOSED>), grid, 256, smem, st, (const T*)in, w4, bias_r, (T*)out, g, M);
  } else {
    const int grid = grid_for(opix * (g.C / (VMAX / 2)), 256, 8);
    QUAN_LAUNCH((qconv_dw_kernel<T, VMAX / 2, TRANSPOSED>), grid, 256, smem, st, (const T*)in, w4, bias_r, (T*)out, g, M);
  }
  QUAN_CHECK_LAUNCH(TRANSPOSED ? "qconv_dw_dgrad" : "qconv_dw_fwd");
  return QUAN_OK;
}

int qconv_dw_fwd(const void* x, const float* const w[4], const float* bias_r, void* y, const quan_conv_dims& d, int dtype,
                 const float* mix, cudaStream_t st) {
  const DwGeom g = make_dw_geom(d);
  const Mix16 M = make_mix(mix);
  if (dtype == QUAN_BF16) return dw_launch_t<__nv_bfloat16, false>(x, w, bias_r, y, g, M, st);
  return dw_launch_t<float, false>(x, w, bias_r, y, g, M, st);
}

// dy is the raw output gradient (mix = forward mixing matrix; M^T is applied on load)
int qconv_dw_dgrad(const void* dy, const float* const w[4], void* dx, const quan_conv_dims& d, int dtype, const float* mix,
                   cudaStream_t st) {
  const DwGeom g = make_dw_geom(d);
  const Mix16 M = make_mix(mix);
  if (dtype == QUAN_BF16) return dw_launch_t<__nv_bfloat16, true>(dy, w, nullptr, dx, g, M, st);
  return dw_launch_t<float, true>(dy, w, nullptr, dx, g, M, st);
}

int qconv_dw_wgrad(const void* dy, const void* x, float* const dw[4], const quan_conv_dims& d, int dtype, const float* mix,
                   cudaStream_t st) {
  const DwGeom g = make_dw_geom(d);
  const Mix16 M = make_mix(mix);
  const int taps = d.kH * d.kW;
  for (int q = 0; q < 4; ++q) QUAN_CUDA(cudaMemsetAsync(dw[q], 0, (size_t)d.Co * taps * sizeof(float), st));
  const int VMAX = dtype == QUAN_BF16 ? 8 : 4;
  const int V = g.C % VMAX == 0 ? VMAX : VMAX / 2;
  const int64_t npix = (int64_t)g.B * g.Ho * g.Wo;
  static const int env_rows = [] { const char* e = getenv("QUAN_DW_WGRAD_ROWS"); return e ? atoi(e) : 1; }();
  // measured (tools/narrow_wgrad_probe.py, B200, bf16, 16 images): rows form 171 vs 241 us at 16 ch x 128^2, 104 vs 112 us at
  // 32 ch x 64^2, but 85 vs 61 us at 64 ch x 32^2 (three times the blocks' fold + atomics tail on few pixels)
  if ((env_rows == 2 || (env_rows == 1 && npix >= 32768)) && d.kH <= 3 && d.kW <= 3 && 4 * (g.C / V) * d.kH <= 256 &&
      npix < (1ll << 31)) {
    const int ppb = 256 / (4 * (g.C / V) * d.kH);
    int64_t blocks = ceil_div64(npix, (int64_t)ppb * 16);        // >= 16 pixels (8 loop trips) per pixel lane
    if (blocks > QUAN_NUM_SMS * 4) blocks = QUAN_NUM_SMS * 4;
    if (blocks < 1) blocks = 1;
    QUAN_TIMED(st);
#define QUAN_DW_ROWS(TT, VV) QUAN_LAUNCH((qconv_dw_wgrad_rows_kernel<TT, VV, 3, 2>), (unsigned)blocks, 256, 0, st,  \
    (const TT*)dy, (const TT*)x, dw[0], dw[1], dw[2], dw[3], g, M)
    if (dtype == QUAN_BF16) { if (V == 8) QUAN_DW_ROWS(__nv_bfloat16, 8); else QUAN_DW_ROWS(__nv_bfloat16, 4); }
    else { if (V == 4) QUAN_DW_ROWS(float, 4); else QUAN_DW_ROWS(float, 2); }
#undef QUAN_DW_ROWS
    QUAN_CHECK_LAUNCH("qconv_dw_wgrad_rows_kernel");
    return QUAN_OK;
  }
  const int tpp = 4 * (g.C / V), ppb = 256 / tpp;
  int64_t blocks = ceil_div64(npix, (int64_t)ppb * 16);          // >= 16 pixels per pixel lane
  if (blocks > QUAN_NUM_SMS * 4) blocks = QUAN_NUM_SMS * 4;
  if (blocks < 1) blocks = 1;
  QUAN_TIMED(st);
  if (dtype == QUAN_BF16) {
    const __nv_bfloat16 *dyp = (const __nv_bfloat16*)dy, *xp = (const __nv_bfloat16*)x;
    if (V == 8) QUAN_LAUNCH((qconv_dw_wgrad_kernel<__nv_bfloat16, 8, 9>), (unsigned)blocks, 256, 0, st, dyp, xp, dw[0], dw[1], dw[2], dw[3], g, M);
    else QUAN_LAUNCH((qconv_dw_wgrad_kernel<__nv_bfloat16, 4, 9>), (unsigned)blocks, 256, 0, st, dyp, xp, dw[0], dw[1], dw[2], dw[3], g, M);
  } else {
    const float *dyp = (const float*)dy, *xp = (const float*)x;
    if (V == 4) QUAN_LAUNCH((qconv_dw_wgrad_kernel<float, 4, 9>), (unsigned)blocks, 256, 0, st, dyp, xp, dw[0], dw[1], dw[2], dw[3], g, M);
    else QUAN_LAUNCH((qconv_dw_wgrad_kernel<float, 2, 9>), (unsigned)blocks, 256, 0, st, dyp, xp, dw[0], dw[1], dw[2], dw[3], g, M);
  }
  QUAN_CHECK_LAUNCH("qconv_dw_wgrad_kernel");
  return QUAN_OK;
}

}  // namespace quan
