// qconv_small.cu — QConv2D for layers with 1..8 quaternion channels on either side (the stem of QUAN-YOLO11n: 1 -> 4 at
// 1024^2, 4 -> 2 and 2 -> 4 at 256^2; ultralytics/cfg/models/11/yolo11-obb-quan.yaml) in the BHWQC layout.  Below the
// tensor core's granularity (a pixel is 8..64 bytes, K rows shorter than a swizzle row) and purely HBM-bound: a thread
// owns one whole pixel — all four components of all channels arrive in one or two 16-byte loads and leave the same way —
// so the mixing matrix is applied in registers (fwd: y = M S; dgrad / wgrad: G = M^T dY on load).
//
// Wide outputs behind a 1..2-channel input (the Q-ResNet-34 stem 1 -> 16, 7x7 stride 2, classification/models/
// quaternion_models.py) run forward and wgrad in chunks of 8 output channels (blockIdx.y): a K row of 8..16 bytes is
// below the tensor core's 32-byte swizzle row, and the direct engine needs 16 ms for that layer's step.
//
// Math (reference semantics): ultralytics/nn/modules/conv.py:472-499; backward = autograd of it.
#include "qconv_internal.cuh"

namespace quan {

struct SmallGeom {
  int B, H, W, Ho, Wo;             // x [B][H][W][4][CI], y / dY [B][Ho][Wo][4][CO]
  int kH, kW, sH, sW, pH, pW, dH, dW;
  int CT;                          // forward / wgrad: C_o of the tensor (>= the COUT a block handles; chunk = blockIdx.y)
};
static SmallGeom make_small_geom(const quan_conv_dims& d) {
  SmallGeom g;
  g.B = d.B; g.H = d.H; g.W = d.W;
  g.Ho = conv_out(d.H, d.kH, d.sH, d.pH, d.dH);
  g.Wo = conv_out(d.W, d.kW, d.sW, d.pW, d.dW);
  g.kH = d.kH; g.kW = d.kW; g.sH = d.sH; g.sW = d.sW; g.pH = d.pH; g.pW = d.pW; g.dH = d.dH; g.dW = d.dW;
  g.CT = d.Co;
  return g;
}
struct W4s {
  const float* w[4];
};

// one pixel = N contiguous elements (N*sizeof(T) in {8, 16, 32, 64} bytes)
template <typename T, int N>
__device__ __forceinline__ void load_pixel(const T* __restrict__ p, float (&v)[N]) {
  constexpr int VW = (N * (int)sizeof(T) >= 16) ? 16 / (int)sizeof(T) : N;
#pragma unroll
  for (int i = 0; i < N; i += VW) {
    float t[VW];
    load_vec<T, VW>(p + i, t);
#pragma unroll
    for (int j = 0; j < VW; ++j) v[i + j] = t[j];
  }
}
template <typename T, int N>
__device__ __forceinline__ void store_pixel(T* __restrict__ p, const float (&v)[N]) {
  constexpr int VW = (N * (int)sizeof(T) >= 16) ? 16 / (int)sizeof(T) : N;
#pragma unroll
  for (int i = 0; i < N; i += VW) {
    float t[VW];
#pragma unroll
    for (int j = 0; j < VW; ++j) t[j] = v[i + j];
    store_vec<T, VW>(p + i, t);
  }
}

// CIN / COUT: channels of the tensor this kernel READS / WRITES.
// TRANSPOSED = false: forward (reads x, writes y = M S).  TRANSPOSED = true: dgrad (reads dY, mixes it with M^T on load,
// writes dX; CIN = C_o, COUT = C_i; taps off the stride grid are skipped).  Weights staged in shared memory as
// [tap][q][cout][cin] — every lane reads the same address (broadcast).
template <typename T, int CIN, int COUT, bool TRANSPOSED>
__global__ void __launch_bounds__(128) qconv_small_kernel(const T* __restrict__ in, W4s w, const float* __restrict__ bias_r,
                                                          T* __restrict__ out, SmallGeom g, Mix16 M) {
  pdl_prologue();
  extern __shared__ float wsm[];   // [taps][4][COUT][CIN]
  const int taps = g.kH * g.kW;
  const int co0 = TRANSPOSED ? 0 : (int)blockIdx.y * COUT;   // forward: this block's chunk of output channels
  const int CT = TRANSPOSED ? COUT : g.CT;
  for (int e = threadIdx.x; e < taps * 4 * COUT * CIN; e += blockDim.x) {
    const int ci = e % CIN, co = (e / CIN) % COUT, q = (e / (CIN * COUT)) & 3, tap = e / (4 * CIN * COUT);
    // master layout W_q[C_o][C_i][tap]
    const int64_t widx = TRANSPOSED ? ((int64_t)ci * COUT + co) * taps + tap : ((int64_t)(co0 + co) * CIN + ci) * taps + tap;
    wsm[e] = __ldg(w.w[q] + widx);
  }
  __syncthreads();
  const int Hout = TRANSPOSED ? g.H : g.Ho, Wout = TRANSPOSED ? g.W : g.Wo;
  const int Hin = TRANSPOSED ? g.Ho : g.H, Win = TRANSPOSED ? g.Wo : g.W;
  const int64_t npix = (int64_t)g.B * Hout * Wout;
  for (int64_t pix = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; pix < npix; pix += (int64_t)gridDim.x * blockDim.x) {
    const int wo = (int)(pix % Wout);
    const int64_t r = pix / Wout;
    const int ho = (int)(r % Hout);
    const int b = (int)(r / Hout);
    float acc[4][COUT];
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int c = 0; c < COUT; ++c) acc[q][c] = 0.f;
    if constexpr (!TRANSPOSED) {
      if (bias_r != nullptr) {
#pragma unroll
        for (int c = 0; c < COUT; ++c) acc[0][c] = __ldg(bias_r + co0 + c);   // conv.py:480: bias joins S_r before the mix
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
        float xv[4 * CIN];
        load_pixel<T, 4 * CIN>(in + (((int64_t)b * Hin + hi) * Win + wi) * 4 * CIN, xv);
        const float* wt = wsm + (size_t)(kh * g.kW + kw) * 4 * COUT * CIN;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float a[CIN];
#pragma unroll
          for (int ci = 0; ci < CIN; ++ci) {
            if constexpr (TRANSPOSED)   // G_q = sum_p M[p][q] dY_p
              a[ci] = M.m[0 * 4 + q] * xv[ci] + M.m[1 * 4 + q] * xv[CIN + ci] + M.m[2 * 4 + q] * xv[2 * CIN + ci] +
                      M.m[3 * 4 + q] * xv[3 * CIN + ci];
            else
              a[ci] = xv[q * CIN + ci];
          }
#pragma unroll
          for (int co = 0; co < COUT; ++co)
#pragma unroll
            for (int ci = 0; ci < CIN; ++ci) acc[q][co] = fmaf(a[ci], wt[(q * COUT + co) * CIN + ci], acc[q][co]);
        }
      }
    }
    float o[4 * COUT];
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
      for (int c = 0; c < COUT; ++c) {
        if constexpr (TRANSPOSED) o[p * COUT + c] = acc[p][c];
        else o[p * COUT + c] = M.m[p * 4 + 0] * acc[0][c] + M.m[p * 4 + 1] * acc[1][c] + M.m[p * 4 + 2] * acc[2][c] + M.m[p * 4 + 3] * acc[3][c];
      }
    if (CT == COUT) {
      store_pixel<T, 4 * COUT>(out + pix * 4 * COUT, o);
    } else {
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        float t[COUT];
#pragma unroll
        for (int c = 0; c < COUT; ++c) t[c] = o[p * COUT + c];
        store_pixel<T, COUT>(out + (pix * 4 + p) * CT + co0, t);
      }
    }
  }
}

// wgrad: dW_q[co][ci][tap] = sum_pix G_q[co](pix) x_q[ci](pix (+) tap).  A thread owns ONE tap and 4*CO*CI accumulators;
// a block is `taps` x (blockDim / taps) pixel lanes and walks tiles of 256 consecutive output pixels:
//   phase A: thread t loads pixel t of the tile (each dY byte leaves HBM once, all loads independent), applies M^T and
//            parks G in shared memory as fp32, with the pixel's input-window origin (no divisions later);
//   phase B: the tap threads of a pixel lane read G from shared memory (broadcast) and their own x pixel from global
//            memory (U of them in flight), 4*CO*CI FMAs per (pixel, tap).
// The first version let every tap thread load and mix dY itself: the taps of a pixel asked for the same bytes, so the
// unique bytes in flight per SM were 1/taps of what the load queues held — 305 us for the QUAN-YOLO11n stem (0.8 TB/s),
// 5.8 ms for the 7x7 Q-ResNet-34 stem.  Block fold in shared memory, then one fp32 atomic per weight and block into the
// zero-initialised dW.  CO is a chunk of the tensor's g.CT output channels (blockIdx.y).
constexpr int SW_TILE = 256;
template <typename T, int CI, int CO, int U>
__global__ void __launch_bounds__(256, 2) qconv_small_wgrad_kernel(const T* __restrict__ dy, const T* __restrict__ x, float* dw0,
                                                                float* dw1, float* dw2, float* dw3, SmallGeom g, Mix16 M) {
  pdl_prologue();
  constexpr int NACC = 4 * CO * CI;
  constexpr int GROW = 4 * CO + 4;                   // padded row: the pixel lanes of a warp hit different banks
  __shared__ __align__(16) float Gs[SW_TILE][GROW];
  __shared__ int org_h[SW_TILE], org_w[SW_TILE], org_b[SW_TILE];
  __shared__ float red[256];
  const int taps = g.kH * g.kW;
  const int lanes = blockDim.x / taps;               // pixel lanes per block
  const int tap = threadIdx.x % taps, pl = threadIdx.x / taps;
  const int kh = tap / g.kW, kw = tap - kh * g.kW;
  const int dh = kh * g.dH, dw_ = kw * g.dW;
  const int co0 = (int)blockIdx.y * CO;
  const int npix = g.B * g.Ho * g.Wo;                // < 2^31 (qconv_small_supported)
  const int ntiles = (npix + SW_TILE - 1) / SW_TILE;
  float acc[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) acc[i] = 0.f;
  // phase A is split in two: `fetch` issues the loads of the block's NEXT tile before phase B of the current one, `park`
  // mixes and stores them afterwards — the dY latency hides behind phase B
  Vec<T, CO> raw[4];
  int f_h = 0, f_w = 0, f_b = 0;
  bool f_ok = false;
  auto fetch = [&](int tile) {
    const int pix = tile * SW_TILE + (int)threadIdx.x;
    f_ok = pix < npix;
    if (f_ok) {
      const int wo = pix % g.Wo, r = pix / g.Wo;
      const int ho = r % g.Ho;
      f_b = r / g.Ho;
      f_h = ho * g.sH - g.pH;
      f_w = wo * g.sW - g.pW;
#pragma unroll
      for (int p = 0; p < 4; ++p) raw[p] = *reinterpret_cast<const Vec<T, CO>*>(dy + ((int64_t)pix * 4 + p) * g.CT + co0);
    }
  };
  auto park = [&]() {
    if (f_ok) {
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int co = 0; co < CO; ++co)
          Gs[threadIdx.x][q * CO + co] = M.m[0 * 4 + q] * to_f32(raw[0].v[co]) + M.m[1 * 4 + q] * to_f32(raw[1].v[co]) +
                                         M.m[2 * 4 + q] * to_f32(raw[2].v[co]) + M.m[3 * 4 + q] * to_f32(raw[3].v[co]);
      org_h[threadIdx.x] = f_h;
      org_w[threadIdx.x] = f_w;
      org_b[threadIdx.x] = f_b;
    }
  };
  if ((int)blockIdx.x < ntiles) fetch(blockIdx.x);
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    __syncthreads();                                 // the previous tile's readers are done
    park();
    __syncthreads();
    if (tile + (int)gridDim.x < ntiles) fetch(tile + gridDim.x);
    if (pl < lanes) {
      const int n = min(SW_TILE, npix - tile * SW_TILE);
      for (int i0 = pl; i0 < n; i0 += U * lanes) {
        float xv[U][4 * CI];
        bool ok[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int i = i0 + u * lanes;
          ok[u] = false;
          if (i < n) {
            const int hi = org_h[i] + dh, wi = org_w[i] + dw_;
            ok[u] = hi >= 0 && hi < g.H && wi >= 0 && wi < g.W;
            if (ok[u]) load_pixel<T, 4 * CI>(x + (((int64_t)org_b[i] * g.H + hi) * g.W + wi) * 4 * CI, xv[u]);
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if (!ok[u]) continue;
          const float* gr = &Gs[i0 + u * lanes][0];
#pragma unroll
          for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int co = 0; co < CO; ++co) {
              const float gq = gr[q * CO + co];
#pragma unroll
              for (int ci = 0; ci < CI; ++ci)
                acc[(q * CO + co) * CI + ci] = fmaf(gq, xv[u][q * CI + ci], acc[(q * CO + co) * CI + ci]);
            }
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < NACC; ++i) {
    __syncthreads();
    red[threadIdx.x] = pl < lanes ? acc[i] : 0.f;
    __syncthreads();
    if (threadIdx.x < taps) {
      float s = 0.f;
      for (int l = 0; l < lanes; ++l) s += red[l * taps + threadIdx.x];
      const int q = i / (CO * CI), co = (i / CI) % CO, ci = i % CI;
      float* dw = q == 0 ? dw0 : q == 1 ? dw1 : q == 2 ? dw2 : dw3;
      atomicAdd(dw + ((int64_t)(co0 + co) * CI + ci) * taps + threadIdx.x, s);
    }
  }
}

// wgrad, first form (QUAN_SMALL_WGRAD=0; every tap thread loads and mixes dY itself): dW_q[co][ci][tap] = sum_pix G_q[co](pix) x_q[ci](pix (+) tap).  A thread owns ONE tap and 4*CO*CI accumulators;
// a block is `taps` x (blockDim / taps) pixel lanes, so the taps of a pixel share its dY load through L1.  Block fold in
// shared memory, then one fp32 atomic per weight and block into the zero-initialised dW.
template <typename T, int CI, int CO>
__global__ void __launch_bounds__(256) qconv_small_wgrad_pertap_kernel(const T* __restrict__ dy, const T* __restrict__ x, float* dw0,
                                                                float* dw1, float* dw2, float* dw3, SmallGeom g, Mix16 M) {
  pdl_prologue();
  constexpr int NACC = 4 * CO * CI;
  __shared__ float red[256];
  const int taps = g.kH * g.kW;
  const int lanes = blockDim.x / taps;               // pixel lanes per block
  const int tap = threadIdx.x % taps, pl = threadIdx.x / taps;
  const int kh = tap / g.kW, kw = tap - kh * g.kW;
  const int64_t npix = (int64_t)g.B * g.Ho * g.Wo;
  float acc[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) acc[i] = 0.f;
  if (pl < lanes) {
    for (int64_t pix = (int64_t)blockIdx.x * lanes + pl; pix < npix; pix += (int64_t)gridDim.x * lanes) {
      const int wo = (int)(pix % g.Wo);
      const int64_t r = pix / g.Wo;
      const int ho = (int)(r % g.Ho);
      const int b = (int)(r / g.Ho);
      const int hi = ho * g.sH - g.pH + kh * g.dH, wi = wo * g.sW - g.pW + kw * g.dW;
      if (hi < 0 || hi >= g.H || wi < 0 || wi >= g.W) continue;
      float gy[4 * CO], xv[4 * CI];
      load_pixel<T, 4 * CO>(dy + pix * 4 * CO, gy);
      load_pixel<T, 4 * CI>(x + (((int64_t)b * g.H + hi) * g.W + wi) * 4 * CI, xv);
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int co = 0; co < CO; ++co) {
          const float gq = M.m[0 * 4 + q] * gy[co] + M.m[1 * 4 + q] * gy[CO + co] + M.m[2 * 4 + q] * gy[2 * CO + co] +
                           M.m[3 * 4 + q] * gy[3 * CO + co];
#pragma unroll
          for (int ci = 0; ci < CI; ++ci) acc[(q * CO + co) * CI + ci] = fmaf(gq, xv[q * CI + ci], acc[(q * CO + co) * CI + ci]);
        }
    }
  }
#pragma unroll
  for (int i = 0; i < NACC; ++i) {
    __syncthreads();
    red[threadIdx.x] = pl < lanes ? acc[i] : 0.f;
    __syncthreads();
    if (threadIdx.x < taps) {
      float s = 0.f;
      for (int l = 0; l < lanes; ++l) s += red[l * taps + threadIdx.x];
      const int q = i / (CO * CI), co = (i / CI) % CO, ci = i % CI;
      float* dw = q == 0 ? dw0 : q == 1 ? dw1 : q == 2 ? dw2 : dw3;
      atomicAdd(dw + ((int64_t)co * CI + ci) * taps + threadIdx.x, s);
    }
  }
}

// ---- host ------------------------------------------------------------------------------------------------------------
static bool small_c(int c) { return c == 1 || c == 2 || c == 4 || c == 8; }

// output channels one block handles: the whole C_o when it is small, else (1..2 input channels, forward / wgrad) chunks of 8
static int small_co_chunk(const quan_conv_dims& d, int pass) {
  if (small_c(d.Co) && d.Ci * d.Co <= 16) return d.Co;   // register budget: 4*CO accumulators x CI weights per tap
  if (pass != PASS_DGRAD && d.Ci <= 2 && d.Co % 8 == 0 && d.Co <= 512) return 8;
  return 0;
}

bool qconv_small_supported(const quan_conv_dims& d, int dtype, int layout, int pass) {
  (void)dtype;
  if (layout != QUAN_LAYOUT_BHWQC || d.groups != 1) return false;
  if (!small_c(d.Ci) || small_co_chunk(d, pass) == 0) return false;
  if (d.kH * d.kW > 49) return false;
  if ((int64_t)d.B * conv_out(d.H, d.kH, d.sH, d.pH, d.dH) * conv_out(d.W, d.kW, d.sW, d.pW, d.dW) >= (1ll << 31) ||
      (int64_t)d.B * d.H * d.W >= (1ll << 31))
    return false;
  return true;
}

#define QUAN_SMALL_DISPATCH(A, B, CALL)                                                                      \
  switch ((A) * 16 + (B)) {                                                                                  \
    case 1 * 16 + 1: { constexpr int kA = 1, kB = 1; CALL; } break;                                          \
    case 1 * 16 + 2: { constexpr int kA = 1, kB = 2; CALL; } break;                                          \
    case 1 * 16 + 4: { constexpr int kA = 1, kB = 4; CALL; } break;                                          \
    case 1 * 16 + 8: { constexpr int kA = 1, kB = 8; CALL; } break;                                          \
    case 2 * 16 + 1: { constexpr int kA = 2, kB = 1; CALL; } break;                                          \
    case 2 * 16 + 2: { constexpr int kA = 2, kB = 2; CALL; } break;                                          \
    case 2 * 16 + 4: { constexpr int kA = 2, kB = 4; CALL; } break;                                          \
    case 2 * 16 + 8: { constexpr int kA = 2, kB = 8; CALL; } break;                                          \
    case 4 * 16 + 1: { constexpr int kA = 4, kB = 1; CALL; } break;                                          \
    case 4 * 16 + 2: { constexpr int kA = 4, kB = 2; CALL; } break;                                          \
    case 4 * 16 + 4: { constexpr int kA = 4, kB = 4; CALL; } break;                                          \
    case 8 * 16 + 1: { constexpr int kA = 8, kB = 1; CALL; } break;                                          \
    case 8 * 16 + 2: { constexpr int kA = 8, kB = 2; CALL; } break;                                          \
    default: set_error("qconv small: channel pair %d -> %d is not instantiated", (A), (B)); return QUAN_E_UNSUPPORTED; \
  }

template <typename T, bool TRANSPOSED>
static int small_launch_t(const void* in, const float* const w[4], const float* bias_r, void* out, const quan_conv_dims& d,
                          const Mix16& M, cudaStream_t st) {
  const SmallGeom g = make_small_geom(d);
  const int chunk = small_co_chunk(d, TRANSPOSED ? PASS_DGRAD : PASS_FWD);
  QUAN_REQUIRE(chunk > 0, QUAN_E_UNSUPPORTED, "qconv small: %d -> %d channels do not qualify", d.Ci, d.Co);
  const int cin = TRANSPOSED ? d.Co : d.Ci, cout = TRANSPOSED ? d.Ci : chunk;
  const size_t smem = (size_t)d.kH * d.kW * 4 * d.Ci * chunk * sizeof(float);
  const int64_t opix = (int64_t)g.B * (TRANSPOSED ? g.H * g.W : g.Ho * g.Wo);
  const dim3 grid((unsigned)grid_for(opix, 128, 16), (unsigned)(d.Co / chunk));
  W4s w4 = {{w[0], w[1], w[2], w[3]}};
  QUAN_TIMED(st);
  QUAN_SMALL_DISPATCH(cin, cout, (QUAN_LAUNCH((qconv_small_kernel<T, kA, kB, TRANSPOSED>), grid, 128, smem, st, 
                                     (const T*)in, w4, bias_r, (T*)out, g, M)));
  QUAN_CHECK_LAUNCH(TRANSPOSED ? "qconv_small_dgrad" : "qconv_small_fwd");
  return QUAN_OK;
}

int qconv_small_fwd(const void* x, const float* const w[4], const float* bias_r, void* y, const quan_conv_dims& d, int dtype,
                    const float* mix, cudaStream_t st) {
  const Mix16 M = make_mix(mix);
  if (dtype == QUAN_BF16) return small_launch_t<__nv_bfloat16, false>(x, w, bias_r, y, d, M, st);
  return small_launch_t<float, false>(x, w, bias_r, y, d, M, st);
}
int qconv_small_dgrad(const void* dy, const float* const w[4], void* dx, const quan_conv_dims& d, int dtype, const float* mix,
                      cudaStream_t st) {
  const Mix16 M = make_mix(mix);
  if (dtype == QUAN_BF16) return small_launch_t<__nv_bfloat16, true>(dy, w, nullptr, dx, d, M, st);
  return small_launch_t<float, true>(dy, w, nullptr, dx, d, M, st);
}

// grid: every block resident at once (occupancy queried once per instantiation), >= 2 tiles per block
template <typename T, int CI, int CO, int U>
static int small_wgrad_launch(const T* dy, const T* x, float* const dw[4], const SmallGeom& g, const Mix16& M, int nchunks,
                              cudaStream_t st) {
  auto kern = qconv_small_wgrad_kernel<T, CI, CO, U>;
  static thread_local int occ = 0;
  if (occ == 0) {
    QUAN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 256, 0));
    if (occ < 1) occ = 1;
  }
  const int64_t npix = (int64_t)g.B * g.Ho * g.Wo;
  int64_t blocks = ceil_div64(npix, (int64_t)SW_TILE * 2);
  int64_t cap = (int64_t)QUAN_NUM_SMS * occ / nchunks;
  if (cap < 1) cap = 1;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  QUAN_LAUNCH((kern), dim3((unsigned)blocks, (unsigned)nchunks), 256, 0, st, dy, x, dw[0], dw[1], dw[2], dw[3], g, M);
  return QUAN_OK;
}

template <typename T>
static int small_wgrad_t(const void* dy, const void* x, float* const dw[4], const quan_conv_dims& d, const Mix16& M,
                         cudaStream_t st) {
  const SmallGeom g = make_small_geom(d);
  const int taps = d.kH * d.kW;
  for (int q = 0; q < 4; ++q) QUAN_CUDA(cudaMemsetAsync(dw[q], 0, (size_t)d.Co * d.Ci * taps * sizeof(float), st));
  const int chunk = small_co_chunk(d, PASS_WGRAD);
  QUAN_REQUIRE(chunk > 0, QUAN_E_UNSUPPORTED, "qconv small wgrad: %d -> %d channels do not qualify", d.Ci, d.Co);
  QUAN_TIMED(st);
  // x pixels in flight per thread: as many as the register file takes next to the 4*CO*CI accumulators
  int rc = QUAN_OK;
  static const int env_form = [] { const char* e = getenv("QUAN_SMALL_WGRAD"); return e ? atoi(e) : 1; }();
  // measured (tools/narrow_wgrad_probe.py, B200, bf16): staged form 248 vs 373 us for 1 -> 8 (YOLO11s stem), 90 vs 105 us for
  // 2 -> 4, but 403 vs 309 us for the YOLO11n stem 1 -> 4: few accumulators per tap thread leave the per-tap form ahead
  if (chunk == d.Co && (env_form == 0 || (env_form == 1 && d.Ci * d.Co < 8))) {
    const int lanes = 256 / taps;
    const int64_t npix = (int64_t)g.B * g.Ho * g.Wo;
    int64_t blocks = ceil_div64(npix, (int64_t)lanes * 32);
    if (blocks > QUAN_NUM_SMS * 4) blocks = QUAN_NUM_SMS * 4;
    if (blocks < 1) blocks = 1;
    QUAN_SMALL_DISPATCH(d.Ci, d.Co, (QUAN_LAUNCH((qconv_small_wgrad_pertap_kernel<T, kA, kB>), (unsigned)blocks, 256, 0, st, 
                                        (const T*)dy, (const T*)x, dw[0], dw[1], dw[2], dw[3], g, M)));
    QUAN_CHECK_LAUNCH("qconv_small_wgrad");
    return QUAN_OK;
  }
  QUAN_SMALL_DISPATCH(d.Ci, chunk, (rc = small_wgrad_launch<T, kA, kB, (kA <= 2 ? 4 : kA == 4 ? 2 : 1)>(
                                       (const T*)dy, (const T*)x, dw, g, M, d.Co / chunk, st)));
  if (rc) return rc;
  QUAN_CHECK_LAUNCH("qconv_small_wgrad");
  return QUAN_OK;
}
int qconv_small_wgrad(const void* dy, const void* x, float* const dw[4], const quan_conv_dims& d, int dtype, const float* mix,
                      cudaStream_t st) {
  const Mix16 M = make_mix(mix);
  if (dtype == QUAN_BF16) return small_wgrad_t<__nv_bfloat16>(dy, x, dw, d, M, st);
  return small_wgrad_t<float>(dy, x, dw, d, M, st);
}

}  // namespace quan
