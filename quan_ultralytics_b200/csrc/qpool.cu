// qpool.cu — QuaternionMaxPool forward / backward.
//
// Reference: ultralytics/nn/modules/block.py:85-109 ≡ classification/models/blocks/quaternion_blocks.py:236-260 — a
// Python loop of four nn.MaxPool2d calls on strided component slices plus a torch.stack copy (≈ 4 slice copies + 4 pool
// kernels + 1 stack: 3 extra full-tensor HBM round trips); backward = autograd of that (4 scatter kernels).  Users: QSPPF
// (block.py:270-302: k=5, s=1, p=2, three times on the P5 map) and the Q-ResNet stems (quaternion_models.py:193,236:
// k=3, s=2, p=1 on 112^2).
//
// Here: the four components of every channel are independent real planes, so in either layout the tensor is
// [outer][H][W][inner] with `inner` contiguous elements pooled element-wise (BCHWQ: outer = B*C, inner = 4;
// BHWQC: outer = B, inner = 4C).  A thread owns one 16-byte vector of `inner` at one output pixel: HBM-bound,
// algorithmic bytes fwd = S_in + S_out (+ S_out/sizeof(T) index bytes when training), bwd = S_dy + idx + S_dx.
//
// nn.MaxPool2d semantics kept exactly (aten/native/cuda/DilatedMaxPool2d.cu behaviour): padding counts as -inf,
// window scanned row-major, a later element replaces the running maximum only if it is strictly greater or NaN — so
// the FIRST maximum wins ties (frequent in bf16) and the gradient goes to that element alone.  The forward stores the
// winning tap (kh*kW + kw, one byte per element); the backward is a gather — each input element sums dy over the
// windows that contain it and whose stored tap points at it — deterministic, no atomics.
#include <cstdint>

#include "common.cuh"
#include "quan_sm100.h"

namespace quan {

struct PoolGeom {
  int64_t outer;
  int H, W, Ho, Wo, inner_vecs;
  int kH, kW, sH, sW, pH, pW;
  int iv_shift;   // log2(inner_vecs) when it is a power of two, else -1
  int rb;         // row kernels: consecutive rows per block
};

// row kernels: one block per (outer, row) — the row / image split is block-uniform scalar work, a thread only splits its
// element index into (pixel, vector) and divides by the stride.  The flat kernels spend ~7 integer divisions per thread.
__device__ __forceinline__ void split_iv(int e, const PoolGeom& g, int& pix, int& iv) {
  if (g.iv_shift >= 0) { pix = e >> g.iv_shift; iv = e & (g.inner_vecs - 1); }
  else { pix = e / g.inner_vecs; iv = e - pix * g.inner_vecs; }
}
__device__ __forceinline__ int div_stride(int a, int s) { return s == 1 ? a : s == 2 ? a >> 1 : a / s; }   // a >= 0

// I = int (tensors below 2^31 vectors: 32-bit index arithmetic) or int64_t.  Loads go out in batches of POOL_CH before
// anything is compared: the first version walked the window with one dependent load per tap (9 serial latencies for the
// 3x3 stem pool: 283 us on 256x16x112^2 bf16, 2.0 TB/s) and did three 64-bit divisions per thread.
constexpr int POOL_CH = 9;

template <typename T, int V, bool WITH_IDX, typename I>
__global__ void __launch_bounds__(256) qmaxpool_fwd_kernel(const T* __restrict__ x, T* __restrict__ y,
                                                           uint8_t* __restrict__ idx, PoolGeom g) {
  pdl_prologue();
  const I total = (I)(g.outer * g.Ho * g.Wo * g.inner_vecs);
  const int taps = g.kH * g.kW;
  for (I i = (I)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (I)gridDim.x * blockDim.x) {
    const int iv = (int)(i % g.inner_vecs);
    I r = i / g.inner_vecs;
    const int wo = (int)(r % g.Wo);
    r /= g.Wo;
    const int ho = (int)(r % g.Ho);
    const I o = r / g.Ho;
    float best[V];
    uint8_t arg[V];
#pragma unroll
    for (int v = 0; v < V; ++v) { best[v] = -INFINITY; arg[v] = 0; }
    bool first = true;
    const int h0 = ho * g.sH - g.pH, w0 = wo * g.sW - g.pW;
    const T* base = x + ((int64_t)o * g.H * g.W * g.inner_vecs + iv) * V;
    int kh = 0, kw = 0;
    for (int t0 = 0; t0 < taps; t0 += POOL_CH) {
      Vec<T, V> raw[POOL_CH];
      bool ok[POOL_CH];
#pragma unroll
      for (int j = 0; j < POOL_CH; ++j) {
        const int hi = h0 + kh, wi = w0 + kw;
        ok[j] = (t0 + j < taps) && hi >= 0 && hi < g.H && wi >= 0 && wi < g.W;
        if (ok[j]) raw[j] = *reinterpret_cast<const Vec<T, V>*>(base + ((int64_t)hi * g.W + wi) * g.inner_vecs * V);
        if (++kw == g.kW) { kw = 0; ++kh; }
      }
#pragma unroll
      for (int j = 0; j < POOL_CH; ++j) {
        if (!ok[j]) continue;
        const uint8_t tap = (uint8_t)(t0 + j);
#pragma unroll
        for (int v = 0; v < V; ++v) {
          const float xv = to_f32(raw[j].v[v]);
          // the first valid element always seeds the maximum (PyTorch starts from it: -inf inputs keep their own index)
          if (first || xv > best[v] || xv != xv) { best[v] = xv; arg[v] = tap; }
        }
        first = false;
      }
    }
    store_vec<T, V>(y + (int64_t)i * V, best);
    if constexpr (WITH_IDX) {
      Vec<uint8_t, V> a;
#pragma unroll
      for (int v = 0; v < V; ++v) a.v[v] = arg[v];
      *reinterpret_cast<Vec<uint8_t, V>*>(idx + (int64_t)i * V) = a;
    }
  }
}

// bf16 forward with packed arithmetic: per bf16x2 pair and tap one compare-mask (strictly greater: the first maximum keeps
// its tap), one max and one LOP3 that merges the tap number under the mask — 3 instructions where the scalar kernel above
// spends ~16 (two conversions, two compares, four selects); 283 -> see DESIGN §4.5.  NaN: __hmax2_nan carries it into the
// value; the tap of a NaN result then comes from the scalar rule (last NaN of the scan) on a rare slow path.
template <int V, bool WITH_IDX, typename I>
__global__ void __launch_bounds__(256) qmaxpool_fwd_bf16_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                                                uint8_t* __restrict__ idx, PoolGeom g) {
  pdl_prologue();
  constexpr int NP = V / 2;
  using T = __nv_bfloat16;
  const I total = (I)(g.outer * g.Ho * g.Wo * g.inner_vecs);
  const int taps = g.kH * g.kW;
  for (I i = (I)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (I)gridDim.x * blockDim.x) {
    const int iv = (int)(i % g.inner_vecs);
    I r = i / g.inner_vecs;
    const int wo = (int)(r % g.Wo);
    r /= g.Wo;
    const int ho = (int)(r % g.Ho);
    const I o = r / g.Ho;
    const int h0 = ho * g.sH - g.pH, w0 = wo * g.sW - g.pW;
    // the first valid tap seeds the index (an all -inf window keeps it, like PyTorch)
    const uint32_t t_first = (uint32_t)((h0 < 0 ? -h0 : 0) * g.kW + (w0 < 0 ? -w0 : 0));
    __nv_bfloat162 best[NP];
    uint32_t arg[NP];
    const __nv_bfloat162 ninf = __float2bfloat162_rn(-INFINITY);
#pragma unroll
    for (int p = 0; p < NP; ++p) { best[p] = ninf; arg[p] = t_first * 0x00010001u; }
    const T* base = x + ((int64_t)o * g.H * g.W * g.inner_vecs + iv) * V;
    int kh = 0, kw = 0;
    for (int t0 = 0; t0 < taps; t0 += POOL_CH) {
      Vec<T, V> raw[POOL_CH];
      bool ok[POOL_CH];
#pragma unroll
      for (int j = 0; j < POOL_CH; ++j) {
        const int hi = h0 + kh, wi = w0 + kw;
        ok[j] = (t0 + j < taps) && hi >= 0 && hi < g.H && wi >= 0 && wi < g.W;
        if (ok[j]) raw[j] = *reinterpret_cast<const Vec<T, V>*>(base + ((int64_t)hi * g.W + wi) * g.inner_vecs * V);
        if (++kw == g.kW) { kw = 0; ++kh; }
      }
#pragma unroll
      for (int j = 0; j < POOL_CH; ++j) {
        if (!ok[j]) continue;
        const uint32_t tap2 = (uint32_t)(t0 + j) * 0x00010001u;
        const __nv_bfloat162* xv = reinterpret_cast<const __nv_bfloat162*>(&raw[j]);
#pragma unroll
        for (int p = 0; p < NP; ++p) {
          const uint32_t gt = __hgt2_mask(xv[p], best[p]);
          best[p] = __hmax2_nan(best[p], xv[p]);
          arg[p] = (tap2 & gt) | (arg[p] & ~gt);
        }
      }
    }
    bool has_nan = false;
#pragma unroll
    for (int p = 0; p < NP; ++p) has_nan |= (__hneu2_mask(best[p], best[p]) != 0u);   // unordered compare: true for NaN
    Vec<uint8_t, V> a;
#pragma unroll
    for (int p = 0; p < NP; ++p) { a.v[2 * p] = (uint8_t)(arg[p] & 0xffu); a.v[2 * p + 1] = (uint8_t)((arg[p] >> 16) & 0xffu); }
    if (WITH_IDX && has_nan) {
      // scalar rule for the lanes whose result is NaN: the last NaN of the row-major scan holds the index
#pragma unroll
      for (int v = 0; v < V; ++v) {
        const float bvf = (v & 1) ? __high2float(best[v / 2]) : __low2float(best[v / 2]);
        if (!(bvf != bvf)) continue;
        uint8_t last = a.v[v];
        for (int t = 0; t < taps; ++t) {
          const int hi = h0 + t / g.kW, wi = w0 + t % g.kW;
          if (hi < 0 || hi >= g.H || wi < 0 || wi >= g.W) continue;
          const float xs = __bfloat162float(base[((int64_t)hi * g.W + wi) * g.inner_vecs * V + v]);
          if (xs != xs) last = (uint8_t)t;
        }
        a.v[v] = last;
      }
    }
    Vec<T, V> out;
#pragma unroll
    for (int p = 0; p < NP; ++p) { out.v[2 * p] = __low2bfloat16(best[p]); out.v[2 * p + 1] = __high2bfloat16(best[p]); }
    *reinterpret_cast<Vec<T, V>*>(y + (int64_t)i * V) = out;
    if constexpr (WITH_IDX) *reinterpret_cast<Vec<uint8_t, V>*>(idx + (int64_t)i * V) = a;
  }
}

template <typename T, int V, typename I>
__global__ void __launch_bounds__(256) qmaxpool_bwd_kernel(const T* __restrict__ dy, const uint8_t* __restrict__ idx,
                                                           T* __restrict__ dx, PoolGeom g) {
  pdl_prologue();
  const I total = (I)(g.outer * g.H * g.W * g.inner_vecs);
  for (I i = (I)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (I)gridDim.x * blockDim.x) {
    const int iv = (int)(i % g.inner_vecs);
    I r = i / g.inner_vecs;
    const int wi = (int)(r % g.W);
    r /= g.W;
    const int hi = (int)(r % g.H);
    const I o = r / g.H;
    float acc[V];
#pragma unroll
    for (int v = 0; v < V; ++v) acc[v] = 0.f;
    // output rows whose window holds hi: ho*sH - pH <= hi <= ho*sH - pH + kH - 1
    int ho_lo = hi + g.pH - g.kH + 1;
    ho_lo = ho_lo <= 0 ? 0 : (ho_lo + g.sH - 1) / g.sH;
    int ho_hi = (hi + g.pH) / g.sH;
    if (ho_hi > g.Ho - 1) ho_hi = g.Ho - 1;
    int wo_lo = wi + g.pW - g.kW + 1;
    wo_lo = wo_lo <= 0 ? 0 : (wo_lo + g.sW - 1) / g.sW;
    int wo_hi = (wi + g.pW) / g.sW;
    if (wo_hi > g.Wo - 1) wo_hi = g.Wo - 1;
    // one window at a time; the index byte decides whether the gradient vector is loaded at all.  Measured on the Q-ResNet
    // stem pool (256x16x112^2, bf16): this loop 636 us; index + gradient of ALL windows requested up front 1359 us; index
    // vectors of a batch first, then the needed gradient vectors 976 us — the simple dependent form stays.
    const int64_t obase = ((int64_t)o * g.Ho * g.Wo * g.inner_vecs + iv) * V;
    for (int ho = ho_lo; ho <= ho_hi; ++ho) {
      const int kh = hi - (ho * g.sH - g.pH);
      for (int wo = wo_lo; wo <= wo_hi; ++wo) {
        const int tap = kh * g.kW + (wi - (wo * g.sW - g.pW));
        const int64_t e = obase + ((int64_t)ho * g.Wo + wo) * g.inner_vecs * V;
        const Vec<uint8_t, V> a = *reinterpret_cast<const Vec<uint8_t, V>*>(idx + e);
        bool any = false;
#pragma unroll
        for (int v = 0; v < V; ++v) any |= (a.v[v] == tap);
        if (!any) continue;
        float gv[V];
        load_vec<T, V>(dy + e, gv);
#pragma unroll
        for (int v = 0; v < V; ++v)
          if (a.v[v] == tap) acc[v] += gv[v];
      }
    }
    store_vec<T, V>(dx + (int64_t)i * V, acc);
  }
}

template <int V, bool WITH_IDX>
__global__ void __launch_bounds__(256) qmaxpool_fwd_bf16_rows_kernel(const __nv_bfloat16* __restrict__ x,
                                                                     __nv_bfloat16* __restrict__ y, uint8_t* __restrict__ idx,
                                                                     PoolGeom g) {
  pdl_prologue();
  constexpr int NP = V / 2;
  using T = __nv_bfloat16;
  const int taps = g.kH * g.kW;
  const int rowlen = g.Wo * g.inner_vecs;
  const __nv_bfloat162 ninf = __float2bfloat162_rn(-INFINITY);
  const int64_t nrows = g.outer * g.Ho;
  // g.rb consecutive output rows per block: consecutive windows share kH - sH input rows through L1
  for (int64_t row = (int64_t)blockIdx.x * g.rb; row < nrows && row < ((int64_t)blockIdx.x + 1) * g.rb; ++row) {
  const int ho = (int)(row % g.Ho);               // row = o * Ho + ho
  const int64_t o = row / g.Ho;
  const int h0 = ho * g.sH - g.pH;
  const T* img = x + (int64_t)o * g.H * g.W * g.inner_vecs * V;
  for (int e = threadIdx.x; e < rowlen; e += blockDim.x) {
    int wo, iv;
    split_iv(e, g, wo, iv);
    const int w0 = wo * g.sW - g.pW;
    const uint32_t t_first = (uint32_t)((h0 < 0 ? -h0 : 0) * g.kW + (w0 < 0 ? -w0 : 0));
    __nv_bfloat162 best[NP];
    uint32_t arg[NP];
#pragma unroll
    for (int p = 0; p < NP; ++p) { best[p] = ninf; arg[p] = t_first * 0x00010001u; }
    const T* base = img + (int64_t)iv * V;
    int kh = 0, kw = 0;
    for (int t0 = 0; t0 < taps; t0 += POOL_CH) {
      Vec<T, V> raw[POOL_CH];
      bool ok[POOL_CH];
#pragma unroll
      for (int j = 0; j < POOL_CH; ++j) {
        const int hi = h0 + kh, wi = w0 + kw;
        ok[j] = (t0 + j < taps) && hi >= 0 && hi < g.H && wi >= 0 && wi < g.W;
        if (ok[j]) raw[j] = *reinterpret_cast<const Vec<T, V>*>(base + ((int64_t)hi * g.W + wi) * g.inner_vecs * V);
        if (++kw == g.kW) { kw = 0; ++kh; }
      }
#pragma unroll
      for (int j = 0; j < POOL_CH; ++j) {
        if (!ok[j]) continue;
        const uint32_t tap2 = (uint32_t)(t0 + j) * 0x00010001u;
        const __nv_bfloat162* xv = reinterpret_cast<const __nv_bfloat162*>(&raw[j]);
#pragma unroll
        for (int p = 0; p < NP; ++p) {
          const uint32_t gt = __hgt2_mask(xv[p], best[p]);
          best[p] = __hmax2_nan(best[p], xv[p]);
          arg[p] = (tap2 & gt) | (arg[p] & ~gt);
        }
      }
    }
    bool has_nan = false;
#pragma unroll
    for (int p = 0; p < NP; ++p) has_nan |= (__hneu2_mask(best[p], best[p]) != 0u);
    Vec<uint8_t, V> a;
#pragma unroll
    for (int p = 0; p < NP; ++p) { a.v[2 * p] = (uint8_t)(arg[p] & 0xffu); a.v[2 * p + 1] = (uint8_t)((arg[p] >> 16) & 0xffu); }
    if (WITH_IDX && has_nan) {
#pragma unroll
      for (int v = 0; v < V; ++v) {
        const float bvf = (v & 1) ? __high2float(best[v / 2]) : __low2float(best[v / 2]);
        if (!(bvf != bvf)) continue;
        uint8_t last = a.v[v];
        for (int t = 0; t < taps; ++t) {
          const int hi = h0 + t / g.kW, wi = w0 + t % g.kW;
          if (hi < 0 || hi >= g.H || wi < 0 || wi >= g.W) continue;
          const float xs = __bfloat162float(base[((int64_t)hi * g.W + wi) * g.inner_vecs * V + v]);
          if (xs != xs) last = (uint8_t)t;
        }
        a.v[v] = last;
      }
    }
    Vec<T, V> out;
#pragma unroll
    for (int p = 0; p < NP; ++p) { out.v[2 * p] = __low2bfloat16(best[p]); out.v[2 * p + 1] = __high2bfloat16(best[p]); }
    const int64_t oi = (row * rowlen + e) * V;
    *reinterpret_cast<Vec<T, V>*>(y + oi) = out;
    if constexpr (WITH_IDX) *reinterpret_cast<Vec<uint8_t, V>*>(idx + oi) = a;
  }
  }
}

constexpr int POOL_RB = 8;

template <typename T, int V>
__global__ void __launch_bounds__(256) qmaxpool_bwd_rows_kernel(const T* __restrict__ dy, const uint8_t* __restrict__ idx,
                                                                T* __restrict__ dx, PoolGeom g) {
  pdl_prologue();
  // g.rb (<= POOL_RB) consecutive input rows per block: the 1 + (kH-1)/sH output rows a row needs are the next row's too, so the
  // index / gradient vectors come from L1 instead of L2 (one row per block: 1348 us on the Q-ResNet stem pool — three
  // blocks on three SMs each pulled the same output rows through L2)
  const int64_t nrows = g.outer * g.H;
  const int rowlen = g.W * g.inner_vecs;
  for (int64_t row = (int64_t)blockIdx.x * g.rb; row < nrows && row < ((int64_t)blockIdx.x + 1) * g.rb; ++row) {
  const int hi = (int)(row % g.H);
  const int64_t o = row / g.H;
  int ho_lo = hi + g.pH - g.kH + 1;
  ho_lo = ho_lo <= 0 ? 0 : (ho_lo + g.sH - 1) / g.sH;
  int ho_hi = (hi + g.pH) / g.sH;
  if (ho_hi > g.Ho - 1) ho_hi = g.Ho - 1;
  const int64_t img = (int64_t)o * g.Ho * g.Wo * g.inner_vecs * V;
  for (int e = threadIdx.x; e < rowlen; e += blockDim.x) {
    int wi, iv;
    split_iv(e, g, wi, iv);
    float acc[V];
#pragma unroll
    for (int v = 0; v < V; ++v) acc[v] = 0.f;
    int wo_lo = wi + g.pW - g.kW + 1;
    wo_lo = wo_lo <= 0 ? 0 : div_stride(wo_lo + g.sW - 1, g.sW);
    int wo_hi = div_stride(wi + g.pW, g.sW);
    if (wo_hi > g.Wo - 1) wo_hi = g.Wo - 1;
    const int64_t obase = img + (int64_t)iv * V;
    for (int ho = ho_lo; ho <= ho_hi; ++ho) {
      const int kh = hi - (ho * g.sH - g.pH);
      for (int wo = wo_lo; wo <= wo_hi; ++wo) {
        const int tap = kh * g.kW + (wi - (wo * g.sW - g.pW));
        const int64_t oe = obase + ((int64_t)ho * g.Wo + wo) * g.inner_vecs * V;
        const Vec<uint8_t, V> a = *reinterpret_cast<const Vec<uint8_t, V>*>(idx + oe);
        bool any = false;
#pragma unroll
        for (int v = 0; v < V; ++v) any |= (a.v[v] == tap);
        if (!any) continue;
        float gv[V];
        load_vec<T, V>(dy + oe, gv);
#pragma unroll
        for (int v = 0; v < V; ++v)
          if (a.v[v] == tap) acc[v] += gv[v];
      }
    }
    store_vec<T, V>(dx + (row * rowlen + e) * V, acc);
  }
  }
}

// Tiled backward: a block owns TH x TW input pixels x cv channel vectors.  The index and gradient vectors of every window that touches
// the tile are staged in shared memory first — all global loads of the block are independent and in flight together — and the gather
// then runs out of shared memory.  The kernels above walk the windows with a dependent (index -> test -> gradient) load pair per
// window: 25 serial L2 round trips for the 5x5 QSPPF pool (36 us for a 10 MB map) and a 9x re-read of index / gradient vectors through
// L2 -> L1 on the stem pool (profiles/r01_pool_probe.txt).  Same arithmetic and summation order (window-major, row by row), so the
// result is bit-identical to the flat kernel.
struct PoolTile {
  int TH, TW, cv, cv_shift;     // input pixels of a tile, channel vectors per block (power of two)
  int oth, otw;                 // staged output rows / columns (upper bound of what a tile can touch)
  int nth, ntw, nchunks;        // tiles per image plane, channel chunks
};

__device__ __forceinline__ void pool_out_range(int i, int n, int k, int s, int p, int no, int& lo, int& hi) {
  // outputs whose window holds input coordinate range [i, i + n): first and last
  lo = i + p - k + 1;
  lo = lo <= 0 ? 0 : div_stride(lo + s - 1, s);
  hi = div_stride(i + n - 1 + p, s);
  if (hi > no - 1) hi = no - 1;
}

// byte-wise equality of the V tap indices of a vector with `tap`: one SIMD compare per four elements instead of V extract-and-compare
// pairs (the scalar form was ~30 instructions per window and vector: the k5 QSPPF pool spent its time there, not on loads)
template <int V> struct TapMask;
template <> struct TapMask<8> {
  uint32_t m0, m1;
  __device__ __forceinline__ TapMask(const void* p, uint32_t tap4) {
    const uint2 a = *reinterpret_cast<const uint2*>(p);
    m0 = __vcmpeq4(a.x, tap4);
    m1 = __vcmpeq4(a.y, tap4);
  }
  __device__ __forceinline__ bool any() const { return (m0 | m1) != 0u; }
  __device__ __forceinline__ bool hit(int v) const { return ((v < 4 ? m0 : m1) >> (8 * (v & 3))) & 1u; }
};
template <> struct TapMask<4> {
  uint32_t m0;
  __device__ __forceinline__ TapMask(const void* p, uint32_t tap4) { m0 = __vcmpeq4(*reinterpret_cast<const uint32_t*>(p), tap4); }
  __device__ __forceinline__ bool any() const { return m0 != 0u; }
  __device__ __forceinline__ bool hit(int v) const { return (m0 >> (8 * v)) & 1u; }
};

template <typename T, int V>
__global__ void __launch_bounds__(256) qmaxpool_bwd_tile_kernel(const T* __restrict__ dy, const uint8_t* __restrict__ idx,
                                                                T* __restrict__ dx, PoolGeom g, PoolTile t) {
  pdl_prologue();
  extern __shared__ __align__(16) uint8_t pool_smem[];
  Vec<T, V>* s_dy = reinterpret_cast<Vec<T, V>*>(pool_smem);
  Vec<uint8_t, V>* s_ix = reinterpret_cast<Vec<uint8_t, V>*>(pool_smem + (size_t)t.oth * t.otw * t.cv * sizeof(Vec<T, V>));
  uint32_t b = blockIdx.x;
  const int chunk = (int)(b % (uint32_t)t.nchunks); b /= (uint32_t)t.nchunks;
  const int tw = (int)(b % (uint32_t)t.ntw); b /= (uint32_t)t.ntw;
  const int th = (int)(b % (uint32_t)t.nth);
  const int64_t o = b / (uint32_t)t.nth;
  const int h0 = th * t.TH, w0 = tw * t.TW, c0 = chunk * t.cv;
  const int nh = min(t.TH, g.H - h0), nw = min(t.TW, g.W - w0), ncv = min(t.cv, g.inner_vecs - c0);
  int ho0, ho1, wo0, wo1;
  pool_out_range(h0, nh, g.kH, g.sH, g.pH, g.Ho, ho0, ho1);
  pool_out_range(w0, nw, g.kW, g.sW, g.pW, g.Wo, wo0, wo1);
  const int nho = ho1 - ho0 + 1, nwo = wo1 - wo0 + 1;
  const int64_t obase = (int64_t)o * g.Ho * g.Wo * g.inner_vecs + c0;
  if (nho > 0 && nwo > 0) {
    const int n = nho * nwo * t.cv;
    for (int e = threadIdx.x; e < n; e += 256) {
      const int c = e & (t.cv - 1);
      const int r = e >> t.cv_shift;
      const int ho = r / nwo, wo = r - ho * nwo;
      if (c < ncv) {
        const int64_t ge = (obase + ((int64_t)(ho0 + ho) * g.Wo + (wo0 + wo)) * g.inner_vecs + c) * V;
        s_dy[(ho * t.otw + wo) * t.cv + c] = *reinterpret_cast<const Vec<T, V>*>(dy + ge);
        s_ix[(ho * t.otw + wo) * t.cv + c] = *reinterpret_cast<const Vec<uint8_t, V>*>(idx + ge);
      }
    }
  }
  __syncthreads();
  // a thread keeps its (pixel column, channel vector) and walks down the tile's rows: the column's window range is computed once
  const int col = threadIdx.x & ((t.TW << t.cv_shift) - 1);       // TW is a power of two
  const int row0 = threadIdx.x / (t.TW << t.cv_shift), rstep = 256 / (t.TW << t.cv_shift);
  const int c = col & (t.cv - 1), wl = col >> t.cv_shift;
  if (wl >= nw || c >= ncv) return;
  const int wi = w0 + wl;
  int b0, b1;
  pool_out_range(wi, 1, g.kW, g.sW, g.pW, g.Wo, b0, b1);
  const int kw_first = wi - (b0 * g.sW - g.pW);                    // tap column in window b0; sW less in every next one
  const int s_col = (b0 - wo0) * t.cv + c;
  const int64_t ibase = (int64_t)o * g.H * g.W * g.inner_vecs + c0 + c;
  for (int hl = row0; hl < nh; hl += rstep) {
    const int hi = h0 + hl;
    int a0, a1;
    pool_out_range(hi, 1, g.kH, g.sH, g.pH, g.Ho, a0, a1);
    float acc[V];
#pragma unroll
    for (int v = 0; v < V; ++v) acc[v] = 0.f;
    int kh = hi - (a0 * g.sH - g.pH);
    for (int ho = a0; ho <= a1; ++ho, kh -= g.sH) {
      int se = (ho - ho0) * t.otw * t.cv + s_col;
      uint32_t tap4 = (uint32_t)(kh * g.kW + kw_first) * 0x01010101u;
      for (int wo = b0; wo <= b1; ++wo, se += t.cv, tap4 -= (uint32_t)g.sW * 0x01010101u) {
        const TapMask<V> m(s_ix + se, tap4);
        if (!m.any()) continue;
        const Vec<T, V> gq = s_dy[se];
#pragma unroll
        for (int v = 0; v < V; ++v)
          if (m.hit(v)) acc[v] += to_f32(gq.v[v]);
      }
    }
    store_vec<T, V>(dx + (ibase + ((int64_t)hi * g.W + wi) * g.inner_vecs) * V, acc);
  }
}

static bool plan_pool_tile(const PoolGeom& g, int V, size_t esz, PoolTile& t, size_t& smem) {
  auto p2floor = [](int v) { int r = 1; while (r * 2 <= v) r *= 2; return r; };
  t.TH = 16;
  t.TW = 16;                                     // powers of two: a thread's column is threadIdx & mask (ragged edges are masked)
  t.cv = p2floor(g.inner_vecs < 8 ? g.inner_vecs : 8);
  for (;;) {
    t.oth = (t.TH + g.kH - 2) / g.sH + 2;
    t.otw = (t.TW + g.kW - 2) / g.sW + 2;
    smem = (size_t)t.oth * t.otw * t.cv * (V * esz + V);
    if (smem <= 40 * 1024) break;
    if (t.cv > 1) t.cv >>= 1;
    else if (t.TH > 4) t.TH >>= 1;
    else if (t.TW > 4) t.TW >>= 1;
    else return false;
  }
  t.cv_shift = 0;
  while ((1 << t.cv_shift) < t.cv) ++t.cv_shift;
  t.nth = (g.H + t.TH - 1) / t.TH;
  t.ntw = (g.W + t.TW - 1) / t.TW;
  t.nchunks = (g.inner_vecs + t.cv - 1) / t.cv;
  const int64_t blocks = g.outer * t.nth * t.ntw * t.nchunks;
  return blocks < (1ll << 31);
}

static int pool_geom(const char* who, int B, int C, int H, int W, int kH, int kW, int sH, int sW, int pH, int pW, int dtype,
                     int layout, PoolGeom& g, int& V) {
  QUAN_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, QUAN_E_ARG, "%s: non-positive dims", who);
  QUAN_REQUIRE(dtype == QUAN_F32 || dtype == QUAN_BF16, QUAN_E_ARG, "%s: bad dtype %d", who, dtype);
  QUAN_REQUIRE(layout == QUAN_LAYOUT_BCHWQ || layout == QUAN_LAYOUT_BHWQC, QUAN_E_ARG, "%s: bad layout %d", who, layout);
  QUAN_REQUIRE(kH >= 1 && kW >= 1 && sH >= 1 && sW >= 1 && pH >= 0 && pW >= 0, QUAN_E_SHAPE, "%s: bad window", who);
  QUAN_REQUIRE(kH * kW <= 255, QUAN_E_UNSUPPORTED, "%s: window %dx%d exceeds the one-byte tap index", who, kH, kW);
  // nn.MaxPool2d: "pad should be at most half of effective kernel size"
  QUAN_REQUIRE(2 * pH <= kH && 2 * pW <= kW, QUAN_E_SHAPE, "%s: padding (%d,%d) exceeds half the window (%d,%d)", who, pH, pW, kH, kW);
  g.H = H; g.W = W;
  g.Ho = (H + 2 * pH - kH) / sH + 1;
  g.Wo = (W + 2 * pW - kW) / sW + 1;
  QUAN_REQUIRE(H + 2 * pH >= kH && W + 2 * pW >= kW && g.Ho > 0 && g.Wo > 0, QUAN_E_SHAPE, "%s: window larger than the padded input", who);
  g.kH = kH; g.kW = kW; g.sH = sH; g.sW = sW; g.pH = pH; g.pW = pW;
  int inner;
  if (layout == QUAN_LAYOUT_BCHWQ) { g.outer = (int64_t)B * C; inner = 4; }
  else { g.outer = B; inner = 4 * C; }
  V = largest_pow2_divisor(inner, dtype == QUAN_BF16 ? 8 : 4);
  g.inner_vecs = inner / V;
  g.iv_shift = -1;
  g.rb = 1;
  for (int b = 0; b < 31; ++b)
    if ((1 << b) == g.inner_vecs) g.iv_shift = b;
  return QUAN_OK;
}

template <typename T>
static int launch_pool_fwd(const void* x, void* y, uint8_t* idx, const PoolGeom& g, int V, cudaStream_t st) {
  const T* xp = reinterpret_cast<const T*>(x);
  T* yp = reinterpret_cast<T*>(y);
  const int grid = grid_for(g.outer * g.Ho * g.Wo * g.inner_vecs, 256, 8);
  QUAN_TIMED(st);
  const bool small = g.outer * g.H * g.W * g.inner_vecs < (1ll << 31) - (1 << 20) && g.outer * g.Ho * g.Wo * g.inner_vecs < (1ll << 31) - (1 << 20);
#define QUAN_POOL_FWD(VV)                                                                                        \
  do {                                                                                                           \
    if (idx != nullptr && small) QUAN_LAUNCH((qmaxpool_fwd_kernel<T, VV, true, int>), grid, 256, 0, st, xp, yp, idx, g);     \
    else if (idx != nullptr) QUAN_LAUNCH((qmaxpool_fwd_kernel<T, VV, true, int64_t>), grid, 256, 0, st, xp, yp, idx, g);     \
    else if (small) QUAN_LAUNCH((qmaxpool_fwd_kernel<T, VV, false, int>), grid, 256, 0, st, xp, yp, idx, g);                 \
    else QUAN_LAUNCH((qmaxpool_fwd_kernel<T, VV, false, int64_t>), grid, 256, 0, st, xp, yp, idx, g);                        \
  } while (0)
#define QUAN_POOL_FWD_BF16(VV)                                                                                   \
  do {                                                                                                           \
    if (idx != nullptr && small) QUAN_LAUNCH((qmaxpool_fwd_bf16_kernel<VV, true, int>), grid, 256, 0, st, xp, yp, idx, g);   \
    else if (idx != nullptr) QUAN_LAUNCH((qmaxpool_fwd_bf16_kernel<VV, true, int64_t>), grid, 256, 0, st, xp, yp, idx, g);   \
    else if (small) QUAN_LAUNCH((qmaxpool_fwd_bf16_kernel<VV, false, int>), grid, 256, 0, st, xp, yp, idx, g);               \
    else QUAN_LAUNCH((qmaxpool_fwd_bf16_kernel<VV, false, int64_t>), grid, 256, 0, st, xp, yp, idx, g);                      \
  } while (0)
  static const int env_packed = [] { const char* e = getenv("QUAN_POOL_PACKED"); return e ? atoi(e) : 1; }();
  static const int env_rows = [] { const char* e = getenv("QUAN_POOL_ROWS"); return e ? atoi(e) : 1; }();
  if constexpr (sizeof(T) == 2) {
    if (env_packed && env_rows && g.Wo * g.inner_vecs >= 128 && g.outer * g.Ho >= 2048 && g.outer * g.Ho < (1ll << 31)) {
      // rows per block: as many as keep >= 8 blocks per SM in the grid (small maps: the flat kernel fills the GPU better —
      // 18.6 vs 23.0 us on the 16 x 32 ch x 32^2 QSPPF pool)
      PoolGeom gr = g;
      gr.rb = (int)(g.outer * g.Ho / (8 * QUAN_NUM_SMS));
      gr.rb = gr.rb < 1 ? 1 : gr.rb > 4 ? 4 : gr.rb;
      const PoolGeom& g = gr;
      const unsigned rows = (unsigned)((g.outer * g.Ho + g.rb - 1) / g.rb);
      if (V == 8) {
        if (idx != nullptr) QUAN_LAUNCH((qmaxpool_fwd_bf16_rows_kernel<8, true>), rows, 256, 0, st, xp, yp, idx, g);
        else QUAN_LAUNCH((qmaxpool_fwd_bf16_rows_kernel<8, false>), rows, 256, 0, st, xp, yp, idx, g);
      } else {
        if (idx != nullptr) QUAN_LAUNCH((qmaxpool_fwd_bf16_rows_kernel<4, true>), rows, 256, 0, st, xp, yp, idx, g);
        else QUAN_LAUNCH((qmaxpool_fwd_bf16_rows_kernel<4, false>), rows, 256, 0, st, xp, yp, idx, g);
      }
      QUAN_CHECK_LAUNCH("qmaxpool_fwd");
      return QUAN_OK;
    }
    if (env_packed) {
      if (V == 8) QUAN_POOL_FWD_BF16(8); else QUAN_POOL_FWD_BF16(4);
      QUAN_CHECK_LAUNCH("qmaxpool_fwd");
      return QUAN_OK;
    }
  }
  if (V == 8) { if constexpr (sizeof(T) == 2) QUAN_POOL_FWD(8); }
  else QUAN_POOL_FWD(4);
#undef QUAN_POOL_FWD
#undef QUAN_POOL_FWD_BF16
  QUAN_CHECK_LAUNCH("qmaxpool_fwd");
  return QUAN_OK;
}

template <typename T>
static int launch_pool_bwd(const void* dy, const uint8_t* idx, void* dx, const PoolGeom& g, int V, cudaStream_t st) {
  const T* gp = reinterpret_cast<const T*>(dy);
  T* dp = reinterpret_cast<T*>(dx);
  const int grid = grid_for(g.outer * g.H * g.W * g.inner_vecs, 256, 8);
  QUAN_TIMED(st);
  const bool small = g.outer * g.H * g.W * g.inner_vecs < (1ll << 31) - (1 << 20);   // i + grid stride stays below 2^31
  static const int env_rows = [] { const char* e = getenv("QUAN_POOL_ROWS"); return e ? atoi(e) : 1; }();
  static const int env_tile = [] { const char* e = getenv("QUAN_POOL_TILE"); return e ? atoi(e) : 1; }();
  {
    PoolTile t;
    size_t smem = 0;
    if (env_tile && plan_pool_tile(g, V, sizeof(T), t, smem)) {
      const unsigned blocks = (unsigned)(g.outer * t.nth * t.ntw * t.nchunks);
      if (V == 8) { if constexpr (sizeof(T) == 2) QUAN_LAUNCH((qmaxpool_bwd_tile_kernel<T, 8>), blocks, 256, smem, st, gp, idx, dp, g, t); }
      else if (V == 4) QUAN_LAUNCH((qmaxpool_bwd_tile_kernel<T, 4>), blocks, 256, smem, st, gp, idx, dp, g, t);
      if (V == 8 || V == 4) {
        QUAN_CHECK_LAUNCH("qmaxpool_bwd");
        return QUAN_OK;
      }
    }
  }
  if (env_rows && g.W * g.inner_vecs >= 128 && g.outer * g.H >= 2048 && g.outer * g.H < (1ll << 31)) {
    // measured on the Q-ResNet stem pool: 8 rows per block 348 us, flat kernel 550 us, 1 row per block 1348 us; small maps
    // (512 rows: 64 blocks of 8 rows, 132 vs 37 us) stay on the flat kernel
    PoolGeom gr = g;
    gr.rb = (int)(g.outer * g.H / (8 * QUAN_NUM_SMS));
    gr.rb = gr.rb < 1 ? 1 : gr.rb > POOL_RB ? POOL_RB : gr.rb;
    const PoolGeom& g = gr;
    const unsigned rows = (unsigned)((g.outer * g.H + g.rb - 1) / g.rb);
    if (V == 8) { if constexpr (sizeof(T) == 2) QUAN_LAUNCH((qmaxpool_bwd_rows_kernel<T, 8>), rows, 256, 0, st, gp, idx, dp, g); }
    else QUAN_LAUNCH((qmaxpool_bwd_rows_kernel<T, 4>), rows, 256, 0, st, gp, idx, dp, g);
    QUAN_CHECK_LAUNCH("qmaxpool_bwd");
    return QUAN_OK;
  }
  if (V == 8) {
    if constexpr (sizeof(T) == 2) {
      if (small) QUAN_LAUNCH((qmaxpool_bwd_kernel<T, 8, int>), grid, 256, 0, st, gp, idx, dp, g);
      else QUAN_LAUNCH((qmaxpool_bwd_kernel<T, 8, int64_t>), grid, 256, 0, st, gp, idx, dp, g);
    }
  } else {
    if (small) QUAN_LAUNCH((qmaxpool_bwd_kernel<T, 4, int>), grid, 256, 0, st, gp, idx, dp, g);
    else QUAN_LAUNCH((qmaxpool_bwd_kernel<T, 4, int64_t>), grid, 256, 0, st, gp, idx, dp, g);
  }
  QUAN_CHECK_LAUNCH("qmaxpool_bwd");
  return QUAN_OK;
}

}  // namespace quan

using namespace quan;

extern "C" {

int quan_qmaxpool_fwd(const void* x, void* y, uint8_t* idx, int32_t B, int32_t C, int32_t H, int32_t W, int32_t kH, int32_t kW,
                      int32_t sH, int32_t sW, int32_t pH, int32_t pW, int dtype, int layout, void* stream) {
  QUAN_REQUIRE(x != nullptr && y != nullptr, QUAN_E_ARG, "qmaxpool_fwd: null pointer");
  PoolGeom g;
  int V = 0;
  int rc = pool_geom("qmaxpool_fwd", B, C, H, W, kH, kW, sH, sW, pH, pW, dtype, layout, g, V);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == QUAN_F32) return launch_pool_fwd<float>(x, y, idx, g, V, st);
  return launch_pool_fwd<__nv_bfloat16>(x, y, idx, g, V, st);
}

int quan_qmaxpool_bwd(const void* dy, const uint8_t* idx, void* dx, int32_t B, int32_t C, int32_t H, int32_t W, int32_t kH,
                      int32_t kW, int32_t sH, int32_t sW, int32_t pH, int32_t pW, int dtype, int layout, void* stream) {
  QUAN_REQUIRE(dy != nullptr && idx != nullptr && dx != nullptr, QUAN_E_ARG, "qmaxpool_bwd: null pointer");
  PoolGeom g;
  int V = 0;
  int rc = pool_geom("qmaxpool_bwd", B, C, H, W, kH, kW, sH, sW, pH, pW, dtype, layout, g, V);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == QUAN_F32) return launch_pool_bwd<float>(dy, idx, dx, g, V, st);
  return launch_pool_bwd<__nv_bfloat16>(dy, idx, dx, g, V, st);
}

}  // extern "C"
