// qer.cu — QER, the quaternion -> real extraction of the detection heads, for sm_100a.
//
// Reference semantics (not code): ultralytics/nn/modules/head.py:26-47 — `x.permute(0, 1, 4, 2, 3).contiguous().view(B, 4C, H, W)`
// followed by a real 1x1 `nn.Conv2d(4C, N)` whose input channel index is c*4 + q.  In the tensor-core layout (BHWQC) a pixel already
// IS a row of K = 4C contiguous values (index q*C + c), so the op is a skinny GEMM over pixel rows:
//     out[pix][n] = bias[n] + sum_k x[pix][k] * W[n][k'],   k = q*C + c  <->  k' = c*4 + q
// with K <= 256 and N <= 64 (the heads: K in {16, 64, 128}, N in {1, 15, 64}): 2N flops per input byte at most, HBM-bound.  Each
// kernel stages a 128-pixel tile and the (re-ordered, bf16-rounded) weight in shared memory and contracts them with warp-level
// mma.sync m16n8k16 — the contraction is small enough that legacy MMA keeps the kernels on the memory roof; tcgen05 / TMEM would
// add a TMEM round trip per 32 KB tile for nothing.  fp32 tensors take exact-fp32 FMA kernels of the same interface (tests, fp32 runs).
//
//   forward : out rows may be `out_ld` elements apart (so the box / class extractions can write straight into the concatenated
//             head tensor of head.py:143) and need no alignment.
//   backward: dy rows `dy_ld` apart, unaligned (the gradient arrives as a channel slice of the concatenated tensor's gradient);
//             dx = dy W (dense BHWQC), dW / db through per-CTA partials and a fold kernel (deterministic).
#include "common.cuh"

namespace quan {

constexpr int QER_TM = 128;        // pixels per tile
constexpr int QER_MAX_K = 256;
constexpr int QER_MAX_N = 64;
constexpr int QER_WG_CTAS = 2 * QUAN_NUM_SMS;   // wgrad: persistent CTAs = partial slots

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x2(uint32_t addr, uint32_t (&r)[2]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x2_t(uint32_t addr, uint32_t (&r)[2]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// reference weight index (c*4 + q) of the layout's contraction index k = q*C + c
__device__ __forceinline__ int ref_k(int k, int C) { return (k % C) * 4 + k / C; }

// ---------------------------------------------------------------------------------------------------------------
// bf16: OUT[128 x N'] = A[128 x K'] * B[N' x K']^T on one CTA (4 warps x 32 rows); A / B in shared memory, k contiguous, pitches
// lda / ldb (elements, multiple of 8, odd multiple of 16 bytes: conflict-free ldmatrix).  N' is walked in chunks of 8 n-tiles.
// `emit(row, col, v0, v1)` receives two adjacent columns of one row.
// ---------------------------------------------------------------------------------------------------------------
template <typename Emit>
__device__ __forceinline__ void cta_gemm_rows(const __nv_bfloat16* As, int lda, const __nv_bfloat16* Bs, int ldb, int kdim, int ndim,
                                              Emit emit) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row0 = warp * 32;
  for (int n0 = 0; n0 < ndim; n0 += 64) {
    const int ntiles = min(8, (ndim - n0) >> 3);
    float acc[2][8][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[mt][nt][e] = 0.f;
    for (int k0 = 0; k0 < kdim; k0 += 16) {
      uint32_t a[2][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
        ldsm_x4(smem_addr(As + (size_t)(row0 + mt * 16 + (lane & 15)) * lda + k0 + (lane >> 4) * 8), a[mt]);
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        if (nt < ntiles) {
          uint32_t b[2];
          ldsm_x2(smem_addr(Bs + (size_t)(n0 + nt * 8 + (lane & 7)) * ldb + k0 + ((lane >> 3) & 1) * 8), b);
          mma_bf16(acc[0][nt], a[0], b);
          mma_bf16(acc[1][nt], a[1], b);
        }
      }
    }
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
        if (nt < ntiles) {
          const int r = row0 + mt * 16 + (lane >> 2), c = n0 + nt * 8 + (lane & 3) * 2;
          emit(r, c, acc[mt][nt][0], acc[mt][nt][1]);
          emit(r + 8, c, acc[mt][nt][2], acc[mt][nt][3]);
        }
  }
}

// e -> (e / d, e % d) without a hardware divide when d is a power of two (the head's widths all are)
__device__ __forceinline__ void divmod(int e, int d, int sh, int& qt, int& rm) {
  if (sh >= 0) { qt = e >> sh; rm = e & (d - 1); }
  else { qt = e / d; rm = e - qt * d; }
}
__device__ __forceinline__ int pow2_shift(int d) { return (d & (d - 1)) == 0 ? 31 - __clz(d) : -1; }

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool live) {   // live = false: 16 bytes of zeros
  const int sz = live ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_addr(smem)), "l"(gmem), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Stage `rows` x `cols` elements of a pixel-row tile (row pitch ld in global) into shared memory [QER_TM][lds], zero padded to
// cols_pad.  The loads of a whole tile are in flight at once — the kernels are latency-bound otherwise (a first version that
// loaded and stored one vector at a time ran at 0.18 of the HBM rate):
//   * rows that start on 16-byte boundaries: cp.async 16-byte copies (no registers; the caller overlaps them with the previous
//     tile's math).  `readable` >= cols columns of every row may be read (row padding), so a ragged width (15 of 16) still takes
//     whole vectors; the kernels clear the extra columns once the tile has landed (zero_cols).
//   * any other pitch / alignment: a warp per row, a lane per element, eight rows of loads in flight per warp.
// Returns true when the copy is asynchronous (the caller must cp_async_wait before reading).
__device__ __forceinline__ bool stage_rows(const __nv_bfloat16* __restrict__ g, int64_t ld, int rows, int cols, int cols_pad, int readable,
                                           __nv_bfloat16* s, int lds) {
  const int cols8 = (cols + 7) & ~7;
  const bool vec = (cols8 <= readable) && (ld % 8 == 0) && ((reinterpret_cast<uintptr_t>(g) & 15) == 0);
  if (vec) {
    const int vpr = cols8 / 8, sh = pow2_shift(vpr);
    for (int e = threadIdx.x; e < QER_TM * vpr; e += blockDim.x) {
      int r, v;
      divmod(e, vpr, sh, r, v);
      const bool live = r < rows;
      cp_async16(s + (size_t)r * lds + v * 8, g + (live ? (int64_t)r * ld : 0) + v * 8, live);
    }
    if (cols_pad > cols8) {
      const int np = cols_pad - cols8, shp = pow2_shift(np);
      for (int e = threadIdx.x; e < QER_TM * np; e += blockDim.x) {
        int r, c;
        divmod(e, np, shp, r, c);
        s[(size_t)r * lds + cols8 + c] = __float2bfloat16_rn(0.f);
      }
    }
    return true;
  }
  const unsigned short* gu = reinterpret_cast<const unsigned short*>(g);
  unsigned short* su = reinterpret_cast<unsigned short*>(s);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int r0 = warp * 8; r0 < QER_TM; r0 += nwarps * 8) {       // cols_pad <= 64: at most two elements per lane and row
    unsigned short v[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = r0 + i;
      const unsigned short* gr = gu + (int64_t)r * ld;
      v[i][0] = (r < rows && lane < cols) ? __ldg(gr + lane) : (unsigned short)0;
      v[i][1] = (r < rows && lane + 32 < cols) ? __ldg(gr + lane + 32) : (unsigned short)0;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      unsigned short* sr = su + (size_t)(r0 + i) * lds;
      if (lane < cols_pad) sr[lane] = v[i][0];
      if (lane + 32 < cols_pad) sr[lane + 32] = v[i][1];
    }
  }
  return false;
}

// columns [c0, c1) of a staged tile := 0 (the over-read padding of a ragged width, after it has landed)
__device__ __forceinline__ void zero_cols(__nv_bfloat16* s, int lds, int c0, int c1) {
  const int n = c1 - c0;
  for (int e = threadIdx.x; e < QER_TM * n; e += blockDim.x) s[(size_t)(e / n) * lds + c0 + e % n] = __float2bfloat16_rn(0.f);
}

// The weight staged for the tensor-core contraction.  Read as float4 = the four components of one (n, c) in the reference's order
// (index c*4 + q), eight vectors per thread in flight (a first version loaded one value at a time: 32 dependent L2 / DRAM round trips
// per thread, ~12 us of every launch), scattered to k = q*C + c.  KN = false: Ws[n][k] (B operand of the forward GEMM, rows n >= N
// zero); KN = true: Wt[k][n] (B operand of the dgrad GEMM, columns n >= N zero).
template <bool KN>
__device__ __forceinline__ void stage_weight(const float* __restrict__ w, int N, int Npad, int K, int C, __nv_bfloat16* s, int lds) {
  const int total = Npad * C;                                  // float4 vectors, (n, c) with c fastest
  const int shc = pow2_shift(C);
  for (int base = threadIdx.x; base < total; base += 8 * blockDim.x) {
    float4 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int e = base + u * blockDim.x;
      int n, c;
      divmod(e, C, shc, n, c);
      v[u] = (e < total && n < N) ? __ldg(reinterpret_cast<const float4*>(w + (size_t)n * K) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int e = base + u * blockDim.x;
      if (e < total) {
        int n, c;
        divmod(e, C, shc, n, c);
        const float q4[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (KN) s[(size_t)(q * C + c) * lds + n] = __float2bfloat16_rn(q4[q]);
          else s[(size_t)n * lds + q * C + c] = __float2bfloat16_rn(q4[q]);
        }
      }
    }
  }
}

// write a [rows][cols] tile from shared memory (pitch lds) to global rows `ld` apart; `writable` >= cols columns of every row may
// be written (row padding: a ragged width then still goes out as whole 16-byte vectors — the extra columns of the tile are zeros)
__device__ __forceinline__ void unstage_rows(const __nv_bfloat16* s, int lds, __nv_bfloat16* __restrict__ g, int64_t ld, int rows, int cols,
                                             int writable) {
  const int cols8 = (cols + 7) & ~7;
  const bool vec = (cols8 <= writable) && (ld % 8 == 0) && ((reinterpret_cast<uintptr_t>(g) & 15) == 0);
  if (vec) {
    const int vpr = cols8 / 8, sh = pow2_shift(vpr);
    for (int e = threadIdx.x; e < rows * vpr; e += blockDim.x) {
      int r, v;
      divmod(e, vpr, sh, r, v);
      *(reinterpret_cast<uint4*>(g + (int64_t)r * ld) + v) = *reinterpret_cast<const uint4*>(s + (size_t)r * lds + v * 8);
    }
  } else {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    for (int r = warp; r < rows; r += nwarps)
      for (int c = lane; c < cols; c += 32) g[(int64_t)r * ld + c] = s[(size_t)r * lds + c];
  }
}

// forward (bf16).  Persistent CTAs: weight and bias are staged once, then 128-pixel tiles are walked with a grid stride, the next
// tile's activation rows in flight (cp.async, two buffers) while the current tile is contracted and written out.
// smem: Xs[2] [128][K+8] | Os [128][Npad+8] | Ws [Npad][K+8] | bias [Npad] floats.
__global__ void __launch_bounds__(128) qer_fwd_bf16_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ w,
                                                            const float* __restrict__ bias, __nv_bfloat16* __restrict__ out,
                                                            int64_t npix, int C, int N, int64_t out_ld, int out_writable) {
  pdl_prologue();
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int K = 4 * C, Npad = (N + 7) & ~7, lda = K + 8, ldo = Npad + 8;
  __nv_bfloat16* Xs = reinterpret_cast<__nv_bfloat16*>(smem_raw);
  __nv_bfloat16* Os = Xs + (size_t)2 * QER_TM * lda;
  __nv_bfloat16* Ws = Os + (size_t)QER_TM * ldo;
  float* bs = reinterpret_cast<float*>(Ws + (size_t)Npad * lda);
  const int64_t tiles = (npix + QER_TM - 1) / QER_TM;
  int64_t t = blockIdx.x;
  if (t < tiles) stage_rows(x + t * QER_TM * K, K, (int)min((int64_t)QER_TM, npix - t * QER_TM), K, K, K, Xs, lda);
  cp_async_commit();
  stage_weight<false>(w, N, Npad, K, C, Ws, lda);
  for (int n = threadIdx.x; n < Npad; n += blockDim.x) bs[n] = (bias != nullptr && n < N) ? __ldg(bias + n) : 0.f;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row0 = warp * 32;
  const int ntiles = Npad >> 3;
  for (int it = 0; t < tiles; t += gridDim.x, ++it) {
    const int64_t pix0 = t * QER_TM, tn = t + gridDim.x;
    const int rows = (int)min((int64_t)QER_TM, npix - pix0);
    const __nv_bfloat16* Xc = Xs + (size_t)(it & 1) * QER_TM * lda;
    if (tn < tiles)                                    // the buffer it is written to was read two tiles ago (barriers below)
      stage_rows(x + tn * QER_TM * K, K, (int)min((int64_t)QER_TM, npix - tn * QER_TM), K, K, K, Xs + (size_t)((it + 1) & 1) * QER_TM * lda, lda);
    cp_async_commit();
    cp_async_wait<1>();                                // this tile has landed (the next may still be in flight)
    __syncthreads();                                   // ... for every thread; also: Os of the previous tile has been written out
    float acc[2][8][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[mt][nt][e] = 0.f;
    for (int k0 = 0; k0 < K; k0 += 16) {
      uint32_t a[2][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) ldsm_x4(smem_addr(Xc + (size_t)(row0 + mt * 16 + (lane & 15)) * lda + k0 + (lane >> 4) * 8), a[mt]);
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
        if (nt < ntiles) {
          uint32_t b[2];
          ldsm_x2(smem_addr(Ws + (size_t)(nt * 8 + (lane & 7)) * lda + k0 + ((lane >> 3) & 1) * 8), b);
          mma_bf16(acc[0][nt], a[0], b);
          mma_bf16(acc[1][nt], a[1], b);
        }
    }
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
        if (nt < ntiles) {
          const int r = row0 + mt * 16 + (lane >> 2), c = nt * 8 + (lane & 3) * 2;
          const float b0 = bs[c], b1 = bs[c + 1];
          *reinterpret_cast<__nv_bfloat162*>(Os + (size_t)r * ldo + c) = __floats2bfloat162_rn(acc[mt][nt][0] + b0, acc[mt][nt][1] + b1);
          *reinterpret_cast<__nv_bfloat162*>(Os + (size_t)(r + 8) * ldo + c) = __floats2bfloat162_rn(acc[mt][nt][2] + b0, acc[mt][nt][3] + b1);
        }
    __syncthreads();                                   // Os complete; every warp is done with Xc
    unstage_rows(Os, ldo, out + pix0 * out_ld, out_ld, rows, N, out_writable);
  }
  cp_async_wait<0>();
}

// dgrad (bf16): dx[pix][k] = sum_n dy[pix][n] W[n][ref_k(k)].  Persistent CTAs, W^T staged once, dy tiles double-buffered.
// smem: DYs[2] [128][Np16+8] | Wt [K][Np16+8] | Ds [128][K+8] (the dx tile, written out with coalesced rows)
__global__ void __launch_bounds__(128) qer_dgrad_bf16_kernel(const __nv_bfloat16* __restrict__ dy, int64_t dy_ld, int dy_readable,
                                                              const float* __restrict__ w, __nv_bfloat16* __restrict__ dx, int64_t npix, int C,
                                                              int N) {
  pdl_prologue();
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int K = 4 * C, Np = (N + 15) & ~15, ldn = Np + 8, ldk = K + 8;
  __nv_bfloat16* DYs = reinterpret_cast<__nv_bfloat16*>(smem_raw);
  __nv_bfloat16* Wt = DYs + (size_t)2 * QER_TM * ldn;
  __nv_bfloat16* Ds = Wt + (size_t)K * ldn;
  const int64_t tiles = (npix + QER_TM - 1) / QER_TM;
  int64_t t = blockIdx.x;
  bool ragged = false;
  if (t < tiles) ragged = stage_rows(dy + t * QER_TM * dy_ld, dy_ld, (int)min((int64_t)QER_TM, npix - t * QER_TM), N, Np, dy_readable, DYs, ldn) && (N & 7);
  cp_async_commit();
  stage_weight<true>(w, N, Np, K, C, Wt, ldn);
  for (int it = 0; t < tiles; t += gridDim.x, ++it) {
    const int64_t pix0 = t * QER_TM, tn = t + gridDim.x;
    const int rows = (int)min((int64_t)QER_TM, npix - pix0);
    const __nv_bfloat16* Dc = DYs + (size_t)(it & 1) * QER_TM * ldn;
    if (tn < tiles)
      stage_rows(dy + tn * QER_TM * dy_ld, dy_ld, (int)min((int64_t)QER_TM, npix - tn * QER_TM), N, Np, dy_readable,
                 DYs + (size_t)((it + 1) & 1) * QER_TM * ldn, ldn);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();                                   // tile landed; Ds of the previous tile has been written out
    if (ragged) {                                      // over-read padding columns -> 0 (uniform branch)
      zero_cols(DYs + (size_t)(it & 1) * QER_TM * ldn, ldn, N, (N + 7) & ~7);
      __syncthreads();
    }
    cta_gemm_rows(Dc, ldn, Wt, ldn, Np, K, [&](int r, int c, float v0, float v1) {
      *reinterpret_cast<__nv_bfloat162*>(Ds + (size_t)r * ldk + c) = __floats2bfloat162_rn(v0, v1);
    });
    __syncthreads();
    unstage_rows(Ds, ldk, dx + pix0 * K, K, rows, K, K);
  }
  cp_async_wait<0>();
}

// wgrad (bf16): persistent CTAs (8 warps) over pixel tiles, both operand tiles double-buffered; acc[n][k] += dy[pix][n] x[pix][k] with
// both operands read transposed (ldmatrix.trans) from their [pix][.] tiles; db[n] through one extra n-tile whose B fragment is the
// constant column of ones.  Warp w owns the k-tiles w, w+8 (KC <= 128 columns per CTA, blockIdx.y selects the chunk) for all (<= 4)
// 16-row n-tiles.  partial: [gridDim.x][Np16][K + 8] floats (column K holds the bias sums).
__global__ void __launch_bounds__(256) qer_wgrad_bf16_kernel(const __nv_bfloat16* __restrict__ dy, int64_t dy_ld, int dy_readable,
                                                              const __nv_bfloat16* __restrict__ x, float* __restrict__ partial, int64_t npix, int C,
                                                              int N, int KC) {
  pdl_prologue();
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int K = 4 * C, Np = (N + 15) & ~15, ldn = Np + 8, ldk = KC + 8;
  __nv_bfloat16* DYs = reinterpret_cast<__nv_bfloat16*>(smem_raw);
  __nv_bfloat16* Xs = DYs + (size_t)2 * QER_TM * ldn;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kc0 = blockIdx.y * KC;
  const int mtiles = Np >> 4, ktiles = KC >> 3;
  const bool bias_warp = (warp == 7) && (blockIdx.y == 0);
  float acc[4][2][4], accb[4][4];
#pragma unroll
  for (int mt = 0; mt < 4; ++mt) {
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[mt][j][e] = 0.f;
#pragma unroll
    for (int e = 0; e < 4; ++e) accb[mt][e] = 0.f;
  }
  const uint32_t ones = (lane >> 2) == 0 ? 0x3F803F80u : 0u;     // B fragment of a k16 x n8 tile whose column 0 is all ones
  const int64_t tiles = (npix + QER_TM - 1) / QER_TM;
  bool ragged = false;
  auto stage = [&](int64_t tt, int buf) {
    const int rows = (int)min((int64_t)QER_TM, npix - tt * QER_TM);
    ragged = stage_rows(dy + tt * QER_TM * dy_ld, dy_ld, rows, N, Np, dy_readable, DYs + (size_t)buf * QER_TM * ldn, ldn) && (N & 7);
    stage_rows(x + tt * QER_TM * K + kc0, K, rows, KC, KC, KC, Xs + (size_t)buf * QER_TM * ldk, ldk);
  };
  int64_t t = blockIdx.x;
  if (t < tiles) stage(t, 0);
  cp_async_commit();
  for (int it = 0; t < tiles; t += gridDim.x, ++it) {
    const int64_t tn = t + gridDim.x;
    __syncthreads();                      // the buffer about to be refilled was fully consumed (previous iteration's math)
    if (tn < tiles) stage(tn, (it + 1) & 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    if (ragged) {                         // over-read padding columns -> 0 (uniform: every tile of a launch takes the same path)
      zero_cols(DYs + (size_t)(it & 1) * QER_TM * ldn, ldn, N, (N + 7) & ~7);
      __syncthreads();
    }
    const __nv_bfloat16* Dc = DYs + (size_t)(it & 1) * QER_TM * ldn;
    const __nv_bfloat16* Xc = Xs + (size_t)(it & 1) * QER_TM * ldk;
    for (int p0 = 0; p0 < QER_TM; p0 += 16) {
      uint32_t a[4][4];
#pragma unroll
      for (int mt = 0; mt < 4; ++mt)
        if (mt < mtiles)   // A[m = n][kk = pix] from DYs[pix][n]: matrices (kk lo, m lo), (kk lo, m hi), (kk hi, m lo), (kk hi, m hi)
          ldsm_x4_t(smem_addr(Dc + (size_t)(p0 + (lane & 7) + ((lane >> 4) & 1) * 8) * ldn + mt * 16 + ((lane >> 3) & 1) * 8), a[mt]);
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int kt = warp + j * 8;
        if (kt < ktiles) {
          uint32_t b[2];     // B[kk = pix][n' = k] from Xs[pix][k]
          ldsm_x2_t(smem_addr(Xc + (size_t)(p0 + (lane & 7) + ((lane >> 3) & 1) * 8) * ldk + kt * 8), b);
#pragma unroll
          for (int mt = 0; mt < 4; ++mt)
            if (mt < mtiles) mma_bf16(acc[mt][j], a[mt], b);
        }
      }
      if (bias_warp) {
        const uint32_t b[2] = {ones, ones};
#pragma unroll
        for (int mt = 0; mt < 4; ++mt)
          if (mt < mtiles) mma_bf16(accb[mt], a[mt], b);
      }
    }
  }
  cp_async_wait<0>();
  float* slot = partial + (size_t)blockIdx.x * Np * (K + 8);
#pragma unroll
  for (int mt = 0; mt < 4; ++mt)
    if (mt < mtiles) {
      const int r = mt * 16 + (lane >> 2);
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int kt = warp + j * 8;
        if (kt < ktiles) {
          const int c = kc0 + kt * 8 + (lane & 3) * 2;
          *reinterpret_cast<float2*>(slot + (size_t)r * (K + 8) + c) = make_float2(acc[mt][j][0], acc[mt][j][1]);
          *reinterpret_cast<float2*>(slot + (size_t)(r + 8) * (K + 8) + c) = make_float2(acc[mt][j][2], acc[mt][j][3]);
        }
      }
      if (bias_warp && (lane & 3) == 0) {
        slot[(size_t)r * (K + 8) + K] = accb[mt][0];
        slot[(size_t)(r + 8) * (K + 8) + K] = accb[mt][2];
      }
    }
}

// fold the wgrad partials: dW[n][ref_k(k)] = sum_slots partial[slot][n][k], db[n] = sum_slots partial[slot][n][K].
// block = 8 outputs x 32 slot lanes (a warp per output: shuffle fold).
__global__ void __launch_bounds__(256) qer_wgrad_fold_kernel(const float* __restrict__ partial, int nslots, int Np, int K, int C, int N,
                                                              float* __restrict__ dw, float* __restrict__ db) {
  pdl_prologue();
  const int lane = threadIdx.x & 31;
  const int idx = blockIdx.x * 8 + (threadIdx.x >> 5);          // over N x (K + 1), k fastest
  if (idx >= N * (K + 1)) return;
  const int n = idx / (K + 1), k = idx - n * (K + 1);
  const float* p = partial + (size_t)n * (K + 8) + k;
  const size_t stride = (size_t)Np * (K + 8);
  float s0 = 0.f, s1 = 0.f;
  int i = lane;
  for (; i + 32 < nslots; i += 64) { s0 += p[(size_t)i * stride]; s1 += p[(size_t)(i + 32) * stride]; }
  if (i < nslots) s0 += p[(size_t)i * stride];
  float s = s0 + s1;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) {
    if (k < K) dw[(size_t)n * K + ref_k(k, C)] = s;
    else if (db != nullptr) db[n] = s;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// fp32 tensors: exact-fp32 FMA kernels with the same interfaces (small workloads: tests, fp32 runs)
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) qer_fwd_f32_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                                                           float* __restrict__ out, int64_t npix, int C, int N, int64_t out_ld) {
  pdl_prologue();
  const int K = 4 * C;
  const int64_t total = npix * N;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t pix = i / N;
    const int n = (int)(i - pix * N);
    const float* xr = x + pix * K;
    const float* wr = w + (size_t)n * K;
    float s = bias != nullptr ? __ldg(bias + n) : 0.f;
    for (int q = 0; q < 4; ++q)
      for (int c = 0; c < C; ++c) s = fmaf(__ldg(xr + q * C + c), __ldg(wr + c * 4 + q), s);
    out[pix * out_ld + n] = s;
  }
}
__global__ void __launch_bounds__(256) qer_dgrad_f32_kernel(const float* __restrict__ dy, int64_t dy_ld, const float* __restrict__ w,
                                                             float* __restrict__ dx, int64_t npix, int C, int N) {
  pdl_prologue();
  const int K = 4 * C;
  const int64_t total = npix * K;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t pix = i / K;
    const int k = (int)(i - pix * K);
    const int kr = ref_k(k, C);
    float s = 0.f;
    for (int n = 0; n < N; ++n) s = fmaf(__ldg(dy + pix * dy_ld + n), __ldg(w + (size_t)n * K + kr), s);
    dx[i] = s;
  }
}
// one CTA per pixel chunk; thread (n, k) pairs strided over the block; same partial format as the bf16 kernel
__global__ void __launch_bounds__(256) qer_wgrad_f32_kernel(const float* __restrict__ dy, int64_t dy_ld, const float* __restrict__ x,
                                                             float* __restrict__ partial, int64_t npix, int C, int N, int Np) {
  pdl_prologue();
  const int K = 4 * C;
  const int64_t per = (npix + gridDim.x - 1) / gridDim.x;
  const int64_t p0 = (int64_t)blockIdx.x * per, p1 = min(npix, p0 + per);
  float* slot = partial + (size_t)blockIdx.x * Np * (K + 8);
  for (int e = threadIdx.x; e < N * (K + 1); e += blockDim.x) {
    const int n = e / (K + 1), k = e - n * (K + 1);
    float s = 0.f;
    if (k < K)
      for (int64_t p = p0; p < p1; ++p) s = fmaf(__ldg(dy + p * dy_ld + n), __ldg(x + p * K + k), s);
    else
      for (int64_t p = p0; p < p1; ++p) s += __ldg(dy + p * dy_ld + n);
    slot[(size_t)n * (K + 8) + k] = s;
  }
}

static int qer_check(int64_t npix, int C, int N, int dtype, const char* who) {
  QUAN_REQUIRE(dtype == QUAN_F32 || dtype == QUAN_BF16, QUAN_E_ARG, "%s: bad dtype %d", who, dtype);
  QUAN_REQUIRE(npix > 0 && C > 0 && N > 0, QUAN_E_ARG, "%s: non-positive sizes", who);
  QUAN_REQUIRE(4 * C <= QER_MAX_K && N <= QER_MAX_N, QUAN_E_UNSUPPORTED, "%s: serves 4C <= %d input and N <= %d output channels (got %d, %d)",
               who, QER_MAX_K, QER_MAX_N, 4 * C, N);
  QUAN_REQUIRE(dtype == QUAN_F32 || C % 4 == 0, QUAN_E_UNSUPPORTED, "%s: bf16 needs C %% 4 == 0 (got %d)", who, C);
  return QUAN_OK;
}

// persistent fwd / dgrad grid: as many 128-thread CTAs as fit an SM's shared memory (<= 8), every SM, never more than the tiles
static unsigned persistent_grid(int64_t npix, size_t smem) {
  const int64_t tiles = (npix + QER_TM - 1) / QER_TM;
  int per_sm = (int)((200 * 1024) / (smem + 1024));
  per_sm = per_sm < 1 ? 1 : per_sm > 8 ? 8 : per_sm;
  const int64_t cap = (int64_t)QUAN_NUM_SMS * per_sm;
  return (unsigned)(tiles < cap ? tiles : cap);
}

static int wgrad_slots(int64_t npix, int dtype) {
  const int64_t tiles = (npix + QER_TM - 1) / QER_TM;
  const int64_t cap = dtype == QUAN_BF16 ? QER_WG_CTAS : 4 * QUAN_NUM_SMS;
  return (int)(tiles < cap ? tiles : cap);
}

}  // namespace quan

using namespace quan;

extern "C" {

size_t quan_qer_workspace_bytes(int64_t npix, int32_t C, int32_t N, int dtype) {
  if (npix <= 0 || C <= 0 || N <= 0) return 0;
  const size_t Np = (size_t)((N + 15) & ~15);
  return (size_t)wgrad_slots(npix, dtype) * Np * (4 * (size_t)C + 8) * sizeof(float);
}

int quan_qer_fwd(const void* x, const float* weight, const float* bias, void* out, int64_t npix, int32_t C, int32_t N, int64_t out_ld,
                 int32_t out_writable, int dtype, void* stream) {
  int rc = qer_check(npix, C, N, dtype, "qer_fwd");
  if (rc) return rc;
  QUAN_REQUIRE(x != nullptr && weight != nullptr && out != nullptr && out_ld >= N, QUAN_E_ARG, "qer_fwd: null pointer or out_ld < N");
  if (out_writable < N) out_writable = N;
  QUAN_REQUIRE(out_writable <= out_ld, QUAN_E_ARG, "qer_fwd: out_writable %d exceeds the row pitch", out_writable);
  cudaStream_t st = (cudaStream_t)stream;
  const int K = 4 * C;
  timing_work("qer_fwd", "", (double)npix * (K + N) * (dtype == QUAN_BF16 ? 2.0 : 4.0), 2.0 * npix * K * N);
  if (dtype == QUAN_BF16) {
    const int Npad = (N + 7) & ~7;
    const size_t smem = ((size_t)2 * QER_TM * (K + 8) + (size_t)QER_TM * (Npad + 8) + (size_t)Npad * (K + 8)) * 2 + Npad * sizeof(float) + 16;
    static DeviceOnce once;
    if (once.first()) QUAN_CUDA(cudaFuncSetAttribute(qer_fwd_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    QUAN_TIMED(st);
    QUAN_LAUNCH(qer_fwd_bf16_kernel, persistent_grid(npix, smem), 128, smem, st, (const __nv_bfloat16*)x, weight, bias,
                (__nv_bfloat16*)out, npix, C, N, out_ld, (int)out_writable);
  } else {
    QUAN_TIMED(st);
    QUAN_LAUNCH(qer_fwd_f32_kernel, grid_for(npix * N, 256, 8), 256, 0, st, (const float*)x, weight, bias, (float*)out, npix, C, N, out_ld);
  }
  QUAN_CHECK_LAUNCH("qer_fwd");
  return QUAN_OK;
}

int quan_qer_bwd(const void* dy, int64_t dy_ld, int32_t dy_readable, const void* x, const float* weight, void* dx, float* dweight, float* dbias,
                 int64_t npix, int32_t C, int32_t N, int dtype, void* workspace, size_t ws_bytes, void* stream) {
  int rc = qer_check(npix, C, N, dtype, "qer_bwd");
  if (rc) return rc;
  QUAN_REQUIRE(dy != nullptr && weight != nullptr && dy_ld >= N, QUAN_E_ARG, "qer_bwd: null pointer or dy_ld < N");
  QUAN_REQUIRE(dweight == nullptr || x != nullptr, QUAN_E_ARG, "qer_bwd: the weight gradient needs x");
  if (dy_readable < N) dy_readable = N;
  QUAN_REQUIRE(dy_readable <= dy_ld, QUAN_E_ARG, "qer_bwd: dy_readable %d exceeds the row pitch", dy_readable);
  cudaStream_t st = (cudaStream_t)stream;
  const int K = 4 * C, Np = (N + 15) & ~15;
  const double esz = dtype == QUAN_BF16 ? 2.0 : 4.0;
  if (dx != nullptr) {
    timing_work("qer_dgrad", "", (double)npix * (K + N) * esz, 2.0 * npix * K * N);
    if (dtype == QUAN_BF16) {
      const size_t smem = ((size_t)(2 * QER_TM + K) * (Np + 8) + (size_t)QER_TM * (K + 8)) * 2 + 16;
      static DeviceOnce once;
      if (once.first()) QUAN_CUDA(cudaFuncSetAttribute(qer_dgrad_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      QUAN_TIMED(st);
      QUAN_LAUNCH(qer_dgrad_bf16_kernel, persistent_grid(npix, smem), 128, smem, st, (const __nv_bfloat16*)dy, dy_ld, (int)dy_readable,
                  weight, (__nv_bfloat16*)dx, npix, C, N);
    } else {
      QUAN_TIMED(st);
      QUAN_LAUNCH(qer_dgrad_f32_kernel, grid_for(npix * K, 256, 8), 256, 0, st, (const float*)dy, dy_ld, weight, (float*)dx, npix, C, N);
    }
    QUAN_CHECK_LAUNCH("qer_dgrad");
  }
  if (dweight != nullptr) {
    const size_t need = quan_qer_workspace_bytes(npix, C, N, dtype);
    QUAN_REQUIRE(workspace != nullptr && ws_bytes >= need, QUAN_E_WORKSPACE, "qer_bwd: workspace needs %zu bytes, got %zu", need, ws_bytes);
    const int slots = wgrad_slots(npix, dtype);
    timing_work("qer_wgrad", "qer_wgrad_fold", (double)npix * (K + N) * esz, 2.0 * npix * K * N);
    if (dtype == QUAN_BF16) {
      const int KC = K <= 128 ? K : 128;
      QUAN_REQUIRE(K % KC == 0, QUAN_E_UNSUPPORTED, "qer_bwd: 4C = %d above 128 must be a multiple of 128", K);
      const size_t smem = (size_t)2 * QER_TM * ((Np + 8) + (KC + 8)) * 2 + 16;
      static DeviceOnce once;
      if (once.first()) QUAN_CUDA(cudaFuncSetAttribute(qer_wgrad_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      QUAN_TIMED(st);
      QUAN_LAUNCH(qer_wgrad_bf16_kernel, dim3((unsigned)slots, (unsigned)(K / KC)), 256, smem, st, (const __nv_bfloat16*)dy, dy_ld,
                  (int)dy_readable, (const __nv_bfloat16*)x, (float*)workspace, npix, C, N, KC);
    } else {
      QUAN_TIMED(st);
      QUAN_LAUNCH(qer_wgrad_f32_kernel, (unsigned)slots, 256, 0, st, (const float*)dy, dy_ld, (const float*)x, (float*)workspace, npix, C, N, Np);
    }
    QUAN_CHECK_LAUNCH("qer_wgrad");
    QUAN_TIMED(st);
    QUAN_LAUNCH(qer_wgrad_fold_kernel, (unsigned)((N * (K + 1) + 7) / 8), 256, 0, st, (const float*)workspace, slots, Np, K, C, N, dweight, dbias);
    QUAN_CHECK_LAUNCH("qer_wgrad_fold");
  }
  return QUAN_OK;
}

}  // extern "C"
