// qconv_tc.cu — tcgen05 / TMEM / TMA implicit-GEMM engine for QConv2D on sm_100a (layout BHWQC, groups = 1).
//
// Math (reference semantics): the separable stage of QConv2D is four independent real convolutions
// S_q = conv2d(x_q, W_q) (ultralytics/nn/modules/conv.py:480-483) followed by the 4x4 mix y = M S (conv.py:485-499).
// Here each S_q is an implicit GEMM  [128 output pixels] x [BN out channels] x [K = taps * C_in]:
//   * A tile  = 128 pixels x BK input channels of component q at filter tap (kh,kw): ONE 5-D TMA box
//               {BK, 1(q), Wt*sW, Ht*sH, Bt} with element strides {1,1,sW,sH,1} and start coordinate shifted by the tap;
//               out-of-bounds coordinates are zero-filled by TMA = the convolution's zero padding.  Rows land
//               K-major (channels contiguous) with the hardware swizzle the UMMA descriptor names.
//   * B tile  = BN x BK slice of the pre-packed weights Wp[q][tap][n][k] (4-D TMA box).
//   * D       = four fp32 accumulators (one per component q) of 128 lanes x BN columns, all resident in TMEM
//               (4*BN <= 512 columns); tcgen05.mma is issued by one thread, operands straight from shared memory.
//   * epilogue: 4 warps tcgen05.ld the four accumulators, apply the mixing matrix (+ bias on S_r, conv.py:480) in
//               registers and store the mixed quaternion outputs — S never touches HBM.
// dgrad (stride 1) is the same kernel on G = M^T dY with tap-flipped, transposed weights and identity mix;
// wgrad is its own kernel below (MN-major operands, split-K over pixels).
#include "qconv_internal.cuh"
#include "tc_ptx.cuh"
#include <mutex>
#include <vector>
#include <stdlib.h>
#include <string.h>
#include <unordered_map>

namespace quan {
static bool depthwise_as_dense(const quan_conv_dims& d);
static size_t packed_weight_bytes_of(const quan_conv_dims& d, int dtype, int dense);

// ---------------------------------------------------------------------------------------------------------------
// driver entry point for TMA descriptor encoding (libcuda is not linked: fetched through the runtime)
// ---------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

constexpr uint32_t SWZ_128B_ATOM32 = 1128;

static int encode_map(CUtensorMap* map, int dtype, int rank, const void* base, const uint64_t* dims, const uint64_t* strides_b,
                      const uint32_t* box, const uint32_t* estr, uint32_t row_bytes) {
  EncodeTiledFn fn = get_encode_fn();
  QUAN_REQUIRE(fn != nullptr, QUAN_E_DRIVER, "cuTensorMapEncodeTiled is not available from this driver");
  // row_bytes 128/64/32 select the plain swizzles; SWZ_128B_ATOM32 selects "128B swizzle with 32B atoms", the only
  // layout the tensor core accepts for MN-major tf32 operands (UMMA layout type SWIZZLE_128B_BASE32B)
  CUtensorMapSwizzle sw = row_bytes == SWZ_128B_ATOM32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B
                          : row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                          : row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                            : CU_TENSOR_MAP_SWIZZLE_32B;
  cuuint64_t gd[5], gs[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = estr[i]; }
  for (int i = 0; i < rank - 1; ++i) gs[i] = strides_b[i];
  CUresult r = fn(map, dtype == QUAN_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank,
                  const_cast<void*>(base), gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  QUAN_REQUIRE(r == CUDA_SUCCESS, QUAN_E_DRIVER, "cuTensorMapEncodeTiled failed with CUresult %d (rank %d, box %u,%u,%u,%u,%u)",
               (int)r, rank, bx[0], bx[1], bx[2], rank > 3 ? bx[3] : 0, rank > 4 ? bx[4] : 0);
  return QUAN_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// weight packing: fp32 master W_q[co][ci][tap] -> T Wp[q][tap][n][k]
//   FWD  : n = co, k = ci, tap kept          DGRAD: n = ci, k = co, tap flipped (taps-1-tap)
// Both are transposes (tap moves from innermost to outermost, DGRAD also swaps co and ci), staged through shared memory
// so that global reads and writes are both contiguous (the first, one-thread-per-element version took 17 us for the
// 2.4 M weights of a C_q = 256 3x3 layer: 6% of the forward pass).
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float round_operand(float v, int esz) {
  if (esz == 4) {
    // the tensor core truncates fp32 operands to tf32; round the weights to nearest here so only the activation
    // operand carries truncation bias (measured: halves the systematic error of the tf32 path)
    uint32_t r32;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r32) : "f"(v));
    v = __uint_as_float(r32);
  }
  return v;
}
constexpr int PACK_SMEM_FLOATS = 8192;   // staging tile upper bound (32 KB); kernels get dynamic smem sized to their tile:
                                         // a 32 KB static array let only ~3 blocks share an SM under the default carveout

// FWD: block = (q, co, ci chunk): reads chunk*taps contiguous floats, writes one contiguous ci-row per tap
template <typename T>
__global__ void __launch_bounds__(256) pack_weights_fwd_kernel(const float* __restrict__ w0, const float* __restrict__ w1,
                                                               const float* __restrict__ w2, const float* __restrict__ w3,
                                                               T* __restrict__ out, int Co, int Ci, int taps, int cchunk) {
  pdl_prologue();
  extern __shared__ float tile[];
  const int q = blockIdx.y / Co, co = blockIdx.y % Co;
  const int ci0 = blockIdx.x * cchunk, nci = min(cchunk, Ci - ci0);
  const float* w = (q == 0 ? w0 : q == 1 ? w1 : q == 2 ? w2 : w3) + ((int64_t)co * Ci + ci0) * taps;
#pragma unroll 8
  for (int e = threadIdx.x; e < nci * taps; e += blockDim.x) tile[e] = __ldg(w + e);
  __syncthreads();
  for (int tap = 0; tap < taps; ++tap) {            // (tap, ci) nested: no integer division per element
    T* orow = out + (((int64_t)q * taps + tap) * Co + co) * Ci + ci0;
    for (int ci = threadIdx.x; ci < nci; ci += blockDim.x) orow[ci] = from_f32<T>(round_operand(tile[ci * taps + tap], sizeof(T)));
  }
}

// DGRAD: block = (q, 32-wide co tile, ci tile): tile[co][ci*taps] in shared memory, output rows contiguous in co
template <typename T>
__global__ void __launch_bounds__(256) pack_weights_dgrad_kernel(const float* __restrict__ w0, const float* __restrict__ w1,
                                                                 const float* __restrict__ w2, const float* __restrict__ w3,
                                                                 T* __restrict__ out, int Co, int Ci, int taps, int tci) {
  pdl_prologue();
  extern __shared__ float tile[];
  const int q = blockIdx.z;
  const int co0 = blockIdx.y * 32, nco = min(32, Co - co0);
  const int ci0 = blockIdx.x * tci, nci = min(tci, Ci - ci0);
  const float* w = q == 0 ? w0 : q == 1 ? w1 : q == 2 ? w2 : w3;
  const int row = nci * taps;                       // contiguous floats per co
  const int pitch = tci * taps + 1;                 // +1: the transposed reads below walk down a column
#pragma unroll 8
  for (int e = threadIdx.x; e < nco * row; e += blockDim.x) {
    const int co = e / row, r = e - co * row;
    tile[co * pitch + r] = __ldg(w + ((int64_t)(co0 + co) * Ci + ci0) * taps + r);
  }
  __syncthreads();
#pragma unroll 4
  for (int e = threadIdx.x; e < nco * row; e += blockDim.x) {
    const int co = e % nco;
    const int r = e / nco;
    const int ci = r % nci, t = r / nci;            // t: packed (flipped) tap index
    out[(((int64_t)q * taps + t) * Ci + ci0 + ci) * Co + co0 + co] =
        from_f32<T>(round_operand(tile[co * pitch + ci * taps + (taps - 1 - t)], sizeof(T)));
  }
}

// ---------------------------------------------------------------------------------------------------------------
// forward / dgrad implicit-GEMM kernel
// ---------------------------------------------------------------------------------------------------------------
// The filter as a table of (input offset, packed-weight index) entries, grouped in output "classes":
//   forward / stride-1 dgrad: one class, entry (kh,kw) -> offset (kh*dH - pH, kw*dW - pW);
//   stride-s dgrad: s*s classes, one per output parity (ph,pw) — dX[i*s+ph][j*s+pw] only receives the taps with
//   (ph + pH - kh*dH) % s == 0, each a plain (stride-1) shifted read of G at offset (ph + pH - kh*dH)/s — so a strided
//   dgrad is s*s small stride-1 implicit GEMMs whose tiles are scattered with stride s by the epilogue.
constexpr int TC_MAX_TAPS = 64;
struct TapTable {
  int ncls;
  int start[5];                        // entries of class c: [start[c], start[c+1])
  int8_t dw[TC_MAX_TAPS], dh[TC_MAX_TAPS];
  uint8_t tap[TC_MAX_TAPS];
  int8_t ow[4], oh[4];                 // output offset of the class
  int zero_fill;                       // host only: some output parity receives no tap (e.g. 1x1 stride 2: three of the four
                                       // classes) — those classes are left out and the output is cleared before the launch
};

// Division by a launch constant in ~4 instructions (Granlund-Montgomery round-up multiplier; exact for every 32-bit numerator).
// The persistent kernels decompose a work-unit index into (class, N tile, pixel tile -> w, h, b) once per unit in EVERY thread of
// EVERY role; with hardware-emulated 32-bit division that was ~175 of the ~230 instructions an epilogue warp spends on a unit of a
// narrow layer (ncu: the 8 -> 8 1x1 layer at 256^2 issues 19 M warp instructions for 134 MB of traffic and runs at 0.45 of the HBM rate).
struct FastDiv {
  uint32_t m, s1, s2, d;
  __device__ __forceinline__ uint32_t div(uint32_t n) const {
    const uint32_t t = __umulhi(m, n);
    return (t + ((n - t) >> s1)) >> s2;
  }
  __device__ __forceinline__ void divmod(uint32_t n, int& q, int& r) const {
    const uint32_t qq = div(n);
    q = (int)qq;
    r = (int)(n - qq * d);
  }
};
static inline FastDiv make_fastdiv(uint32_t d) {
  FastDiv f;
  uint32_t l = 0;
  while ((1ull << l) < d) ++l;
  f.m = (uint32_t)(((1ull << 32) * ((1ull << l) - d)) / d + 1);
  f.s1 = l < 1 ? l : 1;
  f.s2 = l > 1 ? l - 1 : 0;
  f.d = d;
  return f;
}

struct TcConvParams {
  int B, Ho, Wo, Cout;                 // output tensor [B][Ho][Wo][NQ][Cout]
  int Wt, Ht, Bt, tiles_w, tiles_h;    // 128-pixel tile = Bt x Ht x Wt (w fastest) of the class grid, tiles per image plane
  int ntiles_n, units, units_per_cls;  // N tiles; work units = (class, pixel tile [pair], N tile), N fastest
  FastDiv fd_upc, fd_ntn, fd_tw, fd_th;   // / units_per_cls, / ntiles_n, / tiles_w, / tiles_h
  int in_sW, in_sH;                    // input coordinate of class-grid pixel (i,j) = (i*in_sH + dh, j*in_sW + dw)
  int out_sW, out_sH;                  // output coordinate = (i*out_sH + oh, j*out_sW + ow)
  TapTable tt;
  int kblocks, bk_elems;               // k-blocks per tap, elements per k-block row
  int sub;                             // (tap, k-block) sub-steps bundled into one pipeline stage (1 or 2)
  int BN, stages;
  int dbuf;                            // NQ = 4 with 8*BN <= 512: two sets of four accumulators alternate between units
  // halo mode (stride-1, 128-byte rows, tile = 16 rows x 8 pixels of one image): ONE TMA box per (component, k-block)
  // brings the tile's whole receptive field (halo_h x 16 pixels) and every filter tap is a UMMA descriptor that starts
  // `arow` rows into it — the activation tile crosses L2 -> SM once instead of once per tap.
  int halo, halo_dw, halo_dh;          // box origin relative to the tile origin (most negative tap offset)
  int a_stages, tg;                    // A ring slots; taps per B ring slot (one filter row)
  int halo_bo;                         // bring-up switch: put the swizzle phase of the window start into the descriptor
  uint32_t a_halo_bytes;
  uint16_t arow[TC_MAX_TAPS];          // tap -> first row of its window inside the halo tile ((dh - halo_dh)*16 + dw - halo_dw)
  uint32_t a_sub_bytes, b_sub_bytes;   // one sub-step's A / B tile (per CTA)
  uint32_t sbo_bytes, layout_type, idesc, tmem_cols;
  const float* bias;                   // NQ = 4: [Cout], joins S_r before the mix; NQ = 1: [Cout] added to the output
  double* stat_part;                   // or NULL: per-CTA partial IQBN sums of the OUTPUT, [grid][2][C_q*4] (index c*4+q)
  double* stat_acc;                    // or NULL: instead of a slot, every CTA ADDS its sums to these [2][C_q*4] fp64 accumulators
  const float* post_scale;             // or NULL: eval-mode IQBN folded into the epilogue, y = act(y * scale + shift); tables
  const float* post_shift;             //   [4][C_o] in (component, channel) order (stats[12C..20C) of quan_iqbn_eval_stats)
  int post_act;
  int stat_cq;                         // quaternion output channels C_o (NQ = 4: Cout; NQ = 1: Cout / 4)
  // TMA-store epilogue (output stride 1): every epilogue warp stages the 16 columns x 32 pixel rows it holds in its own two
  // shared-memory slots (swizzled like the tensor map, conflict-free 16-byte writes) and one lane sends the slot with
  // cp.async.bulk.tensor; the box is the warp's 32 rows of the tile = ts_wb x ts_hb x ts_bb pixels, clipped at the tensor edge.
  int tstore;                          // 0: per-lane stores; else slots per warp (2 or 4)
  Mix16 mix;
};

// Fused IQBN statistics in the conv epilogue: a thread holds 16 output values of one pixel row; the warp needs the 16
// column sums (and sums of squares) over its 32 rows.  Halving butterfly: 32 values per lane -> 5 exchange steps in which
// every lane sends half of what it still holds (16+8+4+2+1 = 31 shuffles instead of 32 x 5), after which lane L holds the
// total of value L (L < 16: sum of column L, L >= 16: sum of squares of column L-16) and adds it to the CTA's shared
// accumulators sacc[which][col] (4 quarter-warps x 2 column halves share an address at most 4 ways).
__device__ __forceinline__ void stat_add(float* sacc, int cq, int col0, const float (&o)[16], int lane) {
  float v[32];
#pragma unroll
  for (int j = 0; j < 16; ++j) { v[j] = o[j]; v[16 + j] = o[j] * o[j]; }
#pragma unroll
  for (int n = 16, off = 16; n >= 1; n >>= 1, off >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < n; ++i) {
      const float send = up ? v[i] : v[i + n];
      const float keep = up ? v[i + n] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  atomicAdd(sacc + (lane >> 4) * 4 * cq + col0 + (lane & 15), v[0]);
}

// same butterfly on per-thread running sums (s: sums of 16 columns over the rows this thread has seen, q: sums of squares)
__device__ __forceinline__ void stat_add_sums(float* sacc, int cq, int col0, const float (&s)[16], const float (&q)[16], int lane) {
  float v[32];
#pragma unroll
  for (int j = 0; j < 16; ++j) { v[j] = s[j]; v[16 + j] = q[j]; }
#pragma unroll
  for (int n = 16, off = 16; n >= 1; n >>= 1, off >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < n; ++i) {
      const float send = up ? v[i] : v[i + n];
      const float keep = up ? v[i + n] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  atomicAdd(sacc + (lane >> 4) * 4 * cq + col0 + (lane & 15), v[0]);
}

constexpr int TC_THREADS = 192;    // wgrad kernel: warp 0 TMA producer, warp 1 TMEM alloc + MMA issuer, warps 2-5 epilogue
constexpr int EPI_WARPS = 8;       // igemm kernel: two epilogue warps per TMEM lane quarter (they split the columns)
// Epilogue groups: the kernel can run one group of EPI_WARPS epilogue warps per TMEM accumulator of the dense form (units alternating
// between them).  Measured on B200 with two groups: 8 -> 8 1x1 at 256^2 88 vs 86 us, 16 -> 16 3x3 at 128^2 68 vs 61 us, whole step
// 12.67 vs 12.35 ms — the epilogue's cost was its instruction count (the statistics butterfly), not its latency; one group it stays.
constexpr int epi_groups(int nq) { return nq == 1 ? 1 : 1; }   // measured: a second group (one per accumulator) gains nothing — see stat_add
constexpr int ig_threads(int nq) { return 64 + 32 * EPI_WARPS * epi_groups(nq); }

// Persistent implicit-GEMM kernel.  One CTA (CG = 1) or CTA pair (CG = 2, tcgen05.mma.cta_group::2, M = 256, the
// BN x BK weight tile split across the pair's shared memory) per SM loops over work units; the three roles run as
// independent pipelines that only meet at mbarriers, so TMA loads of unit u+1 and, as far as TMEM allows, its MMAs
// overlap the epilogue of unit u:
//   NQ = 4 (separable form): TMEM holds the four component accumulators S_q (4*BN <= 512 columns — no room for a second
//     set).  The epilogue first copies S_0 into registers and hands accumulator 0 back, so the next unit's q = 0
//     mainloop (a quarter of its MMAs) runs while S_1..S_3 are drained, mixed with the stashed S_0 and stored.
//   NQ = 1 (dense Hamilton form for narrow layers: channels = 4*C_q, mixing matrix folded into the packed weights):
//     two accumulators of BN <= 256 columns alternate between units.
// KSTEPS = UMMAs per sub-step (row bytes / 32), compile time so the issue sequence is branch-free.
//
// What bounded the non-persistent version (clock64 / debug-switch experiments, profiles/r01_conv_issue_experiments.md):
// not the loads but the MMA-issuing thread — a UTCHMMA M128 N128 K16 holds the issuing thread for about its 64-cycle
// execution time, so the tensor pipe is busy only while that thread issues.  Hence (1) role loops are warp-uniform with
// `elect.sync` around the single-thread instructions, (2) a pipeline stage bundles `sub` (tap, k-block) steps, and
// (3) this persistent form removes the per-CTA prologue and all but the S_0 copy of the epilogue from the MMA
// thread's critical path.
template <typename T, bool MIX, int CG, int KSTEPS, int NQ, bool RSTAT = false>
__global__ void __launch_bounds__(ig_threads(NQ), 1)
qconv_igemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                   const __grid_constant__ CUtensorMap map_y, T* __restrict__ y, const TcConvParams p) {
  pdl_prologue();
  constexpr int NACC = NQ == 4 ? 4 : 2;   // TMEM accumulators of BN columns
  constexpr int NTB = NACC / NQ;          // units that can be in flight in TMEM (tile_full barriers)
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-byte alignment is required by the 128B swizzle atoms
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t a_stage_bytes = p.a_sub_bytes * p.sub, b_stage_bytes = p.b_sub_bytes * p.sub;
  constexpr uint32_t TS_RB = 16 * sizeof(T);           // bytes of one staged row (16 columns)
  constexpr uint32_t TS_SLOT = 32 * TS_RB;             // one warp's box: 32 pixel rows
  uint8_t* ts_base = smem;                             // [EPI_WARPS][2] slots when p.tstore
  uint8_t* smem_a = smem + (size_t)EPI_WARPS * epi_groups(NQ) * p.tstore * TS_SLOT;
  uint8_t* smem_b = smem_a + (p.halo ? (size_t)p.a_stages * p.a_halo_bytes : (size_t)p.stages * a_stage_bytes);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_b + (size_t)p.stages * b_stage_bytes);
  uint64_t* empty_bar = full_bar + p.stages;
  uint64_t* a_full = empty_bar + p.stages;        // [4]  halo mode: the A ring has its own barriers (B ring uses full/empty_bar)
  uint64_t* a_empty = a_full + 4;                 // [4]
  uint64_t* tile_full = a_empty + 4;              // [2]  MMA -> epilogue: all accumulators of a unit are complete
  uint64_t* acc_empty = tile_full + 2;            // [8]  epilogue -> MMA: accumulator a has been read out
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_empty + 8);
  float* sacc = reinterpret_cast<float*>(tmem_ptr + 4);   // [2][4][C_o] sum / sum of squares of this CTA's outputs (fused IQBN stats)

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // provably warp-uniform
  const int lane = threadIdx.x & 31;
  constexpr int KIND = sizeof(T) == 2 ? 0 : 1;
  const uint32_t cta_rank = CG == 2 ? ptx::cluster_ctarank() : 0u;
  const int cluster = CG == 2 ? (int)ptx::cluster_id_x() : (int)blockIdx.x;
  const int nclusters = CG == 2 ? (int)ptx::cluster_nctaid_x() : (int)gridDim.x;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&map_a);
    ptx::prefetch_tensormap(&map_b);
    if (p.tstore) ptx::prefetch_tensormap(&map_y);
    for (int s = 0; s < p.stages; ++s) {
      ptx::mbar_init(full_bar + s, 1);
      ptx::mbar_init(empty_bar + s, 1);
    }
    for (int i = 0; i < 4; ++i) {
      ptx::mbar_init(a_full + i, 1);
      ptx::mbar_init(a_empty + i, 1);
    }
    for (int i = 0; i < 2; ++i) ptx::mbar_init(tile_full + i, 1);
    for (int i = 0; i < 8; ++i) ptx::mbar_init(acc_empty + i, EPI_WARPS * CG);
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    if constexpr (CG == 2) ptx::tmem_alloc_2cta(tmem_ptr, p.tmem_cols);
    else ptx::tmem_alloc(tmem_ptr, p.tmem_cols);
  }
  if (p.stat_part != nullptr && warp >= 2)
    for (int e = threadIdx.x - 64; e < 8 * p.stat_cq; e += 32 * EPI_WARPS * epi_groups(NQ)) sacc[e] = 0.f;
  ptx::tc_fence_before();
  if constexpr (CG == 2) ptx::cluster_sync_all();   // peer's barriers must be initialised before anything signals them
  else __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;


  if (warp == 0) {
    // ===== TMA producer: warp-uniform loop; one elected lane issues.  Counters instead of div/mod. =====
    const uint32_t sub_tx = (p.a_sub_bytes + p.b_sub_bytes) * CG;   // CG = 2: the leader's barrier counts both CTAs' bytes
    int s = 0;
    uint32_t phase = 0;
    if (p.halo) {
      int sa = 0;
      uint32_t pa = 0;
      const int ntaps = p.tt.start[1];
      for (int unit = cluster; unit < p.units; unit += nclusters) {
        int nt, tile, tw, th, tb;
        p.fd_ntn.divmod((uint32_t)unit, tile, nt);
        tile = tile * CG + (int)cta_rank;
        p.fd_tw.divmod((uint32_t)tile, th, tw);
        p.fd_th.divmod((uint32_t)th, tb, th);
        const int wc = tw * p.Wt + p.halo_dw, hc = th * p.Ht + p.halo_dh;
        const int n0 = nt * p.BN;
        const int nb0 = n0 + (int)cta_rank * (p.BN / CG);
        for (int q = 0; q < NQ; ++q) {
          for (int kb = 0; kb < p.kblocks; ++kb) {
            const int kc = kb * p.bk_elems;
            ptx::mbar_wait(a_empty + sa, pa ^ 1);
            if (ptx::elect_one()) {
              if (CG == 1 || cta_rank == 0) ptx::mbar_arrive_expect_tx(a_full + sa, p.a_halo_bytes * CG);
              uint8_t* a_dst = smem_a + (size_t)sa * p.a_halo_bytes;
              if constexpr (CG == 2) ptx::tma_load_5d_2cta(a_dst, &map_a, a_full + sa, kc, q, wc, hc, tb);
              else ptx::tma_load_5d(a_dst, &map_a, a_full + sa, kc, q, wc, hc, tb);
            }
            __syncwarp();
            if (++sa == p.a_stages) { sa = 0; pa ^= 1; }
            for (int t0 = 0; t0 < ntaps; t0 += p.tg) {
              const int nt_g = min(p.tg, ntaps - t0);
              ptx::mbar_wait(empty_bar + s, phase ^ 1);
              if (ptx::elect_one()) {
                if (CG == 1 || cta_rank == 0) ptx::mbar_arrive_expect_tx(full_bar + s, p.b_sub_bytes * CG * nt_g);
                for (int t = 0; t < nt_g; ++t) {
                  uint8_t* b_dst = smem_b + (size_t)s * b_stage_bytes + (size_t)t * p.b_sub_bytes;
                  const int tap = p.tt.tap[t0 + t];
                  if constexpr (CG == 2) ptx::tma_load_4d_2cta(b_dst, &map_b, full_bar + s, kc, nb0, tap, q);
                  else ptx::tma_load_4d(b_dst, &map_b, full_bar + s, kc, n0, tap, q);
                }
              }
              __syncwarp();
              if (++s == p.stages) { s = 0; phase ^= 1; }
            }
          }
        }
      }
    } else
    for (int unit = cluster; unit < p.units; unit += nclusters) {
      int cls, ucls, nt, tile, tw, th, tb;
      p.fd_upc.divmod((uint32_t)unit, cls, ucls);
      p.fd_ntn.divmod((uint32_t)ucls, tile, nt);
      tile = tile * CG + (int)cta_rank;
      p.fd_tw.divmod((uint32_t)tile, th, tw);
      p.fd_th.divmod((uint32_t)th, tb, th);
      const int b0 = tb * p.Bt;
      const int wbase = tw * p.Wt * p.in_sW, hbase = th * p.Ht * p.in_sH;
      const int n0 = nt * p.BN;
      const int nb0 = n0 + (int)cta_rank * (p.BN / CG);             // this CTA's slice of the weight tile
      const int e0 = p.tt.start[cls];
      const int iters_per_q = (p.tt.start[cls + 1] - e0) * p.kblocks;   // (tap, k-block) steps per component
      for (int q = 0; q < NQ; ++q) {
        int e = e0, kb = 0;
        for (int i = 0; i < iters_per_q; i += p.sub) {
          const int nsub = min(p.sub, iters_per_q - i);
          ptx::mbar_wait(empty_bar + s, phase ^ 1);
          const bool leader = ptx::elect_one();
          if (leader && (CG == 1 || cta_rank == 0)) ptx::mbar_arrive_expect_tx(full_bar + s, sub_tx * nsub);
          for (int u = 0; u < nsub; ++u) {
            if (leader) {
              uint8_t* a_dst = smem_a + (size_t)s * a_stage_bytes + (size_t)u * p.a_sub_bytes;
              uint8_t* b_dst = smem_b + (size_t)s * b_stage_bytes + (size_t)u * p.b_sub_bytes;
              const int wc = wbase + p.tt.dw[e], hc = hbase + p.tt.dh[e], kc = kb * p.bk_elems, tap = p.tt.tap[e];
              if constexpr (CG == 2) {
                ptx::tma_load_5d_2cta(a_dst, &map_a, full_bar + s, kc, q, wc, hc, b0);
                ptx::tma_load_4d_2cta(b_dst, &map_b, full_bar + s, kc, nb0, tap, q);
              } else {
                ptx::tma_load_5d(a_dst, &map_a, full_bar + s, kc, q, wc, hc, b0);
                ptx::tma_load_4d(b_dst, &map_b, full_bar + s, kc, n0, tap, q);
              }
            }
            if (++kb == p.kblocks) { kb = 0; ++e; }
          }
          __syncwarp();
          if (++s == p.stages) { s = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: warp-uniform loop; one elected lane issues (with CG = 2 only in the pair's leader CTA) =====
    if (cta_rank == 0) {
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint64_t da0 = ptx::make_smem_desc(ptx::smem_u32(smem_a), 16, p.sbo_bytes, p.layout_type);
      const uint64_t db0 = ptx::make_smem_desc(ptx::smem_u32(smem_b), 16, p.sbo_bytes, p.layout_type);
      const uint64_t a_step = (uint64_t)(a_stage_bytes >> 4), b_step = (uint64_t)(b_stage_bytes >> 4);
      const uint64_t a_sub = (uint64_t)(p.a_sub_bytes >> 4), b_sub = (uint64_t)(p.b_sub_bytes >> 4);
      int s = 0;
      uint32_t phase = 0;
      uint64_t da = da0, db = db0;
      uint32_t t_local = 0;
      if (p.halo) {
        // halo mode: A descriptors walk over ONE resident tile — tap (dh,dw) starts arow = dh'*16 + dw' rows into it; the
        // 8-row groups of the 16 image rows are 16 rows (2048 B) apart; the hardware swizzle works on absolute address bits,
        // so an unaligned window start needs nothing else (base-offset field stays 0)
        const uint64_t da_h = ptx::make_smem_desc(ptx::smem_u32(smem_a), 16, 16u * 128u, p.layout_type);
        const uint64_t a_slot = (uint64_t)(p.a_halo_bytes >> 4);
        const int ntaps = p.tt.start[1];
        int sa = 0;
        uint32_t pa = 0;
        for (int unit = cluster; unit < p.units; unit += nclusters, ++t_local) {
          for (int q = 0; q < NQ; ++q) {
            const uint32_t a = NQ == 4 ? (uint32_t)q + (p.dbuf ? 4u * (t_local & 1u) : 0u) : (t_local & 1u);
            const uint32_t use = (NQ == 4 && !p.dbuf) ? t_local : (t_local >> 1);
            ptx::mbar_wait(acc_empty + a, (use & 1u) ^ 1u);
            ptx::tc_fence_after();
            const uint32_t d_tmem = tmem_u + a * (uint32_t)p.BN;
            for (int kb = 0; kb < p.kblocks; ++kb) {
              ptx::mbar_wait(a_full + sa, pa);
              const uint64_t da_s = da_h + (uint64_t)sa * a_slot;
              for (int t0 = 0; t0 < ntaps; t0 += p.tg) {
                const int nt_g = min(p.tg, ntaps - t0);
                ptx::mbar_wait(full_bar + s, phase);
                ptx::tc_fence_after();
                if (ptx::elect_one()) {
                  for (int t = 0; t < nt_g; ++t) {
                    const uint32_t ar = p.arow[t0 + t];
                    const uint64_t dat = da_s + (uint64_t)(ar * 8u) + ((uint64_t)((ar & 7u) * (uint32_t)p.halo_bo) << 49);
                    const uint64_t dbt = db + (uint64_t)t * b_sub;
#pragma unroll
                    for (int k = 0; k < KSTEPS; ++k) {
                      if constexpr (CG == 2)
                        ptx::umma_2cta<KIND>(d_tmem, dat + (uint64_t)(2 * k), dbt + (uint64_t)(2 * k), p.idesc, (kb | t0 | t | k) ? 1u : 0u);
                      else
                        ptx::umma<KIND>(d_tmem, dat + (uint64_t)(2 * k), dbt + (uint64_t)(2 * k), p.idesc, (kb | t0 | t | k) ? 1u : 0u);
                    }
                  }
                  if constexpr (CG == 2) ptx::umma_commit_2cta(empty_bar + s, 3);
                  else ptx::umma_commit(empty_bar + s);
                }
                __syncwarp();
                db += b_step;
                if (++s == p.stages) { s = 0; phase ^= 1; db = db0; }
              }
              if (ptx::elect_one()) {           // every tap of this (q, k-block) has been issued: the A tile may be refilled
                if constexpr (CG == 2) ptx::umma_commit_2cta(a_empty + sa, 3);
                else ptx::umma_commit(a_empty + sa);
              }
              __syncwarp();
              if (++sa == p.a_stages) { sa = 0; pa ^= 1; }
            }
          }
          if (ptx::elect_one()) {
            uint64_t* tf = tile_full + ((NTB == 2 || p.dbuf) ? (t_local & 1u) : 0u);
            if constexpr (CG == 2) ptx::umma_commit_2cta(tf, 3);
            else ptx::umma_commit(tf);
          }
          __syncwarp();
        }
      } else
      for (int unit = cluster; unit < p.units; unit += nclusters, ++t_local) {
        const int cls = (int)p.fd_upc.div((uint32_t)unit);
        const int iters_per_q = (p.tt.start[cls + 1] - p.tt.start[cls]) * p.kblocks;
        for (int q = 0; q < NQ; ++q) {
          // accumulator a is reused every NACC/NQ units: wait until the epilogue warps (of both CTAs) have read it out
          const uint32_t a = NQ == 4 ? (uint32_t)q + (p.dbuf ? 4u * (t_local & 1u) : 0u) : (t_local & 1u);
          const uint32_t use = (NQ == 4 && !p.dbuf) ? t_local : (t_local >> 1);
          ptx::mbar_wait(acc_empty + a, (use & 1u) ^ 1u);
          ptx::tc_fence_after();
          const uint32_t d_tmem = tmem_u + a * (uint32_t)p.BN;
          for (int i = 0; i < iters_per_q; i += p.sub) {
            const int nsub = min(p.sub, iters_per_q - i);
            ptx::mbar_wait(full_bar + s, phase);
            ptx::tc_fence_after();
            if (ptx::elect_one()) {
              for (int u = 0; u < nsub; ++u) {
                const uint64_t dau = da + (uint64_t)u * a_sub, dbu = db + (uint64_t)u * b_sub;
#pragma unroll
                for (int k = 0; k < KSTEPS; ++k) {
                  // advance 32 bytes (one UMMA_K slice) inside the swizzle row: +2 in the 16-byte-unit start-address field
                  if constexpr (CG == 2)
                    ptx::umma_2cta<KIND>(d_tmem, dau + (uint64_t)(2 * k), dbu + (uint64_t)(2 * k), p.idesc, (i | u | k) ? 1u : 0u);
                  else
                    ptx::umma<KIND>(d_tmem, dau + (uint64_t)(2 * k), dbu + (uint64_t)(2 * k), p.idesc, (i | u | k) ? 1u : 0u);
                }
              }
              // frees the smem slot (in both CTAs when CG = 2) once these MMAs have read it
              if constexpr (CG == 2) ptx::umma_commit_2cta(empty_bar + s, 3);
              else ptx::umma_commit(empty_bar + s);
            }
            __syncwarp();
            da += a_step;
            db += b_step;
            if (++s == p.stages) { s = 0; phase ^= 1; da = da0; db = db0; }
          }
        }
        // every accumulator of this unit is complete once the MMAs issued so far have retired
        if (ptx::elect_one()) {
          uint64_t* tf = tile_full + ((NTB == 2 || p.dbuf) ? (t_local & 1u) : 0u);
          if constexpr (CG == 2) ptx::umma_commit_2cta(tf, 3);
          else ptx::umma_commit(tf);
        }
        __syncwarp();
      }
    }
  } else {
    // ===== epilogue: warps 2..9; TMEM lane quarter = warp % 4, the two warps of a quarter take alternate 16-column chunks =====
    const int quarter = warp & 3;
    const int half = ((warp - 2) >> 2) & 1;
    const uint32_t grp = (uint32_t)(warp - 2) >> 3;    // with two epilogue groups: the TMEM accumulator this one drains
    (void)grp;
    const int m = quarter * 32 + lane;                 // accumulator row = pixel within the tile
    const int wt = m % p.Wt, ht = (m / p.Wt) % p.Ht, bt = m / (p.Wt * p.Ht);
    const uint32_t lane_base = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const int nchunks = p.BN >> 4;
    constexpr int VW = 16 / sizeof(T);                 // elements per 16-byte store
    uint32_t t_local = 0;
    // Fused IQBN statistics of the dense form.  A thread always holds the same 16 (or 2 x 16) columns of ITS pixel row when the
    // layer is one N tile wide, so the column sums can run in registers across every unit of the CTA and meet the other rows once at
    // the end.  The per-unit butterfly (31 shuffles + a shared-memory atomic per 16 outputs) was 45 % of the forward kernel's stall
    // samples on the narrow layers (ncu source page, 16 -> 16 3x3 at 128^2: profiles/r02_ncu_igemm_narrow_source.txt).
    // RSTAT is its own instantiation: the 64 accumulator registers (155 per thread instead of 86) keep the plain kernels from sharing an
    // SM with the concurrent wgrad of the narrow layers' backward (measured: dgrad 2.06 -> 2.25 ms per step with one fat kernel).
    constexpr bool reg_stats = RSTAT;
    float rs0[RSTAT ? 16 : 1], rq0[RSTAT ? 16 : 1], rs1[RSTAT ? 16 : 1], rq1[RSTAT ? 16 : 1];
#pragma unroll
    for (int j = 0; j < (RSTAT ? 16 : 1); ++j) rs0[j] = rq0[j] = rs1[j] = rq1[j] = 0.f;
    // TMA-store epilogue (p.tstore): this warp's 32 accumulator rows are a box of the output tensor
    uint8_t* ts_slots = ts_base + (size_t)(warp - 2) * p.tstore * TS_SLOT;
    uint32_t ts_n = 0;                                 // stores issued by this warp; slot = ts_n % p.tstore
    constexpr uint32_t TS_NV = TS_RB / 16;             // 16-byte pieces of a staged row
    const uint32_t ts_row = (uint32_t)lane * TS_RB;
    const uint32_t ts_xor = (ts_row >> 7) & (TS_NV - 1);   // the tensor map's 32B / 64B swizzle: piece ^= address bits 7..
    const int ts_m0 = quarter * 32;
    const int ts_wt0 = ts_m0 % p.Wt, ts_ht0 = (ts_m0 / p.Wt) % p.Ht, ts_bt0 = ts_m0 / (p.Wt * p.Ht);
    auto tma_out = [&](const float (&o)[16], const int col, const int pc, const int w0, const int h0, const int b0) {
      if (w0 >= p.Wo || h0 >= p.Ho || b0 >= p.B) return;   // the whole box lies outside (warp-uniform)
      uint8_t* slot = ts_slots + (ts_n & (uint32_t)(p.tstore - 1)) * TS_SLOT;
      if (lane == 0) {                                 // the store that used this slot `tstore` stores ago has left shared memory
        if (p.tstore == 4) ptx::bulk_wait_read<3>();
        else ptx::bulk_wait_read<1>();
      }
      __syncwarp();
      const uint32_t row_addr = ptx::smem_u32(slot) + ts_row;
#pragma unroll
      for (int v = 0; v < (int)TS_NV; ++v) {
        Vec<T, VW> t;
#pragma unroll
        for (int j = 0; j < VW; ++j) t.v[j] = from_f32<T>(o[v * VW + j]);
        const uint4 u = *reinterpret_cast<const uint4*>(&t);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row_addr + (((uint32_t)v ^ ts_xor) << 4)), "r"(u.x), "r"(u.y),
                     "r"(u.z), "r"(u.w)
                     : "memory");
      }
      ptx::fence_proxy_async();                        // generic-proxy writes -> visible to the async proxy (TMA)
      __syncwarp();
      if (lane == 0) {
        ptx::tma_store_5d(&map_y, slot, col, pc, w0, h0, b0);
        ptx::bulk_commit();
      }
      ++ts_n;
    };
    auto release = [&](uint64_t* bar) {                // one arrival per epilogue warp (TMEM reads of this warp are done)
      ptx::tc_fence_before();
      __syncwarp();
      if (ptx::elect_one()) {
        if constexpr (CG == 2) ptx::mbar_arrive_cluster(bar, 0);
        else ptx::mbar_arrive(bar);
      }
      __syncwarp();
    };
    for (int unit = cluster; unit < p.units; unit += nclusters, ++t_local) {
      if constexpr (epi_groups(NQ) == 2) {
        if ((t_local & 1u) != grp) continue;           // the other group's unit (warp-uniform)
      }
      int cls, ucls, nt, tile, tw, th, tb;
      p.fd_upc.divmod((uint32_t)unit, cls, ucls);
      p.fd_ntn.divmod((uint32_t)ucls, tile, nt);
      tile = tile * CG + (int)cta_rank;
      p.fd_tw.divmod((uint32_t)tile, th, tw);
      p.fd_th.divmod((uint32_t)th, tb, th);
      const int wo = (tw * p.Wt + wt) * p.out_sW + p.tt.ow[cls], ho = (th * p.Ht + ht) * p.out_sH + p.tt.oh[cls];
      const int b = tb * p.Bt + bt;
      const int n0 = nt * p.BN;
      const bool valid = (wo < p.Wo) && (ho < p.Ho) && (b < p.B);
      const int bw0 = tw * p.Wt + ts_wt0, bh0 = th * p.Ht + ts_ht0, bb0 = tb * p.Bt + ts_bt0;   // origin of this warp's box (tstore)
      T* yrow = y + ((((int64_t)b * p.Ho + ho) * p.Wo + wo) * NQ) * p.Cout + n0;
      if constexpr (NQ == 4) {
        // dbuf (BN <= 64): the unit's accumulator set alternates, the whole epilogue overlaps the next unit's mainloop
        const uint32_t set = p.dbuf ? (t_local & 1u) : 0u;
        const uint32_t set_base = lane_base + set * 4u * (uint32_t)p.BN;
        uint64_t* set_empty = acc_empty + 4 * set;
        ptx::mbar_wait(tile_full + set, (p.dbuf ? (t_local >> 1) : t_local) & 1u);
        ptx::tc_fence_after();
        // stash this warp's chunks of S_0, then hand accumulator 0 back to the MMA thread
        float st[4][16];
#pragma unroll
        for (int ci = 0; ci < 4; ++ci)
          if (ci * 2 + half < nchunks) ptx::tmem_ld16(set_base + (uint32_t)((ci * 2 + half) * 16), st[ci]);
        ptx::tmem_ld_wait();
        release(set_empty + 0);
#pragma unroll
        for (int ci = 0; ci < 4; ++ci) {
          const int c0 = (ci * 2 + half) * 16;
          if (ci * 2 + half < nchunks) {
            float acc[3][16];
#pragma unroll
            for (int q = 1; q < 4; ++q) ptx::tmem_ld16(set_base + (uint32_t)(q * p.BN + c0), acc[q - 1]);
            ptx::tmem_ld_wait();
            if (p.bias != nullptr) {
#pragma unroll
              for (int j = 0; j < 16; ++j) st[ci][j] += __ldg(p.bias + n0 + c0 + j);
            }
            if (valid || p.tstore) {
#pragma unroll
              for (int pc = 0; pc < 4; ++pc) {
                float o[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  if constexpr (MIX)
                    o[j] = p.mix.m[pc * 4 + 0] * st[ci][j] + p.mix.m[pc * 4 + 1] * acc[0][j] + p.mix.m[pc * 4 + 2] * acc[1][j] +
                           p.mix.m[pc * 4 + 3] * acc[2][j];
                  else
                    o[j] = pc == 0 ? st[ci][j] : acc[pc == 0 ? 0 : pc - 1][j];
                }
                if constexpr (MIX) {
                  if (p.post_scale != nullptr) {      // inference: IQBN (running statistics) + activation in the epilogue
                    const float* sc = p.post_scale + pc * p.Cout + n0 + c0;
                    const float* sh = p.post_shift + pc * p.Cout + n0 + c0;
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                      const float z = fmaf(o[j], __ldg(sc + j), __ldg(sh + j));
                      o[j] = p.post_act == QUAN_ACT_SILU ? act_fwd<QUAN_ACT_SILU, sizeof(T) == 2>(z) : z;
                    }
                  }
                }
                if (p.tstore) {
                  tma_out(o, n0 + c0, pc, bw0, bh0, bb0);
                } else {
                  T* dst = yrow + (int64_t)pc * p.Cout + c0;
#pragma unroll
                  for (int v = 0; v < 16 / VW; ++v) {
                    float part[VW];
#pragma unroll
                    for (int j = 0; j < VW; ++j) part[j] = o[v * VW + j];
                    store_vec<T, VW>(dst + v * VW, part);
                  }
                }
              }
            }
            if constexpr (MIX) {
              if (p.stat_part != nullptr) {   // fused IQBN statistics: column sums of this warp's 32 rows (warp-uniform branch)
#pragma unroll
                for (int pc = 0; pc < 4; ++pc) {
                  float o[16];
#pragma unroll
                  for (int j = 0; j < 16; ++j)
                    o[j] = valid ? p.mix.m[pc * 4 + 0] * st[ci][j] + p.mix.m[pc * 4 + 1] * acc[0][j] + p.mix.m[pc * 4 + 2] * acc[1][j] +
                                       p.mix.m[pc * 4 + 3] * acc[2][j]
                                 : 0.f;
                  stat_add(sacc, p.stat_cq, pc * p.stat_cq + n0 + c0, o, lane);
                }
              }
            }
          }
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (ptx::elect_one()) {
#pragma unroll
          for (int a = 1; a < 4; ++a) {
            if constexpr (CG == 2) ptx::mbar_arrive_cluster(set_empty + a, 0);
            else ptx::mbar_arrive(set_empty + a);
          }
        }
        __syncwarp();
      } else {
        const uint32_t a = t_local & 1u;
        ptx::mbar_wait(tile_full + a, (t_local >> 1) & 1u);
        ptx::tc_fence_after();
        // one 16-column chunk of this thread's row: TMEM -> (+bias, eval-mode IQBN + act) -> global, and the IQBN statistics
        auto chunk = [&](const int c, float (&rsum)[RSTAT ? 16 : 1], float (&rsq)[RSTAT ? 16 : 1], const bool in_regs) {
          const int c0 = c * 16;
          float acc[16];
          ptx::tmem_ld16(lane_base + a * (uint32_t)p.BN + (uint32_t)c0, acc);
          ptx::tmem_ld_wait();
          if (p.bias != nullptr) {
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[j] += __ldg(p.bias + n0 + c0 + j);
          }
          if (p.post_scale != nullptr) {              // inference: column n = p*C_o + co is the table index
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float z = fmaf(acc[j], __ldg(p.post_scale + n0 + c0 + j), __ldg(p.post_shift + n0 + c0 + j));
              acc[j] = p.post_act == QUAN_ACT_SILU ? act_fwd<QUAN_ACT_SILU, sizeof(T) == 2>(z) : z;
            }
          }
          if (p.tstore) {
            tma_out(acc, n0 + c0, 0, bw0, bh0, bb0);
          } else if (valid) {
#pragma unroll
            for (int v = 0; v < 16 / VW; ++v) {
              float part[VW];
#pragma unroll
              for (int j = 0; j < VW; ++j) part[j] = acc[v * VW + j];
              store_vec<T, VW>(yrow + c0 + v * VW, part);
            }
          }
          if (p.stat_part != nullptr) {      // dense form: column n = p*C_o + co is already the (component, channel) pair
            if constexpr (RSTAT) {           // this thread's running column sums: folded across rows ONCE, after the last unit
              if (in_regs && valid) {
#pragma unroll
                for (int j = 0; j < 16; ++j) { rsum[j] += acc[j]; rsq[j] = fmaf(acc[j], acc[j], rsq[j]); }
              }
            } else {
              if (!valid) {
#pragma unroll
                for (int j = 0; j < 16; ++j) acc[j] = 0.f;
              }
              stat_add(sacc, p.stat_cq, n0 + c0, acc, lane);
            }
          }
        };
        if (nchunks <= 4) {                  // <= 2 chunks per warp (BN <= 64): compile-time register sets
          if (half < nchunks) chunk(half, rs0, rq0, reg_stats);
          if (half + 2 < nchunks) chunk(half + 2, rs1, rq1, reg_stats);
        } else {
          for (int c = half; c < nchunks; c += 2) chunk(c, rs0, rq0, false);
        }
        release(acc_empty + a);
      }
    }
    if constexpr (RSTAT) {                             // rows of the warp meet here, once per kernel
      if (half < nchunks) stat_add_sums(sacc, p.stat_cq, half * 16, rs0, rq0, lane);
      if (half + 2 < nchunks) stat_add_sums(sacc, p.stat_cq, (half + 2) * 16, rs1, rq1, lane);
    }
    if (p.stat_part != nullptr) {
      // this CTA's slot of the IQBN partials buffer (what iqbn_reduce_b writes): [2][C_o*4], index c*4 + q
      asm volatile("bar.sync 1, %0;" ::"n"(32 * EPI_WARPS * epi_groups(NQ)) : "memory");   // the epilogue warps only
      const int n4 = 4 * p.stat_cq;
      double* slot = p.stat_part + (size_t)blockIdx.x * 2 * n4;
      for (int e = threadIdx.x - 64; e < 2 * n4; e += 32 * EPI_WARPS * epi_groups(NQ)) {
        const int which = e / n4, r = e - which * n4;
        const int pc = r / p.stat_cq, co = r - pc * p.stat_cq;
        if (p.stat_acc != nullptr) atomicAdd(p.stat_acc + which * n4 + co * 4 + pc, (double)sacc[e]);   // <= 148 x 8C adds per launch
        else slot[which * n4 + co * 4 + pc] = (double)sacc[e];
      }
    }
    if (p.tstore && lane == 0) ptx::bulk_wait_read<0>();   // staged rows must outlive the stores that read them
    ptx::tc_fence_before();
  }
  if constexpr (CG == 2) ptx::cluster_sync_all();   // neither CTA may release TMEM / exit while the pair is still using it
  else __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    if constexpr (CG == 2) ptx::tmem_dealloc_2cta(tmem_base, p.tmem_cols);
    else ptx::tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// wgrad kernel:  dW_q[co][ci][tap] = sum_pixels G_q[pix][co] * x_q[pix (+) tap][ci]
// GEMM per (q, tap): D[M = 128 co][N ci] += A^T B with K = pixels.  Both operands are "MN-major" for the tensor core
// (channels contiguous, K = pixel rows of 128 bytes), which is exactly how the same TMA boxes as the forward pass land
// in shared memory — no transposes.  One CTA owns (q, one kh row of taps, a 128-wide co block, one ci atom) and a
// contiguous range of 64-pixel chunks (split-K); its TG accumulators (one per kw) live in TMEM.  Partials go to a
// [split][q][tap][co][ci] fp32 workspace and a small deterministic reduce kernel writes dW in the master layout.
// ---------------------------------------------------------------------------------------------------------------
struct TcWgradParams {
  int Wt, Ht, Bt, tiles_w, tiles_h, nchunks;   // 64-pixel chunk = Bt x Ht x Wt over the OUTPUT (G) grid
  int kH, kW, sH, sW, pH, pW, dH, dW;
  int Co, Ci, taps;            // channel counts of the GEMM (dense form: 4*C_o, 4*C_i)
  int nq;                      // 4: separable form (one GEMM per component); 1: dense Hamilton form
  int TG;                      // taps per CTA (= kW)
  int NA;                      // channels per swizzle atom row (128 / 64 / 32 bytes)
  int NB;                      // ci atoms per CTA: N = NB * NA
  int MA;                      // co atoms per CTA (M = 128 -> 128*es/128)
  int co_blocks, ci_blocks, tap_groups;
  int chunks_per_split;
  int ksteps;                  // UMMAs per chunk per tap
  uint32_t kadv;               // start-address advance per UMMA (in 16-byte units)
  int stages;
  uint32_t atom_bytes;         // 64 rows x 128 B
  uint32_t sbo_bytes, layout_type;   // bf16: 8-row groups, SWIZZLE_128B; tf32: 4-row groups, SWIZZLE_128B_BASE32B
  uint32_t a_stage_bytes, b_stage_bytes, idesc, tmem_cols;
  float* partial;              // [splits][nq][taps][Co][Ci]
  int atomic;                  // 1: every split adds into ONE zero-initialised partial set with vector atomics (narrow
                               // layers: many splits of a tiny dW — the fold of 64 partial sets cost more than the GEMM)
  // halo mode (stride 1 along W): ONE box per ci atom holds the chunk's pixels for all kW taps (Wt + (kW-1)*dW wide rows);
  // tap kw / K-step k are UMMA descriptors starting kw*dW + brow[k] pixel rows into it (absolute-address swizzle)
  int halo;
  uint32_t b_atom_bytes;       // slot pitch of one ci atom in the B stage (1024-byte multiple)
  uint16_t brow[8];            // K-step -> first pixel row inside the halo box
};

constexpr int WG_PIX = 64;     // pixels (K) per pipeline stage

template <typename T, int KSTEPS>
__global__ void __launch_bounds__(TC_THREADS, 1)
qconv_wgrad_kernel(const __grid_constant__ CUtensorMap map_g, const __grid_constant__ CUtensorMap map_x, const TcWgradParams p) {
  pdl_prologue();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + (size_t)p.stages * p.a_stage_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_b + (size_t)p.stages * p.b_stage_bytes);
  uint64_t* empty_bar = full_bar + p.stages;
  uint64_t* tmem_full_bar = empty_bar + p.stages;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // provably warp-uniform
  const int lane = threadIdx.x & 31;
  constexpr int KIND = sizeof(T) == 2 ? 0 : 1;
  const int N = p.NB * p.NA;

  // work decomposition
  int combo = blockIdx.y;
  const int nb = combo % p.ci_blocks; combo /= p.ci_blocks;
  const int mb = combo % p.co_blocks; combo /= p.co_blocks;
  const int tg = combo % p.tap_groups;
  const int q = combo / p.tap_groups;
  const int split = blockIdx.x;
  const int chunk0 = split * p.chunks_per_split;
  const int chunk1 = min(chunk0 + p.chunks_per_split, p.nchunks);
  const int nch = chunk1 - chunk0;
  const int co0 = mb * 128, ci0 = nb * N;
  const int kh = tg;   // tap group = one filter row

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&map_g);
    ptx::prefetch_tensormap(&map_x);
    for (int s = 0; s < p.stages; ++s) {
      ptx::mbar_init(full_bar + s, 1);
      ptx::mbar_init(empty_bar + s, 1);
    }
    ptx::mbar_init(tmem_full_bar, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc(tmem_ptr, p.tmem_cols);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===== TMA producer: warp-uniform loop, one elected lane issues =====
    // co atoms past C_o are not loaded (their accumulator rows are never stored)
    const int ma_load = min(p.MA, (p.Co - co0 + p.NA - 1) / p.NA);
    const uint32_t halo_rows = (uint32_t)(p.Bt * p.Ht * (p.Wt + (p.kW - 1) * p.dW));
    const uint32_t tx = (uint32_t)ma_load * p.atom_bytes +
                        (p.halo ? (uint32_t)p.NB * halo_rows * (p.atom_bytes / WG_PIX) : p.b_stage_bytes);
    // chunk -> (tb, th, tw) once, then incrementally (no div/mod on the producer's critical path)
    int tw = chunk0 % p.tiles_w;
    int th = (chunk0 / p.tiles_w) % p.tiles_h;
    int tb = chunk0 / (p.tiles_w * p.tiles_h);
    int s = 0;
    uint32_t phase = 0;
    for (int it = 0; it < nch; ++it) {
      ptx::mbar_wait(empty_bar + s, phase ^ 1);
      if (ptx::elect_one()) {
        const int w0 = tw * p.Wt, h0 = th * p.Ht, b0 = tb * p.Bt;
        ptx::mbar_arrive_expect_tx(full_bar + s, tx);
        uint8_t* a_dst = smem_a + (size_t)s * p.a_stage_bytes;
        for (int a = 0; a < ma_load; ++a)
          ptx::tma_load_5d(a_dst + (size_t)a * p.atom_bytes, &map_g, full_bar + s, co0 + a * p.NA, q, w0, h0, b0);
        uint8_t* b_dst = smem_b + (size_t)s * p.b_stage_bytes;
        const int wc = w0 * p.sW - p.pW, hc = h0 * p.sH - p.pH + kh * p.dH;
        if (p.halo) {
          for (int a = 0; a < p.NB; ++a)
            ptx::tma_load_5d(b_dst + (size_t)a * p.b_atom_bytes, &map_x, full_bar + s, ci0 + a * p.NA, q, wc, hc, b0);
        } else {
          for (int t = 0; t < p.TG; ++t)
            for (int a = 0; a < p.NB; ++a)
              ptx::tma_load_5d(b_dst + (size_t)(t * p.NB + a) * p.atom_bytes, &map_x, full_bar + s, ci0 + a * p.NA, q,
                               wc + t * p.dW, hc, b0);
        }
      }
      __syncwarp();
      if (++s == p.stages) { s = 0; phase ^= 1; }
      if (++tw == p.tiles_w) { tw = 0; if (++th == p.tiles_h) { th = 0; ++tb; } }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: warp-uniform loop, one elected lane issues =====
    // MN-major, 128B swizzle: LBO = distance between 128-byte column atoms, SBO = one K-row group
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint64_t da0 = ptx::make_smem_desc(ptx::smem_u32(smem_a), p.atom_bytes, p.sbo_bytes, p.layout_type);
    const uint64_t db0 = ptx::make_smem_desc(ptx::smem_u32(smem_b), p.halo ? p.b_atom_bytes : p.atom_bytes, p.sbo_bytes, p.layout_type);
    const uint64_t a_step = (uint64_t)(p.a_stage_bytes >> 4), b_step = (uint64_t)(p.b_stage_bytes >> 4);
    // per tap: the next NB atoms (plain mode) or kw*dW pixel rows further into the halo box (halo mode)
    const uint32_t row16 = (p.atom_bytes / WG_PIX) >> 4;                  // one pixel row in 16-byte units
    const uint64_t tap_step = p.halo ? (uint64_t)(p.dW * row16) : (uint64_t)((p.atom_bytes * p.NB) >> 4);
    int s = 0;
    uint32_t phase = 0;
    uint64_t da = da0, db = db0;
    for (int it = 0; it < nch; ++it) {
      ptx::mbar_wait(full_bar + s, phase);
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
        uint64_t dbt = db;
        for (int t = 0; t < p.TG; ++t, dbt += tap_step) {
          const uint32_t d_tmem = tmem_u + (uint32_t)(t * N);
#pragma unroll
          for (int k = 0; k < KSTEPS; ++k)
            ptx::umma<KIND>(d_tmem, da + (uint64_t)(k * p.kadv), dbt + (uint64_t)(p.halo ? p.brow[k] * row16 : k * p.kadv), p.idesc,
                            (it | k) ? 1u : 0u);
        }
        ptx::umma_commit(empty_bar + s);
      }
      __syncwarp();
      da += a_step;
      db += b_step;
      if (++s == p.stages) { s = 0; phase ^= 1; da = da0; db = db0; }
    }
    if (ptx::elect_one()) ptx::umma_commit(tmem_full_bar);
    __syncwarp();
  } else {
    const int quarter = warp & 3;
    const int m = quarter * 32 + lane;            // co row inside the 128 block
    const int co = co0 + m;
    ptx::mbar_wait(tmem_full_bar, 0);
    ptx::tc_fence_after();
    const uint32_t lane_base = tmem_base + ((uint32_t)(quarter * 32) << 16);
    for (int t = 0; t < p.TG; ++t) {
      const int tap = kh * p.kW + t;
      float* dst = p.partial + ((((int64_t)(p.atomic ? 0 : split) * p.nq + q) * p.taps + tap) * p.Co + co) * p.Ci + ci0;
      for (int c0 = 0; c0 < N; c0 += 16) {
        float acc[16];
        ptx::tmem_ld16(lane_base + (uint32_t)(t * N + c0), acc);
        ptx::tmem_ld_wait();
        if (co < p.Co && nch > 0) {
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            float part[4] = {acc[v * 4], acc[v * 4 + 1], acc[v * 4 + 2], acc[v * 4 + 3]};
            if (p.atomic) atomicAdd(reinterpret_cast<float4*>(dst + c0 + v * 4), make_float4(part[0], part[1], part[2], part[3]));
            else store_vec<float, 4>(dst + c0 + v * 4, part);
          }
        }
      }
    }
    ptx::tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// Split-K fold + transpose to the master layout.  Block = (q, co, ci chunk): for every tap it sums the splits of a
// contiguous ci-row of the partials (SL split lanes per element when the row is short and the splits are many — narrow
// layers), stages [ci][tap] in shared memory and writes the chunk*taps contiguous floats of dW_q[co].  Fixed summation
// order: deterministic.
//   separable: dW_q[co][ci][tap] = sum_split partial[split][q][tap][co][ci]
//   dense    : partial[split][tap][p*Co + co][q*Ci + ci] holds sum_pix dY_p[co] x_q[ci]; the mixing matrix is applied
//              here: dW_q[co][ci][tap] = sum_p M[p][q] sum_split partial[...]   (G = M^T dY never materialises)
template <bool DENSE>
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dw0,
                                                           float* __restrict__ dw1, float* __restrict__ dw2,
                                                           float* __restrict__ dw3, int splits, int taps, int Co, int Ci,
                                                           int cchunk, int SL, const Mix16 mix) {
  pdl_prologue();
  extern __shared__ float tile[];                     // [SL][nci*taps] partial sums, folded into lane 0's slice
  const int q = blockIdx.y / Co, co = blockIdx.y % Co;
  const int ci0 = blockIdx.x * cchunk, nci = min(cchunk, Ci - ci0);
  const int items = nci * taps;
  const int64_t split_stride = (int64_t)(DENSE ? 16 : 4) * taps * Co * Ci;
  // small blocks (items * SL <= blockDim: narrow layers): one (split lane, tap, ci) per thread, one division per thread;
  // large blocks (wide layers, SL = 1): nested tap / ci loops, no division per element
  const bool small = items * SL <= (int)blockDim.x;
  const int lanes_ci = small ? nci : (int)blockDim.x;
  const int sl = small ? (int)threadIdx.x / items : 0;
  const int rem = small ? (int)threadIdx.x % items : 0;
  for (int tap = small ? rem / nci : 0; tap < taps; tap += small ? taps : 1) {
    if (small && sl >= SL) break;
    for (int ci = small ? rem % nci : (int)threadIdx.x; ci < nci; ci += lanes_ci) {
      float s = 0.f;
      if constexpr (DENSE) {
#pragma unroll
        for (int pc = 0; pc < 4; ++pc) {
          const float* src = partial + (((int64_t)tap * 4 * Co + pc * Co + co) * 4 * Ci + q * Ci + ci0 + ci);
          float t0 = 0.f, t1 = 0.f;
          int sp = sl;
          for (; sp + SL < splits; sp += 2 * SL) {
            t0 += __ldg(src + sp * split_stride);
            t1 += __ldg(src + (sp + SL) * split_stride);
          }
          if (sp < splits) t0 += __ldg(src + sp * split_stride);
          s += mix.m[pc * 4 + q] * (t0 + t1);
        }
      } else {
        const float* src = partial + ((((int64_t)q * taps + tap) * Co + co) * Ci + ci0 + ci);
        float t0 = 0.f, t1 = 0.f;
        int sp = sl;
        for (; sp + SL < splits; sp += 2 * SL) {
          t0 += __ldg(src + sp * split_stride);
          t1 += __ldg(src + (sp + SL) * split_stride);
        }
        if (sp < splits) t0 += __ldg(src + sp * split_stride);
        s = t0 + t1;
      }
      tile[sl * items + ci * taps + tap] = s;
    }
  }
  __syncthreads();
  float* dw = (q == 0 ? dw0 : q == 1 ? dw1 : q == 2 ? dw2 : dw3) + ((int64_t)co * Ci + ci0) * taps;
  for (int e = threadIdx.x; e < items; e += blockDim.x) {
    float s = tile[e];
    for (int g = 1; g < SL; ++g) s += tile[g * items + e];
    dw[e] = s;
  }
}

// Fold for small dW (narrow layers, few splits left after the atomic accumulation): one thread per dW element in output
// order (q, co, ci, tap), all of its loads independent — the blocked kernel above spends ~25 us of load latency per
// launch there (ncu launch list of the yolo11n trace: 23-28 us for 3x3 layers whatever the grid size).
template <bool DENSE>
__global__ void __launch_bounds__(256) wgrad_reduce_flat_kernel(const float* __restrict__ partial, float* __restrict__ dw0,
                                                                float* __restrict__ dw1, float* __restrict__ dw2,
                                                                float* __restrict__ dw3, int splits, int taps, int Co, int Ci,
                                                                const Mix16 mix) {
  pdl_prologue();
  const int per_q = Co * Ci * taps;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= 4 * per_q) return;
  const int q = e / per_q, r = e - q * per_q;
  const int tap = r % taps, cc = r / taps;
  const int ci = cc % Ci, co = cc / Ci;
  const int64_t split_stride = (int64_t)(DENSE ? 16 : 4) * taps * Co * Ci;
  float s;
  if constexpr (DENSE) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    const float* src = partial + (((int64_t)tap * 4 * Co + co) * 4 * Ci + q * Ci + ci);
    const int64_t pc_stride = (int64_t)Co * 4 * Ci;
    for (int sp = 0; sp < splits; ++sp) {
#pragma unroll
      for (int pc = 0; pc < 4; ++pc) acc[pc] += __ldg(src + sp * split_stride + pc * pc_stride);
    }
    s = mix.m[q] * acc[0] + mix.m[4 + q] * acc[1] + mix.m[8 + q] * acc[2] + mix.m[12 + q] * acc[3];
  } else {
    const float* src = partial + ((((int64_t)q * taps + tap) * Co + co) * Ci + ci);
    float t0 = 0.f, t1 = 0.f;
    int sp = 0;
    for (; sp + 1 < splits; sp += 2) {
      t0 += __ldg(src + sp * split_stride);
      t1 += __ldg(src + (sp + 1) * split_stride);
    }
    if (sp < splits) t0 += __ldg(src + sp * split_stride);
    s = t0 + t1;
  }
  (q == 0 ? dw0 : q == 1 ? dw1 : q == 2 ? dw2 : dw3)[r] = s;
}

// depthwise layer run as the dense form: only the block diagonal of the dense weight gradient is a parameter gradient,
// dW_q[c][tap] = sum_p M[p][q] * partial[tap][p*C + c][q*C + c]; one thread per (q, c, tap)
__global__ void __launch_bounds__(256) wgrad_reduce_dw_kernel(const float* __restrict__ partial, float* __restrict__ dw0, float* __restrict__ dw1,
                                                              float* __restrict__ dw2, float* __restrict__ dw3, int splits, int taps, int C,
                                                              const Mix16 mix) {
  pdl_prologue();
  const int per_q = C * taps;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= 4 * per_q) return;
  const int q = e / per_q, r = e - q * per_q;
  const int tap = r % taps, c = r / taps;
  const int64_t split_stride = (int64_t)16 * taps * C * C;
  const float* src = partial + (((int64_t)tap * 4 * C + c) * 4 * C + q * C + c);
  const int64_t pc_stride = (int64_t)C * 4 * C;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int sp = 0; sp < splits; ++sp) {
#pragma unroll
    for (int pc = 0; pc < 4; ++pc) acc[pc] += __ldg(src + sp * split_stride + pc * pc_stride);
  }
  (q == 0 ? dw0 : q == 1 ? dw1 : q == 2 ? dw2 : dw3)[r] = mix.m[q] * acc[0] + mix.m[4 + q] * acc[1] + mix.m[8 + q] * acc[2] + mix.m[12 + q] * acc[3];
}

// dense Hamilton weights for narrow layers: one real conv over all 4*C_q channels with the mixing matrix folded in,
//   FWD  : Wp[tap][n = p*Co + co][k = q*Ci + ci] = M[p][q] W_q[co][ci][tap];  bias'[p*Co + co] = M[p][0] b_r[co]
//   DGRAD: Wp[tap][n = q*Ci + ci][k = p*Co + co] = M[p][q] W_q[co][ci][taps-1-tap]      (input is dY itself)
template <typename T, bool DGRAD>
__global__ void __launch_bounds__(256) pack_weights_dense_kernel(const float* __restrict__ w0, const float* __restrict__ w1,
                                                                 const float* __restrict__ w2, const float* __restrict__ w3,
                                                                 const float* __restrict__ bias_r, T* __restrict__ out,
                                                                 float* __restrict__ bias_out, int Co, int Ci, int taps,
                                                                 const Mix16 mix, int depthwise) {
  pdl_prologue();
  const int N = DGRAD ? 4 * Ci : 4 * Co, K = DGRAD ? 4 * Co : 4 * Ci;
  const int64_t total = (int64_t)taps * N * K;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(i % K);
    int64_t r = i / K;
    const int n = (int)(r % N);
    const int t = (int)(r / N);
    int pc, q, co, ci, tap;
    if constexpr (DGRAD) { q = n / Ci; ci = n % Ci; pc = k / Co; co = k % Co; tap = taps - 1 - t; }
    else { pc = n / Co; co = n % Co; q = k / Ci; ci = k % Ci; tap = t; }
    const float* w = q == 0 ? w0 : q == 1 ? w1 : q == 2 ? w2 : w3;
    // depthwise: the master weight is [Co][1][taps]; the dense matrix is its block diagonal
    float v = depthwise ? (ci == co ? mix.m[pc * 4 + q] * __ldg(w + (int64_t)co * taps + tap) : 0.f)
                        : mix.m[pc * 4 + q] * __ldg(w + ((int64_t)co * Ci + ci) * taps + tap);
    if constexpr (sizeof(T) == 4) {
      uint32_t r32;
      asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r32) : "f"(v));
      v = __uint_as_float(r32);
    }
    out[i] = from_f32<T>(v);
    if (!DGRAD && bias_out != nullptr && i < 4 * Co) bias_out[i] = bias_r ? mix.m[(i / Co) * 4] * __ldg(bias_r + i % Co) : 0.f;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
static int pow2_ceil(int v) {
  int r = 1;
  while (r < v) r <<= 1;
  return r;
}
static int pick_bn(int n, int cap) {   // largest multiple of 16 that is <= cap and divides n
  for (int bn = cap; bn >= 16; bn -= 16)
    if (n % bn == 0) return bn;
  return 0;
}
static int pick_row_bytes(int k_elems, int esz) {   // widest swizzle row that tiles the K extent
  const int kb = k_elems * esz;
  if (kb % 128 == 0) return 128;
  if (kb % 64 == 0) return 64;
  if (kb % 32 == 0) return 32;
  return 0;
}

struct TilePlan {
  int Wt, Ht, Bt, tiles_w, tiles_h, tiles_b;
};
static bool plan_tiles(int B, int Ho, int Wo, int sH, int sW, TilePlan& t) {
  t.Wt = pow2_ceil(Wo) < 128 ? pow2_ceil(Wo) : 128;
  while (t.Wt * sW > 256) t.Wt >>= 1;                 // TMA box extent limit
  int rem = 128 / t.Wt;
  t.Ht = pow2_ceil(Ho) < rem ? pow2_ceil(Ho) : rem;
  while (t.Ht * sH > 256) t.Ht >>= 1;
  t.Bt = 128 / (t.Wt * t.Ht);
  if (t.Bt > B || t.Bt > 256) return false;           // tiny problems stay on the direct engine
  t.tiles_w = (Wo + t.Wt - 1) / t.Wt;
  t.tiles_h = (Ho + t.Ht - 1) / t.Ht;
  t.tiles_b = (B + t.Bt - 1) / t.Bt;
  return true;
}

// geometry of a conv expressed as "output [B,Ho,Wo,N] from input [B,Hi,Wi,K]" (dgrad swaps the roles)
// nq = 4: separable form, tensors [B][H][W][4][K or N]; nq = 1: dense Hamilton form, K and N count all 4*C_q real channels
// scat > 1: stride-`scat` dgrad — (sH,sW) are 1, the output grid is visited in scat*scat parity classes (TapTable)
struct IgemmShape {
  int B, Hi, Wi, K, Ho, Wo, N, kH, kW, sH, sW, pH, pW, dH, dW, nq, scatH, scatW;
  const char* name;   // kernel label for the launch log / timing table
};
static bool build_taps(const IgemmShape& s, TapTable& t) {
  if (s.kH * s.kW > TC_MAX_TAPS || s.scatH * s.scatW > 4) return false;
  t = TapTable{};
  int n = 0;
  if (s.scatH == 1 && s.scatW == 1) {
    t.ncls = 1;
    for (int kh = 0; kh < s.kH; ++kh)
      for (int kw = 0; kw < s.kW; ++kw) {
        const int dh = kh * s.dH - s.pH, dw = kw * s.dW - s.pW;
        if (dh < -128 || dh > 127 || dw < -128 || dw > 127) return false;
        t.dh[n] = (int8_t)dh; t.dw[n] = (int8_t)dw; t.tap[n] = (uint8_t)(kh * s.kW + kw);
        ++n;
      }
    t.start[1] = n;
    return true;
  }
  // strided dgrad: s.pH/pW hold the conv's own padding, weights are packed tap-flipped (pack_weights*<DGRAD>)
  int c = 0;
  for (int ph = 0; ph < s.scatH; ++ph)
    for (int pw = 0; pw < s.scatW; ++pw) {
      t.start[c] = n;
      t.oh[c] = (int8_t)ph; t.ow[c] = (int8_t)pw;
      for (int kh = 0; kh < s.kH; ++kh)
        for (int kw = 0; kw < s.kW; ++kw) {
          const int nh = ph + s.pH - kh * s.dH, nw = pw + s.pW - kw * s.dW;
          if (((nh % s.scatH) + s.scatH) % s.scatH != 0 || ((nw % s.scatW) + s.scatW) % s.scatW != 0) continue;
          const int dh = nh / s.scatH, dw = nw / s.scatW;   // exact
          if (dh < -128 || dh > 127 || dw < -128 || dw > 127) return false;
          t.dh[n] = (int8_t)dh; t.dw[n] = (int8_t)dw;
          t.tap[n] = (uint8_t)((s.kH - 1 - kh) * s.kW + (s.kW - 1 - kw));
          ++n;
        }
      if (n == t.start[c]) t.zero_fill = 1;   // a parity without taps (e.g. 1x1 stride 2) is all zeros: no class for it
      else ++c;
    }
  if (c == 0) return false;
  t.ncls = c;
  t.start[t.ncls] = n;
  return true;
}
static int bn_cap(int nq) { return nq == 4 ? 128 : 256; }   // TMEM: 4 accumulators of BN, or 2 of BN

static bool igemm_supported(const IgemmShape& s, int dtype) {
  const int esz = dtype == QUAN_BF16 ? 2 : 4;
  if (pick_row_bytes(s.K, esz) == 0) return false;
  if (pick_bn(s.N, bn_cap(s.nq)) == 0) return false;
  if ((s.N * esz) % 16 != 0 || (s.K * esz) % 16 != 0) return false;
  if (s.sH > 8 || s.sW > 8) return false;             // TMA element-stride limit
  TilePlan t;
  if (!plan_tiles(s.B, (s.Ho + s.scatH - 1) / s.scatH, (s.Wo + s.scatW - 1) / s.scatW, s.sH, s.sW, t)) return false;
  if ((int64_t)t.tiles_w * t.tiles_h * t.tiles_b * s.scatH * s.scatW > 0x3fffffff) return false;
  TapTable tt;
  return build_taps(s, tt);
}


template <typename T, bool MIX, int CG, int KSTEPS, int NQ, bool RSTAT = false>
static int launch_igemm_inst(const CUtensorMap& map_a, const CUtensorMap& map_b, const CUtensorMap& map_y, void* out,
                             const TcConvParams& p, size_t smem,
                             const char* name, cudaStream_t st, int* ctas) {
  auto kern = qconv_igemm_kernel<T, MIX, CG, KSTEPS, NQ, RSTAT>;
  // persistent grid: one CTA (pair) per SM, as many as can be co-resident (queried once per instantiation)
  static thread_local int max_groups = 0;
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(ig_threads(NQ));
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;      // PDL, see common.cuh pdl_prologue()
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  static thread_local DeviceOnce attr_set;
  if (attr_set.first()) QUAN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  if (max_groups == 0) {
    int n = 0;
    if (CG == 2) {
      cfg.gridDim = dim3(QUAN_NUM_SMS);
      QUAN_CUDA(cudaOccupancyMaxActiveClusters(&n, kern, &cfg));
    } else {
      int dev = 0, sms = 0;
      QUAN_CUDA(cudaGetDevice(&dev));
      QUAN_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
      n = sms;
    }
    QUAN_REQUIRE(n > 0, QUAN_E_DRIVER, "tcgen05 conv: no co-resident CTA %s fits on this device", CG == 2 ? "pair" : "");
    max_groups = n;
  }
  const int groups = p.units < max_groups ? p.units : max_groups;
  cfg.gridDim = dim3((unsigned)(groups * CG));
  *ctas = groups * CG;
  QUAN_TIMED(st);
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  QUAN_CUDA(cudaLaunchKernelEx(&cfg, kern, map_a, map_b, map_y, reinterpret_cast<T*>(out), p));
  QUAN_CHECK_LAUNCH(name);
  return QUAN_OK;
}

struct PostOp {                        // eval-mode IQBN + activation folded into the forward epilogue
  const float* scale = nullptr;
  const float* shift = nullptr;
  int act = QUAN_ACT_NONE;
};
template <typename T, bool MIX, int NQ>
static int launch_igemm(const void* in, const void* wpacked, const float* bias, void* out, const IgemmShape& s, int dtype,
                        const Mix16& mix, cudaStream_t st, double* stat_part = nullptr, int* stat_nparts = nullptr,
                        const PostOp* post = nullptr) {
  const int esz = sizeof(T);
  const int row_bytes = pick_row_bytes(s.K, esz);
  TilePlan t;
  TcConvParams p = {};
  QUAN_REQUIRE(row_bytes != 0 && s.nq == NQ && build_taps(s, p.tt) &&
                   plan_tiles(s.B, (s.Ho + s.scatH - 1) / s.scatH, (s.Wo + s.scatW - 1) / s.scatW, s.sH, s.sW, t),
               QUAN_E_UNSUPPORTED, "tcgen05 conv: shape does not qualify");
  // halo mode (see TcConvParams): stride 1, one class, 128-byte K rows, 16 x 8-pixel tiles of one image that cover the
  // output without much waste, receptive field within a 16-pixel-wide box.  QUAN_TC_HALO=0 disables, =1 default.
  int halo_h = 0;
  {
    static const int env_halo = [] { const char* e = getenv("QUAN_TC_HALO"); return e ? atoi(e) : 1; }();
    int dwmin = 0, dwmax = 0, dhmin = 0, dhmax = 0;
    for (int i = 0; i < p.tt.start[1]; ++i) {
      dwmin = i ? (p.tt.dw[i] < dwmin ? p.tt.dw[i] : dwmin) : p.tt.dw[i];
      dwmax = i ? (p.tt.dw[i] > dwmax ? p.tt.dw[i] : dwmax) : p.tt.dw[i];
      dhmin = i ? (p.tt.dh[i] < dhmin ? p.tt.dh[i] : dhmin) : p.tt.dh[i];
      dhmax = i ? (p.tt.dh[i] > dhmax ? p.tt.dh[i] : dhmax) : p.tt.dh[i];
    }
    const int tw8 = (s.Wo + 7) / 8, th16 = (s.Ho + 15) / 16;
    const double cover = (double)s.Ho * s.Wo / ((double)tw8 * 8 * th16 * 16);
    halo_h = 16 + dhmax - dhmin;
    const size_t a_bytes = (size_t)halo_h * 16 * 128;
    // NQ = 1 (dense form of the narrow layers, 4*C_q real channels per 128-byte row = C_q 16 in bf16, k-blocks beyond that): the
    // same kernel path; per-tap mode re-reads the activation tile once per filter tap through L2 -> SM (9x for 3x3), which is
    // what bounds these HBM-sized layers.  QUAN_TC_HALO_DENSE=0 keeps them on the per-tap path.
    static const int env_halo_dense = [] { const char* e = getenv("QUAN_TC_HALO_DENSE"); return e ? atoi(e) : 1; }();
    p.halo = env_halo && (NQ == 4 || env_halo_dense) && row_bytes == 128 && p.tt.ncls == 1 && s.sH == 1 && s.sW == 1 && p.tt.start[1] > 1 &&
             8 + dwmax - dwmin <= 16 && halo_h <= 256 && 2 * a_bytes <= 96 * 1024 && cover >= 0.8;
    if (p.halo) {
      t.Wt = 8; t.Ht = 16; t.Bt = 1; t.tiles_w = tw8; t.tiles_h = th16; t.tiles_b = s.B;
      p.halo_dw = dwmin; p.halo_dh = dhmin;
      p.a_stages = 2;
      p.a_halo_bytes = (uint32_t)a_bytes;
      p.tg = s.kW;                                      // one filter row of taps per B ring slot
      // measured on B200: the 128B swizzle XOR is taken from the absolute shared-memory address bits, for TMA writes and
      // UMMA reads alike — a window that starts on any 128-byte row is read correctly with base offset 0 (setting the
      // descriptor's base-offset field to the row phase gives wrong results: tests/tc_probe.py, QUAN_TC_HALO_BO=1)
      p.halo_bo = 0;
      if (const char* e = getenv("QUAN_TC_HALO_BO")) p.halo_bo = atoi(e) != 0;
      for (int i = 0; i < p.tt.start[1]; ++i) p.arow[i] = (uint16_t)((p.tt.dh[i] - dhmin) * 16 + (p.tt.dw[i] - dwmin));
    }
  }
  p.B = s.B; p.Ho = s.Ho; p.Wo = s.Wo; p.Cout = s.N;
  p.Wt = t.Wt; p.Ht = t.Ht; p.Bt = t.Bt; p.tiles_w = t.tiles_w; p.tiles_h = t.tiles_h;
  p.in_sW = s.sW; p.in_sH = s.sH; p.out_sW = s.scatW; p.out_sH = s.scatH;
  p.bk_elems = row_bytes / esz;
  p.kblocks = s.K / p.bk_elems;
  const int ksteps = row_bytes / 32;
  p.BN = pick_bn(s.N, bn_cap(NQ));
  const int64_t mtiles = (int64_t)t.tiles_w * t.tiles_h * t.tiles_b;
  // CTA pairs (cta_group::2) when the weight tile splits into two halves that are themselves legal UMMA-N slices.
  // Measured on B200 (C=256 3x3, non-persistent kernel): 283 us with pairs vs 312 us without.
  int cg = (p.BN % 32 == 0 && mtiles >= 2) ? 2 : 1;
  if (const char* e = getenv("QUAN_TC_CG")) { if (atoi(e) == 1) cg = 1; }
  p.ntiles_n = s.N / p.BN;
  p.units_per_cls = (int)((mtiles + cg - 1) / cg) * p.ntiles_n;   // whole pairs; a spare tile is masked (TMA zero fill + row mask)
  p.units = p.units_per_cls * p.tt.ncls;
  p.fd_upc = make_fastdiv((uint32_t)p.units_per_cls); p.fd_ntn = make_fastdiv((uint32_t)p.ntiles_n);
  p.fd_tw = make_fastdiv((uint32_t)t.tiles_w); p.fd_th = make_fastdiv((uint32_t)t.tiles_h);
  p.a_sub_bytes = 128u * row_bytes;
  p.b_sub_bytes = (uint32_t)(p.BN / cg) * row_bytes;
  p.sbo_bytes = 8u * row_bytes;
  p.layout_type = row_bytes == 128 ? 2u : row_bytes == 64 ? 4u : 6u;
  p.idesc = ptx::make_idesc(sizeof(T) == 2 ? 1u : 2u, 0u, 0u, 128u * cg, (uint32_t)p.BN);
  p.dbuf = (NQ == 4 && 8 * p.BN <= 512) ? 1 : 0;
  if (const char* e = getenv("QUAN_TC_DBUF")) { if (atoi(e) == 0) p.dbuf = 0; }
  const int acc_cols = (NQ == 4 ? (p.dbuf ? 8 : 4) : 2) * p.BN;
  p.tmem_cols = (uint32_t)pow2_ceil(acc_cols < 32 ? 32 : acc_cols);
  p.bias = bias;
  if (post != nullptr) { p.post_scale = post->scale; p.post_shift = post->shift; p.post_act = post->act; }
  p.stat_cq = NQ == 4 ? s.N : s.N / 4;
  // fused IQBN partial statistics: needs the [2][4][C_o] fp32 accumulators next to the pipeline stages in shared memory
  // Measured on B200 (bench shape, separable form): the 31-shuffle column reduction per 16 outputs makes the epilogue
  // longer than the q = 0 mainloop it hides behind — igemm 220 -> 293 us, more than the 38 + 11 us statistics pass it
  // replaces — so the wide layers keep the streaming statistics kernel; the dense form (narrow, HBM-bound layers with an
  // idle tensor pipe) emits them.  QUAN_TC_EPI_STATS=2 forces, =0 disables.
  static const int epi_stats = [] { const char* e = getenv("QUAN_TC_EPI_STATS"); return e ? atoi(e) : 1; }();
  const bool want_stats = epi_stats == 2 || (epi_stats == 1 && NQ == 1);
  p.stat_part = (stat_part != nullptr && want_stats && (size_t)32 * p.stat_cq <= 24 * 1024) ? stat_part : nullptr;
  p.mix = mix;
  int iters_per_q = 1 << 30;                         // of the smallest class
  for (int c = 0; c < p.tt.ncls; ++c) {
    const int it = (p.tt.start[c + 1] - p.tt.start[c]) * p.kblocks;
    if (it < iters_per_q) iters_per_q = it;
  }
  p.sub = iters_per_q >= 2 ? 2 : 1;                  // 8 MMAs per barrier round trip when there is enough K
  if (const char* e = getenv("QUAN_TC_SUB")) { int v = atoi(e); if (v >= 1 && v <= 4 && v <= iters_per_q) p.sub = v; }
  if (p.halo) p.sub = p.tg;                          // halo mode: a B ring slot holds one filter row of taps; A has its own ring
  const size_t a_ring = p.halo ? (size_t)p.a_stages * p.a_halo_bytes : 0;
  const size_t stage_bytes = p.halo ? (size_t)p.b_sub_bytes * p.sub : (size_t)(p.a_sub_bytes + p.b_sub_bytes) * p.sub;
  // TMA-store epilogue: whenever the output is written with stride 1 (forward of any stride, dgrad of stride-1 layers); the strided
  // scatter of the parity-class dgrad keeps per-lane stores (its pixel dims would need a rank-7 map).  QUAN_TC_TSTORE=0 disables.
  static const int env_tstore = [] { const char* e = getenv("QUAN_TC_TSTORE"); return e ? atoi(e) : 1; }();
  p.tstore = (env_tstore && s.scatH == 1 && s.scatW == 1 && !p.tt.zero_fill && (s.N * esz) % 16 == 0 &&
              (reinterpret_cast<uintptr_t>(out) & 15) == 0)
                 ? (env_tstore == 4 || env_tstore == 2 ? env_tstore : 2)
                 : 0;
  const size_t stat_bytes = p.stat_part != nullptr ? (size_t)8 * p.stat_cq * sizeof(float) : 0;
  size_t ts_bytes = 0;
  int stages = 0;
  for (;;) {
    ts_bytes = (size_t)EPI_WARPS * epi_groups(NQ) * p.tstore * 32 * 16 * esz;   // 2 slots per warp: 16 KB (bf16) / 32 KB (fp32)
    size_t budget = 226 * 1024 - 1024 - 512 - stat_bytes - ts_bytes;
    if (budget > 200 * 1024) budget = 200 * 1024;
    // dense (narrow-layer) form: the pipeline's shared memory is capped so that two kernels of the step (the main chain's igemm / IQBN
    // kernels and a deferred wgrad on the side stream) can be co-resident on an SM instead of taking turns — with 200 KB
    // rings every persistent kernel owned its SM and the side stream only ran in the gaps.  Measured on B200 (QUAN-YOLO11n step, three
    // alternating runs each): no caps 11.20 ms, wgrad ring <= 100 KB 10.98, both <= 100 KB 10.81-10.95, 72 / 72 10.86-11.03, wgrad <= 64 KB
    // 11.0-11.2 (its own prefetch depth starts to matter).  QUAN_TC_DENSE_SMEM_KB / QUAN_TC_WG_SMEM_KB = 0 lift the caps.
    if (NQ == 1) {
      static const int cap_kb = [] { const char* e = getenv("QUAN_TC_DENSE_SMEM_KB"); return e ? atoi(e) : 96; }();
      const size_t cap = (size_t)cap_kb * 1024;
      if (cap_kb > 0 && budget > cap && cap >= a_ring + 2 * stage_bytes) budget = cap;    // never below a two-stage ring
    }
    stages = budget > a_ring ? (int)((budget - a_ring) / stage_bytes) : 0;
    if (stages >= 2 || p.tstore == 0) break;
    p.tstore = 0;                                      // a two-stage ring of big tiles needs the staging slots' room: per-lane stores
  }
  if (stages > 8) stages = 8;
  if (const char* e = getenv("QUAN_TC_STAGES")) { int v = atoi(e); if (v >= 1 && v < stages) stages = v; }
  QUAN_REQUIRE(stages >= 2, QUAN_E_UNSUPPORTED, "tcgen05 conv: stage too large");
  p.stages = stages;
  const size_t smem = 1024 + ts_bytes + a_ring + stages * stage_bytes + (2 * stages + 18) * sizeof(uint64_t) + 16 + stat_bytes;

  // A: input activations [B][Hi][Wi][nq][K] -> 5-D map {K, nq, Wi, Hi, B}
  CUtensorMap map_a, map_b;
  {
    const uint64_t dims[5] = {(uint64_t)s.K, (uint64_t)NQ, (uint64_t)s.Wi, (uint64_t)s.Hi, (uint64_t)s.B};
    const uint64_t str[4] = {(uint64_t)s.K * esz, (uint64_t)NQ * s.K * esz, (uint64_t)s.Wi * NQ * s.K * esz,
                             (uint64_t)s.Hi * s.Wi * NQ * s.K * esz};
    uint32_t box[5] = {(uint32_t)p.bk_elems, 1, (uint32_t)(t.Wt * s.sW), (uint32_t)(t.Ht * s.sH), (uint32_t)t.Bt};
    if (p.halo) { box[2] = 16; box[3] = (uint32_t)halo_h; box[4] = 1; }
    const uint32_t est[5] = {1, 1, (uint32_t)s.sW, (uint32_t)s.sH, 1};
    int rc = encode_map(&map_a, dtype, 5, in, dims, str, box, est, row_bytes);
    if (rc) return rc;
  }
  // B: packed weights [nq][taps][N][K] -> 4-D map {K, N, taps, nq}
  {
    const int taps = s.kH * s.kW;
    const uint64_t dims[4] = {(uint64_t)s.K, (uint64_t)s.N, (uint64_t)taps, (uint64_t)NQ};
    const uint64_t str[3] = {(uint64_t)s.K * esz, (uint64_t)s.N * s.K * esz, (uint64_t)taps * s.N * s.K * esz};
    const uint32_t box[4] = {(uint32_t)p.bk_elems, (uint32_t)(p.BN / cg), 1, 1};
    const uint32_t est[4] = {1, 1, 1, 1};
    int rc = encode_map(&map_b, dtype, 4, wpacked, dims, str, box, est, row_bytes);
    if (rc) return rc;
  }
  // Y: output [B][Ho][Wo][nq][N] -> 5-D map {N, nq, Wo, Ho, B}; box = 16 columns of one component x a warp's 32 pixel rows
  CUtensorMap map_y = map_a;
  if (p.tstore) {
    const int wb = t.Wt < 32 ? t.Wt : 32;
    const int hb = t.Ht < 32 / wb ? t.Ht : 32 / wb;
    const int bb = 32 / (wb * hb);
    const uint64_t dims[5] = {(uint64_t)s.N, (uint64_t)NQ, (uint64_t)s.Wo, (uint64_t)s.Ho, (uint64_t)s.B};
    const uint64_t str[4] = {(uint64_t)s.N * esz, (uint64_t)NQ * s.N * esz, (uint64_t)s.Wo * NQ * s.N * esz,
                             (uint64_t)s.Ho * s.Wo * NQ * s.N * esz};
    const uint32_t box[5] = {16, 1, (uint32_t)wb, (uint32_t)hb, (uint32_t)bb};
    const uint32_t est[5] = {1, 1, 1, 1, 1};
    int rc = encode_map(&map_y, dtype, 5, out, dims, str, box, est, 16 * esz);
    if (rc) return rc;
  }
  if (p.tt.zero_fill) QUAN_CUDA(cudaMemsetAsync(out, 0, (size_t)s.B * s.Ho * s.Wo * NQ * s.N * esz, st));
  int ctas = 0, rc = QUAN_OK;
  // the dense form's statistics run in registers (RSTAT instantiation) when a thread keeps its columns for the whole kernel
  static const int env_rstat = [] { const char* e = getenv("QUAN_TC_RSTAT"); return e ? atoi(e) : 1; }();
  const bool rstat = env_rstat && NQ == 1 && p.stat_part != nullptr && p.ntiles_n == 1 && p.BN <= 64;
  // the caller opted in (*stat_nparts < 0 on entry) and the layer is narrow: accumulate in L2 instead of writing slots — the IQBN apply
  // kernel then finishes the statistics itself (iqbn_apply_fwd_from_acc) and the fold launch disappears.  Measured on B200 (QUAN-YOLO11n
  // step, same box, alternating): 54 fewer launches, 11.215 vs 11.227 ms/step — the fold's time moves into the igemm tail (atomics) and
  // the apply prologue (fp64 finish per block); neutral, and the summation order is no longer fixed, so it is OFF unless
  // QUAN_TC_STAT_ACC=1.
  static const int env_acc = [] { const char* e = getenv("QUAN_TC_STAT_ACC"); return e ? atoi(e) : 0; }();
  const bool accumulate = env_acc && rstat && stat_nparts != nullptr && *stat_nparts < 0;
  p.stat_acc = accumulate ? stat_part + (size_t)QUAN_IQBN_MAX_PARTS * 8 * p.stat_cq : nullptr;
#define QUAN_IGEMM_CASE(CGV, KS)                                                                                           \
  do {                                                                                                                     \
    if constexpr (NQ == 1) {                                                                                               \
      if (rstat) { rc = launch_igemm_inst<T, MIX, CGV, KS, NQ, true>(map_a, map_b, map_y, out, p, smem, s.name, st, &ctas); break; } \
    }                                                                                                                      \
    rc = launch_igemm_inst<T, MIX, CGV, KS, NQ, false>(map_a, map_b, map_y, out, p, smem, s.name, st, &ctas);                       \
  } while (0)
  if (cg == 2) {
    if (ksteps == 4) { QUAN_IGEMM_CASE(2, 4); }
    else if (ksteps == 2) { QUAN_IGEMM_CASE(2, 2); }
    else { QUAN_IGEMM_CASE(2, 1); }
  } else {
    if (ksteps == 4) { QUAN_IGEMM_CASE(1, 4); }
    else if (ksteps == 2) { QUAN_IGEMM_CASE(1, 2); }
    else { QUAN_IGEMM_CASE(1, 1); }
  }
  if (stat_nparts != nullptr) *stat_nparts = (rc == QUAN_OK && p.stat_part != nullptr) ? (p.stat_acc != nullptr ? -ctas : ctas) : 0;
  return rc;
#undef QUAN_IGEMM_CASE
}

template <typename T, bool DGRAD>
static int pack_weights(const float* const w[4], void* out, const quan_conv_dims& d, cudaStream_t st) {
  const int taps = d.kH * d.kW;
  QUAN_REQUIRE(taps <= PACK_SMEM_FLOATS / 32 && d.Co <= 16383, QUAN_E_UNSUPPORTED, "tcgen05 conv: kernel %dx%d too large to pack", d.kH, d.kW);
  if (DGRAD) {
    int tci = PACK_SMEM_FLOATS / (32 * taps + 32);        // 32 co rows of tci*taps + 1 floats
    if (tci > 8) tci = 8;                                 // ~2.3 K elements per block: many blocks, few serial steps
    if (tci < 1) tci = 1;
    dim3 grid((unsigned)((d.Ci + tci - 1) / tci), (unsigned)((d.Co + 31) / 32), 4);
    QUAN_TIMED(st);
    QUAN_LAUNCH((pack_weights_dgrad_kernel<T>), grid, 256, (size_t)32 * (tci * taps + 1) * sizeof(float), st, w[0], w[1], w[2], w[3], reinterpret_cast<T*>(out), d.Co, d.Ci, taps, tci);
  } else {
    int cchunk = 2304 / taps;                             // ~9 KB tiles: many resident blocks
    if (cchunk < 1) cchunk = 1;
    if (cchunk > d.Ci) cchunk = d.Ci;
    dim3 grid((unsigned)((d.Ci + cchunk - 1) / cchunk), (unsigned)(4 * d.Co));
    QUAN_TIMED(st);
    QUAN_LAUNCH((pack_weights_fwd_kernel<T>), grid, 256, (size_t)cchunk * taps * sizeof(float), st, w[0], w[1], w[2], w[3], reinterpret_cast<T*>(out), d.Co, d.Ci, taps, cchunk);
  }
  QUAN_CHECK_LAUNCH("pack_weights_kernel");
  return QUAN_OK;
}
template <typename T, bool DGRAD>
static int pack_weights_dense(const float* const w[4], const float* bias_r, void* out, float* bias_out, const quan_conv_dims& d,
                              const Mix16& mix, cudaStream_t st) {
  const int taps = d.kH * d.kW;
  const int64_t total = (int64_t)16 * d.Co * d.Ci * taps;
  int grid = grid_for(total, 256, 4);
  QUAN_TIMED(st);
  QUAN_LAUNCH((pack_weights_dense_kernel<T, DGRAD>), grid, 256, 0, st, w[0], w[1], w[2], w[3], bias_r, reinterpret_cast<T*>(out), bias_out,
                                                            d.Co, d.Ci, taps, mix, depthwise_as_dense(d) ? 1 : 0);
  QUAN_CHECK_LAUNCH("pack_weights_dense_kernel");
  return QUAN_OK;
}

// ---- pack plan: every layer's packed weights in ONE launch per optimizer step ------------------------------------------------
// Weights change once per step but a narrow layer's chain used to start with its own 2.5 us pack kernel, in forward AND in dgrad
// (127 launches of the QUAN-YOLO11n step).  A train-step driver (graphs.GraphedTrainStep) records the packs of a warm-up step,
// commits them to one caller-owned arena and runs quan_pack_plan_run at the top of the step: while the plan is active every
// tc_fwd_t / tc_dgrad_t whose (weights, form, dtype, shape, mix) is in the plan reads its slot of the arena and launches nothing.
enum PackForm { PACK_SEP_FWD = 0, PACK_SEP_DGRAD = 1, PACK_DENSE_FWD = 2, PACK_DENSE_DGRAD = 3 };
struct PackJob {
  const float* w[4];
  const float* bias_r;
  void* out;
  float* bias_out;
  int Co, Ci, taps, form, esz, depthwise;
  int block0, nblocks;
  int64_t total;
  Mix16 mix;
};
struct PackPlan {
  std::mutex mu;
  bool recording = false, active = false;
  std::vector<PackJob> jobs;
  std::vector<size_t> offset, bytes;
  const PackJob* table_dev = nullptr;
  int total_blocks = 0;
};
static PackPlan& pack_plan() {
  static PackPlan p;
  return p;
}
static bool same_job(const PackJob& a, const float* const w[4], const float* bias_r, int form, int esz, const quan_conv_dims& d, const Mix16& M) {
  return a.w[0] == w[0] && a.w[1] == w[1] && a.w[2] == w[2] && a.w[3] == w[3] && a.bias_r == bias_r && a.form == form && a.esz == esz &&
         a.Co == d.Co && a.Ci == d.Ci && a.taps == d.kH * d.kW && a.depthwise == (depthwise_as_dense(d) ? 1 : 0) &&
         memcmp(a.mix.m, M.m, sizeof(M.m)) == 0;
}
// the arena slot of this pack when the plan is active (else nullptr; while recording, the pack is noted for the next commit)
static void* planned_pack(const float* const w[4], const float* bias_r, int form, int dtype, const quan_conv_dims& d, const Mix16& M) {
  PackPlan& pl = pack_plan();
  if (!pl.recording && !pl.active) return nullptr;
  std::lock_guard<std::mutex> lock(pl.mu);
  const int esz = dtype == QUAN_BF16 ? 2 : 4;
  for (size_t i = 0; i < pl.jobs.size(); ++i)
    if (same_job(pl.jobs[i], w, bias_r, form, esz, d, M)) return pl.active ? pl.jobs[i].out : nullptr;
  if (pl.recording) {
    PackJob j = {};
    for (int q = 0; q < 4; ++q) j.w[q] = w[q];
    j.bias_r = bias_r; j.Co = d.Co; j.Ci = d.Ci; j.taps = d.kH * d.kW; j.form = form; j.esz = esz; j.mix = M;
    j.depthwise = depthwise_as_dense(d) ? 1 : 0;
    const bool dense = form >= PACK_DENSE_FWD;
    j.total = (int64_t)(dense ? 16 : 4) * j.taps * d.Co * d.Ci;
    pl.jobs.push_back(j);
    pl.bytes.push_back(packed_weight_bytes_of(d, dtype, dense));
  }
  return nullptr;
}

// one element of job j (index i in the packed order of its form; same arithmetic as the per-layer kernels above)
template <typename T>
__device__ __forceinline__ void pack_one(const PackJob& j, int64_t i) {
  const int Co = j.Co, Ci = j.Ci, taps = j.taps;
  float v;
  if (j.form == PACK_SEP_FWD) {                    // [q][tap][co][ci]
    const int ci = (int)(i % Ci);
    int64_t r = i / Ci;
    const int co = (int)(r % Co); r /= Co;
    const int tap = (int)(r % taps), q = (int)(r / taps);
    v = round_operand(__ldg(j.w[q] + ((int64_t)co * Ci + ci) * taps + tap), sizeof(T));
  } else if (j.form == PACK_SEP_DGRAD) {           // [q][t][ci][co], flipped taps
    const int co = (int)(i % Co);
    int64_t r = i / Co;
    const int ci = (int)(r % Ci); r /= Ci;
    const int t = (int)(r % taps), q = (int)(r / taps);
    v = round_operand(__ldg(j.w[q] + ((int64_t)co * Ci + ci) * taps + (taps - 1 - t)), sizeof(T));
  } else {
    const bool dg = j.form == PACK_DENSE_DGRAD;
    const int N = dg ? 4 * Ci : 4 * Co, K = dg ? 4 * Co : 4 * Ci;
    const int k = (int)(i % K);
    int64_t r = i / K;
    const int n = (int)(r % N), t = (int)(r / N);
    int pc, q, co, ci, tap;
    if (dg) { q = n / Ci; ci = n % Ci; pc = k / Co; co = k % Co; tap = taps - 1 - t; }
    else { pc = n / Co; co = n % Co; q = k / Ci; ci = k % Ci; tap = t; }
    v = j.depthwise ? (ci == co ? j.mix.m[pc * 4 + q] * __ldg(j.w[q] + (int64_t)co * taps + tap) : 0.f)
                    : j.mix.m[pc * 4 + q] * __ldg(j.w[q] + ((int64_t)co * Ci + ci) * taps + tap);
    v = round_operand(v, sizeof(T));
    if (!dg && j.bias_out != nullptr && i < 4 * Co) j.bias_out[i] = j.bias_r ? j.mix.m[(i / Co) * 4] * __ldg(j.bias_r + i % Co) : 0.f;
  }
  reinterpret_cast<T*>(j.out)[i] = from_f32<T>(v);
}

__global__ void __launch_bounds__(256) pack_plan_kernel(const PackJob* __restrict__ jobs, int njobs) {
  pdl_prologue();
  int lo = 0, hi = njobs - 1;                      // the job whose block range holds blockIdx.x
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (jobs[mid].block0 <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
  }
  const PackJob& j = jobs[lo];
  const int64_t step = (int64_t)j.nblocks * 256;
  for (int64_t i = (int64_t)((int)blockIdx.x - j.block0) * 256 + threadIdx.x; i < j.total; i += step) {
    if (j.esz == 2) pack_one<__nv_bfloat16>(j, i);
    else pack_one<float>(j, i);
  }
}

// ---- shapes -------------------------------------------------------------------------------------------------------
// dense = 1: the conv over all 4*C_q real channels (nq = 1); dense = 0: one conv per component (nq = 4)
static IgemmShape fwd_shape(const quan_conv_dims& d, int dense) {
  IgemmShape s;
  const int f = dense ? 4 : 1;
  s.B = d.B; s.Hi = d.H; s.Wi = d.W; s.K = d.Ci * f; s.N = d.Co * f;
  s.Ho = conv_out(d.H, d.kH, d.sH, d.pH, d.dH);
  s.Wo = conv_out(d.W, d.kW, d.sW, d.pW, d.dW);
  s.kH = d.kH; s.kW = d.kW; s.sH = d.sH; s.sW = d.sW; s.pH = d.pH; s.pW = d.pW; s.dH = d.dH; s.dW = d.dW;
  s.nq = dense ? 1 : 4;
  s.scatH = s.scatW = 1;
  s.name = dense ? "qconv_igemm_fwd_dense" : "qconv_igemm_fwd";
  return s;
}
// stride-1 dgrad as a forward conv of G (Co channels, Ho x Wo) with flipped kernels and padding d*(k-1)-p;
// strided dgrad as stride*stride parity classes of stride-1 reads (build_taps)
static IgemmShape dgrad_shape(const quan_conv_dims& d, int dense) {
  IgemmShape s;
  const int f = dense ? 4 : 1;
  s.B = d.B; s.K = d.Co * f; s.N = d.Ci * f;
  s.Hi = conv_out(d.H, d.kH, d.sH, d.pH, d.dH);
  s.Wi = conv_out(d.W, d.kW, d.sW, d.pW, d.dW);
  s.Ho = d.H; s.Wo = d.W;
  s.kH = d.kH; s.kW = d.kW; s.sH = 1; s.sW = 1;
  s.dH = d.dH; s.dW = d.dW;
  s.nq = dense ? 1 : 4;
  s.scatH = d.sH; s.scatW = d.sW;
  if (d.sH == 1 && d.sW == 1) { s.pH = d.dH * (d.kH - 1) - d.pH; s.pW = d.dW * (d.kW - 1) - d.pW; }
  else { s.pH = d.pH; s.pW = d.pW; }
  s.name = dense ? "qconv_igemm_dgrad_dense" : "qconv_igemm_dgrad";
  return s;
}

static size_t packed_weight_bytes(const quan_conv_dims& d, int dtype, int dense);
static size_t packed_weight_bytes_of(const quan_conv_dims& d, int dtype, int dense) { return packed_weight_bytes(d, dtype, dense); }
static size_t packed_weight_bytes(const quan_conv_dims& d, int dtype, int dense) {
  const size_t w = ((size_t)(dense ? 16 : 4) * d.kH * d.kW * d.Co * d.Ci * (dtype == QUAN_BF16 ? 2 : 4) + 1023) / 1024 * 1024;
  return w + (dense ? (size_t)4 * d.Co * sizeof(float) + 1024 : 0);   // dense: + the mixed bias vector
}

// ---- wgrad plan ----------------------------------------------------------------------------------------------------
struct WgradPlan {
  TilePlan t;
  int Co, Ci;                  // channel counts of the GEMM (dense: 4x the quaternion counts)
  int row_bytes;               // swizzle atom row: 128 / 64 / 32 bytes of channels
  int nchunks, TG, NA, NB, MA, co_blocks, ci_blocks, tap_groups, splits, chunks_per_split, stages;
  int halo;                    // B operand as one halo box per ci atom (stride 1 along W)
  bool atomic;                 // the splits add into one partial set (dense form, > 4 splits)
  size_t b_atom_bytes;         // halo: slot pitch of a ci atom
  size_t smem, partial_bytes;
};

static bool plan_tiles_n(int B, int Ho, int Wo, int sH, int sW, int npix, TilePlan& t) {
  t.Wt = pow2_ceil(Wo) < npix ? pow2_ceil(Wo) : npix;
  while (t.Wt * sW > 256) t.Wt >>= 1;
  int rem = npix / t.Wt;
  t.Ht = pow2_ceil(Ho) < rem ? pow2_ceil(Ho) : rem;
  while (t.Ht * sH > 256) t.Ht >>= 1;
  t.Bt = npix / (t.Wt * t.Ht);
  if (t.Bt > B || t.Bt > 256) return false;
  t.tiles_w = (Wo + t.Wt - 1) / t.Wt;
  t.tiles_h = (Ho + t.Ht - 1) / t.Ht;
  t.tiles_b = (B + t.Bt - 1) / t.Bt;
  return true;
}

static bool plan_wgrad(const quan_conv_dims& d, int dtype, int dense, WgradPlan& w) {
  const int esz = dtype == QUAN_BF16 ? 2 : 4;
  const int nq = dense ? 1 : 4;
  w.Co = d.Co * (dense ? 4 : 1);
  w.Ci = d.Ci * (dense ? 4 : 1);
  // MN-major operands: rows of `row_bytes` channel bytes per pixel.  tf32 only has the 128-byte form
  // (SWIZZLE_128B_BASE32B); bf16 also takes 64- and 32-byte rows, which is what narrow layers need.
  w.row_bytes = 0;
  for (int rb = 128; rb >= (esz == 2 ? 32 : 128); rb >>= 1)
    if ((w.Ci * esz) % rb == 0 && (w.Co * esz) % rb == 0) { w.row_bytes = rb; break; }
  if (w.row_bytes == 0) return false;
  w.NA = w.row_bytes / esz;
  if (d.sH > 8 || d.sW > 8) return false;
  const int Ho = conv_out(d.H, d.kH, d.sH, d.pH, d.dH), Wo = conv_out(d.W, d.kW, d.sW, d.pW, d.dW);
  if (!plan_tiles_n(d.B, Ho, Wo, d.sH, d.sW, WG_PIX, w.t)) return false;
  const int64_t nchunks = (int64_t)w.t.tiles_w * w.t.tiles_h * w.t.tiles_b;
  if (nchunks > 0x3fffffff) return false;
  w.nchunks = (int)nchunks;
  w.TG = d.kW;
  if (w.TG * 16 > 512) return false;
  w.tap_groups = d.kH;
  w.MA = 128 / w.NA;                                 // atom slots that make M = 128
  // N = NB atoms of ci per CTA: as wide as TMEM (TG accumulators of N columns), UMMA (N <= 256, multiple of 16) and a
  // >= 2-stage smem ring allow — wider N means fewer, longer MMAs per barrier round trip and fewer re-reads of the G tile
  const size_t atom = (size_t)WG_PIX * w.row_bytes;
  // halo mode: the kW taps of a filter row read the same pixels shifted by dW — one box per ci atom, (kW-1)*dW pixels
  // wider than the chunk, serves them all (B operand traffic / kW); needs stride 1 along W and K-steps that stay inside
  // an image row.  QUAN_TC_WG_HALO=0 disables.
  {
    static const int env_halo = [] { const char* e = getenv("QUAN_TC_WG_HALO"); return e ? atoi(e) : 1; }();
    const int umma_k = 32 / esz, wbox = w.t.Wt + (d.kW - 1) * d.dW;
    w.halo = env_halo && d.sW == 1 && d.kW > 1 && w.t.Wt % umma_k == 0 && wbox <= 256;
    w.b_atom_bytes = w.halo ? ((size_t)w.t.Bt * w.t.Ht * wbox * w.row_bytes + 1023) / 1024 * 1024 : atom;
  }
  auto b_stage = [&](int nb) { return w.halo ? (size_t)nb * w.b_atom_bytes : (size_t)w.TG * nb * atom; };
  w.NB = 0;
  for (int nb = 1; nb * w.NA <= 256; ++nb) {
    const int n = nb * w.NA;
    const size_t stage_nb = (size_t)w.MA * atom + b_stage(nb);
    if (w.Ci % n != 0 || n % 16 != 0 || w.TG * n > 512) continue;
    static const int min_stages = [] { const char* e = getenv("QUAN_TC_WG_MINSTAGES"); return e ? atoi(e) : 2; }();
    // measured (tf32, C_q = 256): N = 128 with 2 stages 489 us vs N = 64 with 4 stages 631 us — wider N wins over ring depth
    if (w.NB != 0 && (int)((200 * 1024) / stage_nb) < min_stages) continue;
    w.NB = nb;
  }
  if (w.NB == 0) return false;
  if (const char* e = getenv("QUAN_TC_WG_NB")) { int v = atoi(e); if (v >= 1 && v <= w.NB && w.Ci % (v * w.NA) == 0 && (v * w.NA) % 16 == 0) w.NB = v; }
  w.co_blocks = (w.Co + 127) / 128;
  w.ci_blocks = w.Ci / (w.NA * w.NB);
  const int64_t combos = (int64_t)nq * w.tap_groups * w.co_blocks * w.ci_blocks;
  if (combos > 65535) return false;
  // split-K factor: 1 CTA/SM is resident, so pick the split count whose grid fills whole waves of 148 best
  // (first ncu capture: 336 CTAs = 2.27 waves, i.e. a third wave at 27% occupancy)
  int64_t splits = 1;
  double best = 1e30;
  const int taps = d.kH * d.kW;
  const double flops = 2.0 * nq * (double)d.B * Ho * Wo * (w.co_blocks * 128.0) * w.Ci * taps;   // what the tensor pipe executes
  const double in_bytes = (double)d.B * Ho * Wo * 4.0 * (d.Co + d.Ci * d.sH * d.sW) * esz;       // one pass over dY and x
  const double part_bytes = (double)nq * taps * w.Co * w.Ci * 4.0;                               // one split's fp32 partials
  // dense form: the splits add into ONE partial set with atomics (launch_wgrad), so nothing grows with the split count and a 1x1 layer
  // (one tap group: combos = 1) could use every SM instead of 64 of them.  Measured on the QUAN-YOLO11n step (three alternating runs):
  // 64 splits 10.61-10.73 ms, 148 10.62-10.72, 296 with two CTAs per SM 10.78 — the wgrads run beside the main chain
  // and more CTAs only take more of its SMs; 64 stays.  QUAN_TC_WG_MAXSPLIT / QUAN_TC_WG_RESID (CTAs per SM the wave model assumes).
  static const int max_split_dense = [] { const char* e = getenv("QUAN_TC_WG_MAXSPLIT"); return e ? atoi(e) : 64; }();
  static const int resid = [] { const char* e = getenv("QUAN_TC_WG_RESID"); return e ? atoi(e) : 1; }();
  const int64_t max_sp = dense ? max_split_dense : 64;
  const int64_t slots = (int64_t)QUAN_NUM_SMS * (dense ? resid : 1);
  for (int64_t sp = 1; sp <= max_sp && sp <= w.nchunks; ++sp) {
    const int64_t ctas = combos * sp;
    const int64_t waves = (ctas + slots - 1) / slots;
    const double eff = (double)ctas / (double)(waves * slots);
    // modelled time: tensor work at ~1 PFLOP/s or the operand stream at ~4 TB/s, scaled by wave fill, + partial
    // write/read at ~4 TB/s + per-CTA prologue
    const double work = flops / 1.0e15 > in_bytes / 4.0e12 ? flops / 1.0e15 : in_bytes / 4.0e12;
    const double t = work / eff + sp * part_bytes * 2.0 / 4.0e12 + waves * 4.0e-6;
    if (t < best) { best = t; splits = sp; }
  }
  w.chunks_per_split = (int)((w.nchunks + splits - 1) / splits);
  w.splits = (w.nchunks + w.chunks_per_split - 1) / w.chunks_per_split;
  const size_t stage = (size_t)w.MA * atom + b_stage(w.NB);
  int stages = (int)((200 * 1024) / stage);
  if (stages > 8) stages = 8;
  if (dense) {                                       // see QUAN_TC_DENSE_SMEM_KB in launch_igemm
    static const int cap_kb = [] { const char* e = getenv("QUAN_TC_WG_SMEM_KB"); return e ? atoi(e) : 100; }();
    if (cap_kb > 0 && (size_t)stages * stage > (size_t)cap_kb * 1024) stages = (int)((size_t)cap_kb * 1024 / stage);
    if (stages < 2) stages = 2;                      // never below a two-stage ring
  }
  if (stages < 2) return false;
  w.stages = stages;
  w.smem = 1024 + stages * stage + (2 * stages + 1) * sizeof(uint64_t) + 16;
  w.atomic = dense && w.splits > 4;                  // many splits of a tiny dW: one partial set, vector atomics (launch_wgrad)
  w.partial_bytes = (size_t)(w.atomic ? 1 : w.splits) * nq * taps * w.Co * w.Ci * sizeof(float);
  return true;
}

// gq: G = M^T dY (separable form) or dY itself (dense form, the mix is applied by the reduce kernel)
template <typename T>
static int launch_wgrad(const void* gq, const void* x, float* const dw[4], const quan_conv_dims& d, int dtype, int dense,
                        const Mix16& mix, void* ws, size_t ws_bytes, cudaStream_t st) {
  WgradPlan w;
  QUAN_REQUIRE(plan_wgrad(d, dtype, dense, w), QUAN_E_UNSUPPORTED, "tcgen05 wgrad: shape does not qualify");
  QUAN_REQUIRE(ws_bytes >= w.partial_bytes, QUAN_E_WORKSPACE, "tcgen05 wgrad: workspace needs %zu bytes, got %zu",
               w.partial_bytes, ws_bytes);
  const int esz = sizeof(T);
  const int nq = dense ? 1 : 4;
  const int Ho = conv_out(d.H, d.kH, d.sH, d.pH, d.dH), Wo = conv_out(d.W, d.kW, d.sW, d.pW, d.dW);
  TcWgradParams p = {};
  p.Wt = w.t.Wt; p.Ht = w.t.Ht; p.Bt = w.t.Bt; p.tiles_w = w.t.tiles_w; p.tiles_h = w.t.tiles_h; p.nchunks = w.nchunks;
  p.kH = d.kH; p.kW = d.kW; p.sH = d.sH; p.sW = d.sW; p.pH = d.pH; p.pW = d.pW; p.dH = d.dH; p.dW = d.dW;
  p.Co = w.Co; p.Ci = w.Ci; p.taps = d.kH * d.kW; p.nq = nq;
  p.TG = w.TG; p.NA = w.NA; p.NB = w.NB; p.MA = w.MA;
  p.co_blocks = w.co_blocks; p.ci_blocks = w.ci_blocks; p.tap_groups = w.tap_groups;
  p.chunks_per_split = w.chunks_per_split;
  const int umma_k = 32 / esz;                       // pixels per UMMA
  p.ksteps = WG_PIX / umma_k;
  p.kadv = (uint32_t)(umma_k * w.row_bytes) >> 4;    // umma_k pixel rows
  p.stages = w.stages;
  p.atom_bytes = (uint32_t)(WG_PIX * w.row_bytes);
  // bf16: 8-row groups, SWIZZLE_128B / 64B / 32B; tf32: 4-row groups, SWIZZLE_128B_BASE32B
  p.sbo_bytes = sizeof(T) == 2 ? 8u * w.row_bytes : 512u;
  p.layout_type = sizeof(T) == 2 ? (w.row_bytes == 128 ? 2u : w.row_bytes == 64 ? 4u : 6u) : 1u;
  const uint32_t swz = sizeof(T) == 2 ? (uint32_t)w.row_bytes : SWZ_128B_ATOM32;
  p.a_stage_bytes = (uint32_t)w.MA * p.atom_bytes;
  p.halo = w.halo;
  p.b_atom_bytes = (uint32_t)w.b_atom_bytes;
  p.b_stage_bytes = w.halo ? (uint32_t)(w.NB * w.b_atom_bytes) : (uint32_t)(w.TG * w.NB) * p.atom_bytes;
  if (w.halo)
    for (int k = 0; k < p.ksteps && k < 8; ++k) {
      const int p0 = k * umma_k;
      p.brow[k] = (uint16_t)((p0 / w.t.Wt) * (w.t.Wt + (d.kW - 1) * d.dW) + p0 % w.t.Wt);
    }
  p.idesc = ptx::make_idesc(sizeof(T) == 2 ? 1u : 2u, 1u, 1u, 128u, (uint32_t)(w.NA * w.NB));
  p.tmem_cols = (uint32_t)pow2_ceil(w.TG * w.NA * w.NB < 32 ? 32 : w.TG * w.NA * w.NB);
  p.partial = reinterpret_cast<float*>(ws);
  // dense form with many splits: accumulate with atomics (summation order is then not deterministic, as in the direct engine)
  p.atomic = w.atomic ? 1 : 0;
  if (p.atomic) QUAN_CUDA(cudaMemsetAsync(ws, 0, w.partial_bytes, st));
  const int fold_splits = p.atomic ? 1 : w.splits;

  CUtensorMap map_g, map_x;
  {
    const uint64_t dims[5] = {(uint64_t)w.Co, (uint64_t)nq, (uint64_t)Wo, (uint64_t)Ho, (uint64_t)d.B};
    const uint64_t str[4] = {(uint64_t)w.Co * esz, (uint64_t)nq * w.Co * esz, (uint64_t)Wo * nq * w.Co * esz,
                             (uint64_t)Ho * Wo * nq * w.Co * esz};
    const uint32_t box[5] = {(uint32_t)w.NA, 1, (uint32_t)w.t.Wt, (uint32_t)w.t.Ht, (uint32_t)w.t.Bt};
    const uint32_t est[5] = {1, 1, 1, 1, 1};
    int rc = encode_map(&map_g, dtype, 5, gq, dims, str, box, est, swz);
    if (rc) return rc;
  }
  {
    const uint64_t dims[5] = {(uint64_t)w.Ci, (uint64_t)nq, (uint64_t)d.W, (uint64_t)d.H, (uint64_t)d.B};
    const uint64_t str[4] = {(uint64_t)w.Ci * esz, (uint64_t)nq * w.Ci * esz, (uint64_t)d.W * nq * w.Ci * esz,
                             (uint64_t)d.H * d.W * nq * w.Ci * esz};
    const uint32_t box[5] = {(uint32_t)w.NA, 1, (uint32_t)(w.halo ? w.t.Wt + (d.kW - 1) * d.dW : w.t.Wt * d.sW),
                             (uint32_t)(w.t.Ht * d.sH), (uint32_t)w.t.Bt};
    const uint32_t est[5] = {1, 1, (uint32_t)d.sW, (uint32_t)d.sH, 1};
    int rc = encode_map(&map_x, dtype, 5, x, dims, str, box, est, swz);
    if (rc) return rc;
  }
  auto kern = qconv_wgrad_kernel<T, WG_PIX / (32 / (int)sizeof(T))>;
  static thread_local DeviceOnce attr_set;
  if (attr_set.first()) QUAN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  dim3 grid((unsigned)w.splits, (unsigned)(nq * w.tap_groups * w.co_blocks * w.ci_blocks));
  QUAN_TIMED(st);
  QUAN_LAUNCH((kern), grid, TC_THREADS, w.smem, st, map_g, map_x, p);
  QUAN_CHECK_LAUNCH(dense ? "qconv_wgrad_kernel_dense" : "qconv_wgrad_kernel");
  {
    int cchunk = 2304 / p.taps;
    if (cchunk < 1) cchunk = 1;
    if (cchunk > d.Ci) cchunk = d.Ci;
    // split lanes per element: short rows with many splits (narrow layers) spread the split loop over the idle threads
    int SL = 256 / (cchunk * p.taps);
    if (SL > 8) SL = 8;
    if (SL > fold_splits) SL = fold_splits;
    if (SL < 1) SL = 1;
    dim3 rgrid((unsigned)((d.Ci + cchunk - 1) / cchunk), (unsigned)(4 * d.Co));
    const size_t rsmem = (size_t)(SL > 1 ? 256 : cchunk * p.taps) * sizeof(float) + 16;
    QUAN_TIMED(st);
    static const int env_flat = [] { const char* e = getenv("QUAN_TC_WG_FLATFOLD"); return e ? atoi(e) : 1; }();
    const int64_t dw_elems = (int64_t)4 * d.Co * d.Ci * p.taps;
    if (dense && depthwise_as_dense(d)) {
      QUAN_LAUNCH((wgrad_reduce_dw_kernel), (unsigned)((4 * d.Co * p.taps + 255) / 256), 256, 0, st, p.partial, dw[0], dw[1], dw[2], dw[3], fold_splits,
                  p.taps, d.Co, mix);
    } else if (env_flat && fold_splits <= 4 && dw_elems <= (1 << 18)) {
      const unsigned fgrid = (unsigned)((dw_elems + 255) / 256);
      if (dense)
        QUAN_LAUNCH((wgrad_reduce_flat_kernel<true>), fgrid, 256, 0, st, p.partial, dw[0], dw[1], dw[2], dw[3], fold_splits, p.taps, d.Co, d.Ci, mix);
      else
        QUAN_LAUNCH((wgrad_reduce_flat_kernel<false>), fgrid, 256, 0, st, p.partial, dw[0], dw[1], dw[2], dw[3], fold_splits, p.taps, d.Co, d.Ci, mix);
    } else if (dense)
      QUAN_LAUNCH((wgrad_reduce_kernel<true>), rgrid, 256, rsmem, st, p.partial, dw[0], dw[1], dw[2], dw[3], fold_splits, p.taps, d.Co, d.Ci, cchunk, SL, mix);
    else
      QUAN_LAUNCH((wgrad_reduce_kernel<false>), rgrid, 256, rsmem, st, p.partial, dw[0], dw[1], dw[2], dw[3], fold_splits, p.taps, d.Co, d.Ci, cchunk, SL, mix);
  }
  QUAN_CHECK_LAUNCH("wgrad_reduce_kernel");
  return QUAN_OK;
}

// ---- form selection --------------------------------------------------------------------------------------------------
// Narrow layers (QUAN-YOLO11n: 4..32 quaternion channels) are HBM-bound and their per-component GEMMs are too thin for
// the tensor core (K rows of C_i*2 bytes < one swizzle row, N = C_o).  For them the Hamilton-structured weight is
// treated as the dense contraction it is: one GEMM over all 4*C_q channels, M folded into the packed weights — 4x the
// separable FLOPs on a pipe with far more than 4x headroom, no mix epilogue and no G = M^T dY pre-pass.
// Depthwise QConv2D (DWConv, conv.py:918-923: groups = C_i = C_o) on the tensor cores: the QUAN heads use it at 16-64 quaternion
// channels, where the layer moves 2-4 bytes per MAC — far below the ridge whatever the engine — and the CUDA-core depthwise kernels
// are bound by their own load / fold instructions (0.3 of the HBM floor forward, 0.15 backward).  The dense Hamilton form with a
// block-diagonal weight (W[co][ci] = 0 for ci != co) spends C times the MACs on a pipe with more than C times the headroom and
// re-uses every piece of the narrow-layer path: halo tiles, epilogue statistics, split-K wgrad (whose fold keeps the diagonal).
// QUAN_TC_DEPTHWISE=0 keeps depthwise layers on the CUDA-core engine.
static bool depthwise_as_dense(const quan_conv_dims& d) {
  static const int on = [] { const char* e = getenv("QUAN_TC_DEPTHWISE"); return e ? atoi(e) : 1; }();
  return on && d.groups > 1 && d.groups == d.Ci && d.Ci == d.Co && 4 * d.Ci <= 256;
}

static int env_dense() {   // QUAN_TC_DENSE = 0: never, 1: whenever the shape allows (tests), unset: by channel count
  static int v = -2;
  if (v == -2) { const char* e = getenv("QUAN_TC_DENSE"); v = e ? atoi(e) : -1; }
  return v;
}
static bool prefer_dense(int k_channels, int dtype) {   // k_channels: per-component contraction width
  return k_channels * (dtype == QUAN_BF16 ? 2 : 4) <= 64;
}

// shape -> answer memo (per host thread): the API layer asks for the mode / workspace of a shape several times per call
// and planning walks the tap table and the split-K cost model each time
struct ShapeKey {
  int v[18];
  bool operator==(const ShapeKey& o) const { return memcmp(v, o.v, sizeof(v)) == 0; }
};
struct ShapeKeyHash {
  size_t operator()(const ShapeKey& k) const {
    uint64_t h = 1469598103934665603ull;
    for (int i = 0; i < 18; ++i) { h ^= (uint32_t)k.v[i]; h *= 1099511628211ull; }
    return (size_t)h;
  }
};
static ShapeKey shape_key(const quan_conv_dims& d, int dtype, int layout, int pass, int what) {
  ShapeKey k = {{d.B, d.Ci, d.Co, d.H, d.W, d.kH, d.kW, d.sH, d.sW, d.pH, d.pW, d.dH, d.dW, d.groups, dtype, layout, pass, what}};
  return k;
}

static int qconv_tc_mode_uncached(const quan_conv_dims& d, int dtype, int layout, int pass) {
  // depthwise layers of up to 64 quaternion channels run as the DENSE form with a block-diagonal packed weight (depthwise_as_dense)
  const bool dwd = depthwise_as_dense(d);
  if (layout != QUAN_LAYOUT_BHWQC || (d.groups != 1 && !dwd)) return TC_NONE;
  if (get_encode_fn() == nullptr) return TC_NONE;
  const int e = dwd ? 1 : env_dense();
  bool sep = false, dense = false;
  int kch = d.Ci;
  if (pass == PASS_FWD) {
    sep = igemm_supported(fwd_shape(d, 0), dtype);
    dense = igemm_supported(fwd_shape(d, 1), dtype);
  } else if (pass == PASS_DGRAD) {
    sep = igemm_supported(dgrad_shape(d, 0), dtype);
    dense = igemm_supported(dgrad_shape(d, 1), dtype);
    kch = d.Co;
  } else {
    WgradPlan w;
    sep = plan_wgrad(d, dtype, 0, w);
    dense = plan_wgrad(d, dtype, 1, w);
    kch = d.Ci < d.Co ? d.Ci : d.Co;
  }
  if (e == 0) dense = false;
  if (dwd) return dense ? TC_DENSE : TC_NONE;
  if (dense && (e == 1 || !sep || prefer_dense(kch, dtype))) return TC_DENSE;
  return sep ? TC_SEPARABLE : TC_NONE;
}

int qconv_tc_mode(const quan_conv_dims& d, int dtype, int layout, int pass) {
  static thread_local std::unordered_map<ShapeKey, int, ShapeKeyHash> memo;
  const ShapeKey k = shape_key(d, dtype, layout, pass, 0);
  auto it = memo.find(k);
  if (it != memo.end()) return it->second;
  const int m = qconv_tc_mode_uncached(d, dtype, layout, pass);
  if (memo.size() > 4096) memo.clear();
  memo[k] = m;
  return m;
}

bool qconv_tc_supported(const quan_conv_dims& d, int dtype, int layout, int pass) {
  return qconv_tc_mode(d, dtype, layout, pass) != TC_NONE;
}

size_t qconv_tc_workspace_bytes(const quan_conv_dims& d, int dtype, int layout, int pass) {
  static thread_local std::unordered_map<ShapeKey, size_t, ShapeKeyHash> memo;
  const ShapeKey k = shape_key(d, dtype, layout, pass, 1);
  auto it = memo.find(k);
  if (it != memo.end()) return it->second;
  const int mode = qconv_tc_mode(d, dtype, layout, pass);
  size_t bytes = 0;
  if (mode != TC_NONE) {
    if (pass == PASS_FWD || pass == PASS_DGRAD) {
      bytes = packed_weight_bytes(d, dtype, mode == TC_DENSE);
    } else {
      WgradPlan w;
      bytes = plan_wgrad(d, dtype, mode == TC_DENSE, w) ? w.partial_bytes : 0;
    }
  }
  if (memo.size() > 4096) memo.clear();
  memo[k] = bytes;
  return bytes;
}

template <typename T>
static int tc_fwd_t(const void* x, const float* const w[4], const float* bias_r, void* y, const quan_conv_dims& d, int dtype,
                    int dense, const Mix16& M, void* ws, cudaStream_t st, double* stat_part, int* stat_nparts, const PostOp* post) {
  if (dense) {
    void* pre = planned_pack(w, bias_r, PACK_DENSE_FWD, dtype, d, M);
    if (pre != nullptr) ws = pre;
    float* bias_out = reinterpret_cast<float*>((char*)ws + packed_weight_bytes(d, dtype, 1) - (size_t)4 * d.Co * sizeof(float) - 1024);
    if (pre == nullptr) {
      int rc = pack_weights_dense<T, false>(w, bias_r, ws, bias_out, d, M, st);
      if (rc) return rc;
    }
    return launch_igemm<T, false, 1>(x, ws, bias_r ? bias_out : nullptr, y, fwd_shape(d, 1), dtype, M, st, stat_part, stat_nparts, post);
  }
  void* pre = planned_pack(w, nullptr, PACK_SEP_FWD, dtype, d, M);
  if (pre != nullptr) ws = pre;
  else {
    int rc = pack_weights<T, false>(w, ws, d, st);
    if (rc) return rc;
  }
  return launch_igemm<T, true, 4>(x, ws, bias_r, y, fwd_shape(d, 0), dtype, M, st, stat_part, stat_nparts, post);
}

int qconv_tc_fwd(const void* x, const float* const w[4], const float* bias_r, void* y, const quan_conv_dims& d, int dtype,
                 int mode, const float* mix, void* ws, size_t ws_bytes, cudaStream_t st, double* stat_part, int* stat_nparts,
                 const float* post_scale, const float* post_shift, int post_act) {
  const int dense = mode == TC_DENSE;
  QUAN_REQUIRE(ws_bytes >= packed_weight_bytes(d, dtype, dense), QUAN_E_WORKSPACE, "tcgen05 fwd: workspace too small");
  const Mix16 M = make_mix(mix);
  PostOp post;
  post.scale = post_scale; post.shift = post_shift; post.act = post_act;
  const PostOp* pp = post_scale != nullptr ? &post : nullptr;
  if (dtype == QUAN_BF16) return tc_fwd_t<__nv_bfloat16>(x, w, bias_r, y, d, dtype, dense, M, ws, st, stat_part, stat_nparts, pp);
  return tc_fwd_t<float>(x, w, bias_r, y, d, dtype, dense, M, ws, st, stat_part, stat_nparts, pp);
}

template <typename T>
static int tc_dgrad_t(const void* g, const float* const w[4], void* dx, const quan_conv_dims& d, int dtype, int dense,
                      const Mix16& M, void* ws, cudaStream_t st) {
  if (dense) {
    void* pre = planned_pack(w, nullptr, PACK_DENSE_DGRAD, dtype, d, M);
    if (pre != nullptr) ws = pre;
    else {
      int rc = pack_weights_dense<T, true>(w, nullptr, ws, nullptr, d, M, st);
      if (rc) return rc;
    }
    return launch_igemm<T, false, 1>(g, ws, nullptr, dx, dgrad_shape(d, 1), dtype, M, st);
  }
  void* pre = planned_pack(w, nullptr, PACK_SEP_DGRAD, dtype, d, M);
  if (pre != nullptr) ws = pre;
  else {
    int rc = pack_weights<T, true>(w, ws, d, st);
    if (rc) return rc;
  }
  return launch_igemm<T, false, 4>(g, ws, nullptr, dx, dgrad_shape(d, 0), dtype, M, st);
}

// g: G = M^T dY for the separable form, dY itself for the dense form (mix: the forward mixing matrix)
int qconv_tc_dgrad(const void* g, const float* const w[4], void* dx, const quan_conv_dims& d, int dtype, int mode,
                   const float* mix, void* ws, size_t ws_bytes, cudaStream_t st) {
  const int dense = mode == TC_DENSE;
  QUAN_REQUIRE(ws_bytes >= packed_weight_bytes(d, dtype, dense), QUAN_E_WORKSPACE, "tcgen05 dgrad: workspace too small");
  const Mix16 M = make_mix(mix);
  if (dtype == QUAN_BF16) return tc_dgrad_t<__nv_bfloat16>(g, w, dx, d, dtype, dense, M, ws, st);
  return tc_dgrad_t<float>(g, w, dx, d, dtype, dense, M, ws, st);
}

int qconv_tc_wgrad(const void* g, const void* x, float* const dw[4], const quan_conv_dims& d, int dtype, int mode,
                   const float* mix, void* ws, size_t ws_bytes, cudaStream_t st) {
  const int dense = mode == TC_DENSE;
  const Mix16 M = make_mix(mix);
  if (dtype == QUAN_BF16) return launch_wgrad<__nv_bfloat16>(g, x, dw, d, dtype, dense, M, ws, ws_bytes, st);
  return launch_wgrad<float>(g, x, dw, d, dtype, dense, M, ws, ws_bytes, st);
}

}  // namespace quan

using namespace quan;

extern "C" {

int quan_pack_plan_record(int on) {
  PackPlan& pl = pack_plan();
  std::lock_guard<std::mutex> lock(pl.mu);
  if (on) {
    pl.jobs.clear(); pl.bytes.clear(); pl.offset.clear();
    pl.active = false; pl.table_dev = nullptr; pl.total_blocks = 0;
  }
  pl.recording = on != 0;
  return (int)pl.jobs.size();
}

size_t quan_pack_plan_bytes(size_t* table_bytes) {
  PackPlan& pl = pack_plan();
  std::lock_guard<std::mutex> lock(pl.mu);
  size_t total = 0;
  for (size_t b : pl.bytes) total += (b + 1023) / 1024 * 1024;
  if (table_bytes != nullptr) *table_bytes = pl.jobs.size() * sizeof(PackJob);
  return total;
}

int quan_pack_plan_commit(void* arena, size_t arena_bytes, void* table, size_t table_bytes, void* stream) {
  PackPlan& pl = pack_plan();
  std::lock_guard<std::mutex> lock(pl.mu);
  QUAN_REQUIRE(!pl.jobs.empty(), QUAN_E_ARG, "pack_plan_commit: nothing recorded (run a step between quan_pack_plan_record(1) and (0))");
  QUAN_REQUIRE(arena != nullptr && table != nullptr && table_bytes >= pl.jobs.size() * sizeof(PackJob), QUAN_E_ARG, "pack_plan_commit: null or small table");
  QUAN_REQUIRE((reinterpret_cast<uintptr_t>(arena) & 1023) == 0, QUAN_E_ARG, "pack_plan_commit: arena must be 1024-byte aligned");
  size_t off = 0;
  int block0 = 0;
  pl.offset.assign(pl.jobs.size(), 0);
  for (size_t i = 0; i < pl.jobs.size(); ++i) {
    PackJob& j = pl.jobs[i];
    pl.offset[i] = off;
    j.out = (char*)arena + off;
    j.bias_out = j.form == PACK_DENSE_FWD ? reinterpret_cast<float*>((char*)j.out + pl.bytes[i] - (size_t)4 * j.Co * sizeof(float) - 1024) : nullptr;
    off += (pl.bytes[i] + 1023) / 1024 * 1024;
    int nb = (int)((j.total + 256 * 8 - 1) / (256 * 8));       // ~8 elements per thread
    j.block0 = block0; j.nblocks = nb < 1 ? 1 : nb;
    block0 += j.nblocks;
  }
  QUAN_REQUIRE(off <= arena_bytes, QUAN_E_WORKSPACE, "pack_plan_commit: arena needs %zu bytes, got %zu", off, arena_bytes);
  pl.total_blocks = block0;
  QUAN_CUDA(cudaMemcpyAsync(table, pl.jobs.data(), pl.jobs.size() * sizeof(PackJob), cudaMemcpyHostToDevice, (cudaStream_t)stream));
  QUAN_CUDA(cudaStreamSynchronize((cudaStream_t)stream));      // the host vector may be re-used by the next record
  pl.table_dev = reinterpret_cast<const PackJob*>(table);
  pl.recording = false;
  pl.active = true;
  return QUAN_OK;
}

int quan_pack_plan_run(void* stream) {
  PackPlan& pl = pack_plan();
  int njobs, blocks;
  const PackJob* table;
  {
    std::lock_guard<std::mutex> lock(pl.mu);
    QUAN_REQUIRE(pl.active && pl.table_dev != nullptr, QUAN_E_ARG, "pack_plan_run: no committed plan");
    njobs = (int)pl.jobs.size(); blocks = pl.total_blocks; table = pl.table_dev;
  }
  cudaStream_t st = (cudaStream_t)stream;
  QUAN_TIMED(st);
  QUAN_LAUNCH(pack_plan_kernel, (unsigned)blocks, 256, 0, st, table, njobs);
  QUAN_CHECK_LAUNCH("pack_plan_kernel");
  return QUAN_OK;
}

int quan_pack_plan_release(void) {
  PackPlan& pl = pack_plan();
  std::lock_guard<std::mutex> lock(pl.mu);
  pl.active = false; pl.recording = false; pl.table_dev = nullptr;
  return QUAN_OK;
}

}  // extern "C"
