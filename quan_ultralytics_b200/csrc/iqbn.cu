// iqbn.cu — IQBN (independent quaternion batch-norm) for sm_100a: bandwidth-bound kernels.
//
// Reference semantics (not code): ultralytics/nn/modules/conv.py:553-571 (training branch, biased variance
// + 1e-8, running-stat update with momentum), :546-552 (eval branch), SiLU from conv.py:789,809.
// Backward is the analytic gradient of that expression (the reference relies on autograd).
//
// Design: every kernel is a grid-stride stream of 16-byte vector loads.  A thread owns a fixed
// "column" (c,q) set for its whole life: per-channel coefficients come from a precomputed table in column order (a few
// 16-byte loads) and per-channel partial sums are private fp32 registers (locally shifted, so cancellation stays
// local), un-shifted in fp64, folded across the block through shared memory and written to the block's own slot of a
// partials buffer; a second small kernel folds the slots and finishes (mean/var/rstd + running stats + coefficient
// tables).  No atomics, deterministic.
//
// Two physical layouts (include/quan_sm100.h): BCHWQ (reference) and BHWQC (channels_last_3d, tensor-core path).
#include "common.cuh"
#include "tc_ptx.cuh"
#include <cooperative_groups.h>
#include <stdlib.h>

namespace quan {

// workspace: per-row-split partial sums part[split][which][c*4+q] in fp64, written with plain stores by the reduction
// kernels and folded by a second, parallel finalize kernel.  Deterministic, nothing to zero.  (The first versions used
// fp64 atomics + a last-block tail: the L2 retires only ~18 G fp64 atomics/s chip-wide, so 1.2 M atomics — 592 blocks x
// 2048 accumulators — cost 69 us on top of a 24 us stream; measured in profiles/r01_iqbn_tune2.log.)
constexpr int IQBN_MAX_PARTS = QUAN_IQBN_MAX_PARTS;
struct IqbnWs {
  double* part;          // [IQBN_MAX_PARTS][2][4C] per-block partial sums (plain stores, folded by iqbn_fold_kernel)
  double* acc;           // [2][4C] accumulators of the single-launch reduction for small tensors: ZERO between launches
  unsigned* counter;     // its block ticket: ZERO between launches
};
static inline IqbnWs carve_ws(void* ws, int C) {
  IqbnWs w;
  w.part = reinterpret_cast<double*>(ws);
  w.acc = w.part + (size_t)IQBN_MAX_PARTS * 8 * C;
  w.counter = reinterpret_cast<unsigned*>(w.acc + 8 * (size_t)C);
  return w;
}

// What the last block does with the accumulated sums.
enum TailMode { TAIL_RAW_SUMS = 0, TAIL_FWD_STATS = 1, TAIL_BWD_SUMS = 2 };

// stats buffer [20C] floats: mean | var(+1e-8) | rstd (index c*4+q)  |  scaleT | shiftT (index q*C+c: BHWQC column order,
//                            so a thread's V coefficients are one or two 16-byte loads instead of 4V scalar ones)
// bwd buffer: [8C] doubles {sum dz, sum dz*xhat} (index c*4+q) followed by [12C] floats k1T | k2T | k3T (index q*C+c)
struct TailArgs {
  int mode;
  int C;
  double count;        // bwd: <= 0 means "sums only" (synced IQBN: coefficients are made after the all-reduce)
  float eps, momentum;
  float* running_mean;
  float* running_var;
  float* stats;        // [20C] (fwd: out; bwd: in)
  double* sums_out;    // [8C] raw sums / [8C + 6C] bwd sums + coefficient table
  const float* gamma;  // fwd stats + bwd coefficients
  const float* beta;
};

// dx = k1*dz + k2*x + k3 with xhat = (x - mean)*rstd:  k1 = g r, k2 = -g r^2 mdzx, k3 = -g r (mdz - mean r mdzx)
__device__ __forceinline__ void bwd_coefficients(float gamma, float mean, float rstd, double sdz, double sdzx, double count,
                                                 float& k1, float& k2, float& k3) {
  const float gr = gamma * rstd;
  const float mdz = (float)(sdz / count), mdzx = (float)(sdzx / count);
  k1 = gr;
  k2 = -gr * rstd * mdzx;
  k3 = -gr * (mdz - mean * rstd * mdzx);
}

__device__ __forceinline__ void write_fwd_stats(const TailArgs& t, int i, double s0, double s1) {
  const int n = 4 * t.C;
  double mean = s0 / t.count;
  double var = s1 / t.count - mean * mean;
  if (var < 0.0) var = 0.0;
  var += 1e-8;  // conv.py:557
  const float rstd = (float)(1.0 / sqrt(var + (double)t.eps));
  t.stats[i] = (float)mean;
  t.stats[n + i] = (float)var;
  t.stats[2 * n + i] = rstd;
  if (t.gamma != nullptr) {
    const int c = i >> 2, q = i & 3;
    const float scale = t.gamma[i] * rstd;
    t.stats[3 * n + q * t.C + c] = scale;
    t.stats[4 * n + q * t.C + c] = t.beta[i] - (float)mean * scale;
  }
  if (t.running_mean != nullptr) {  // conv.py:561-562
    t.running_mean[i] = (1.0f - t.momentum) * t.running_mean[i] + t.momentum * (float)mean;
    t.running_var[i] = (1.0f - t.momentum) * t.running_var[i] + t.momentum * (float)var;
  }
}

__device__ __forceinline__ void write_bwd_sums(const TailArgs& t, int i, double s0, double s1) {
  // s0 = sum dz, s1 = sum dz*x  ->  sum dz*xhat = rstd*(s1 - mean*s0)
  const int n = 4 * t.C;
  const double mean = (double)t.stats[i], rstd = (double)t.stats[2 * n + i];
  const double sdzx = rstd * (s1 - mean * s0);
  t.sums_out[i] = s0;
  t.sums_out[n + i] = sdzx;
  if (t.count > 0.0) {
    float* coef = reinterpret_cast<float*>(t.sums_out + 2 * n);
    const int c = i >> 2, q = i & 3, j = q * t.C + c;
    float k1, k2, k3;
    bwd_coefficients(t.gamma[i], (float)mean, (float)rstd, s0, sdzx, t.count, k1, k2, k3);
    coef[j] = k1;
    coef[n + j] = k2;
    coef[2 * n + j] = k3;
  }
}

__device__ __forceinline__ void finish_accumulator(const TailArgs& t, int i, double s0, double s1) {
  if (t.mode == TAIL_RAW_SUMS) {
    t.sums_out[i] = s0;
    t.sums_out[4 * t.C + i] = s1;
  } else if (t.mode == TAIL_FWD_STATS) {
    write_fwd_stats(t, i, s0, s1);
  } else {
    write_bwd_sums(t, i, s0, s1);
  }
}

// Second kernel of every reduction: fold the per-block partials of accumulator i = c*4+q and finish.  Pure latency:
// block = 2 accumulators x 128 split lanes, so a thread walks only nparts / 128 (<= 5) partials, two at a time; the lanes
// fold by shuffle (16 per warp) and a 16-entry shared-memory pass.  grid = ceil(4C / 2).
constexpr int FOLD_ACC = 2, FOLD_LANES = 128;
__global__ void __launch_bounds__(FOLD_ACC * FOLD_LANES) iqbn_fold_kernel(const double* __restrict__ part, int nparts, TailArgs t) {
  pdl_prologue();
  __shared__ double red[2][FOLD_ACC * FOLD_LANES / 32][FOLD_ACC];
  const int n = 4 * t.C;
  const int il = threadIdx.x % FOLD_ACC, gl = threadIdx.x / FOLD_ACC;
  const int i = blockIdx.x * FOLD_ACC + il;
  double s0 = 0.0, s1 = 0.0;
  if (i < n) {
    int sp = gl;
    for (; sp + FOLD_LANES < nparts; sp += 2 * FOLD_LANES) {
      const double a0 = part[((size_t)sp * 2 + 0) * n + i], a1 = part[((size_t)sp * 2 + 1) * n + i];
      const double b0 = part[((size_t)(sp + FOLD_LANES) * 2 + 0) * n + i], b1 = part[((size_t)(sp + FOLD_LANES) * 2 + 1) * n + i];
      s0 += a0 + b0;
      s1 += a1 + b1;
    }
    if (sp < nparts) {
      s0 += part[((size_t)sp * 2 + 0) * n + i];
      s1 += part[((size_t)sp * 2 + 1) * n + i];
    }
  }
  // lanes of a warp: tid = gl*2 + il -> xor over bits 1..4 folds the warp's 16 split lanes of each accumulator
#pragma unroll
  for (int o = FOLD_ACC; o < 32; o <<= 1) {
    s0 += __shfl_xor_sync(0xffffffffu, s0, o);
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane < FOLD_ACC) {
    red[0][warp][lane] = s0;
    red[1][warp][lane] = s1;
  }
  __syncthreads();
  if (threadIdx.x < FOLD_ACC && i < n) {
    s0 = s1 = 0.0;
#pragma unroll
    for (int w = 0; w < FOLD_ACC * FOLD_LANES / 32; ++w) {
      s0 += red[0][w][threadIdx.x];
      s1 += red[1][w][threadIdx.x];
    }
    finish_accumulator(t, i, s0, s1);
  }
}

// =================================================================================================
// Layout BHWQC: rows R = B*H*W, row length L = 4C.  A column vector cv covers V consecutive elements of a row
// (col = cv*V -> q = col / C, c = col % C; V | C so a vector never straddles q).
// A block is rpb row lanes x cvpg column vectors with the column index fastest, so one block-iteration reads
// rpb whole rows (or a <=4 KB contiguous piece of a wide row: blockIdx.y = column group) — long contiguous bursts
// keep HBM pages open (a 128-byte-wide column-slice mapping measured 2x slower).  A thread owns its column vector for
// the whole kernel: per-channel parameters and partial sums stay in registers, U row loads are in flight at once.
// Reductions: fp64 through shared memory across the row lanes, then ONE fp64 atomic per column element per block and
// at most 2 blocks per SM, so an accumulator address sees <= 296 atomics (the first version issued one per thread —
// 2368 per address — and the same-address serialisation in L2 held the kernel at 15% of HBM bandwidth).
// =================================================================================================
struct GeomB {
  int64_t R;      // rows
  int L;          // 4C
  int C;
  int cvpg;       // column vectors per group (per block row)
  int rpb;        // row lanes per block
  int rev;        // 1: visit the rows last-to-first (reductions: the producer's most recent rows are still in L2)
};

// V consecutive fp32 coefficients (V <= 8) as 16-byte loads where alignment allows
template <typename TT, int V>
__device__ __forceinline__ void load_coef(const float* __restrict__ p, float (&out)[V]) {
  if constexpr (V == 8) {
    float a[4], b[4];
    load_vec<float, 4>(p, a);
    load_vec<float, 4>(p + 4, b);
#pragma unroll
    for (int i = 0; i < 4; ++i) { out[i] = a[i]; out[4 + i] = b[i]; }
  } else {
    load_vec<float, V>(p, out);
  }
}

template <int V>
__device__ __forceinline__ void colvec_param_index(int cv, int C, int (&idx)[V]) {
  int col = cv * V;
  int q = col / C, c = col - q * C;
#pragma unroll
  for (int i = 0; i < V; ++i) idx[i] = (c + i) * 4 + q;
}

// MODE 0: sums of x and x^2.  MODE 1: sums of dz and dz*x (dz = dy*act'(x*scale+shift)).
// Memory-level parallelism is what bounds a read-only stream (Little: ~40 KB must be in flight per SM for 6.5 TB/s):
// U raw 16-byte loads per stream are issued back to back before any of them is converted, 1024 threads per SM.
constexpr int IQBN_RED_THREADS = 512;
// FUSE: cooperative launch (all blocks co-resident) — after a grid-wide barrier every warp of the grid folds one
// accumulator's partials and finishes it, so the statistics cost one launch and no atomics.
template <typename T, int V, int MODE, int ACT, int U, bool FUSE>
__global__ void __launch_bounds__(IQBN_RED_THREADS, 2) iqbn_reduce_b(const T* __restrict__ x, const T* __restrict__ dy, GeomB g,
                                                                      const float* __restrict__ gamma,
                                                                      const float* __restrict__ beta, IqbnWs ws, TailArgs tail) {
  pdl_prologue();
  using VecT = Vec<T, V>;
  const int cvl = threadIdx.x % g.cvpg;
  const int rl = threadIdx.x / g.cvpg;
  const int cv = blockIdx.y * g.cvpg + cvl;
  const int64_t coloff = (int64_t)cv * V;
  float scale[V], shift[V];
  if constexpr (MODE == 1 && ACT != QUAN_ACT_NONE) {   // coefficient table in column order: vector loads
    load_coef<float, V>(tail.stats + 12 * g.C + coloff, scale);
    load_coef<float, V>(tail.stats + 16 * g.C + coloff, shift);
  }

  float s0[V], s1[V], k[V];
#pragma unroll
  for (int i = 0; i < V; ++i) s0[i] = s1[i] = k[i] = 0.f;

  const VecT* xr = reinterpret_cast<const VecT*>(x + coloff);
  const VecT* gr = reinterpret_cast<const VecT*>(dy + coloff);
  const int64_t rsv = g.L / V;                                  // row stride in vectors
  const int64_t rs = (int64_t)gridDim.x * g.rpb;                // rows between two visits of this thread
  int64_t r = (int64_t)blockIdx.x * g.rpb + rl;
  const bool lane_on = rl < g.rpb;
  // Reductions walk the tensor back to front: the kernel that produced it (conv epilogue, the next layer's dgrad) wrote
  // front to back, so its last ~100 MB are still in the 126 MB L2, and the apply kernel that follows walks front to back
  // again over what this kernel touched last.
  auto rowi = [&](int64_t rr) { return g.rev ? g.R - 1 - rr : rr; };
  int cnt = 0;
  if (lane_on && r < g.R) {
    cnt = (int)((g.R - 1 - r) / rs) + 1;
    if constexpr (MODE == 0) {                                  // local shift: keeps fp32 partials well conditioned
      const VecT t = xr[rowi(r) * rsv];
#pragma unroll
      for (int i = 0; i < V; ++i) k[i] = to_f32(t.v[i]);
    }
  }
  auto accumulate = [&](const VecT& xa, const VecT& ga) {
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const float xv = to_f32(xa.v[i]);
      if constexpr (MODE == 0) {
        const float d = xv - k[i];
        s0[i] += d;
        s1[i] = fmaf(d, d, s1[i]);
      } else {
        float dz = to_f32(ga.v[i]);
        if constexpr (ACT != QUAN_ACT_NONE) dz *= act_grad<ACT, sizeof(T) == 2>(fmaf(xv, scale[i], shift[i]));
        s0[i] += dz;
        s1[i] = fmaf(dz, xv, s1[i]);
      }
    }
  };
  if (lane_on) {
    for (; r + (U - 1) * rs < g.R; r += U * rs) {
      VecT xa[U], ga[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        xa[u] = xr[rowi(r + u * rs) * rsv];
        if constexpr (MODE == 1) ga[u] = gr[rowi(r + u * rs) * rsv];
      }
#pragma unroll
      for (int u = 0; u < U; ++u) accumulate(xa[u], ga[u]);
    }
    for (; r < g.R; r += rs) {
      VecT xa = xr[rowi(r) * rsv], ga;
      if constexpr (MODE == 1) ga = gr[rowi(r) * rsv];
      accumulate(xa, ga);
    }
  }

  // thread partials -> fp64 raw sums, folded over the row lanes through shared memory one column element at a time
  // (2 doubles per thread: 8 KB), then the block's own slot of the partials buffer: plain stores, no atomics
  __shared__ double red[IQBN_RED_THREADS][2];
  int half = 1;
  while (half * 2 < g.rpb) half *= 2;          // largest power of two below rpb (>= rpb/2)
#pragma unroll
  for (int i = 0; i < V; ++i) {
    double a0, a1;
    if constexpr (MODE == 0) {  // un-shift in fp64: sum x = sum d + n k ; sum x^2 = sum d^2 + 2k sum d + n k^2
      const double kd = (double)k[i], nd = (double)cnt;
      a0 = (double)s0[i] + nd * kd;
      a1 = (double)s1[i] + 2.0 * kd * (double)s0[i] + nd * kd * kd;
    } else {
      a0 = (double)s0[i];
      a1 = (double)s1[i];
    }
    __syncthreads();
    red[threadIdx.x][0] = lane_on ? a0 : 0.0;
    red[threadIdx.x][1] = lane_on ? a1 : 0.0;
    __syncthreads();
    for (int st = half; st >= 1; st >>= 1) {      // tree over the row lanes (fixed order: deterministic)
      if (rl < st && rl + st < g.rpb) {
        red[threadIdx.x][0] += red[threadIdx.x + st * g.cvpg][0];
        red[threadIdx.x][1] += red[threadIdx.x + st * g.cvpg][1];
      }
      __syncthreads();
    }
    if (threadIdx.x < g.cvpg) {
      const int col = (blockIdx.y * g.cvpg + threadIdx.x) * V + i;
      const int q = col / g.C, c = col - q * g.C;
      ws.part[((size_t)blockIdx.x * 2 + 0) * 4 * g.C + c * 4 + q] = red[threadIdx.x][0];
      ws.part[((size_t)blockIdx.x * 2 + 1) * 4 * g.C + c * 4 + q] = red[threadIdx.x][1];
    }
  }
  if constexpr (FUSE) {
    __threadfence();
    cooperative_groups::this_grid().sync();
    const int n = 4 * g.C, lane = threadIdx.x & 31;
    const int full_warps = blockDim.x >> 5;                      // a trailing partial warp sits this out
    const int wib = threadIdx.x >> 5;
    if (wib < full_warps) {
      const int total_warps = gridDim.x * gridDim.y * full_warps;
      for (int i = (blockIdx.y * gridDim.x + blockIdx.x) * full_warps + wib; i < n; i += total_warps) {
        double a0 = 0.0, a1 = 0.0;
        for (int sp = lane; sp < (int)gridDim.x; sp += 32) {     // written by other blocks of this launch: L2 loads
          a0 += __ldcg(ws.part + ((size_t)sp * 2 + 0) * n + i);
          a1 += __ldcg(ws.part + ((size_t)sp * 2 + 1) * n + i);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          a0 += __shfl_xor_sync(0xffffffffu, a0, o);
          a1 += __shfl_xor_sync(0xffffffffu, a1, o);
        }
        if (lane == 0) finish_accumulator(tail, i, a0, a1);
      }
    }
  }
}

// -------------------------------------------------------------------------------------------------------------------
// Reductions fed by the TMA engine.  ncu on iqbn_reduce_b (profiles/r01_iqbn_tune5.log): 64 registers x 1024 threads fill
// the register file, so at most 4 x 16 bytes per thread are in flight and a warp alternates between waiting for its loads
// and reducing them (issue-active 52 %, long-scoreboard stalls) — 3.4-4.4 TB/s.  Here memory-level parallelism does not
// live in registers: one producer lane streams contiguous row tiles (cp.async.bulk, 1-D) into a shared-memory ring of
// NS stages per block, two blocks per SM (~190 KB in flight per SM), and 256 consumer threads reduce from shared memory
// with the same thread -> column mapping (16-byte LDS, conflict-free).  Same partials format and fold kernel.
// -------------------------------------------------------------------------------------------------------------------
constexpr int IQBN_TMA_CONSUMERS = 256;
constexpr int IQBN_TMA_THREADS = IQBN_TMA_CONSUMERS + 32;

__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(ptx::smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(ptx::smem_u32(bar))
               : "memory");
}

struct GeomT {
  int64_t R;        // rows
  int L, C;
  int cvpg, rpb;    // column vectors per row, rows per consumer pass (256 / cvpg)
  int tile_rows;    // rows per stage (multiple of rpb)
  int ntiles;       // ceil(R / tile_rows)
  int stages;
};

template <typename T, int V, int MODE, int ACT>
__global__ void __launch_bounds__(IQBN_TMA_THREADS, 2) iqbn_reduce_tma(const T* __restrict__ x, const T* __restrict__ dy,
                                                                        GeomT g, IqbnWs ws, TailArgs tail) {
  pdl_prologue();
  extern __shared__ __align__(128) uint8_t tsm[];
  using VecT = Vec<T, V>;
  constexpr int NSTREAM = MODE == 1 ? 2 : 1;
  const uint32_t tile_bytes = (uint32_t)g.tile_rows * g.L * sizeof(T);
  uint8_t* ring = tsm;                                                        // [stages][NSTREAM][tile_bytes]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(ring + (size_t)g.stages * NSTREAM * tile_bytes);
  uint64_t* empty_bar = full_bar + g.stages;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    for (int i = 0; i < g.stages; ++i) {
      ptx::mbar_init(full_bar + i, 1);
      ptx::mbar_init(empty_bar + i, IQBN_TMA_CONSUMERS / 32);
    }
    ptx::fence_barrier_init();
  }
  __syncthreads();
  // tiles of this block, last to first (the producer of x wrote front to back: its tail is still in L2)
  const int my_tiles = (g.ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + ((int)blockIdx.x < g.ntiles ? 1 : 0);

  if (warp == IQBN_TMA_CONSUMERS / 32) {
    // ===== producer =====
    int s = 0;
    uint32_t phase = 0;
    for (int i = 0; i < my_tiles; ++i) {
      const int tile = g.ntiles - 1 - ((int)blockIdx.x + i * (int)gridDim.x);
      const int64_t r0 = (int64_t)tile * g.tile_rows;
      const int64_t rows = g.R - r0 < g.tile_rows ? g.R - r0 : g.tile_rows;
      const uint32_t bytes = (uint32_t)(rows * g.L * sizeof(T));
      ptx::mbar_wait(empty_bar + s, phase ^ 1);
      if (ptx::elect_one()) {
        ptx::mbar_arrive_expect_tx(full_bar + s, bytes * NSTREAM);
        uint8_t* dst = ring + (size_t)s * NSTREAM * tile_bytes;
        bulk_load_1d(dst, x + r0 * g.L, bytes, full_bar + s);
        if constexpr (MODE == 1) bulk_load_1d(dst + tile_bytes, dy + r0 * g.L, bytes, full_bar + s);
      }
      __syncwarp();
      if (++s == g.stages) { s = 0; phase ^= 1; }
    }
    return;
  }

  // ===== consumers: thread owns column vector cvl of every row it visits =====
  const int cvl = threadIdx.x % g.cvpg;
  const int rl = threadIdx.x / g.cvpg;
  const bool lane_on = rl < g.rpb;
  const int64_t coloff = (int64_t)cvl * V;
  static_assert(V % 2 == 0, "packed fp32x2 lanes");
  constexpr int P = V / 2;
  // all per-element arithmetic runs on fp32 pairs (add / mul / fma .f32x2)
  uint64_t hscale[P], hshift[P];            // 0.5*scale, 0.5*shift: z/2 feeds tanh directly (sigmoid = 0.5 tanh(z/2) + 0.5)
  if constexpr (MODE == 1 && ACT != QUAN_ACT_NONE) {
    float sc[V], sh[V];
    load_coef<float, V>(tail.stats + 12 * g.C + coloff, sc);
    load_coef<float, V>(tail.stats + 16 * g.C + coloff, sh);
#pragma unroll
    for (int i = 0; i < P; ++i) {
      hscale[i] = f2_pack(0.5f * sc[2 * i], 0.5f * sc[2 * i + 1]);
      hshift[i] = f2_pack(0.5f * sh[2 * i], 0.5f * sh[2 * i + 1]);
    }
  }
  const uint64_t c_half = f2_pack(0.5f, 0.5f), c_nhalf = f2_pack(-0.5f, -0.5f), c_one = f2_pack(1.f, 1.f), c_two = f2_pack(2.f, 2.f);
  uint64_t s0p[P], s1p[P], nk[P];           // sums, and MINUS the local shift k
#pragma unroll
  for (int i = 0; i < P; ++i) s0p[i] = s1p[i] = nk[i] = 0ull;
  int cnt = 0;
  bool have_k = false;
  int s = 0;
  uint32_t phase = 0;
  for (int it = 0; it < my_tiles; ++it) {
    const int tile = g.ntiles - 1 - ((int)blockIdx.x + it * (int)gridDim.x);
    const int64_t r0 = (int64_t)tile * g.tile_rows;
    const int rows = (int)(g.R - r0 < g.tile_rows ? g.R - r0 : g.tile_rows);
    ptx::mbar_wait(full_bar + s, phase);
    const T* xt = reinterpret_cast<const T*>(ring + (size_t)s * NSTREAM * tile_bytes);
    const T* gt = xt + (size_t)g.tile_rows * g.L;
    if (lane_on) {
      // UR rows per trip, all shared-memory loads first: ncu showed the one-row loop stalled on the LDS it had just issued
      // (short-scoreboard on the first use, 4 warps per scheduler cannot hide 30 cycles per row)
      constexpr int UR = MODE == 0 ? 4 : 2;
      for (int r = rl; r < rows; r += UR * g.rpb) {
        VecT xs[UR], gs[UR];
#pragma unroll
        for (int u = 0; u < UR; ++u) {
          const int ru = r + u * g.rpb;
          if (ru < rows) {
            xs[u] = *reinterpret_cast<const VecT*>(xt + (size_t)ru * g.L + coloff);
            if constexpr (MODE == 1) gs[u] = *reinterpret_cast<const VecT*>(gt + (size_t)ru * g.L + coloff);
          }
        }
#pragma unroll
        for (int u = 0; u < UR; ++u) {
          if (r + u * g.rpb >= rows) break;
          const VecT& xa = xs[u];
          const VecT& ga = gs[u];
          if constexpr (MODE == 0) {
            if (!have_k) {                       // local shift: keeps fp32 partials well conditioned
#pragma unroll
              for (int i = 0; i < P; ++i) nk[i] = f2_pack(-to_f32(xa.v[2 * i]), -to_f32(xa.v[2 * i + 1]));
              have_k = true;
            }
          }
#pragma unroll
          for (int i = 0; i < P; ++i) {
            const uint64_t xv = f2_from(&xa.v[2 * i]);
            if constexpr (MODE == 0) {
              const uint64_t d = f2_add(xv, nk[i]);
              s0p[i] = f2_add(s0p[i], d);
              s1p[i] = f2_fma(d, d, s1p[i]);
            } else {
              uint64_t dz = f2_from(&ga.v[2 * i]);
              if constexpr (ACT != QUAN_ACT_NONE) {
                // dz *= s (1 + z (1 - s)),  s = 0.5 t + 0.5,  t = tanh(z/2):  1 - s = 0.5 - 0.5 t,  z (1 - s) = 2 zh (1 - s)
                const uint64_t zh = f2_fma(xv, hscale[i], hshift[i]);
                float z0, z1, t0, t1;
                f2_unpack(zh, z0, z1);
                if constexpr (sizeof(T) == 2) {
                  asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(z0));
                  asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(z1));
                } else {                              // fp32 tensors keep the exact sigmoid (ex2 + rcp)
                  t0 = 2.f * sigmoid_f(2.f * z0) - 1.f;
                  t1 = 2.f * sigmoid_f(2.f * z1) - 1.f;
                }
                const uint64_t t = f2_pack(t0, t1);
                const uint64_t sg = f2_fma(c_half, t, c_half);
                const uint64_t oms = f2_fma(c_nhalf, t, c_half);
                const uint64_t u2 = f2_fma(f2_mul(zh, oms), c_two, c_one);
                dz = f2_mul(dz, f2_mul(sg, u2));
              }
              s0p[i] = f2_add(s0p[i], dz);
              s1p[i] = f2_fma(dz, xv, s1p[i]);
            }
          }
          ++cnt;
        }
      }
    }
    __syncwarp();
    if (ptx::elect_one()) ptx::mbar_arrive(empty_bar + s);
    __syncwarp();
    if (++s == g.stages) { s = 0; phase ^= 1; }
  }
  float s0[V], s1[V], k[V];
#pragma unroll
  for (int i = 0; i < P; ++i) {
    f2_unpack(s0p[i], s0[2 * i], s0[2 * i + 1]);
    f2_unpack(s1p[i], s1[2 * i], s1[2 * i + 1]);
    float a, b;
    f2_unpack(nk[i], a, b);
    k[2 * i] = -a;
    k[2 * i + 1] = -b;
  }

  // fold over the row lanes (consumer threads only: named barrier 1) and write this block's slot of the partials.  All V column
  // elements of a thread go through the tree together, in the ring's memory (free once every consumer has left the tile loop): the
  // first version folded one element at a time through a [256][2] array — 7 barrier rounds x V = 56 per block, ~3 us of every launch,
  // a third of the kernel on the 2-8 MB tensors of the narrow layers.
  int half = 1;
  while (half * 2 < g.rpb) half *= 2;
  double (*redv)[2 * V] = reinterpret_cast<double (*)[2 * V]>(ring);          // [256][2V] doubles <= 32 KB <= one ring stage pair
  asm volatile("bar.sync 1, %0;" ::"n"(IQBN_TMA_CONSUMERS) : "memory");      // every consumer is done with the ring
#pragma unroll
  for (int i = 0; i < V; ++i) {
    double a0, a1;
    if constexpr (MODE == 0) {
      const double kd = (double)k[i], nd = (double)cnt;
      a0 = (double)s0[i] + nd * kd;
      a1 = (double)s1[i] + 2.0 * kd * (double)s0[i] + nd * kd * kd;
    } else {
      a0 = (double)s0[i];
      a1 = (double)s1[i];
    }
    redv[threadIdx.x][2 * i] = lane_on ? a0 : 0.0;
    redv[threadIdx.x][2 * i + 1] = lane_on ? a1 : 0.0;
  }
  asm volatile("bar.sync 1, %0;" ::"n"(IQBN_TMA_CONSUMERS) : "memory");
  for (int st = half; st >= 1; st >>= 1) {
    if (rl < st && rl + st < g.rpb) {
#pragma unroll
      for (int e = 0; e < 2 * V; ++e) redv[threadIdx.x][e] += redv[threadIdx.x + st * g.cvpg][e];
    }
    asm volatile("bar.sync 1, %0;" ::"n"(IQBN_TMA_CONSUMERS) : "memory");
  }
  if ((int)threadIdx.x < g.cvpg) {
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const int col = threadIdx.x * V + i;
      const int q = col / g.C, c = col - q * g.C;
      ws.part[((size_t)blockIdx.x * 2 + 0) * 4 * g.C + c * 4 + q] = redv[threadIdx.x][2 * i];
      ws.part[((size_t)blockIdx.x * 2 + 1) * 4 * g.C + c * 4 + q] = redv[threadIdx.x][2 * i + 1];
    }
  }
}

// Small tensors (the 32^2 / 64^2 maps of the narrow layers: 2-16 MB): ONE launch.  The TMA-ring kernel above plus the fold kernel cost
// 9-14 us + 4.6 us + a launch gap for ~2 us of traffic, 150 times per QUAN-YOLO11n step.  Here a block streams its rows with plain
// 16-byte loads (U rows in flight), folds its row lanes through shared memory once, adds its 8C sums to fp64 accumulators in L2 with
// atomics (<= 64 K of them per launch: ~3 us at the L2's fp64-atomic rate, overlapped with the other blocks' streams) and the LAST
// block (ticket) finishes: mean / var / rstd / running statistics / coefficient tables, then re-zeroes accumulators and ticket for
// the next launch on this stream.  Summation order across blocks is not fixed; the sums are fp64 (differences ~1e-16 relative).
template <typename T, int V, int MODE, int ACT>
__global__ void __launch_bounds__(256) iqbn_reduce_small(const T* __restrict__ x, const T* __restrict__ dy, int64_t R, int C, int cvpg,
                                                          int rpb, IqbnWs ws, TailArgs tail) {
  pdl_prologue();
  extern __shared__ __align__(16) uint8_t sm_raw[];
  double (*sred)[2 * V] = reinterpret_cast<double (*)[2 * V]>(sm_raw);        // [256][2V]
  __shared__ int is_last;
  using VecT = Vec<T, V>;
  const int L = 4 * C;
  const int cvl = threadIdx.x % cvpg, rl = threadIdx.x / cvpg;
  const bool on = rl < rpb;
  const int64_t coloff = (int64_t)cvl * V;
  float sc[V], sh[V];
  if constexpr (MODE == 1 && ACT != QUAN_ACT_NONE) {
    load_coef<float, V>(tail.stats + 12 * C + coloff, sc);
    load_coef<float, V>(tail.stats + 16 * C + coloff, sh);
  }
  float s0[V], s1[V], k[V];
#pragma unroll
  for (int i = 0; i < V; ++i) s0[i] = s1[i] = k[i] = 0.f;
  int cnt = 0;
  bool have_k = false;
  constexpr int U = 4;
  const int64_t step = (int64_t)gridDim.x * rpb;
  if (on) {
    for (int64_t r = (int64_t)blockIdx.x * rpb + rl; r < R; r += U * step) {
      VecT xs[U], gs[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t ru = r + u * step;
        if (ru < R) {
          xs[u] = *reinterpret_cast<const VecT*>(x + ru * L + coloff);
          if constexpr (MODE == 1) gs[u] = *reinterpret_cast<const VecT*>(dy + ru * L + coloff);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (r + u * step >= R) break;
        if constexpr (MODE == 0) {
          if (!have_k) {                               // local shift: keeps the fp32 partials well conditioned
#pragma unroll
            for (int i = 0; i < V; ++i) k[i] = to_f32(xs[u].v[i]);
            have_k = true;
          }
        }
#pragma unroll
        for (int i = 0; i < V; ++i) {
          const float xv = to_f32(xs[u].v[i]);
          if constexpr (MODE == 0) {
            const float d = xv - k[i];
            s0[i] += d;
            s1[i] = fmaf(d, d, s1[i]);
          } else {
            float dz = to_f32(gs[u].v[i]);
            if constexpr (ACT != QUAN_ACT_NONE) dz *= act_grad<ACT, sizeof(T) == 2>(fmaf(xv, sc[i], sh[i]));
            s0[i] += dz;
            s1[i] = fmaf(dz, xv, s1[i]);
          }
        }
        ++cnt;
      }
    }
  }
  double v[2 * V];                                     // this thread's sums, exact in fp64 from here on
#pragma unroll
  for (int i = 0; i < V; ++i) {
    double a0, a1;
    if constexpr (MODE == 0) {
      const double kd = (double)k[i], nd = (double)cnt;
      a0 = (double)s0[i] + nd * kd;
      a1 = (double)s1[i] + 2.0 * kd * (double)s0[i] + nd * kd * kd;
    } else {
      a0 = (double)s0[i];
      a1 = (double)s1[i];
    }
    v[2 * i] = on ? a0 : 0.0;
    v[2 * i + 1] = on ? a1 : 0.0;
  }
  if ((cvpg & (cvpg - 1)) == 0) {
    // power-of-two column groups (every model shape): the row lanes of a warp meet by shuffle, the warps (or, for rows wider than a
    // warp, the row lanes) in ONE shared-memory round, and each of the 8C sums is finished and sent to L2 by its own thread.  The
    // five-level fp64 tree through shared memory it replaces (barrier + 16 dependent DADDs per level) was 57 % of this kernel's stall
    // samples on the 2 MB tensors (ncu source page, profiles/r02_ncu_iqbn_reduce_small.txt): 13.4-14.9 us for a 4 us stream.
    const int lane = threadIdx.x & 31;
    for (int off = cvpg; off < 32; off <<= 1) {
#pragma unroll
      for (int e = 0; e < 2 * V; ++e) v[e] += __shfl_xor_sync(0xffffffffu, v[e], off);
    }
    double* flat = &sred[0][0];                        // [groups][cvpg][2V]
    const int groups = cvpg < 32 ? 256 / 32 : rpb;     // partial sets left: one per warp, or one per row lane
    const int grp = cvpg < 32 ? (int)(threadIdx.x >> 5) : rl;
    if (cvpg >= 32 || lane < cvpg) {
#pragma unroll
      for (int e = 0; e < 2 * V; ++e) flat[((size_t)grp * cvpg + cvl) * 2 * V + e] = v[e];
    }
    __syncthreads();
    const int nsum = cvpg * 2 * V;                     // = 8C
    for (int t = threadIdx.x; t < nsum; t += 256) {
      double sum = 0.0;
      for (int gi = 0; gi < groups; ++gi) sum += flat[(size_t)gi * nsum + t];
      const int e = t % (2 * V), cv = t / (2 * V);
      const int col = cv * V + (e >> 1);
      const int q = col / C, c = col - q * C;
      atomicAdd(ws.acc + (e & 1) * 4 * C + c * 4 + q, sum);
    }
  } else {
#pragma unroll
  for (int e = 0; e < 2 * V; ++e) sred[threadIdx.x][e] = v[e];
  __syncthreads();
  int half = 1;
  while (half * 2 < rpb) half *= 2;
  for (int st = half; st >= 1; st >>= 1) {
    if (on && rl < st && rl + st < rpb) {
#pragma unroll
      for (int e = 0; e < 2 * V; ++e) sred[threadIdx.x][e] += sred[threadIdx.x + st * cvpg][e];
    }
    __syncthreads();
  }
  if ((int)threadIdx.x < cvpg) {
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const int col = threadIdx.x * V + i;
      const int q = col / C, c = col - q * C;
      atomicAdd(ws.acc + c * 4 + q, sred[threadIdx.x][2 * i]);
      atomicAdd(ws.acc + 4 * C + c * 4 + q, sred[threadIdx.x][2 * i + 1]);
    }
  }
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = atomicAdd(ws.counter, 1u) == gridDim.x - 1 ? 1 : 0;
  __syncthreads();
  if (is_last) {
    __threadfence();
    for (int i = threadIdx.x; i < 4 * C; i += blockDim.x) {
      const double t0 = __ldcg(ws.acc + i), t1 = __ldcg(ws.acc + 4 * C + i);
      finish_accumulator(tail, i, t0, t1);
      ws.acc[i] = 0.0;
      ws.acc[4 * C + i] = 0.0;
    }
    if (threadIdx.x == 0) *ws.counter = 0u;
  }
}

// Elementwise coefficient form shared by fwd apply / eval / bwd apply:
//   FWD : y  = act(x*scale + shift)
//   BWD : dx = k1*dz + k2*x + k3,  dz = dy*act'(x*scale + shift)
struct ApplyArgs {
  const float* gamma;
  const float* beta;
  const float* stats;         // train: [12C]; NULL in eval mode
  const float* running_mean;  // eval mode
  const float* running_var;
  float eps;
  const double* sums;         // bwd train: [8C]; NULL => eval-mode backward (statistics are constants)
  double count;
  float* dgamma;              // bwd train: optional outputs (written by block 0)
  float* dbeta;
  int C;
  const float* coefT;         // bwd train, BHWQC: k1T | k2T | k3T (index q*C+c), made by the bwd-reduce tail / bwd_coef
  // forward, training, statistics still RAW: [8C] fp64 sums in L2 left by the conv epilogue (fsums != NULL).  Every thread finishes
  // the statistics of its own columns, block (0,0) writes the stats table / running statistics, the last block to have read (ticket)
  // re-zeroes the accumulators.  No fold launch between the conv and this kernel.
  double* fsums;
  unsigned* ticket;
  double fcount;
  float momentum;
  float* stats_w;
  float* running_mean_w;
  float* running_var_w;
};

__device__ __forceinline__ void coeff_fwd(const ApplyArgs& a, int idx, float& scale, float& shift) {
  float mean, rstd;
  if (a.stats != nullptr) {
    mean = a.stats[idx];
    rstd = a.stats[8 * a.C + idx];
  } else {
    mean = a.running_mean[idx];
    rstd = 1.0f / sqrtf(a.running_var[idx] + a.eps);  // conv.py:550
  }
  scale = a.gamma[idx] * rstd;
  shift = a.beta[idx] - mean * scale;
}

__device__ __forceinline__ void coeff_bwd(const ApplyArgs& a, int idx, float& k1, float& k2, float& k3) {
  float mean, rstd;
  if (a.stats != nullptr) {
    mean = a.stats[idx];
    rstd = a.stats[8 * a.C + idx];
  } else {
    mean = a.running_mean[idx];
    rstd = 1.0f / sqrtf(a.running_var[idx] + a.eps);
  }
  const float gr = a.gamma[idx] * rstd;
  if (a.sums != nullptr) {
    // dx = g*r*(dz - mdz - xhat*mdzx), xhat = x*r - mean*r
    const float mdz = (float)(a.sums[idx] / a.count);
    const float mdzx = (float)(a.sums[4 * a.C + idx] / a.count);
    k1 = gr;
    k2 = -gr * rstd * mdzx;
    k3 = -gr * (mdz - mean * rstd * mdzx);
  } else {
    k1 = gr;
    k2 = 0.f;
    k3 = 0.f;
  }
}

__device__ __forceinline__ void write_param_grads(const ApplyArgs& a) {
  if (a.dgamma != nullptr && a.sums != nullptr && blockIdx.x == 0 && blockIdx.y == 0) {
    for (int i = threadIdx.x; i < 4 * a.C; i += blockDim.x) {
      a.dbeta[i] = (float)a.sums[i];
      a.dgamma[i] = (float)a.sums[4 * a.C + i];
    }
  }
}

template <typename T, int V, int ACT, bool BWD, int U>
__global__ void __launch_bounds__(256, BWD ? 3 : 4) iqbn_apply_b(const T* __restrict__ x, const T* __restrict__ dy,
                                                    T* __restrict__ out, GeomB g, ApplyArgs a) {
  pdl_prologue();
  const int cvl = threadIdx.x % g.cvpg;
  const int rl = threadIdx.x / g.cvpg;
  const int cv = blockIdx.y * g.cvpg + cvl;
  const int64_t coloff = (int64_t)cv * V;
  float scale[V], shift[V], k1[V], k2[V], k3[V];
  if (!BWD && a.fsums != nullptr) {
    // one thread per column finishes that column's statistics (fp64 division + square root: ~100 instructions — done once per block,
    // not once per thread and column: the first version, V columns in every thread, was slower than the fold launch it replaced),
    // the block's threads then pick their V columns up from shared memory.  4C <= 512 (narrow layers only).
    __shared__ float s_scale[512], s_shift[512];
    for (int col = threadIdx.x; col < 4 * g.C; col += blockDim.x) {
      const int q = col / g.C, c = col - q * g.C, idx = c * 4 + q;
      const double s0 = __ldcg(a.fsums + idx), s1 = __ldcg(a.fsums + 4 * g.C + idx);
      const double mean = s0 / a.fcount;
      double var = s1 / a.fcount - mean * mean;
      if (var < 0.0) var = 0.0;
      var += 1e-8;                                       // conv.py:557
      const float rstd = (float)(1.0 / sqrt(var + (double)a.eps));
      const float sc = a.gamma[idx] * rstd;
      s_scale[col] = sc;
      s_shift[col] = a.beta[idx] - (float)mean * sc;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < V; ++i) {
      scale[i] = s_scale[coloff + i];
      shift[i] = s_shift[coloff + i];
    }
    if (blockIdx.x == 0 && blockIdx.y == 0) {            // the table the backward reads + the running statistics (conv.py:561-562)
      TailArgs t = {};
      t.mode = TAIL_FWD_STATS; t.C = g.C; t.count = a.fcount; t.eps = a.eps; t.momentum = a.momentum;
      t.running_mean = a.running_mean_w; t.running_var = a.running_var_w; t.stats = a.stats_w; t.gamma = a.gamma; t.beta = a.beta;
      for (int i = threadIdx.x; i < 4 * g.C; i += blockDim.x) write_fwd_stats(t, i, __ldcg(a.fsums + i), __ldcg(a.fsums + 4 * g.C + i));
    }
    __syncthreads();                                     // every read of the accumulators by this block is done
    if (threadIdx.x == 0) {
      const unsigned nblocks = gridDim.x * gridDim.y;
      if (atomicAdd(a.ticket, 1u) == nblocks - 1) {      // all blocks have read: leave accumulators and ticket zero for the next layer
        __threadfence();
        for (int i = 0; i < 8 * g.C; ++i) a.fsums[i] = 0.0;
        *a.ticket = 0u;
      }
    }
  } else if (a.stats != nullptr && (!BWD || a.coefT != nullptr)) {
    // training path: coefficient tables in column order (stats[12C..20C), coefT) — a few 16-byte loads per thread
    load_coef<float, V>(a.stats + 12 * g.C + coloff, scale);
    load_coef<float, V>(a.stats + 16 * g.C + coloff, shift);
    if constexpr (BWD) {
      load_coef<float, V>(a.coefT + coloff, k1);
      load_coef<float, V>(a.coefT + 4 * g.C + coloff, k2);
      load_coef<float, V>(a.coefT + 8 * g.C + coloff, k3);
    }
  } else {
    int pidx[V];
    colvec_param_index<V>(cv, g.C, pidx);
#pragma unroll
    for (int i = 0; i < V; ++i) {
      coeff_fwd(a, pidx[i], scale[i], shift[i]);
      if constexpr (BWD) coeff_bwd(a, pidx[i], k1[i], k2[i], k3[i]);
    }
  }
  if constexpr (BWD) write_param_grads(a);

  using VecT = Vec<T, V>;
  const VecT* xr = reinterpret_cast<const VecT*>(x + coloff);
  const VecT* gr = reinterpret_cast<const VecT*>(dy + coloff);
  VecT* orow = reinterpret_cast<VecT*>(out + coloff);
  const int64_t rsv = g.L / V;
  const int64_t rs = (int64_t)gridDim.x * g.rpb;
  auto apply = [&](const VecT& xa, const VecT& ga) {
    VecT o;
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const float xv = to_f32(xa.v[i]);
      if constexpr (!BWD) {
        o.v[i] = from_f32<T>(act_fwd<ACT, sizeof(T) == 2>(fmaf(xv, scale[i], shift[i])));
      } else {
        float dz = to_f32(ga.v[i]);
        if constexpr (ACT != QUAN_ACT_NONE) dz *= act_grad<ACT, sizeof(T) == 2>(fmaf(xv, scale[i], shift[i]));
        o.v[i] = from_f32<T>(fmaf(k1[i], dz, fmaf(k2[i], xv, k3[i])));
      }
    }
    return o;
  };
  if (rl < g.rpb) {
    int64_t r = (int64_t)blockIdx.x * g.rpb + rl;
    for (; r + (U - 1) * rs < g.R; r += U * rs) {
      VecT xa[U], ga[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        xa[u] = xr[(r + u * rs) * rsv];
        if constexpr (BWD) ga[u] = gr[(r + u * rs) * rsv];
      }
#pragma unroll
      for (int u = 0; u < U; ++u) orow[(r + u * rs) * rsv] = apply(xa[u], ga[u]);
    }
    for (; r < g.R; r += rs) {
      VecT xa = xr[r * rsv], ga;
      if constexpr (BWD) ga = gr[r * rsv];
      orow[r * rsv] = apply(xa, ga);
    }
  }
}

// Backward apply that emits G = M^T dx for the QConv2D that produced x (the Conv block's conv -> IQBN order, conv.py:805-
// 809), so the separate mix pass over the gradient (one full read + write) disappears and G is rounded to bf16 once.
// Thread mapping: the four components of a channel vector sit in ONE warp — a slot of 4*gl lanes = 4 components x gl
// consecutive column vectors, lane = q*gl + j — so every global access still covers gl*16 contiguous bytes per component
// and the other three components of a thread's channels arrive by three xor-shuffles per element (lane ^ gl, ^ 2gl,
// ^ 3gl).  No shared memory, no block barrier; U rows per thread in flight as in iqbn_apply_b.
struct GeomW {
  int64_t R;      // rows
  int L, C;
  int cpq;        // column vectors per component (C / V)
  int gl;         // column vectors per component and slot (power of two, <= 8, divides cpq)
  int nb;         // channel blocks per row (cpq / gl)
  int rpb;        // rows per block step ((256 / (4*gl)) / nb)
};
template <typename T, int V, int ACT, int U>
__global__ void __launch_bounds__(256, 2) iqbn_apply_bwd_mix_b(const T* __restrict__ x, const T* __restrict__ dy,
                                                              T* __restrict__ out, GeomW g, ApplyArgs a, Mix16 mt) {
  pdl_prologue();
  using VecT = Vec<T, V>;
  const int slot = threadIdx.x / (4 * g.gl), ls = threadIdx.x % (4 * g.gl);
  const int q = ls / g.gl, j = ls - q * g.gl;
  const int rl = slot / g.nb, cb = slot - rl * g.nb;
  const bool lane_on = rl < g.rpb;
  const int cv = q * g.cpq + cb * g.gl + j;            // this thread's column vector (fixed for the whole kernel)
  const int64_t coloff = lane_on ? (int64_t)cv * V : 0;
  float scale[V], shift[V], k1[V], k2[V], k3[V];
  load_coef<float, V>(a.stats + 12 * g.C + coloff, scale);
  load_coef<float, V>(a.stats + 16 * g.C + coloff, shift);
  load_coef<float, V>(a.coefT + coloff, k1);
  load_coef<float, V>(a.coefT + 4 * g.C + coloff, k2);
  load_coef<float, V>(a.coefT + 8 * g.C + coloff, k3);
  write_param_grads(a);
  float m1 = mt.m[q * 4 + q], mx1 = mt.m[q * 4 + (q ^ 1)], mx2 = mt.m[q * 4 + (q ^ 2)], mx3 = mt.m[q * 4 + (q ^ 3)];
  const VecT* xr = reinterpret_cast<const VecT*>(x + coloff);
  const VecT* gr = reinterpret_cast<const VecT*>(dy + coloff);
  VecT* orow = reinterpret_cast<VecT*>(out + coloff);
  const int64_t rsv = g.L / V;
  const int64_t rs = (int64_t)gridDim.x * g.rpb;
  for (int64_t base = (int64_t)blockIdx.x * g.rpb; base < g.R; base += U * rs) {   // warp-uniform trip count (shuffles below)
    VecT xa[U], ga[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t r = base + u * rs + rl;
      if (lane_on && r < g.R) {
        xa[u] = xr[r * rsv];
        ga[u] = gr[r * rsv];
      } else {
        xa[u] = VecT{};
        ga[u] = VecT{};
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t r = base + u * rs + rl;
      VecT o;
#pragma unroll
      for (int i = 0; i < V; ++i) {
        const float xv = to_f32(xa[u].v[i]);
        float dz = to_f32(ga[u].v[i]);
        if constexpr (ACT != QUAN_ACT_NONE) dz *= act_grad<ACT, sizeof(T) == 2>(fmaf(xv, scale[i], shift[i]));
        const float d = fmaf(k1[i], dz, fmaf(k2[i], xv, k3[i]));
        const float d1 = __shfl_xor_sync(0xffffffffu, d, g.gl);
        const float d2 = __shfl_xor_sync(0xffffffffu, d, 2 * g.gl);
        const float d3 = __shfl_xor_sync(0xffffffffu, d, 3 * g.gl);
        o.v[i] = from_f32<T>(m1 * d + mx1 * d1 + mx2 * d2 + mx3 * d3);
      }
      if (lane_on && r < g.R) orow[r * rsv] = o;
    }
  }
}

// TMA-fed backward apply (same ring as iqbn_reduce_tma, two input streams) with the warp-local quaternion mapping of
// iqbn_apply_bwd_mix_b: consumers read x / dy rows from shared memory, form dx, optionally exchange the four components
// by xor-shuffles to emit G = M^T dx, and store 16-byte vectors straight to global memory.
struct GeomTW {
  int64_t R;
  int L, C;
  int cpq, gl, nb, rpb;   // as GeomW: column vectors per component, per slot; channel blocks per row; rows per pass
  int tile_rows, ntiles, stages;
};
template <typename T, int V, int ACT, bool MIX>
__global__ void __launch_bounds__(IQBN_TMA_THREADS, 2) iqbn_apply_bwd_tma(const T* __restrict__ x, const T* __restrict__ dy,
                                                                           T* __restrict__ out, GeomTW g, ApplyArgs a, Mix16 mt) {
  pdl_prologue();
  extern __shared__ __align__(128) uint8_t tsm[];
  using VecT = Vec<T, V>;
  const uint32_t tile_bytes = (uint32_t)g.tile_rows * g.L * sizeof(T);
  uint8_t* ring = tsm;                                                        // [stages][2][tile_bytes]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(ring + (size_t)g.stages * 2 * tile_bytes);
  uint64_t* empty_bar = full_bar + g.stages;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    for (int i = 0; i < g.stages; ++i) {
      ptx::mbar_init(full_bar + i, 1);
      ptx::mbar_init(empty_bar + i, IQBN_TMA_CONSUMERS / 32);
    }
    ptx::fence_barrier_init();
  }
  __syncthreads();
  const int my_tiles = (g.ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + ((int)blockIdx.x < g.ntiles ? 1 : 0);
  if (warp == IQBN_TMA_CONSUMERS / 32) {
    int s = 0;
    uint32_t phase = 0;
    for (int i = 0; i < my_tiles; ++i) {
      const int tile = (int)blockIdx.x + i * (int)gridDim.x;      // front to back (the reduction before it went back to front)
      const int64_t r0 = (int64_t)tile * g.tile_rows;
      const int64_t rows = g.R - r0 < g.tile_rows ? g.R - r0 : g.tile_rows;
      const uint32_t bytes = (uint32_t)(rows * g.L * sizeof(T));
      ptx::mbar_wait(empty_bar + s, phase ^ 1);
      if (ptx::elect_one()) {
        ptx::mbar_arrive_expect_tx(full_bar + s, bytes * 2);
        uint8_t* dst = ring + (size_t)s * 2 * tile_bytes;
        bulk_load_1d(dst, x + r0 * g.L, bytes, full_bar + s);
        bulk_load_1d(dst + tile_bytes, dy + r0 * g.L, bytes, full_bar + s);
      }
      __syncwarp();
      if (++s == g.stages) { s = 0; phase ^= 1; }
    }
    return;
  }
  const int slot = threadIdx.x / (4 * g.gl), ls = threadIdx.x % (4 * g.gl);
  const int q = ls / g.gl, j = ls - q * g.gl;
  const int rl = slot / g.nb, cb = slot - rl * g.nb;
  const bool lane_on = rl < g.rpb;
  const int cv = q * g.cpq + cb * g.gl + j;
  const int64_t coloff = lane_on ? (int64_t)cv * V : 0;
  float scale[V], shift[V], k1[V], k2[V], k3[V];
  load_coef<float, V>(a.stats + 12 * g.C + coloff, scale);
  load_coef<float, V>(a.stats + 16 * g.C + coloff, shift);
  load_coef<float, V>(a.coefT + coloff, k1);
  load_coef<float, V>(a.coefT + 4 * g.C + coloff, k2);
  load_coef<float, V>(a.coefT + 8 * g.C + coloff, k3);
  if (a.dgamma != nullptr && a.sums != nullptr && blockIdx.x == 0) {   // consumers only: stride 256, not blockDim.x
    for (int i = threadIdx.x; i < 4 * a.C; i += IQBN_TMA_CONSUMERS) {
      a.dbeta[i] = (float)a.sums[i];
      a.dgamma[i] = (float)a.sums[4 * a.C + i];
    }
  }
  const float m1 = mt.m[q * 4 + q], mx1 = mt.m[q * 4 + (q ^ 1)], mx2 = mt.m[q * 4 + (q ^ 2)], mx3 = mt.m[q * 4 + (q ^ 3)];
  int s = 0;
  uint32_t phase = 0;
  for (int it = 0; it < my_tiles; ++it) {
    const int tile = (int)blockIdx.x + it * (int)gridDim.x;
    const int64_t r0 = (int64_t)tile * g.tile_rows;
    const int rows = (int)(g.R - r0 < g.tile_rows ? g.R - r0 : g.tile_rows);
    ptx::mbar_wait(full_bar + s, phase);
    const T* xt = reinterpret_cast<const T*>(ring + (size_t)s * 2 * tile_bytes);
    const T* gt = xt + (size_t)g.tile_rows * g.L;
    for (int rb = 0; rb < rows; rb += g.rpb) {                    // warp-uniform trip count (shuffles below)
      const int r = rb + rl;
      const bool on = lane_on && r < rows;
      VecT xa = VecT{}, ga = VecT{};
      if (on) {
        xa = *reinterpret_cast<const VecT*>(xt + (size_t)r * g.L + coloff);
        ga = *reinterpret_cast<const VecT*>(gt + (size_t)r * g.L + coloff);
      }
      VecT o;
#pragma unroll
      for (int i = 0; i < V; ++i) {
        const float xv = to_f32(xa.v[i]);
        float dz = to_f32(ga.v[i]);
        if constexpr (ACT != QUAN_ACT_NONE) dz *= act_grad<ACT, sizeof(T) == 2>(fmaf(xv, scale[i], shift[i]));
        const float d = fmaf(k1[i], dz, fmaf(k2[i], xv, k3[i]));
        if constexpr (MIX) {
          const float d1 = __shfl_xor_sync(0xffffffffu, d, g.gl);
          const float d2 = __shfl_xor_sync(0xffffffffu, d, 2 * g.gl);
          const float d3 = __shfl_xor_sync(0xffffffffu, d, 3 * g.gl);
          o.v[i] = from_f32<T>(m1 * d + mx1 * d1 + mx2 * d2 + mx3 * d3);
        } else {
          o.v[i] = from_f32<T>(d);
        }
      }
      if (on) *reinterpret_cast<VecT*>(out + (r0 + r) * g.L + coloff) = o;
    }
    __syncwarp();
    if (ptx::elect_one()) ptx::mbar_arrive(empty_bar + s);
    __syncwarp();
    if (++s == g.stages) { s = 0; phase ^= 1; }
  }
}

// =================================================================================================
// Layout BCHWQ: for a fixed (b,c) the plane is HW*4 contiguous elements, q = element & 3.
// grid = (splits, C); a thread walks vectors v of channel c: b = v / vpp, i = v % vpp.
// V is 4 (one quaternion) or 8 (two quaternions); params depend on q only -> 4 registers each.
// =================================================================================================
struct GeomA {
  int B, C;
  int64_t plane;  // H*W*4 elements
  int64_t vpp;    // vectors per plane
};

template <typename T, int V, int MODE, int ACT>
__global__ void __launch_bounds__(256) iqbn_reduce_a(const T* __restrict__ x, const T* __restrict__ dy, GeomA g,
                                                     const float* __restrict__ gamma,
                                                     const float* __restrict__ beta, IqbnWs ws, TailArgs tail) {
  pdl_prologue();
  static_assert(V == 4 || V == 8, "a vector holds whole quaternions");
  const int c = blockIdx.y;
  float scale[4], shift[4];
  if constexpr (MODE == 1 && ACT != QUAN_ACT_NONE) {
    const int n = 4 * g.C;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float mean = tail.stats[c * 4 + q], rstd = tail.stats[2 * n + c * 4 + q];
      scale[q] = gamma[c * 4 + q] * rstd;
      shift[q] = beta[c * 4 + q] - mean * scale[q];
    }
  }
  float s0[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f}, k[4] = {0.f, 0.f, 0.f, 0.f};
  int cnt = 0;  // quaternions seen by this thread
  const int64_t total = (int64_t)g.B * g.vpp;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < total; v += stride) {
    const int64_t b = v / g.vpp, i = v - b * g.vpp;
    const int64_t off = (b * g.C + c) * g.plane + i * V;
    float xv[V];
    load_vec<T, V>(x + off, xv);
    if constexpr (MODE == 0) {
      if (cnt == 0) {
#pragma unroll
        for (int q = 0; q < 4; ++q) k[q] = xv[q];
      }
#pragma unroll
      for (int e = 0; e < V; ++e) {
        float d = xv[e] - k[e & 3];
        s0[e & 3] += d;
        s1[e & 3] = fmaf(d, d, s1[e & 3]);
      }
    } else {
      float gv[V];
      load_vec<T, V>(dy + off, gv);
#pragma unroll
      for (int e = 0; e < V; ++e) {
        float dz = gv[e];
        if constexpr (ACT != QUAN_ACT_NONE) dz *= act_grad<ACT>(fmaf(xv[e], scale[e & 3], shift[e & 3]));
        s0[e & 3] += dz;
        s1[e & 3] = fmaf(dz, xv[e], s1[e & 3]);
      }
    }
    cnt += V / 4;
  }

  // thread -> fp64 raw sums, then warp shuffle, then one atomic per warp per value
  double a0[4], a1[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    if constexpr (MODE == 0) {
      double kd = (double)k[q], nd = (double)cnt;
      a0[q] = (double)s0[q] + nd * kd;
      a1[q] = (double)s1[q] + 2.0 * kd * (double)s0[q] + nd * kd * kd;
    } else {
      a0[q] = (double)s0[q];
      a1[q] = (double)s1[q];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a0[q] += __shfl_xor_sync(0xffffffffu, a0[q], o);
      a1[q] += __shfl_xor_sync(0xffffffffu, a1[q], o);
    }
  }
  __shared__ double wred[8][8];
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      wred[threadIdx.x >> 5][q] = a0[q];
      wred[threadIdx.x >> 5][4 + q] = a1[q];
    }
  }
  __syncthreads();
  if (threadIdx.x < 8) {
    double acc = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) acc += wred[w][threadIdx.x];
    const int n = 4 * g.C, which = threadIdx.x >> 2, q = threadIdx.x & 3;
    ws.part[((size_t)blockIdx.x * 2 + which) * n + c * 4 + q] = acc;
  }
}

template <typename T, int V, int ACT, bool BWD, bool MIX>
__global__ void __launch_bounds__(256) iqbn_apply_a(const T* __restrict__ x, const T* __restrict__ dy,
                                                    T* __restrict__ out, GeomA g, ApplyArgs a, Mix16 mix) {
  pdl_prologue();
  const int c = blockIdx.y;
  float scale[4], shift[4], k1[4], k2[4], k3[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    coeff_fwd(a, c * 4 + q, scale[q], shift[q]);
    if constexpr (BWD) coeff_bwd(a, c * 4 + q, k1[q], k2[q], k3[q]);
  }
  if constexpr (BWD) write_param_grads(a);
  const int64_t total = (int64_t)g.B * g.vpp;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < total; v += stride) {
    const int64_t b = v / g.vpp, i = v - b * g.vpp;
    const int64_t off = (b * g.C + c) * g.plane + i * V;
    float xv[V], ov[V];
    load_vec<T, V>(x + off, xv);
    if constexpr (!BWD) {
#pragma unroll
      for (int e = 0; e < V; ++e) ov[e] = act_fwd<ACT>(fmaf(xv[e], scale[e & 3], shift[e & 3]));
    } else {
      float gv[V];
      load_vec<T, V>(dy + off, gv);
#pragma unroll
      for (int e = 0; e < V; ++e) {
        float dz = gv[e];
        if constexpr (ACT != QUAN_ACT_NONE) dz *= act_grad<ACT>(fmaf(xv[e], scale[e & 3], shift[e & 3]));
        ov[e] = fmaf(k1[e & 3], dz, fmaf(k2[e & 3], xv[e], k3[e & 3]));
      }
      if constexpr (MIX) {
#pragma unroll
        for (int e0 = 0; e0 < V; e0 += 4) {
          float in4[4] = {ov[e0], ov[e0 + 1], ov[e0 + 2], ov[e0 + 3]}, o4[4];
          apply_mix(mix, in4, o4);
          ov[e0] = o4[0]; ov[e0 + 1] = o4[1]; ov[e0 + 2] = o4[2]; ov[e0 + 3] = o4[3];
        }
      }
    }
    store_vec<T, V>(out + off, ov);
  }
}

// ---- generic scalar fallback for layout BCHWQ when the plane is not 16-B/8-B vectorisable (odd H*W in bf16):
// handled by choosing V=4 with bf16 (8-B vectors), always legal since a plane is a multiple of 4 elements.

// =================================================================================================
// host-side launch helpers
// =================================================================================================
struct LaunchB {
  GeomB g;
  dim3 grid, block;
  int V;
};

// Tuning overrides (bring-up only): QUAN_IQBN_BPS = blocks per SM for the BHWQC kernels, QUAN_IQBN_U = rows in flight.
static int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}

template <typename T>
static bool plan_b(int B, int C, int H, int W, int unroll, int blocks_per_sm, int threads, LaunchB& p) {
  p.V = largest_pow2_divisor(C, VecTraits<T>::kMaxVec);
  const int colvecs = 4 * C / p.V;
  int cg = (colvecs + 255) / 256;                   // column groups (wide rows only)
  while (colvecs % cg) ++cg;
  const int cvpg = colvecs / cg;
  if (cvpg > 256 || cg > 65535) return false;
  const int rpb = threads / cvpg;
  p.g.R = (int64_t)B * H * W;
  p.g.L = 4 * C;
  p.g.C = C;
  p.g.cvpg = cvpg;
  p.g.rpb = rpb;
  p.g.rev = 0;
  p.block = dim3(rpb * cvpg);
  int64_t want = ceil_div64(p.g.R, (int64_t)rpb * unroll);
  int64_t cap = ceil_div64((int64_t)QUAN_NUM_SMS * blocks_per_sm, cg);
  if (want < 1) want = 1;
  if (cap < 1) cap = 1;
  p.grid = dim3((unsigned)(want < cap ? want : cap), (unsigned)cg);
  return true;
}

struct LaunchA {
  GeomA g;
  dim3 grid, block;
  int V;
};
template <typename T>
static void plan_a(int B, int C, int H, int W, LaunchA& p) {
  int64_t plane = (int64_t)H * W * 4;
  int V = VecTraits<T>::kMaxVec;       // 4 (fp32) or 8 (bf16)
  if (plane % V) V = 4;                // odd H*W in bf16: 8-byte vectors
  p.V = V;
  p.g.B = B;
  p.g.C = C;
  p.g.plane = plane;
  p.g.vpp = plane / V;
  p.block = dim3(256);
  int64_t per_c = ceil_div64((int64_t)B * p.g.vpp, 256 * 2);
  int64_t cap = ceil_div64((int64_t)QUAN_NUM_SMS * 8, C);
  if (per_c < 1) per_c = 1;
  if (cap < 1) cap = 1;
  p.grid = dim3((unsigned)(per_c < cap ? per_c : cap), C);
}

#define QUAN_DISPATCH_V(V, ...)                    \
  switch (V) {                                     \
    case 8: { constexpr int kV = 8; __VA_ARGS__; } break; \
    case 4: { constexpr int kV = 4; __VA_ARGS__; } break; \
    case 2: { constexpr int kV = 2; __VA_ARGS__; } break; \
    default: { constexpr int kV = 1; __VA_ARGS__; } break; \
  }

// cooperative (fused fold) launch when the whole grid is co-resident, else the plain kernel (the caller adds the fold)
template <typename T, int V, int MODE, int ACT, int U>
static int launch_reduce_b(dim3 grid, dim3 block, cudaStream_t st, const T* xp, const T* dyp, GeomB g, const float* gamma,
                           const float* beta, IqbnWs ws, TailArgs tail, bool* fused) {
  static thread_local int blocks_per_sm = -1, sms = 0;
  auto coop = iqbn_reduce_b<T, V, MODE, ACT, U, true>;
  if (blocks_per_sm < 0) {
    int dev = 0;
    QUAN_CUDA(cudaGetDevice(&dev));
    QUAN_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    int supported = 0;
    QUAN_CUDA(cudaDeviceGetAttribute(&supported, cudaDevAttrCooperativeLaunch, dev));
    int nb = 0;
    QUAN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, coop, IQBN_RED_THREADS, 0));
    blocks_per_sm = supported ? nb : 0;
    // measured on B200 (C=256, 134 MB): one cooperative launch 46.9 us vs reduce + fold kernels 41.9 us — a cooperative
    // launch cannot overlap the previous kernel's tail and its grid barrier costs more than the second launch: opt-in only
    if (env_int("QUAN_IQBN_FUSE", 0) == 0) blocks_per_sm = 0;
  }
  if ((int64_t)grid.x * grid.y <= (int64_t)blocks_per_sm * sms) {
    void* args[] = {(void*)&xp, (void*)&dyp, (void*)&g, (void*)&gamma, (void*)&beta, (void*)&ws, (void*)&tail};
    QUAN_CUDA(cudaLaunchCooperativeKernel((const void*)coop, grid, block, args, 0, st));
    count_launch();
    *fused = true;
    return QUAN_OK;
  }
  QUAN_LAUNCH((iqbn_reduce_b<T, V, MODE, ACT, U, false>), grid, block, 0, st, xp, dyp, g, gamma, beta, ws, tail);
  *fused = false;
  return QUAN_OK;
}

template <typename T, int MODE, int ACT>
static int launch_reduce(const void* x, const void* dy, int B, int C, int H, int W, int layout,
                         const float* gamma, const float* beta, IqbnWs ws, TailArgs tail, cudaStream_t st) {
  const T* xp = reinterpret_cast<const T*>(x);
  const T* dyp = reinterpret_cast<const T*>(dy);
  int nparts = 1;
  if (layout == QUAN_LAYOUT_BHWQC && C > 1) {
    LaunchB p;
    {
      // small tensors: one launch (iqbn_reduce_small), no fold kernel.  QUAN_IQBN_SMALL_MB = 0 disables.
      const int Vs = largest_pow2_divisor(C, VecTraits<T>::kMaxVec);
      const int colvecs = 4 * C / Vs;
      const int64_t R = (int64_t)B * H * W;
      const double mb = (double)R * 4 * C * sizeof(T) * (MODE == 1 ? 2 : 1) / 1e6;
      static const int small_mb = env_int("QUAN_IQBN_SMALL_MB", 40);
      if (mb <= small_mb && colvecs <= 256 && Vs * sizeof(T) >= 8) {
        const int rpb = 256 / colvecs;
        int64_t blocks = ceil_div64(R, (int64_t)rpb * 4);
        static const int small_bps = env_int("QUAN_IQBN_SMALL_BPS", 1);
        int64_t cap = (int64_t)QUAN_NUM_SMS * small_bps;
        while (cap > QUAN_NUM_SMS && cap * 8 * C > 65536) cap -= QUAN_NUM_SMS;
        if (blocks > cap) blocks = cap;
        if (blocks * 8 * C <= 65536) {
          const size_t smem = (size_t)256 * 2 * Vs * sizeof(double);
          QUAN_TIMED(st);
#define QUAN_REDUCE_S(VV) QUAN_LAUNCH((iqbn_reduce_small<T, VV, MODE, ACT>), (unsigned)blocks, 256, smem, st, xp, dyp, R, C, colvecs, rpb, ws, tail)
          if (Vs * sizeof(T) == 16) { if constexpr (sizeof(T) == 2) QUAN_REDUCE_S(8); else QUAN_REDUCE_S(4); }
          else { if constexpr (sizeof(T) == 2) QUAN_REDUCE_S(4); else QUAN_REDUCE_S(2); }
#undef QUAN_REDUCE_S
          QUAN_CHECK_LAUNCH(MODE == 0 ? "iqbn_reduce_fwd_small" : "iqbn_reduce_bwd_small");
          return QUAN_OK;
        }
      }
    }
    {
      // TMA-fed variant: whole rows per block (<= 256 column vectors), 16-byte multiples
      const int Vt = largest_pow2_divisor(C, VecTraits<T>::kMaxVec);
      const int colvecs = 4 * C / Vt;
      const size_t row_bytes = (size_t)4 * C * sizeof(T);
      if (env_int("QUAN_IQBN_TMA", 1) && colvecs <= IQBN_TMA_CONSUMERS && row_bytes % 16 == 0 && Vt * sizeof(T) >= 8) {
        GeomT g;
        g.R = (int64_t)B * H * W; g.L = 4 * C; g.C = C; g.cvpg = colvecs; g.rpb = IQBN_TMA_CONSUMERS / colvecs;
        const size_t target = MODE == 1 ? 8 * 1024 : 16 * 1024;
        int tr = (int)(target / row_bytes) / g.rpb * g.rpb;
        if (tr < g.rpb) tr = g.rpb;
        g.tile_rows = tr;
        g.ntiles = (int)ceil_div64(g.R, tr);
        const size_t tile_bytes = (size_t)tr * row_bytes * (MODE == 1 ? 2 : 1);
        // ncu (profiles/r01_iqbn_tune5.log): fed by TMA the reduction is issue-bound (62-68 % issue-active, 9.4 thread
        // instructions per element); 3 or 4 smaller blocks per SM measured no better (42.6 / 50.1 vs 40.3 us with fold)
        const int bps = env_int("QUAN_IQBN_TMA_BPS", 2);
        // ring: 96 KB per block for the tensors of the wide layers; 56 KB below 300 MB, where the kernel runs inside the narrow models'
        // step beside a deferred wgrad (100 KB ring) — two blocks + that CTA fit an SM.  Measured on the QUAN-YOLO11n
        // step (two alternating runs each): 96 KB 10.87 / 10.94 ms, 56 KB 10.73 / 10.74 ms.
        const double stream_mb = (double)g.R * row_bytes * (MODE == 1 ? 2 : 1) / 1e6;
        const size_t ring = (size_t)env_int("QUAN_IQBN_TMA_KB", bps >= 3 ? 64 : stream_mb <= 300.0 ? 56 : 96) * 1024;
        int stages = (int)(ring / tile_bytes);
        if (stages > 8) stages = 8;
        if (stages >= 2) {
          g.stages = stages;
          // the block fold re-uses the ring: [256 consumers][2 x 8 column elements] doubles = 32 KB at most
          const size_t fold_bytes = (size_t)IQBN_TMA_CONSUMERS * 16 * sizeof(double);
          const size_t smem = (stages * tile_bytes > fold_bytes ? stages * tile_bytes : fold_bytes) + 2 * stages * sizeof(uint64_t) + 128;
          int grid = g.ntiles < bps * QUAN_NUM_SMS ? g.ntiles : bps * QUAN_NUM_SMS;
          nparts = grid;
          QUAN_TIMED(st);
#define QUAN_REDUCE_T(VV) { auto kern = iqbn_reduce_tma<T, VV, MODE, ACT>;                                                  \
            static thread_local DeviceOnce attr;                                                                        \
            if (attr.first()) { QUAN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024)); } \
            QUAN_LAUNCH((kern), grid, IQBN_TMA_THREADS, smem, st, xp, dyp, g, ws, tail); }
          if (Vt * sizeof(T) == 16) { if constexpr (sizeof(T) == 2) QUAN_REDUCE_T(8) else QUAN_REDUCE_T(4) }
          else { if constexpr (sizeof(T) == 2) QUAN_REDUCE_T(4) else QUAN_REDUCE_T(2) }
#undef QUAN_REDUCE_T
          QUAN_CHECK_LAUNCH(MODE == 0 ? "iqbn_reduce_fwd" : "iqbn_reduce_bwd");
          QUAN_TIMED(st);
          QUAN_LAUNCH((iqbn_fold_kernel), (4 * C + FOLD_ACC - 1) / FOLD_ACC, FOLD_ACC * FOLD_LANES, 0, st, ws.part, nparts, tail);
          QUAN_CHECK_LAUNCH("iqbn_fold");
          return QUAN_OK;
        }
      }
    }
    // 2 blocks of 512 threads per SM, U independent 16-byte loads per stream and thread in flight (>= 64 KB per SM)
    const int U = env_int("QUAN_IQBN_RU", MODE == 0 ? 4 : 2);
    const int bps = env_int("QUAN_IQBN_RBPS", 2);
    if (!plan_b<T>(B, C, H, W, U, bps, IQBN_RED_THREADS, p)) {
      set_error("iqbn: C=%d too large for the BHWQC kernels", C);
      return QUAN_E_UNSUPPORTED;
    }
    if (p.grid.x > (unsigned)IQBN_MAX_PARTS) p.grid.x = IQBN_MAX_PARTS;
    nparts = (int)p.grid.x;
    p.g.rev = env_int("QUAN_IQBN_REV", 1);
    bool fused = false;
    QUAN_TIMED(st);
#define QUAN_REDUCE_B(UU) QUAN_DISPATCH_V(p.V, { int rc_ = launch_reduce_b<T, (sizeof(T) == 4 && kV == 8) ? 4 : kV, MODE, ACT, UU>( \
                          p.grid, p.block, st, xp, dyp, p.g, gamma, beta, ws, tail, &fused); if (rc_) return rc_; })
    switch (U) {
      case 1: QUAN_REDUCE_B(1); break;
      case 2: QUAN_REDUCE_B(2); break;
      default: QUAN_REDUCE_B(4); break;
    }
#undef QUAN_REDUCE_B
    if (fused) return QUAN_OK;
  } else {
    LaunchA p;
    plan_a<T>(B, C, H, W, p);
    if (p.grid.x > (unsigned)IQBN_MAX_PARTS) p.grid.x = IQBN_MAX_PARTS;
    nparts = (int)p.grid.x;
    if (p.V == 8) {
      if constexpr (sizeof(T) == 2)
        QUAN_LAUNCH((iqbn_reduce_a<T, 8, MODE, ACT>), p.grid, p.block, 0, st, xp, dyp, p.g, gamma, beta, ws, tail);
    } else {
      QUAN_LAUNCH((iqbn_reduce_a<T, 4, MODE, ACT>), p.grid, p.block, 0, st, xp, dyp, p.g, gamma, beta, ws, tail);
    }
  }
  QUAN_CHECK_LAUNCH(MODE == 0 ? "iqbn_reduce_fwd" : "iqbn_reduce_bwd");
  QUAN_TIMED(st);
  QUAN_LAUNCH((iqbn_fold_kernel), (4 * C + FOLD_ACC - 1) / FOLD_ACC, FOLD_ACC * FOLD_LANES, 0, st, ws.part, nparts, tail);
  QUAN_CHECK_LAUNCH("iqbn_fold");
  return QUAN_OK;
}

template <typename T, int ACT, bool BWD>
static int launch_apply(const void* x, const void* dy, void* out, int B, int C, int H, int W, int layout,
                        const ApplyArgs& a, const float* mix_t, cudaStream_t st) {
  const T* xp = reinterpret_cast<const T*>(x);
  const T* dyp = reinterpret_cast<const T*>(dy);
  T* op = reinterpret_cast<T*>(out);
  if (layout == QUAN_LAYOUT_BHWQC && C > 1) {
    LaunchB p;
    if constexpr (BWD) {
      // train-mode backward that also emits G = M^T dx (fused Conv block): one kernel when a row fits one block
      const int Vw = largest_pow2_divisor(C, VecTraits<T>::kMaxVec);
      const int cpq = C / Vw;
      int gl = 1;
      while (gl < 8 && cpq % (gl * 2) == 0) gl *= 2;
      const int nb = cpq / gl, spb = 256 / (4 * gl);
      const size_t row_bytes_t = (size_t)4 * C * sizeof(T);
      if (env_int("QUAN_IQBN_TMA", 1) && a.stats != nullptr && a.coefT != nullptr && nb <= spb && row_bytes_t % 16 == 0 &&
          Vw * sizeof(T) == 16) {
        // TMA-fed variant (with or without the fused mix)
        GeomTW g;
        g.R = (int64_t)B * H * W; g.L = 4 * C; g.C = C; g.cpq = cpq; g.gl = gl; g.nb = nb; g.rpb = spb / nb;
        int tr = (int)((8 * 1024) / row_bytes_t) / g.rpb * g.rpb;
        if (tr < g.rpb) tr = g.rpb;
        g.tile_rows = tr;
        g.ntiles = (int)ceil_div64(g.R, tr);
        const size_t tile_bytes = (size_t)tr * row_bytes_t * 2;
        int stages = (int)(((size_t)env_int("QUAN_IQBN_APPLY_TMA_KB", 96) * 1024) / tile_bytes);
        if (stages > 8) stages = 8;
        if (stages >= 2) {
          g.stages = stages;
          const size_t smem = stages * tile_bytes + 2 * stages * sizeof(uint64_t) + 128;
          const int grid = g.ntiles < 2 * QUAN_NUM_SMS ? g.ntiles : 2 * QUAN_NUM_SMS;
          const Mix16 mt = mix_t != nullptr ? make_mix(mix_t) : Mix16{};
          constexpr int VT = 16 / (int)sizeof(T);
          QUAN_TIMED(st);
          if (mix_t != nullptr) {
            auto kern = iqbn_apply_bwd_tma<T, VT, ACT, true>;
            static thread_local DeviceOnce attr;
            if (attr.first()) { QUAN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024)); }
            QUAN_LAUNCH((kern), grid, IQBN_TMA_THREADS, smem, st, xp, dyp, op, g, a, mt);
          } else {
            auto kern = iqbn_apply_bwd_tma<T, VT, ACT, false>;
            static thread_local DeviceOnce attr;
            if (attr.first()) { QUAN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024)); }
            QUAN_LAUNCH((kern), grid, IQBN_TMA_THREADS, smem, st, xp, dyp, op, g, a, mt);
          }
          QUAN_CHECK_LAUNCH(mix_t != nullptr ? "iqbn_apply_bwd_mix" : "iqbn_apply_bwd");
          return QUAN_OK;
        }
      }
      if (mix_t != nullptr && a.stats != nullptr && a.coefT != nullptr && nb <= spb) {
        GeomW gw;
        gw.R = (int64_t)B * H * W; gw.L = 4 * C; gw.C = C; gw.cpq = cpq; gw.gl = gl; gw.nb = nb; gw.rpb = spb / nb;
        const Mix16 mt = make_mix(mix_t);
        const int bps = env_int("QUAN_IQBN_MBPS", 2);
        int64_t want = ceil_div64(gw.R, (int64_t)gw.rpb * 2);
        const int64_t cap = (int64_t)QUAN_NUM_SMS * bps;
        if (want > cap) want = cap;
        if (want < 1) want = 1;
        QUAN_TIMED(st);
        QUAN_DISPATCH_V(Vw, QUAN_LAUNCH((iqbn_apply_bwd_mix_b<T, (sizeof(T) == 4 && kV == 8) ? 4 : kV, ACT, 2>), (unsigned)want, 256, 0, st,
                                        xp, dyp, op, gw, a, mt));
        QUAN_CHECK_LAUNCH("iqbn_apply_bwd_mix");
        return QUAN_OK;
      }
    }
    // 4 blocks of 256 threads per SM; U rows (16-byte vectors) per stream in flight per thread
    const int U = env_int("QUAN_IQBN_U", BWD ? 2 : 4);
    if (!plan_b<T>(B, C, H, W, U, env_int("QUAN_IQBN_BPS", BWD ? 3 : 4), 256, p)) {
      set_error("iqbn: C=%d too large for the BHWQC kernels", C);
      return QUAN_E_UNSUPPORTED;
    }
    QUAN_TIMED(st);
#define QUAN_APPLY_B(UU) QUAN_DISPATCH_V(p.V, QUAN_LAUNCH((iqbn_apply_b<T, (sizeof(T) == 4 && kV == 8) ? 4 : kV, ACT, BWD, UU>), p.grid, \
                                              p.block, 0, st, xp, dyp, op, p.g, a))
    switch (U) {
      case 1: QUAN_APPLY_B(1); break;
      case 2: QUAN_APPLY_B(2); break;
      default: QUAN_APPLY_B(4); break;
    }
#undef QUAN_APPLY_B
    QUAN_CHECK_LAUNCH(BWD ? "iqbn_apply_bwd" : "iqbn_apply_fwd");
    if (BWD && mix_t != nullptr) {  // rows wider than one block: G = M^T dx as a second, in-place pass
      int rc = quan_mix(out, out, B, C, H, W, sizeof(T) == 4 ? QUAN_F32 : QUAN_BF16, layout, mix_t, st);
      if (rc) return rc;
    }
  } else {
    LaunchA p;
    plan_a<T>(B, C, H, W, p);
    Mix16 m = {};
    const bool use_mix = BWD && mix_t != nullptr;
    if (use_mix) m = make_mix(mix_t);
    if (p.V == 8) {
      if constexpr (sizeof(T) == 2) {
        if (use_mix) QUAN_LAUNCH((iqbn_apply_a<T, 8, ACT, BWD, BWD>), p.grid, p.block, 0, st, xp, dyp, op, p.g, a, m);
        else QUAN_LAUNCH((iqbn_apply_a<T, 8, ACT, BWD, false>), p.grid, p.block, 0, st, xp, dyp, op, p.g, a, m);
      }
    } else {
      if (use_mix) QUAN_LAUNCH((iqbn_apply_a<T, 4, ACT, BWD, BWD>), p.grid, p.block, 0, st, xp, dyp, op, p.g, a, m);
      else QUAN_LAUNCH((iqbn_apply_a<T, 4, ACT, BWD, false>), p.grid, p.block, 0, st, xp, dyp, op, p.g, a, m);
    }
    QUAN_CHECK_LAUNCH("iqbn_apply_a");
  }
  return QUAN_OK;
}

static int check_common(const void* x, int B, int C, int H, int W, int dtype, int layout) {
  QUAN_REQUIRE(x != nullptr, QUAN_E_ARG, "iqbn: null tensor pointer");
  QUAN_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, QUAN_E_ARG, "iqbn: non-positive dims B=%d C=%d H=%d W=%d", B, C, H, W);
  QUAN_REQUIRE(dtype == QUAN_F32 || dtype == QUAN_BF16, QUAN_E_ARG, "iqbn: bad dtype %d", dtype);
  QUAN_REQUIRE(layout == QUAN_LAYOUT_BCHWQ || layout == QUAN_LAYOUT_BHWQC, QUAN_E_ARG, "iqbn: bad layout %d", layout);
  return QUAN_OK;
}

}  // namespace quan

using namespace quan;

extern "C" {

size_t quan_iqbn_workspace_bytes(int32_t C) { return ((size_t)IQBN_MAX_PARTS * 8 * C + 8 * (size_t)C + 2) * sizeof(double); }

static int reduce_entry(int mode, const void* x, const void* dy, int B, int C, int H, int W, int dtype, int layout,
                        const float* gamma, const float* beta, int act, TailArgs tail, void* workspace,
                        size_t ws_bytes, void* stream) {
  int rc = check_common(x, B, C, H, W, dtype, layout);
  if (rc) return rc;
  QUAN_REQUIRE(workspace != nullptr && ws_bytes >= quan_iqbn_workspace_bytes(C), QUAN_E_WORKSPACE,
               "iqbn: workspace needs %zu bytes, got %zu", quan_iqbn_workspace_bytes(C), ws_bytes);
  IqbnWs ws = carve_ws(workspace, C);
  cudaStream_t st = (cudaStream_t)stream;
  tail.C = C;
  if (mode == 0) {
    if (dtype == QUAN_F32) return launch_reduce<float, 0, QUAN_ACT_NONE>(x, nullptr, B, C, H, W, layout, nullptr, nullptr, ws, tail, st);
    return launch_reduce<__nv_bfloat16, 0, QUAN_ACT_NONE>(x, nullptr, B, C, H, W, layout, nullptr, nullptr, ws, tail, st);
  }
  QUAN_REQUIRE(act == QUAN_ACT_NONE || act == QUAN_ACT_SILU, QUAN_E_ARG, "iqbn: bad act %d", act);
  if (dtype == QUAN_F32) {
    if (act == QUAN_ACT_SILU) return launch_reduce<float, 1, QUAN_ACT_SILU>(x, dy, B, C, H, W, layout, gamma, beta, ws, tail, st);
    return launch_reduce<float, 1, QUAN_ACT_NONE>(x, dy, B, C, H, W, layout, gamma, beta, ws, tail, st);
  }
  if (act == QUAN_ACT_SILU) return launch_reduce<__nv_bfloat16, 1, QUAN_ACT_SILU>(x, dy, B, C, H, W, layout, gamma, beta, ws, tail, st);
  return launch_reduce<__nv_bfloat16, 1, QUAN_ACT_NONE>(x, dy, B, C, H, W, layout, gamma, beta, ws, tail, st);
}

int quan_iqbn_finalize_partials(const void* workspace, int32_t nparts, double count, int32_t C, const float* gamma,
                                const float* beta, float eps, float momentum, float* running_mean, float* running_var,
                                float* stats, void* stream) {
  QUAN_REQUIRE(workspace != nullptr && stats != nullptr && gamma != nullptr && beta != nullptr, QUAN_E_ARG,
               "iqbn_finalize_partials: null pointer");
  QUAN_REQUIRE(nparts > 0 && nparts <= IQBN_MAX_PARTS && C > 0 && count > 0, QUAN_E_ARG,
               "iqbn_finalize_partials: bad nparts=%d / C=%d / count", nparts, C);
  QUAN_REQUIRE((running_mean == nullptr) == (running_var == nullptr), QUAN_E_ARG,
               "iqbn_finalize_partials: running_mean/var must both be given or both NULL");
  TailArgs t = {};
  t.mode = TAIL_FWD_STATS;
  t.C = C;
  t.count = count;
  t.eps = eps;
  t.momentum = momentum;
  t.running_mean = running_mean;
  t.running_var = running_var;
  t.stats = stats;
  t.gamma = gamma;
  t.beta = beta;
  cudaStream_t st = (cudaStream_t)stream;
  QUAN_TIMED(st);
  QUAN_LAUNCH((iqbn_fold_kernel), (4 * C + FOLD_ACC - 1) / FOLD_ACC, FOLD_ACC * FOLD_LANES, 0, st, reinterpret_cast<const double*>(workspace), nparts, t);
  QUAN_CHECK_LAUNCH("iqbn_fold");
  return QUAN_OK;
}

// algorithmic bytes of an IQBN pass for the kernel-timing table: `passes` streams of the activation tensor
static inline void announce_iqbn_work(int32_t B, int32_t C, int32_t H, int32_t W, int dtype, int passes) {
  quan::timing_work("iqbn_", "iqbn_fold", (double)passes * 4.0 * B * C * (double)H * W * (dtype == QUAN_BF16 ? 2.0 : 4.0), 0.0);
}

int quan_iqbn_train_stats(const void* x, int32_t B, int32_t C, int32_t H, int32_t W, int dtype, int layout,
                          const float* gamma, const float* beta, float eps, float momentum, float* running_mean,
                          float* running_var, float* stats, void* workspace, size_t ws_bytes, void* stream) {
  QUAN_REQUIRE(stats != nullptr && gamma != nullptr && beta != nullptr, QUAN_E_ARG, "iqbn_train_stats: null pointer");
  QUAN_REQUIRE((running_mean == nullptr) == (running_var == nullptr), QUAN_E_ARG,
               "iqbn_train_stats: running_mean/var must both be given or both NULL");
  TailArgs t = {};
  t.mode = TAIL_FWD_STATS;
  t.count = (double)B * H * W;
  t.eps = eps;
  t.momentum = momentum;
  t.running_mean = running_mean;
  t.running_var = running_var;
  t.stats = stats;
  t.gamma = gamma;
  t.beta = beta;
  announce_iqbn_work(B, C, H, W, dtype, 1);
  return reduce_entry(0, x, nullptr, B, C, H, W, dtype, layout, nullptr, nullptr, 0, t, workspace, ws_bytes, stream);
}

int quan_iqbn_partial_sums(const void* x, int32_t B, int32_t C, int32_t H, int32_t W, int dtype, int layout,
                           double* sums, void* workspace, size_t ws_bytes, void* stream) {
  QUAN_REQUIRE(sums != nullptr, QUAN_E_ARG, "iqbn_partial_sums: null sums");
  TailArgs t = {};
  t.mode = TAIL_RAW_SUMS;
  t.sums_out = sums;
  return reduce_entry(0, x, nullptr, B, C, H, W, dtype, layout, nullptr, nullptr, 0, t, workspace, ws_bytes, stream);
}

namespace quan {
// mode 0: finalize forward stats from all-reduced raw sums; 1: eval-mode table from running stats;
// 2: backward coefficient table from all-reduced backward sums
__global__ void iqbn_small_kernel(int mode, const double* __restrict__ sums, const float* __restrict__ rm,
                                  const float* __restrict__ rv, TailArgs t) {
  pdl_prologue();
  const int n = 4 * t.C;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    if (mode == 0) {
      write_fwd_stats(t, i, sums[i], sums[n + i]);
    } else if (mode == 1) {
      const float mean = rm[i], var = rv[i];
      const float rstd = 1.0f / sqrtf(var + t.eps);            // conv.py:550 (no +1e-8 in eval)
      const int c = i >> 2, q = i & 3;
      const float scale = t.gamma[i] * rstd;
      t.stats[i] = mean;
      t.stats[n + i] = var;
      t.stats[2 * n + i] = rstd;
      t.stats[3 * n + q * t.C + c] = scale;
      t.stats[4 * n + q * t.C + c] = t.beta[i] - mean * scale;
    } else {
      float* coef = reinterpret_cast<float*>(t.sums_out + 2 * n);
      const int c = i >> 2, q = i & 3, j = q * t.C + c;
      float k1, k2, k3;
      bwd_coefficients(t.gamma[i], t.stats[i], t.stats[2 * n + i], sums[i], sums[n + i], t.count, k1, k2, k3);
      coef[j] = k1;
      coef[n + j] = k2;
      coef[2 * n + j] = k3;
    }
  }
}
static int launch_small(int mode, const double* sums, const float* rm, const float* rv, const TailArgs& t, void* stream) {
  int threads = 128, blocks = (4 * t.C + threads - 1) / threads;
  QUAN_LAUNCH((iqbn_small_kernel), blocks, threads, 0, (cudaStream_t)stream, mode, sums, rm, rv, t);
  QUAN_CHECK_LAUNCH("iqbn_small_kernel");
  return QUAN_OK;
}
}  // namespace quan

int quan_iqbn_finalize_stats(const double* sums, double count, int32_t C, const float* gamma, const float* beta, float eps,
                             float momentum, float* running_mean, float* running_var, float* stats, void* stream) {
  QUAN_REQUIRE(sums != nullptr && stats != nullptr && gamma != nullptr && beta != nullptr && C > 0 && count > 0, QUAN_E_ARG,
               "iqbn_finalize_stats: bad args");
  TailArgs t = {};
  t.C = C; t.count = count; t.eps = eps; t.momentum = momentum;
  t.running_mean = running_mean; t.running_var = running_var; t.stats = stats; t.gamma = gamma; t.beta = beta;
  return launch_small(0, sums, nullptr, nullptr, t, stream);
}

int quan_iqbn_eval_stats(const float* gamma, const float* beta, const float* running_mean, const float* running_var,
                         float eps, int32_t C, float* stats, void* stream) {
  QUAN_REQUIRE(gamma != nullptr && beta != nullptr && running_mean != nullptr && running_var != nullptr && stats != nullptr && C > 0,
               QUAN_E_ARG, "iqbn_eval_stats: bad args");
  TailArgs t = {};
  t.C = C; t.eps = eps; t.stats = stats; t.gamma = gamma; t.beta = beta;
  return launch_small(1, nullptr, running_mean, running_var, t, stream);
}

int quan_iqbn_bwd_coef(double* sums, double count, int32_t C, const float* stats, const float* gamma, void* stream) {
  QUAN_REQUIRE(sums != nullptr && stats != nullptr && gamma != nullptr && C > 0 && count > 0, QUAN_E_ARG,
               "iqbn_bwd_coef: bad args");
  TailArgs t = {};
  t.C = C; t.count = count; t.stats = const_cast<float*>(stats); t.gamma = gamma; t.sums_out = sums;
  return launch_small(2, sums, nullptr, nullptr, t, stream);
}

static int apply_entry(bool bwd, const void* x, const void* dy, void* out, int B, int C, int H, int W, int dtype,
                       int layout, const ApplyArgs& a, int act, const float* mix_t, void* stream) {
  int rc = check_common(x, B, C, H, W, dtype, layout);
  if (rc) return rc;
  QUAN_REQUIRE(out != nullptr && a.gamma != nullptr && a.beta != nullptr, QUAN_E_ARG, "iqbn apply: null pointer");
  QUAN_REQUIRE(act == QUAN_ACT_NONE || act == QUAN_ACT_SILU, QUAN_E_ARG, "iqbn: bad act %d", act);
  cudaStream_t st = (cudaStream_t)stream;
#define QUAN_APPLY_CASE(T, ACT)                                                                           \
  return bwd ? launch_apply<T, ACT, true>(x, dy, out, B, C, H, W, layout, a, mix_t, st)                    \
             : launch_apply<T, ACT, false>(x, dy, out, B, C, H, W, layout, a, mix_t, st)
  if (dtype == QUAN_F32) {
    if (act == QUAN_ACT_SILU) { QUAN_APPLY_CASE(float, QUAN_ACT_SILU); }
    QUAN_APPLY_CASE(float, QUAN_ACT_NONE);
  }
  if (act == QUAN_ACT_SILU) { QUAN_APPLY_CASE(__nv_bfloat16, QUAN_ACT_SILU); }
  QUAN_APPLY_CASE(__nv_bfloat16, QUAN_ACT_NONE);
#undef QUAN_APPLY_CASE
}

int quan_iqbn_apply_fwd(const void* x, void* y, int32_t B, int32_t C, int32_t H, int32_t W, int dtype, int layout,
                        const float* stats, const float* gamma, const float* beta, int act, void* stream) {
  QUAN_REQUIRE(stats != nullptr, QUAN_E_ARG, "iqbn_apply_fwd: null stats");
  ApplyArgs a = {};
  a.gamma = gamma; a.beta = beta; a.stats = stats; a.C = C;
  announce_iqbn_work(B, C, H, W, dtype, 2);
  return apply_entry(false, x, nullptr, y, B, C, H, W, dtype, layout, a, act, nullptr, stream);
}

}  // extern "C"

namespace quan {
// y = act(IQBN(x)) with the batch statistics still as raw sums in the workspace's accumulators (left there by the conv epilogue,
// qconv_tc.cu stat_acc): statistics, table, running buffers and the normalisation in ONE launch (block_api.cu)
int iqbn_apply_fwd_from_acc(const void* x, void* y, int B, int C, int H, int W, int dtype, int layout, void* iqbn_ws, double count,
                            const float* gamma, const float* beta, float eps, float momentum, float* running_mean, float* running_var,
                            float* stats, int act, void* stream) {
  QUAN_REQUIRE(iqbn_ws != nullptr && stats != nullptr && layout == QUAN_LAYOUT_BHWQC && C > 1 && 4 * C <= 512, QUAN_E_ARG,
               "iqbn_apply_fwd_from_acc: bad argument");
  const IqbnWs ws = carve_ws(iqbn_ws, C);
  ApplyArgs a = {};
  a.gamma = gamma; a.beta = beta; a.C = C; a.eps = eps;
  a.fsums = ws.acc; a.ticket = ws.counter + 1; a.fcount = count; a.momentum = momentum;
  a.stats_w = stats; a.running_mean_w = running_mean; a.running_var_w = running_var;
  announce_iqbn_work(B, C, H, W, dtype, 2);
  return apply_entry(false, x, nullptr, y, B, C, H, W, dtype, layout, a, act, nullptr, stream);
}
}  // namespace quan

extern "C" {

int quan_iqbn_eval_fwd(const void* x, void* y, int32_t B, int32_t C, int32_t H, int32_t W, int dtype, int layout,
                       const float* gamma, const float* beta, const float* running_mean, const float* running_var,
                       float eps, int act, void* stream) {
  QUAN_REQUIRE(running_mean != nullptr && running_var != nullptr, QUAN_E_ARG, "iqbn_eval_fwd: null running stats");
  ApplyArgs a = {};
  a.gamma = gamma; a.beta = beta; a.running_mean = running_mean; a.running_var = running_var; a.eps = eps; a.C = C;
  return apply_entry(false, x, nullptr, y, B, C, H, W, dtype, layout, a, act, nullptr, stream);
}

int quan_iqbn_bwd_reduce(const void* dy, const void* x, int32_t B, int32_t C, int32_t H, int32_t W, int dtype,
                         int layout, const float* stats, const float* gamma, const float* beta, int act, double count,
                         double* sums, void* workspace, size_t ws_bytes, void* stream) {
  QUAN_REQUIRE(dy != nullptr && stats != nullptr && sums != nullptr && gamma != nullptr && beta != nullptr,
               QUAN_E_ARG, "iqbn_bwd_reduce: null pointer");
  TailArgs t = {};
  t.mode = TAIL_BWD_SUMS;
  t.stats = const_cast<float*>(stats);
  t.sums_out = sums;
  t.count = count;
  t.gamma = gamma;
  t.beta = beta;
  announce_iqbn_work(B, C, H, W, dtype, 2);
  return reduce_entry(1, x, dy, B, C, H, W, dtype, layout, gamma, beta, act, t, workspace, ws_bytes, stream);
}

int quan_iqbn_bwd_apply(const void* dy, const void* x, void* dx, int32_t B, int32_t C, int32_t H, int32_t W,
                        int dtype, int layout, const float* stats, const float* gamma, const float* beta, int act,
                        const double* sums, double count, float* dgamma, float* dbeta, const float* mix_t,
                        void* stream) {
  QUAN_REQUIRE(dy != nullptr && stats != nullptr && sums != nullptr && count > 0, QUAN_E_ARG,
               "iqbn_bwd_apply: null pointer / bad count");
  QUAN_REQUIRE((dgamma == nullptr) == (dbeta == nullptr), QUAN_E_ARG, "iqbn_bwd_apply: dgamma/dbeta both or neither");
  ApplyArgs a = {};
  a.gamma = gamma; a.beta = beta; a.stats = stats; a.sums = sums; a.count = count;
  a.dgamma = dgamma; a.dbeta = dbeta; a.C = C;
  a.coefT = reinterpret_cast<const float*>(sums + 8 * (size_t)C);   // written by the bwd-reduce tail or quan_iqbn_bwd_coef
  announce_iqbn_work(B, C, H, W, dtype, 3);
  return apply_entry(true, x, dy, dx, B, C, H, W, dtype, layout, a, act, mix_t, stream);
}

int quan_iqbn_eval_bwd(const void* dy, const void* x, void* dx, int32_t B, int32_t C, int32_t H, int32_t W,
                       int dtype, int layout, const float* gamma, const float* beta, const float* running_mean,
                       const float* running_var, float eps, int act, void* stream) {
  QUAN_REQUIRE(dy != nullptr && running_mean != nullptr && running_var != nullptr, QUAN_E_ARG,
               "iqbn_eval_bwd: null pointer");
  ApplyArgs a = {};
  a.gamma = gamma; a.beta = beta; a.running_mean = running_mean; a.running_var = running_var; a.eps = eps; a.C = C;
  return apply_entry(true, x, dy, dx, B, C, H, W, dtype, layout, a, act, nullptr, stream);
}

}  // extern "C"
