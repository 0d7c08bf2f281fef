// iqbn.cu — IQBN (independent quaternion batch-norm) for sm_100a: bandwidth-bound kernels.
//
// Reference semantics (not code): ultralytics/nn/modules/conv.py:553-571 (training branch, biased variance
// + 1e-8, running-stat update with momentum), :546-552 (eval branch), SiLU from conv.py:789,809.
// Backward is the analytic gradient of that expression (the reference relies on autograd).
//
// Design: every kernel is a grid-stride stream of 16-byte vector loads.  A thread owns a
// fixed "column" (c,q) set for its whole life, so per-channel parameters live in registers and per-channel
// partial sums are private fp32 registers (locally shifted, so cancellation stays local), flushed once per
// thread as fp64 atomics into a [8C] accumulator; the last block to finish (threadfence + counter) turns the
// sums into mean/var/rstd (+ running-stat update) and re-zeroes the workspace, so stats take ONE launch.
//
// Two physical layouts (include/quan_sm100.h): BCHWQ (reference) and BHWQC (channels_last_3d, tensor-core path).
#include "common.cuh"
#include <stdlib.h>

namespace quan {

// workspace: [8C] doubles of accumulators + one unsigned counter (padded to 16 B)
struct IqbnWs {
  double* acc;
  unsigned int* counter;
};
static inline IqbnWs carve_ws(void* ws, int C) {
  IqbnWs w;
  w.acc = reinterpret_cast<double*>(ws);
  w.counter = reinterpret_cast<unsigned int*>(w.acc + 8 * (size_t)C);
  return w;
}

// What the last block does with the accumulated sums.
enum TailMode { TAIL_RAW_SUMS = 0, TAIL_FWD_STATS = 1, TAIL_BWD_SUMS = 2 };

struct TailArgs {
  int mode;
  int C;
  double count;
  float eps, momentum;
  float* running_mean;
  float* running_var;
  float* stats;        // [12C] (fwd: out; bwd: in)
  double* sums_out;    // [8C] (raw sums or bwd sums)
};

// Executed by every thread of the LAST block.  acc[0..4C) = first sum, acc[4C..8C) = second sum.
__device__ void tail_finalize(const TailArgs& t, double* acc) {
  const int n = 4 * t.C;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    // atomics landed in L2; read through L2 (volatile) — L1 may hold nothing for these but be explicit
    double s0 = __ldcg(acc + i), s1 = __ldcg(acc + n + i);
    if (t.mode == TAIL_RAW_SUMS) {
      t.sums_out[i] = s0;
      t.sums_out[n + i] = s1;
    } else if (t.mode == TAIL_FWD_STATS) {
      double mean = s0 / t.count;
      double var = s1 / t.count - mean * mean;
      if (var < 0.0) var = 0.0;
      var += 1e-8;  // conv.py:557
      float rstd = (float)(1.0 / sqrt(var + (double)t.eps));
      t.stats[i] = (float)mean;
      t.stats[n + i] = (float)var;
      t.stats[2 * n + i] = rstd;
      if (t.running_mean != nullptr) {  // conv.py:561-562
        t.running_mean[i] = (1.0f - t.momentum) * t.running_mean[i] + t.momentum * (float)mean;
        t.running_var[i] = (1.0f - t.momentum) * t.running_var[i] + t.momentum * (float)var;
      }
    } else {  // TAIL_BWD_SUMS: s0 = sum dz, s1 = sum dz*x  ->  sum dz*xhat = rstd*(s1 - mean*s0)
      double mean = (double)t.stats[i], rstd = (double)t.stats[2 * n + i];
      t.sums_out[i] = s0;
      t.sums_out[n + i] = rstd * (s1 - mean * s0);
    }
    acc[i] = 0.0;  // leave the workspace zeroed for the next call
    acc[n + i] = 0.0;
  }
}

// Block-level epilogue shared by the reduction kernels: returns true in the last block.
__device__ bool last_block_arrive(unsigned int* counter) {
  __shared__ bool is_last;
  __threadfence();  // make this block's atomics visible before the counter bump
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int total = gridDim.x * gridDim.y;
    unsigned int prev = atomicAdd(counter, 1u);
    is_last = (prev == total - 1);
    if (is_last) *counter = 0u;
  }
  __syncthreads();
  if (is_last) __threadfence();
  return is_last;
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
};

template <int V>
__device__ __forceinline__ void colvec_param_index(int cv, int C, int (&idx)[V]) {
  int col = cv * V;
  int q = col / C, c = col - q * C;
#pragma unroll
  for (int i = 0; i < V; ++i) idx[i] = (c + i) * 4 + q;
}

// MODE 0: sums of x and x^2.  MODE 1: sums of dz and dz*x (dz = dy*act'(x*scale+shift)).
template <typename T, int V, int MODE, int ACT, int U>
__global__ void __launch_bounds__(256) iqbn_reduce_b(const T* __restrict__ x, const T* __restrict__ dy, GeomB g,
                                                     const float* __restrict__ gamma,
                                                     const float* __restrict__ beta, IqbnWs ws, TailArgs tail) {
  const int cvl = threadIdx.x % g.cvpg;
  const int rl = threadIdx.x / g.cvpg;
  const int cv = blockIdx.y * g.cvpg + cvl;
  int pidx[V];
  colvec_param_index<V>(cv, g.C, pidx);

  float scale[V], shift[V];
  if constexpr (MODE == 1 && ACT != QUAN_ACT_NONE) {
    const int n = 4 * g.C;
#pragma unroll
    for (int i = 0; i < V; ++i) {
      float mean = tail.stats[pidx[i]], rstd = tail.stats[2 * n + pidx[i]];
      scale[i] = gamma[pidx[i]] * rstd;
      shift[i] = beta[pidx[i]] - mean * scale[i];
    }
  }

  float s0[V], s1[V], k[V];
  int cnt = 0;
#pragma unroll
  for (int i = 0; i < V; ++i) s0[i] = s1[i] = k[i] = 0.f;

  const int64_t rstride = (int64_t)gridDim.x * g.rpb;
  const int64_t coloff = (int64_t)cv * V;
  for (int64_t r = (int64_t)blockIdx.x * g.rpb + rl; r < g.R; r += rstride * U) {
    float xv[U][V], gv[U][V];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t rr = r + u * rstride;
      if (rr < g.R) {
        load_vec<T, V>(x + rr * g.L + coloff, xv[u]);
        if constexpr (MODE == 1) load_vec<T, V>(dy + rr * g.L + coloff, gv[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t rr = r + u * rstride;
      if (rr < g.R) {
        if constexpr (MODE == 0) {
          if (cnt == 0) {
#pragma unroll
            for (int i = 0; i < V; ++i) k[i] = xv[u][i];  // local shift: keeps fp32 partials well conditioned
          }
#pragma unroll
          for (int i = 0; i < V; ++i) {
            float d = xv[u][i] - k[i];
            s0[i] += d;
            s1[i] = fmaf(d, d, s1[i]);
          }
        } else {
#pragma unroll
          for (int i = 0; i < V; ++i) {
            float dz = gv[u][i];
            if constexpr (ACT != QUAN_ACT_NONE) dz *= act_grad<ACT>(fmaf(xv[u][i], scale[i], shift[i]));
            s0[i] += dz;
            s1[i] = fmaf(dz, xv[u][i], s1[i]);
          }
        }
        ++cnt;
      }
    }
  }

  // thread partials -> fp64 raw sums -> shared memory; row lane 0 of every column vector folds the other lanes and
  // issues the block's single atomic per accumulator
  extern __shared__ double red[];   // [blockDim.x][2V]
  double* mine = red + (size_t)threadIdx.x * (2 * V);
#pragma unroll
  for (int i = 0; i < V; ++i) {
    if constexpr (MODE == 0) {  // un-shift in fp64: sum x = sum d + n k ; sum x^2 = sum d^2 + 2k sum d + n k^2
      double kd = (double)k[i], nd = (double)cnt;
      mine[2 * i] = (double)s0[i] + nd * kd;
      mine[2 * i + 1] = (double)s1[i] + 2.0 * kd * (double)s0[i] + nd * kd * kd;
    } else {
      mine[2 * i] = (double)s0[i];
      mine[2 * i + 1] = (double)s1[i];
    }
  }
  __syncthreads();
  // 2V values per column vector, cvpg column vectors: spread the folding over all threads of the block
  const int nval = g.cvpg * 2 * V;
  for (int e = threadIdx.x; e < nval; e += blockDim.x) {
    const int c_v = e / (2 * V), j = e - c_v * (2 * V);      // column vector, value index (2*i + which)
    double acc = 0.0;
    for (int l = 0; l < g.rpb; ++l) acc += red[(size_t)(l * g.cvpg + c_v) * (2 * V) + j];
    const int col = (blockIdx.y * g.cvpg + c_v) * V + (j >> 1);
    const int q = col / g.C, c = col - q * g.C;
    atomicAdd(ws.acc + (j & 1) * 4 * g.C + c * 4 + q, acc);
  }
  if (last_block_arrive(ws.counter)) tail_finalize(tail, ws.acc);
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
__global__ void __launch_bounds__(256) iqbn_apply_b(const T* __restrict__ x, const T* __restrict__ dy,
                                                    T* __restrict__ out, GeomB g, ApplyArgs a) {
  const int cvl = threadIdx.x % g.cvpg;
  const int rl = threadIdx.x / g.cvpg;
  const int cv = blockIdx.y * g.cvpg + cvl;
  int pidx[V];
  colvec_param_index<V>(cv, g.C, pidx);
  float scale[V], shift[V], k1[V], k2[V], k3[V];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    coeff_fwd(a, pidx[i], scale[i], shift[i]);
    if constexpr (BWD) coeff_bwd(a, pidx[i], k1[i], k2[i], k3[i]);
  }
  if constexpr (BWD) write_param_grads(a);

  const int64_t rstride = (int64_t)gridDim.x * g.rpb;
  const int64_t coloff = (int64_t)cv * V;
  for (int64_t r = (int64_t)blockIdx.x * g.rpb + rl; r < g.R; r += rstride * U) {
    float xv[U][V], gv[U][V];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t rr = r + u * rstride;
      if (rr < g.R) {
        load_vec<T, V>(x + rr * g.L + coloff, xv[u]);
        if constexpr (BWD) load_vec<T, V>(dy + rr * g.L + coloff, gv[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t rr = r + u * rstride;
      if (rr < g.R) {
        float ov[V];
        if constexpr (!BWD) {
#pragma unroll
          for (int i = 0; i < V; ++i) ov[i] = act_fwd<ACT>(fmaf(xv[u][i], scale[i], shift[i]));
        } else {
#pragma unroll
          for (int i = 0; i < V; ++i) {
            float dz = gv[u][i];
            if constexpr (ACT != QUAN_ACT_NONE) dz *= act_grad<ACT>(fmaf(xv[u][i], scale[i], shift[i]));
            ov[i] = fmaf(k1[i], dz, fmaf(k2[i], xv[u][i], k3[i]));
          }
        }
        store_vec<T, V>(out + rr * g.L + coloff, ov);
      }
    }
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
  if ((threadIdx.x & 31) == 0) {
    const int n = 4 * g.C;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      atomicAdd(ws.acc + c * 4 + q, a0[q]);
      atomicAdd(ws.acc + n + c * 4 + q, a1[q]);
    }
  }
  if (last_block_arrive(ws.counter)) tail_finalize(tail, ws.acc);
}

template <typename T, int V, int ACT, bool BWD, bool MIX>
__global__ void __launch_bounds__(256) iqbn_apply_a(const T* __restrict__ x, const T* __restrict__ dy,
                                                    T* __restrict__ out, GeomA g, ApplyArgs a, Mix16 mix) {
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
static bool plan_b(int B, int C, int H, int W, int unroll, int blocks_per_sm, LaunchB& p) {
  blocks_per_sm = env_int("QUAN_IQBN_BPS", blocks_per_sm);
  p.V = largest_pow2_divisor(C, VecTraits<T>::kMaxVec);
  const int colvecs = 4 * C / p.V;
  int cg = (colvecs + 255) / 256;                   // column groups (wide rows only)
  while (colvecs % cg) ++cg;
  const int cvpg = colvecs / cg;
  if (cvpg > 256 || cg > 65535) return false;
  const int rpb = 256 / cvpg;
  p.g.R = (int64_t)B * H * W;
  p.g.L = 4 * C;
  p.g.C = C;
  p.g.cvpg = cvpg;
  p.g.rpb = rpb;
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

template <typename T, int MODE, int ACT>
static int launch_reduce(const void* x, const void* dy, int B, int C, int H, int W, int layout,
                         const float* gamma, const float* beta, IqbnWs ws, TailArgs tail, cudaStream_t st) {
  const T* xp = reinterpret_cast<const T*>(x);
  const T* dyp = reinterpret_cast<const T*>(dy);
  if (layout == QUAN_LAYOUT_BHWQC && C > 1) {
    LaunchB p;
    constexpr int U = MODE == 1 ? 4 : 8;   // rows in flight per thread (two tensors are streamed in MODE 1)
    if (!plan_b<T>(B, C, H, W, U, 2, p)) {
      set_error("iqbn: C=%d too large for the BHWQC kernels", C);
      return QUAN_E_UNSUPPORTED;
    }
    QUAN_DISPATCH_V(p.V, (iqbn_reduce_b<T, (sizeof(T) == 4 && kV == 8) ? 4 : kV, MODE, ACT, U>
                          <<<p.grid, p.block, (size_t)p.block.x * 2 * kV * sizeof(double), st>>>(xp, dyp, p.g, gamma, beta,
                                                                                                  ws, tail)));
  } else {
    LaunchA p;
    plan_a<T>(B, C, H, W, p);
    if (p.V == 8) {
      if constexpr (sizeof(T) == 2)
        iqbn_reduce_a<T, 8, MODE, ACT><<<p.grid, p.block, 0, st>>>(xp, dyp, p.g, gamma, beta, ws, tail);
    } else {
      iqbn_reduce_a<T, 4, MODE, ACT><<<p.grid, p.block, 0, st>>>(xp, dyp, p.g, gamma, beta, ws, tail);
    }
  }
  QUAN_CHECK_LAUNCH("iqbn_reduce");
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
    // measured on B200 (profiles/r01_iqbn_tune.log): the per-thread coefficient prologue makes many light blocks lose to
    // few blocks; U = 1 with 4 (fwd) / 2 (bwd) blocks per SM is the best point of the sweep
    const int U = env_int("QUAN_IQBN_U", 1);
    if (!plan_b<T>(B, C, H, W, U, BWD ? 2 : 4, p)) {
      set_error("iqbn: C=%d too large for the BHWQC kernels", C);
      return QUAN_E_UNSUPPORTED;
    }
#define QUAN_APPLY_B(UU) QUAN_DISPATCH_V(p.V, (iqbn_apply_b<T, (sizeof(T) == 4 && kV == 8) ? 4 : kV, ACT, BWD, UU> \
                          <<<p.grid, p.block, 0, st>>>(xp, dyp, op, p.g, a)))
    switch (U) {
      case 1: QUAN_APPLY_B(1); break;
      case 2: QUAN_APPLY_B(2); break;
      case 8: QUAN_APPLY_B(8); break;
      default: QUAN_APPLY_B(4); break;
    }
#undef QUAN_APPLY_B
    QUAN_CHECK_LAUNCH("iqbn_apply_b");
    if (BWD && mix_t != nullptr) {  // G = M^T dY for the producing QConv2D: second in-place pass in this layout
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
        if (use_mix) iqbn_apply_a<T, 8, ACT, BWD, BWD><<<p.grid, p.block, 0, st>>>(xp, dyp, op, p.g, a, m);
        else iqbn_apply_a<T, 8, ACT, BWD, false><<<p.grid, p.block, 0, st>>>(xp, dyp, op, p.g, a, m);
      }
    } else {
      if (use_mix) iqbn_apply_a<T, 4, ACT, BWD, BWD><<<p.grid, p.block, 0, st>>>(xp, dyp, op, p.g, a, m);
      else iqbn_apply_a<T, 4, ACT, BWD, false><<<p.grid, p.block, 0, st>>>(xp, dyp, op, p.g, a, m);
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

size_t quan_iqbn_workspace_bytes(int32_t C) { return (size_t)8 * C * sizeof(double) + 16; }

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

int quan_iqbn_train_stats(const void* x, int32_t B, int32_t C, int32_t H, int32_t W, int dtype, int layout,
                          float eps, float momentum, float* running_mean, float* running_var, float* stats,
                          void* workspace, size_t ws_bytes, void* stream) {
  QUAN_REQUIRE(stats != nullptr, QUAN_E_ARG, "iqbn_train_stats: null stats");
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
__global__ void iqbn_finalize_kernel(const double* __restrict__ sums, TailArgs t) {
  const int n = 4 * t.C;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    double mean = sums[i] / t.count;
    double var = sums[n + i] / t.count - mean * mean;
    if (var < 0.0) var = 0.0;
    var += 1e-8;
    t.stats[i] = (float)mean;
    t.stats[n + i] = (float)var;
    t.stats[2 * n + i] = (float)(1.0 / sqrt(var + (double)t.eps));
    if (t.running_mean != nullptr) {
      t.running_mean[i] = (1.0f - t.momentum) * t.running_mean[i] + t.momentum * (float)mean;
      t.running_var[i] = (1.0f - t.momentum) * t.running_var[i] + t.momentum * (float)var;
    }
  }
}
}  // namespace quan

int quan_iqbn_finalize_stats(const double* sums, double count, int32_t C, float eps, float momentum,
                             float* running_mean, float* running_var, float* stats, void* stream) {
  QUAN_REQUIRE(sums != nullptr && stats != nullptr && C > 0 && count > 0, QUAN_E_ARG, "iqbn_finalize_stats: bad args");
  TailArgs t = {};
  t.C = C;
  t.count = count;
  t.eps = eps;
  t.momentum = momentum;
  t.running_mean = running_mean;
  t.running_var = running_var;
  t.stats = stats;
  int threads = 128, blocks = (4 * C + threads - 1) / threads;
  iqbn_finalize_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(sums, t);
  QUAN_CHECK_LAUNCH("iqbn_finalize");
  return QUAN_OK;
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
  return apply_entry(false, x, nullptr, y, B, C, H, W, dtype, layout, a, act, nullptr, stream);
}

int quan_iqbn_eval_fwd(const void* x, void* y, int32_t B, int32_t C, int32_t H, int32_t W, int dtype, int layout,
                       const float* gamma, const float* beta, const float* running_mean, const float* running_var,
                       float eps, int act, void* stream) {
  QUAN_REQUIRE(running_mean != nullptr && running_var != nullptr, QUAN_E_ARG, "iqbn_eval_fwd: null running stats");
  ApplyArgs a = {};
  a.gamma = gamma; a.beta = beta; a.running_mean = running_mean; a.running_var = running_var; a.eps = eps; a.C = C;
  return apply_entry(false, x, nullptr, y, B, C, H, W, dtype, layout, a, act, nullptr, stream);
}

int quan_iqbn_bwd_reduce(const void* dy, const void* x, int32_t B, int32_t C, int32_t H, int32_t W, int dtype,
                         int layout, const float* stats, const float* gamma, const float* beta, int act,
                         double* sums, void* workspace, size_t ws_bytes, void* stream) {
  QUAN_REQUIRE(dy != nullptr && stats != nullptr && sums != nullptr && gamma != nullptr && beta != nullptr,
               QUAN_E_ARG, "iqbn_bwd_reduce: null pointer");
  TailArgs t = {};
  t.mode = TAIL_BWD_SUMS;
  t.stats = const_cast<float*>(stats);
  t.sums_out = sums;
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
