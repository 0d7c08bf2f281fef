// qattn.cu — QAttention core: per-quaternion-component multi-head attention over the H*W tokens of a feature map,
// fused (scores, softmax and the value product never touch HBM), forward and backward.
//
// Reference: ultralytics/nn/modules/block.py:1511-1540 (QAttention.forward) — after the qkv QConv2D:
//   q, k, v = split(qkv, [h*K, h*K, h*V], dim=1);  per (batch b, head, component p):  A = softmax(q k^T * K^-0.5) [N x N],  o = A v
// with N = H*W tokens, key_dim K and head_dim V per head.  The reference materialises A ([B, h, 4, N, N]: 537 M elements at
// 16 x 1024^2 input for QUAN-YOLO11n, moved through HBM by two bmm, a softmax, a scale and the autocast copies, forward and backward:
// ~9 ms of a 30 ms step, tools/yolo_step_profile.py).  The heads of the QUAN models are tiny (K = 2, V = 4: the yaml fixes
// heads = C/16), so the work is exp-bound CUDA-core arithmetic, not a tensor-core GEMM: 6 FMAs + one ex2 per (query, key) pair.
//
// Layout: BHWQC only (the library's internal layout).  qkv[b][n][p][c], c in [0, 2hK + hV): q = head*K + kd, k = hK + head*K + kd,
// v = 2hK + head*V + vd;  o[b][n][p][head*V + vd].
//   forward : one thread per query, keys/values of the head staged in shared memory (broadcast reads), online softmax over key tiles;
//             saves lse[b][p][head][n] (log2 domain) for the backward.
//   backward: kernel A (thread per query) -> dq;  kernel B (thread per key, queries staged in shared memory) -> dk, dv.
//             P is recomputed from lse; D_i = sum_v dO_iv O_iv is computed on the fly.
#include "common.cuh"

namespace quan {

constexpr int QA_THREADS = 256;
// keys (fwd, bwd A) / queries (bwd B) staged per shared-memory tile: <= 32 KB of static shared memory for every instantiation
template <int K, int V> struct QaTile { static constexpr int n = (K + V <= 6) ? 1024 : (K + V <= 12) ? 512 : 256; };

template <typename T> __device__ __forceinline__ float qa_ld(const T* p) { return (float)*p; }
template <> __device__ __forceinline__ float qa_ld<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <typename T> __device__ __forceinline__ void qa_st(T* p, float v) { *p = (T)v; }
template <> __device__ __forceinline__ void qa_st<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

struct QaGeom {
  int B, N, heads, Cqkv, Co;   // Cqkv = heads*(2K+V), Co = heads*V
  float scale_log2e;           // K^-0.5 * log2(e)
  float scale;
};

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// blockIdx.y = (b*4 + p)*heads + head; blockIdx.x = query tile
template <typename T, int K, int V>
__global__ void __launch_bounds__(QA_THREADS) qattn_fwd_kernel(const T* __restrict__ qkv, T* __restrict__ o, float* __restrict__ lse,
                                                               QaGeom g) {
  pdl_prologue();
  constexpr int QA_TILE = QaTile<K, V>::n;
  __shared__ float Ks[QA_TILE * K];
  __shared__ float Vs[QA_TILE * V];
  const int head = blockIdx.y % g.heads, bp = blockIdx.y / g.heads, p = bp & 3, b = bp >> 2;
  const int64_t row = (int64_t)4 * g.Cqkv;                         // elements per token
  const T* base = qkv + (int64_t)b * g.N * row + (int64_t)p * g.Cqkv;
  const int qoff = head * K, koff = g.heads * K + head * K, voff = 2 * g.heads * K + head * V;
  const int i = blockIdx.x * QA_THREADS + threadIdx.x;
  const bool live = i < g.N;
  float q[K];
#pragma unroll
  for (int k = 0; k < K; ++k) q[k] = live ? qa_ld(base + (int64_t)i * row + qoff + k) * g.scale_log2e : 0.f;
  float m = -INFINITY, l = 0.f, acc[V];
#pragma unroll
  for (int v = 0; v < V; ++v) acc[v] = 0.f;
  for (int j0 = 0; j0 < g.N; j0 += QA_TILE) {
    const int nt = min(QA_TILE, g.N - j0);
    __syncthreads();
    for (int e = threadIdx.x; e < nt * K; e += QA_THREADS) Ks[e] = qa_ld(base + (int64_t)(j0 + e / K) * row + koff + e % K);
    for (int e = threadIdx.x; e < nt * V; e += QA_THREADS) Vs[e] = qa_ld(base + (int64_t)(j0 + e / V) * row + voff + e % V);
    __syncthreads();
    int j = 0;
    for (; j + 8 <= nt; j += 8) {
      float s[8], mx = m;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < K; ++k) t = fmaf(q[k], Ks[(j + u) * K + k], t);
        s[u] = t;
        mx = fmaxf(mx, t);
      }
      const float corr = fast_exp2(m - mx);
      m = mx;
      l *= corr;
#pragma unroll
      for (int v = 0; v < V; ++v) acc[v] *= corr;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const float pe = fast_exp2(s[u] - mx);
        l += pe;
#pragma unroll
        for (int v = 0; v < V; ++v) acc[v] = fmaf(pe, Vs[(j + u) * V + v], acc[v]);
      }
    }
    for (; j < nt; ++j) {
      float t = 0.f;
#pragma unroll
      for (int k = 0; k < K; ++k) t = fmaf(q[k], Ks[j * K + k], t);
      const float mx = fmaxf(m, t), corr = fast_exp2(m - mx), pe = fast_exp2(t - mx);
      m = mx;
      l = fmaf(l, corr, pe);
#pragma unroll
      for (int v = 0; v < V; ++v) acc[v] = fmaf(pe, Vs[j * V + v], acc[v] * corr);
    }
  }
  if (live) {
    const float inv = 1.f / l;
    T* op = o + ((int64_t)b * g.N + i) * 4 * g.Co + (int64_t)p * g.Co + head * V;
#pragma unroll
    for (int v = 0; v < V; ++v) qa_st(op + v, acc[v] * inv);
    lse[(int64_t)blockIdx.y * g.N + i] = m + log2f(l);
  }
}

// backward A: thread per query -> dq (and nothing else)
template <typename T, int K, int V>
__global__ void __launch_bounds__(QA_THREADS) qattn_bwd_dq_kernel(const T* __restrict__ qkv, const T* __restrict__ o,
                                                                  const T* __restrict__ d_o, const float* __restrict__ lse,
                                                                  T* __restrict__ dqkv, QaGeom g) {
  pdl_prologue();
  constexpr int QA_TILE = QaTile<K, V>::n;
  __shared__ float Ks[QA_TILE * K];
  __shared__ float Vs[QA_TILE * V];
  const int head = blockIdx.y % g.heads, bp = blockIdx.y / g.heads, p = bp & 3, b = bp >> 2;
  const int64_t row = (int64_t)4 * g.Cqkv, orow = (int64_t)4 * g.Co;
  const T* base = qkv + (int64_t)b * g.N * row + (int64_t)p * g.Cqkv;
  const int qoff = head * K, koff = g.heads * K + head * K, voff = 2 * g.heads * K + head * V;
  const int i = blockIdx.x * QA_THREADS + threadIdx.x;
  const bool live = i < g.N;
  float q[K], dq[K], dO[V], D = 0.f, L = 0.f;
#pragma unroll
  for (int k = 0; k < K; ++k) { q[k] = live ? qa_ld(base + (int64_t)i * row + qoff + k) * g.scale_log2e : 0.f; dq[k] = 0.f; }
  if (live) {
    const int64_t oi = ((int64_t)b * g.N + i) * orow + (int64_t)p * g.Co + head * V;
#pragma unroll
    for (int v = 0; v < V; ++v) { dO[v] = qa_ld(d_o + oi + v); D = fmaf(dO[v], qa_ld(o + oi + v), D); }
    L = lse[(int64_t)blockIdx.y * g.N + i];
  } else {
#pragma unroll
    for (int v = 0; v < V; ++v) dO[v] = 0.f;
  }
  for (int j0 = 0; j0 < g.N; j0 += QA_TILE) {
    const int nt = min(QA_TILE, g.N - j0);
    __syncthreads();
    for (int e = threadIdx.x; e < nt * K; e += QA_THREADS) Ks[e] = qa_ld(base + (int64_t)(j0 + e / K) * row + koff + e % K);
    for (int e = threadIdx.x; e < nt * V; e += QA_THREADS) Vs[e] = qa_ld(base + (int64_t)(j0 + e / V) * row + voff + e % V);
    __syncthreads();
#pragma unroll 4
    for (int j = 0; j < nt; ++j) {
      float t = -L, dp = -D;
#pragma unroll
      for (int k = 0; k < K; ++k) t = fmaf(q[k], Ks[j * K + k], t);
#pragma unroll
      for (int v = 0; v < V; ++v) dp = fmaf(dO[v], Vs[j * V + v], dp);
      const float ds = fast_exp2(t) * dp;
#pragma unroll
      for (int k = 0; k < K; ++k) dq[k] = fmaf(ds, Ks[j * K + k], dq[k]);
    }
  }
  if (live) {
    T* dp_ = dqkv + (int64_t)b * g.N * row + (int64_t)p * g.Cqkv + (int64_t)i * row + qoff;
#pragma unroll
    for (int k = 0; k < K; ++k) qa_st(dp_ + k, dq[k] * g.scale);
  }
}

// backward B: thread per key -> dk, dv; queries (q, dO, lse, D) staged in shared memory
template <typename T, int K, int V>
__global__ void __launch_bounds__(QA_THREADS) qattn_bwd_dkv_kernel(const T* __restrict__ qkv, const T* __restrict__ o,
                                                                   const T* __restrict__ d_o, const float* __restrict__ lse,
                                                                   T* __restrict__ dqkv, QaGeom g) {
  pdl_prologue();
  constexpr int R = K + V + 2;
  constexpr int QA_TILE = QaTile<K, V>::n;
  __shared__ float Qs[QA_TILE * R];      // per query: q[K] (pre-scaled), dO[V], lse, D
  const int head = blockIdx.y % g.heads, bp = blockIdx.y / g.heads, p = bp & 3, b = bp >> 2;
  const int64_t row = (int64_t)4 * g.Cqkv, orow = (int64_t)4 * g.Co;
  const T* base = qkv + (int64_t)b * g.N * row + (int64_t)p * g.Cqkv;
  const int qoff = head * K, koff = g.heads * K + head * K, voff = 2 * g.heads * K + head * V;
  const int j = blockIdx.x * QA_THREADS + threadIdx.x;
  const bool live = j < g.N;
  float kk[K], vv[V], dk[K], dv[V];
#pragma unroll
  for (int k = 0; k < K; ++k) { kk[k] = live ? qa_ld(base + (int64_t)j * row + koff + k) : 0.f; dk[k] = 0.f; }
#pragma unroll
  for (int v = 0; v < V; ++v) { vv[v] = live ? qa_ld(base + (int64_t)j * row + voff + v) : 0.f; dv[v] = 0.f; }
  for (int i0 = 0; i0 < g.N; i0 += QA_TILE) {
    const int nt = min(QA_TILE, g.N - i0);
    __syncthreads();
    for (int e = threadIdx.x; e < nt; e += QA_THREADS) {
      const int i = i0 + e;
      float* r = Qs + e * R;
      const int64_t oi = ((int64_t)b * g.N + i) * orow + (int64_t)p * g.Co + head * V;
      float D = 0.f;
#pragma unroll
      for (int k = 0; k < K; ++k) r[k] = qa_ld(base + (int64_t)i * row + qoff + k) * g.scale_log2e;
#pragma unroll
      for (int v = 0; v < V; ++v) { const float d = qa_ld(d_o + oi + v); r[K + v] = d; D = fmaf(d, qa_ld(o + oi + v), D); }
      r[K + V] = lse[(int64_t)blockIdx.y * g.N + i];
      r[K + V + 1] = D;
    }
    __syncthreads();
#pragma unroll 4
    for (int e = 0; e < nt; ++e) {
      const float* r = Qs + e * R;
      float t = -r[K + V], dp = -r[K + V + 1];
#pragma unroll
      for (int k = 0; k < K; ++k) t = fmaf(r[k], kk[k], t);
      const float pe = fast_exp2(t);
#pragma unroll
      for (int v = 0; v < V; ++v) { dp = fmaf(r[K + v], vv[v], dp); dv[v] = fmaf(pe, r[K + v], dv[v]); }
      const float ds = pe * dp;
#pragma unroll
      for (int k = 0; k < K; ++k) dk[k] = fmaf(ds, r[k], dk[k]);      // r[k] carries scale*log2e: undone below
    }
  }
  if (live) {
    T* dbase = dqkv + (int64_t)b * g.N * row + (int64_t)p * g.Cqkv + (int64_t)j * row;
    const float un = g.scale / g.scale_log2e;                            // q was staged as q*scale*log2e: dk = scale * sum ds q
#pragma unroll
    for (int k = 0; k < K; ++k) qa_st(dbase + koff + k, dk[k] * un);
#pragma unroll
    for (int v = 0; v < V; ++v) qa_st(dbase + voff + v, dv[v]);
  }
}

template <typename T, int K, int V>
static int qattn_launch(bool bwd, const void* qkv, const void* o, const void* d_o, float* lse, void* out, const QaGeom& g, cudaStream_t st) {
  dim3 grid((g.N + QA_THREADS - 1) / QA_THREADS, g.B * 4 * g.heads);
  if (!bwd) {
    QUAN_TIMED(st);
    QUAN_LAUNCH((qattn_fwd_kernel<T, K, V>), grid, QA_THREADS, 0, st, (const T*)qkv, (T*)out, lse, g);
    QUAN_CHECK_LAUNCH("qattn_fwd");
  } else {
    QUAN_TIMED(st);
    QUAN_LAUNCH((qattn_bwd_dq_kernel<T, K, V>), grid, QA_THREADS, 0, st, (const T*)qkv, (const T*)o, (const T*)d_o, lse, (T*)out, g);
    QUAN_CHECK_LAUNCH("qattn_bwd_dq");
    QUAN_TIMED(st);
    QUAN_LAUNCH((qattn_bwd_dkv_kernel<T, K, V>), grid, QA_THREADS, 0, st, (const T*)qkv, (const T*)o, (const T*)d_o, lse, (T*)out, g);
    QUAN_CHECK_LAUNCH("qattn_bwd_dkv");
  }
  return QUAN_OK;
}

static int qattn_dispatch(bool bwd, const void* qkv, const void* o, const void* d_o, float* lse, void* out, int B, int H, int W, int heads,
                          int key_dim, int head_dim, float scale, int dtype, int layout, void* stream) {
  QUAN_REQUIRE(qkv && lse && out && (!bwd || (o && d_o)), QUAN_E_ARG, "qattention: null pointer");
  QUAN_REQUIRE(B > 0 && H > 0 && W > 0 && heads > 0, QUAN_E_ARG, "qattention: non-positive size");
  QUAN_REQUIRE(layout == QUAN_LAYOUT_BHWQC, QUAN_E_UNSUPPORTED, "qattention: BHWQC layout only (convert with quan_layout_convert)");
  QUAN_REQUIRE((int64_t)B * 4 * heads <= 65535, QUAN_E_UNSUPPORTED, "qattention: B*4*heads = %lld exceeds the grid limit", (long long)B * 4 * heads);
  QaGeom g;
  g.B = B; g.N = H * W; g.heads = heads; g.Cqkv = heads * (2 * key_dim + head_dim); g.Co = heads * head_dim;
  g.scale = scale; g.scale_log2e = scale * 1.4426950408889634f;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
#define QA_CASE(KK, VV)                                                                                                      \
  if (key_dim == KK && head_dim == VV)                                                                                       \
    return dtype == QUAN_BF16 ? qattn_launch<__nv_bfloat16, KK, VV>(bwd, qkv, o, d_o, lse, out, g, st)                 \
                                    : qattn_launch<float, KK, VV>(bwd, qkv, o, d_o, lse, out, g, st);
  QA_CASE(2, 4) QA_CASE(4, 8) QA_CASE(8, 16) QA_CASE(1, 2)
#undef QA_CASE
  QUAN_REQUIRE(false, QUAN_E_UNSUPPORTED, "qattention: (key_dim, head_dim) = (%d, %d) not instantiated (built: (1,2) (2,4) (4,8) (8,16))",
               key_dim, head_dim);
  return QUAN_OK;
}

}  // namespace quan

extern "C" {

int quan_qattention_fwd(const void* qkv, void* o, float* lse, int32_t B, int32_t H, int32_t W, int32_t heads, int32_t key_dim,
                        int32_t head_dim, float scale, int dtype, int layout, void* stream) {
  return quan::qattn_dispatch(false, qkv, nullptr, nullptr, lse, o, B, H, W, heads, key_dim, head_dim, scale, dtype, layout, stream);
}

int quan_qattention_bwd(const void* qkv, const void* o, const void* d_o, const float* lse, void* dqkv, int32_t B, int32_t H, int32_t W,
                        int32_t heads, int32_t key_dim, int32_t head_dim, float scale, int dtype, int layout, void* stream) {
  return quan::qattn_dispatch(true, qkv, o, d_o, const_cast<float*>(lse), dqkv, B, H, W, heads, key_dim, head_dim, scale, dtype, layout, stream);
}

}  // extern "C"
