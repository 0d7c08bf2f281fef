// obbloss.cu — the differentiable half of v8OBBLoss as ONE kernel: loss terms AND their gradients with respect
// to the head outputs, read and written in the head's own memory layout.
//
// Reference: ultralytics/utils/loss.py:941-1033 (v8OBBLoss.__call__ after the assigner): BCE-with-logits class loss (:995-998),
// RotatedBboxLoss :364-378 (ProbIoU of metrics.py:198-233 + DFLoss :306-329 on the foreground anchors), the quaternion angular loss
// :870-921 / :1008-1025, the decode of :1035-1050 + tal.py:366-385 (dist2rbox), gains :1027-1030, `loss.sum() * batch_size`.
// In the reference (and in loss.OBBLossStatic's torch tail) this is ~350 elementwise / reduction kernels forward + backward over
// [B, A, .] tensors (2.5 ms of a 17 ms QUAN-YOLO11n step) behind a cat / split / permute / contiguous of the head outputs.
//
// One thread per (image, anchor).  The head emits, per level l, a channels-last [B][H_l][W_l][no] tensor (no = 4*reg_max + nc; our QER
// produces exactly this memory) — element (b, a_l, ch) is one contiguous row of `no` values per anchor — and theta as [B][1][A].
// Forward and backward in one pass: the box term is differentiated with forward-mode dual numbers over the five decoded quantities
// (d_l, d_t, d_r, d_b, theta) — ~200 flops per FOREGROUND anchor (a few thousand per step) — and chained analytically through the
// DFL soft-arg-max; class / DFL / angle terms have closed-form gradients.  Sums go to four double atomics; a one-thread tail applies
// the gains.  Gradients are written for d(total)/d(head output) with total = (box + cls + dfl + angle) * B.
#include "common.cuh"
#include <math.h>

namespace quan {

constexpr int OL_THREADS = 128;
constexpr int OL_MAXREG = 16;

struct D5 {            // value + partials w.r.t. (d_l, d_t, d_r, d_b, theta)
  float v, d[5];
};
__device__ __forceinline__ D5 d5c(float c) { D5 r; r.v = c; for (int i = 0; i < 5; ++i) r.d[i] = 0.f; return r; }
__device__ __forceinline__ D5 d5var(float v, int k) { D5 r = d5c(v); r.d[k] = 1.f; return r; }
__device__ __forceinline__ D5 operator+(const D5& a, const D5& b) { D5 r; r.v = a.v + b.v; for (int i = 0; i < 5; ++i) r.d[i] = a.d[i] + b.d[i]; return r; }
__device__ __forceinline__ D5 operator-(const D5& a, const D5& b) { D5 r; r.v = a.v - b.v; for (int i = 0; i < 5; ++i) r.d[i] = a.d[i] - b.d[i]; return r; }
__device__ __forceinline__ D5 operator*(const D5& a, const D5& b) { D5 r; r.v = a.v * b.v; for (int i = 0; i < 5; ++i) r.d[i] = a.d[i] * b.v + a.v * b.d[i]; return r; }
__device__ __forceinline__ D5 operator/(const D5& a, const D5& b) {
  D5 r; const float inv = 1.f / b.v; r.v = a.v * inv;
  for (int i = 0; i < 5; ++i) r.d[i] = (a.d[i] - r.v * b.d[i]) * inv;
  return r;
}
__device__ __forceinline__ D5 operator+(const D5& a, float c) { D5 r = a; r.v += c; return r; }
__device__ __forceinline__ D5 operator-(const D5& a, float c) { D5 r = a; r.v -= c; return r; }
__device__ __forceinline__ D5 operator*(const D5& a, float c) { D5 r; r.v = a.v * c; for (int i = 0; i < 5; ++i) r.d[i] = a.d[i] * c; return r; }
__device__ __forceinline__ D5 d5chain(const D5& a, float v, float dv) { D5 r; r.v = v; for (int i = 0; i < 5; ++i) r.d[i] = a.d[i] * dv; return r; }
__device__ __forceinline__ D5 d5cos(const D5& a) { return d5chain(a, cosf(a.v), -sinf(a.v)); }
__device__ __forceinline__ D5 d5sin(const D5& a) { return d5chain(a, sinf(a.v), cosf(a.v)); }
__device__ __forceinline__ D5 d5sqrt(const D5& a) { const float s = sqrtf(a.v); return d5chain(a, s, s > 0.f ? 0.5f / s : 0.f); }
__device__ __forceinline__ D5 d5log(const D5& a) { return d5chain(a, logf(a.v), 1.f / a.v); }
__device__ __forceinline__ D5 d5exp(const D5& a) { const float e = expf(a.v); return d5chain(a, e, e); }
__device__ __forceinline__ D5 d5clamp(const D5& a, float lo, float hi) {       // torch.clamp: gradient passes inside [lo, hi] only
  if (a.v < lo) return d5c(lo);
  if (a.v > hi) return d5c(hi);
  return a;
}

__device__ __forceinline__ void d5cov(const D5& w, const D5& h, const D5& r, D5& a, D5& b, D5& c) {   // metrics.py:178-195
  const D5 A = w * w * (1.f / 12.f), B = h * h * (1.f / 12.f);
  const D5 cs = d5cos(r), sn = d5sin(r);
  const D5 c2 = cs * cs, s2 = sn * sn;
  a = A * c2 + B * s2;
  b = A * s2 + B * c2;
  c = (A - B) * cs * sn;
}

template <typename T> __device__ __forceinline__ float ol_ld(const T* p) { return (float)*p; }
template <> __device__ __forceinline__ float ol_ld<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <typename T> __device__ __forceinline__ void ol_st(T* p, float v) { *p = (T)v; }
template <> __device__ __forceinline__ void ol_st<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

struct ObbLossArgs {
  const void* feat[3];
  void* dfeat[3];
  const void* angle;
  void* dangle;
  int H[3], W[3];
  float stride[3];
  int B, A, nc, reg_max, no;
  int ld[3];                       // row pitch (elements) of feat / dfeat per level: >= no (padded rows of the fused head tensor)
  const float* t_boxes;            // [B][A][5] pixels
  const float* t_scores;           // [B][A][nc]
  const unsigned char* fg;         // [B][A]
  const double* tss;               // sum of target_scores
  double* sums;                    // [4] raw sums (box, cls, dfl, angle), zero on entry
  float gain[4];                   // hyp.box, hyp.cls, hyp.dfl, lambda_angular
  int nblk[3];                     // staged kernels: blocks per image on each level (ceil(H*W / OL_THREADS))
};

// One anchor of the loss: `row` / `drow` are the anchor's head-output and gradient rows (global memory, or the block's staged copy in
// shared memory), `ts` / `tb` its target scores [nc] and box [5]; returns d(total)/d(theta).
template <typename T>
__device__ __forceinline__ float obb_loss_anchor(const ObbLossArgs& p, const T* row, T* drow, const float* ts, const float* tb, bool fg,
                                                 float theta, float st, float ax, float ay, float (&part)[4]) {
  const float tss = fmaxf((float)*p.tss, 1.f);
  const float scale = (float)p.B / tss;                       // d(total)/d(raw sum term) before the gain
  const int R = p.reg_max;
  // ---- class term: BCE with logits over all anchors (loss.py:998) ----------------------------------------------------------
  float wgt = 0.f;
  for (int c = 0; c < p.nc; ++c) {
    const float s = ol_ld(row + 4 * R + c), t = ts[c];
    wgt += t;
    part[1] += fmaxf(s, 0.f) - s * t + log1pf(expf(-fabsf(s)));
    const float sg = 1.f / (1.f + expf(-s));
    ol_st(drow + 4 * R + c, (sg - t) * scale * p.gain[1]);
  }
  float dtheta = 0.f;
  if (!fg) {
    for (int e = 0; e < 4 * R; ++e) ol_st(drow + e, 0.f);
  } else {
    // ---- decode (loss.py:1046-1050, tal.py:379-385) ---------------------------------------------------------------------------
    float prob[4][OL_MAXREG], lse[4], dist[4];
    for (int k = 0; k < 4; ++k) {
      float mx = -INFINITY;
      for (int j = 0; j < R; ++j) { prob[k][j] = ol_ld(row + k * R + j); mx = fmaxf(mx, prob[k][j]); }
      float sum = 0.f, ex = 0.f;
      for (int j = 0; j < R; ++j) { const float e = expf(prob[k][j] - mx); prob[k][j] = e; sum += e; }
      lse[k] = mx + logf(sum);
      const float inv = 1.f / sum;
      for (int j = 0; j < R; ++j) { prob[k][j] *= inv; ex += (float)j * prob[k][j]; }
      dist[k] = ex;
    }
    const float tx = tb[0] / st, ty = tb[1] / st, tw = tb[2] / st, th = tb[3] / st, tr = tb[4];
    // ---- box term: (1 - probiou(pred, target)) * weight, duals over (d_l, d_t, d_r, d_b, theta) -------------------------------
    const D5 d0 = d5var(dist[0], 0), d1 = d5var(dist[1], 1), d2 = d5var(dist[2], 2), d3 = d5var(dist[3], 3), dr = d5var(theta, 4);
    const D5 cs = d5cos(dr), sn = d5sin(dr);
    const D5 xf = (d2 - d0) * 0.5f, yf = (d3 - d1) * 0.5f;
    const D5 x1 = xf * cs - yf * sn + ax, y1 = xf * sn + yf * cs + ay, w1 = d0 + d2, h1 = d1 + d3;
    D5 a1, b1, c1, a2, b2, c2;
    d5cov(w1, h1, dr, a1, b1, c1);
    d5cov(d5c(tw), d5c(th), d5c(tr), a2, b2, c2);
    const float eps = 1e-7f;
    const D5 sa = a1 + a2, sb = b1 + b2, sc = c1 + c2;
    const D5 det = sa * sb - sc * sc;
    const D5 den = det + eps;
    const D5 dx = x1 - tx, dy = y1 - ty;                                   // (x1 - x2), (y1 - y2)
    const D5 t1 = ((sa * dy * dy + sb * dx * dx) / den) * 0.25f;
    const D5 t2 = ((sc * (dx * -1.f) * dy) / den) * 0.5f;                  // (c1+c2)(x2-x1)(y1-y2)
    const D5 q1 = d5clamp(a1 * b1 - c1 * c1, 0.f, INFINITY), q2 = d5clamp(a2 * b2 - c2 * c2, 0.f, INFINITY);
    const D5 t3 = d5log(det / (d5sqrt(q1 * q2) * 4.f + eps) + eps) * 0.5f;
    const D5 bd = d5clamp(t1 + t2 + t3, eps, 100.f);
    const D5 hd = d5sqrt(d5c(1.f + eps) - d5exp(bd * -1.f));
    part[0] = hd.v * wgt;                                                  // 1 - iou = hd
    float gd[4];                                                           // d(total)/d(dist_k) from the box term
    for (int k = 0; k < 4; ++k) gd[k] = hd.d[k] * wgt * scale * p.gain[0];
    dtheta = hd.d[4] * wgt * scale * p.gain[0];
    // ---- DFL term (loss.py:306-329, :372-374) + chain of the box term through the soft-arg-max ---------------------------------
    const float lim = (float)R - 1.f - 0.01f;
    const float tl4[4] = {ax - (tx - 0.5f * tw), ay - (ty - 0.5f * th), (tx + 0.5f * tw) - ax, (ty + 0.5f * th) - ay};
    for (int k = 0; k < 4; ++k) {
      const float tv = fminf(fmaxf(tl4[k], 0.f), lim);
      const int il = (int)tv;
      const float wl = (float)(il + 1) - tv, wr = 1.f - wl;
      const float zl = ol_ld(row + k * R + il), zr = ol_ld(row + k * R + il + 1);
      part[2] += ((lse[k] - zl) * wl + (lse[k] - zr) * wr) * 0.25f * wgt;
      const float gdfl = 0.25f * wgt * scale * p.gain[2];
      for (int j = 0; j < R; ++j) {
        const float pj = prob[k][j];
        const float g = gdfl * (pj - (j == il ? wl : 0.f) - (j == il + 1 ? wr : 0.f)) + gd[k] * pj * ((float)j - dist[k]);
        ol_st(drow + k * R + j, g);
      }
    }
    // ---- quaternion angular term (loss.py:870-903, :1019-1021): geodesic distance between rotations about z -------------------
    const float hdlt = 0.5f * (theta - tr);
    const float u = cosf(hdlt);
    const float uc = fminf(fmaxf(u, -1.f + 1e-7f), 1.f - 1e-7f);
    part[3] = 2.f * acosf(fabsf(uc)) * wgt;
    if (u == uc) dtheta += (uc >= 0.f ? 1.f : -1.f) * sinf(hdlt) * rsqrtf(1.f - uc * uc) * wgt * scale * p.gain[3];
  }
  return dtheta;
}

// block sums -> four double atomics
__device__ __forceinline__ void obb_loss_block_sums(const ObbLossArgs& p, const float (&part)[4], double (*red)[OL_THREADS / 32]) {
  for (int k = 0; k < 4; ++k) {
    double v = (double)part[k];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) red[k][threadIdx.x >> 5] = v;
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    double v = 0.0;
    for (int w = 0; w < OL_THREADS / 32; ++w) v += red[threadIdx.x][w];
    if (v != 0.0) atomicAdd(p.sums + threadIdx.x, v);
  }
}

// direct form: a thread reads and writes its anchor's rows in global memory (any row pitch)
template <typename T>
__global__ void __launch_bounds__(OL_THREADS) obb_loss_kernel(ObbLossArgs p) {
  pdl_prologue();
  __shared__ double red[4][OL_THREADS / 32];
  const int64_t i = (int64_t)blockIdx.x * OL_THREADS + threadIdx.x;
  float part[4] = {0.f, 0.f, 0.f, 0.f};
  if (i < (int64_t)p.B * p.A) {
    const int b = (int)(i / p.A), a = (int)(i % p.A);
    int l = 0, al = a;
    if (al >= p.H[0] * p.W[0]) { al -= p.H[0] * p.W[0]; l = 1; if (al >= p.H[1] * p.W[1]) { al -= p.H[1] * p.W[1]; l = 2; } }
    const int Wl = p.W[l], Al = p.H[l] * Wl;
    const float ax = (float)(al % Wl) + 0.5f, ay = (float)(al / Wl) + 0.5f;
    const T* row = reinterpret_cast<const T*>(p.feat[l]) + ((int64_t)b * Al + al) * p.ld[l];
    T* drow = reinterpret_cast<T*>(p.dfeat[l]) + ((int64_t)b * Al + al) * p.ld[l];
    const float theta = ol_ld(reinterpret_cast<const T*>(p.angle) + (int64_t)b * p.A + a);
    const float dtheta = obb_loss_anchor<T>(p, row, drow, p.t_scores + i * p.nc, p.t_boxes + i * 5, p.fg[i] != 0, theta, p.stride[l], ax, ay, part);
    ol_st(reinterpret_cast<T*>(p.dangle) + (int64_t)b * p.A + a, dtheta);
  }
  obb_loss_block_sums(p, part, red);
}

// ---- staged form ------------------------------------------------------------------------------------------------------------------
// The direct kernels move every anchor's row (no = 79 values) with one 2-byte access per thread and element: 32 lanes touch 32
// different sectors per instruction — 158 such instructions per anchor made the loss kernel 267 us for 110 MB of traffic.  Here a
// block owns up to OL_THREADS consecutive anchors of ONE level of one image: their rows (and target rows) are one contiguous piece of
// global memory, copied with 16-byte accesses into shared memory at an odd word pitch (a thread's row walk is conflict-free), worked
// on there, and the gradient rows go back the same way.
__device__ __forceinline__ void ol_block_anchor(const ObbLossArgs& p, int& b, int& l, int& al0) {
  const int bpi = p.nblk[0] + p.nblk[1] + p.nblk[2];
  b = (int)(blockIdx.x / (unsigned)bpi);
  int r = (int)(blockIdx.x - (unsigned)b * bpi);
  l = 0;
  if (r >= p.nblk[0]) { r -= p.nblk[0]; l = 1; if (r >= p.nblk[1]) { r -= p.nblk[1]; l = 2; } }
  al0 = r * OL_THREADS;
}
// rows of `ldw` 32-bit words, contiguous in global memory <-> shared memory at pitch `ldpw`
__device__ __forceinline__ void ol_rows_in(const uint32_t* __restrict__ g, uint32_t* s, int nwords, int ldw, int ldpw) {
  if ((ldw & 3) == 0 && (reinterpret_cast<uintptr_t>(g) & 15) == 0) {
    for (int v = threadIdx.x; v < (nwords >> 2); v += OL_THREADS) {
      const uint4 q = reinterpret_cast<const uint4*>(g)[v];
      const int w = v << 2, r = w / ldw, c = w - r * ldw;
      uint32_t* d = s + r * ldpw + c;
      d[0] = q.x; d[1] = q.y; d[2] = q.z; d[3] = q.w;
    }
  } else {
    for (int w = threadIdx.x; w < nwords; w += OL_THREADS) { const int r = w / ldw, c = w - r * ldw; s[r * ldpw + c] = g[w]; }
  }
}
__device__ __forceinline__ void ol_rows_out(uint32_t* __restrict__ g, const uint32_t* s, int nwords, int ldw, int ldpw) {
  if ((ldw & 3) == 0 && (reinterpret_cast<uintptr_t>(g) & 15) == 0) {
    for (int v = threadIdx.x; v < (nwords >> 2); v += OL_THREADS) {
      const int w = v << 2, r = w / ldw, c = w - r * ldw;
      const uint32_t* d = s + r * ldpw + c;
      reinterpret_cast<uint4*>(g)[v] = make_uint4(d[0], d[1], d[2], d[3]);
    }
  } else {
    for (int w = threadIdx.x; w < nwords; w += OL_THREADS) { const int r = w / ldw, c = w - r * ldw; g[w] = s[r * ldpw + c]; }
  }
}

template <typename T>
__global__ void __launch_bounds__(OL_THREADS) obb_loss_staged_kernel(ObbLossArgs p) {
  pdl_prologue();
  extern __shared__ __align__(16) uint32_t ol_smem[];
  __shared__ double red[4][OL_THREADS / 32];
  int b, l, al0;
  ol_block_anchor(p, b, l, al0);
  const int Wl = p.W[l], Al = p.H[l] * Wl;
  const int n = min(OL_THREADS, Al - al0);
  const int ldw = p.ld[l] * (int)sizeof(T) / 4, ldpw = ldw | 1;
  const int ncp = p.nc | 1;                               // odd pitch of the staged target-score rows
  // the gradient row overwrites the staged input row: obb_loss_anchor reads every element it needs before it writes that element
  // (class logits one by one, the DFL logits into registers first) — one buffer, 32 KB per block instead of 52 KB, 7 blocks per SM
  uint32_t* s_in = ol_smem;                               // [OL_THREADS][ldpw]
  uint32_t* s_out = s_in;
  float* s_ts = reinterpret_cast<float*>(s_in + OL_THREADS * ldpw);    // [OL_THREADS][ncp]
  float* s_tb = s_ts + OL_THREADS * ncp;                  // [OL_THREADS][5]
  int aoff = 0;
  for (int k = 0; k < l; ++k) aoff += p.H[k] * p.W[k];
  const int64_t i0 = (int64_t)b * p.A + aoff + al0;       // first anchor of the block in [B][A] order
  const int64_t e0 = ((int64_t)b * Al + al0) * p.ld[l];   // its first element in the level's tensor
  ol_rows_in(reinterpret_cast<const uint32_t*>(reinterpret_cast<const T*>(p.feat[l]) + e0), s_in, n * ldw, ldw, ldpw);
  for (int e = threadIdx.x; e < n * p.nc; e += OL_THREADS) { const int r = e / p.nc, c = e - r * p.nc; s_ts[r * ncp + c] = p.t_scores[i0 * p.nc + e]; }
  for (int e = threadIdx.x; e < n * 5; e += OL_THREADS) s_tb[e] = p.t_boxes[i0 * 5 + e];
  __syncthreads();
  float part[4] = {0.f, 0.f, 0.f, 0.f};
  if ((int)threadIdx.x < n) {
    const int al = al0 + threadIdx.x;
    const int64_t i = i0 + threadIdx.x;
    const float ax = (float)(al % Wl) + 0.5f, ay = (float)(al / Wl) + 0.5f;
    const T* row = reinterpret_cast<const T*>(s_in + threadIdx.x * ldpw);
    T* drow = reinterpret_cast<T*>(s_out + threadIdx.x * ldpw);
    for (int e = p.no; e < p.ld[l]; ++e) ol_st(drow + e, 0.f);        // padding columns of the gradient rows
    const float theta = ol_ld(reinterpret_cast<const T*>(p.angle) + i);
    const float dtheta = obb_loss_anchor<T>(p, row, drow, s_ts + threadIdx.x * ncp, s_tb + threadIdx.x * 5, p.fg[i] != 0, theta, p.stride[l], ax,
                                            ay, part);
    ol_st(reinterpret_cast<T*>(p.dangle) + i, dtheta);
  }
  __syncthreads();
  ol_rows_out(reinterpret_cast<uint32_t*>(reinterpret_cast<T*>(p.dfeat[l]) + e0), s_out, n * ldw, ldw, ldpw);
  obb_loss_block_sums(p, part, red);
}

// predictions for the assigner (loss.py:978-993, no gradient): sigmoid class scores [B][A][nc] and decoded boxes [B][A][5] in pixels
template <typename T>
__global__ void __launch_bounds__(OL_THREADS) obb_decode_kernel(ObbLossArgs p, float* __restrict__ scores, float* __restrict__ boxes) {
  pdl_prologue();
  const int64_t i = (int64_t)blockIdx.x * OL_THREADS + threadIdx.x;
  if (i >= (int64_t)p.B * p.A) return;
  const int b = (int)(i / p.A), a = (int)(i % p.A);
  int l = 0, al = a;
  if (al >= p.H[0] * p.W[0]) { al -= p.H[0] * p.W[0]; l = 1; if (al >= p.H[1] * p.W[1]) { al -= p.H[1] * p.W[1]; l = 2; } }
  const int Wl = p.W[l], Al = p.H[l] * Wl, R = p.reg_max;
  const float st = p.stride[l];
  const float ax = (float)(al % Wl) + 0.5f, ay = (float)(al / Wl) + 0.5f;
  const T* row = reinterpret_cast<const T*>(p.feat[l]) + ((int64_t)b * Al + al) * p.ld[l];
  float dist[4];
  for (int k = 0; k < 4; ++k) {
    float mx = -INFINITY, sum = 0.f, ex = 0.f;
    for (int j = 0; j < R; ++j) mx = fmaxf(mx, ol_ld(row + k * R + j));
    for (int j = 0; j < R; ++j) { const float e = expf(ol_ld(row + k * R + j) - mx); sum += e; ex += (float)j * e; }
    dist[k] = ex / sum;
  }
  const float theta = ol_ld(reinterpret_cast<const T*>(p.angle) + (int64_t)b * p.A + a);
  const float cs = cosf(theta), sn = sinf(theta);
  const float xf = 0.5f * (dist[2] - dist[0]), yf = 0.5f * (dist[3] - dist[1]);
  float* bx = boxes + i * 5;
  bx[0] = (xf * cs - yf * sn + ax) * st;
  bx[1] = (xf * sn + yf * cs + ay) * st;
  bx[2] = (dist[0] + dist[2]) * st;
  bx[3] = (dist[1] + dist[3]) * st;
  bx[4] = theta;
  float* sc = scores + i * p.nc;
  for (int c = 0; c < p.nc; ++c) sc[c] = 1.f / (1.f + expf(-ol_ld(row + 4 * R + c)));
}

// staged form of the decode (see obb_loss_staged_kernel): rows in through shared memory, scores / boxes out through shared memory
template <typename T>
__global__ void __launch_bounds__(OL_THREADS) obb_decode_staged_kernel(ObbLossArgs p, float* __restrict__ scores, float* __restrict__ boxes) {
  pdl_prologue();
  extern __shared__ __align__(16) uint32_t ol_smem[];
  int b, l, al0;
  ol_block_anchor(p, b, l, al0);
  const int Wl = p.W[l], Al = p.H[l] * Wl, R = p.reg_max;
  const int n = min(OL_THREADS, Al - al0);
  const int ldw = p.ld[l] * (int)sizeof(T) / 4, ldpw = ldw | 1;
  const int ncp = p.nc | 1;
  uint32_t* s_in = ol_smem;                               // [OL_THREADS][ldpw]
  float* s_sc = reinterpret_cast<float*>(s_in + OL_THREADS * ldpw);   // [OL_THREADS][ncp]
  float* s_bx = s_sc + OL_THREADS * ncp;                  // [OL_THREADS][5]
  int aoff = 0;
  for (int k = 0; k < l; ++k) aoff += p.H[k] * p.W[k];
  const int64_t i0 = (int64_t)b * p.A + aoff + al0;
  const int64_t e0 = ((int64_t)b * Al + al0) * p.ld[l];
  ol_rows_in(reinterpret_cast<const uint32_t*>(reinterpret_cast<const T*>(p.feat[l]) + e0), s_in, n * ldw, ldw, ldpw);
  __syncthreads();
  if ((int)threadIdx.x < n) {
    const int al = al0 + threadIdx.x;
    const float st = p.stride[l];
    const float ax = (float)(al % Wl) + 0.5f, ay = (float)(al / Wl) + 0.5f;
    const T* row = reinterpret_cast<const T*>(s_in + threadIdx.x * ldpw);
    float dist[4];
    for (int k = 0; k < 4; ++k) {
      float mx = -INFINITY, sum = 0.f, ex = 0.f;
      for (int j = 0; j < R; ++j) mx = fmaxf(mx, ol_ld(row + k * R + j));
      for (int j = 0; j < R; ++j) { const float e = expf(ol_ld(row + k * R + j) - mx); sum += e; ex += (float)j * e; }
      dist[k] = ex / sum;
    }
    const float theta = ol_ld(reinterpret_cast<const T*>(p.angle) + i0 + threadIdx.x);
    const float cs = cosf(theta), sn = sinf(theta);
    const float xf = 0.5f * (dist[2] - dist[0]), yf = 0.5f * (dist[3] - dist[1]);
    float* bx = s_bx + threadIdx.x * 5;
    bx[0] = (xf * cs - yf * sn + ax) * st;
    bx[1] = (xf * sn + yf * cs + ay) * st;
    bx[2] = (dist[0] + dist[2]) * st;
    bx[3] = (dist[1] + dist[3]) * st;
    bx[4] = theta;
    float* sc = s_sc + threadIdx.x * ncp;
    for (int c = 0; c < p.nc; ++c) sc[c] = 1.f / (1.f + expf(-ol_ld(row + 4 * R + c)));
  }
  __syncthreads();
  for (int e = threadIdx.x; e < n * p.nc; e += OL_THREADS) { const int r = e / p.nc, c = e - r * p.nc; scores[i0 * p.nc + e] = s_sc[r * ncp + c]; }
  for (int e = threadIdx.x; e < n * 5; e += OL_THREADS) boxes[i0 * 5 + e] = s_bx[e];
}

// staged kernels: rows must be whole 32-bit words; shared memory per block; grid = B x (blocks per image)
template <typename T>
static bool ol_stage_plan(ObbLossArgs& p, bool loss, size_t& smem, unsigned& blocks) {
  static const int on = [] { const char* e = getenv("QUAN_OBB_STAGED"); return e ? atoi(e) : 1; }();
  if (!on) return false;
  int ldpw_max = 0, bpi = 0;
  for (int l = 0; l < 3; ++l) {
    if ((p.ld[l] * (int)sizeof(T)) % 4 != 0) return false;
    const int ldpw = (p.ld[l] * (int)sizeof(T) / 4) | 1;
    ldpw_max = ldpw > ldpw_max ? ldpw : ldpw_max;
    p.nblk[l] = (p.H[l] * p.W[l] + OL_THREADS - 1) / OL_THREADS;
    bpi += p.nblk[l];
  }
  smem = (size_t)OL_THREADS * (ldpw_max + (p.nc | 1) + 5) * 4;
  (void)loss;
  blocks = (unsigned)((int64_t)p.B * bpi);
  return smem <= 160 * 1024 && (int64_t)p.B * bpi < (1ll << 31);
}

__global__ void obb_loss_tail_kernel(const double* sums, const double* tss, float g0, float g1, float g2, float g3, int B, float* items,
                                     float* total) {
  pdl_prologue();
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    const double t = *tss > 1.0 ? *tss : 1.0;
    const float it[4] = {(float)(sums[0] / t) * g0, (float)(sums[1] / t) * g1, (float)(sums[2] / t) * g2, (float)(sums[3] / t) * g3};
    for (int k = 0; k < 4; ++k) items[k] = it[k];
    *total = (it[0] + it[1] + it[2] + it[3]) * (float)B;
  }
}

__global__ void obb_tss_kernel(const float* __restrict__ ts, int64_t n, double* __restrict__ tss, double* __restrict__ sums) {
  pdl_prologue();
  __shared__ double red[8];
  double v = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) v += (double)ts[i];
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
    if (s != 0.0) atomicAdd(tss, s);
  }
  (void)sums;
}

}  // namespace quan

extern "C" {

int quan_obb_decode(const void* const feats[3], const void* pred_angle, const int32_t* hw, const float* strides, int32_t B, int32_t nc,
                    int32_t reg_max, const int32_t* feat_ld, float* pd_scores, float* pd_bboxes, int dtype, void* stream) {
  using namespace quan;
  QUAN_REQUIRE(feats && pred_angle && hw && strides && pd_scores && pd_bboxes, QUAN_E_ARG, "obb_decode: null pointer");
  QUAN_REQUIRE(B > 0 && nc > 0 && reg_max >= 2 && reg_max <= OL_MAXREG, QUAN_E_ARG, "obb_decode: B=%d nc=%d reg_max=%d", B, nc, reg_max);
  ObbLossArgs p = {};
  int64_t A = 0;
  for (int l = 0; l < 3; ++l) {
    QUAN_REQUIRE(feats[l] && hw[2 * l] > 0 && hw[2 * l + 1] > 0, QUAN_E_ARG, "obb_decode: level %d", l);
    p.feat[l] = feats[l]; p.H[l] = hw[2 * l]; p.W[l] = hw[2 * l + 1]; p.stride[l] = strides[l];
    p.ld[l] = feat_ld != nullptr ? feat_ld[l] : 4 * reg_max + nc;
    QUAN_REQUIRE(p.ld[l] >= 4 * reg_max + nc, QUAN_E_ARG, "obb_decode: feat_ld[%d] = %d below the row width", l, p.ld[l]);
    A += (int64_t)hw[2 * l] * hw[2 * l + 1];
  }
  QUAN_REQUIRE(A * B < (1ll << 31), QUAN_E_UNSUPPORTED, "obb_decode: too many anchors");
  p.angle = pred_angle; p.B = B; p.A = (int)A; p.nc = nc; p.reg_max = reg_max; p.no = 4 * reg_max + nc;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const unsigned blocks = (unsigned)(((int64_t)B * A + OL_THREADS - 1) / OL_THREADS);
  size_t smem = 0;
  unsigned sblocks = 0;
  QUAN_TIMED(st);
  if (dtype == QUAN_BF16) {
    if (ol_stage_plan<__nv_bfloat16>(p, false, smem, sblocks)) {
      static thread_local DeviceOnce attr;
      if (attr.first()) QUAN_CUDA(cudaFuncSetAttribute(obb_decode_staged_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
      QUAN_LAUNCH((obb_decode_staged_kernel<__nv_bfloat16>), sblocks, OL_THREADS, smem, st, p, pd_scores, pd_bboxes);
    } else QUAN_LAUNCH((obb_decode_kernel<__nv_bfloat16>), blocks, OL_THREADS, 0, st, p, pd_scores, pd_bboxes);
  } else {
    if (ol_stage_plan<float>(p, false, smem, sblocks)) {
      static thread_local DeviceOnce attr;
      if (attr.first()) QUAN_CUDA(cudaFuncSetAttribute(obb_decode_staged_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
      QUAN_LAUNCH((obb_decode_staged_kernel<float>), sblocks, OL_THREADS, smem, st, p, pd_scores, pd_bboxes);
    } else QUAN_LAUNCH((obb_decode_kernel<float>), blocks, OL_THREADS, 0, st, p, pd_scores, pd_bboxes);
  }
  QUAN_CHECK_LAUNCH("obb_decode");
  return QUAN_OK;
}

int quan_obb_loss_fwd_bwd(const void* const feats[3], const void* pred_angle, const int32_t* hw, const float* strides, int32_t B, int32_t nc,
                          int32_t reg_max, const float* target_bboxes, const float* target_scores, const uint8_t* fg_mask, float box_gain,
                          float cls_gain, float dfl_gain, float angle_gain, const int32_t* feat_ld, void* const d_feats[3], void* d_angle,
                          double* scratch, float* items, float* total, int dtype, void* stream) {
  using namespace quan;
  QUAN_REQUIRE(feats && d_feats && pred_angle && d_angle && hw && strides && target_bboxes && target_scores && fg_mask && scratch && items && total,
               QUAN_E_ARG, "obb_loss: null pointer");
  QUAN_REQUIRE(B > 0 && nc > 0 && reg_max >= 2 && reg_max <= OL_MAXREG, QUAN_E_ARG, "obb_loss: B=%d nc=%d reg_max=%d (reg_max <= %d)", B, nc, reg_max,
               OL_MAXREG);
  ObbLossArgs p = {};
  int64_t A = 0;
  for (int l = 0; l < 3; ++l) {
    QUAN_REQUIRE(feats[l] && d_feats[l] && hw[2 * l] > 0 && hw[2 * l + 1] > 0, QUAN_E_ARG, "obb_loss: level %d", l);
    p.feat[l] = feats[l]; p.dfeat[l] = d_feats[l]; p.H[l] = hw[2 * l]; p.W[l] = hw[2 * l + 1]; p.stride[l] = strides[l];
    p.ld[l] = feat_ld != nullptr ? feat_ld[l] : 4 * reg_max + nc;
    QUAN_REQUIRE(p.ld[l] >= 4 * reg_max + nc, QUAN_E_ARG, "obb_loss: feat_ld[%d] = %d below the row width", l, p.ld[l]);
    A += (int64_t)hw[2 * l] * hw[2 * l + 1];
  }
  QUAN_REQUIRE(A * B < (1ll << 31), QUAN_E_UNSUPPORTED, "obb_loss: too many anchors");
  p.angle = pred_angle; p.dangle = d_angle; p.B = B; p.A = (int)A; p.nc = nc; p.reg_max = reg_max; p.no = 4 * reg_max + nc;
  p.t_boxes = target_bboxes; p.t_scores = target_scores; p.fg = fg_mask;
  p.tss = scratch; p.sums = scratch + 1;
  p.gain[0] = box_gain; p.gain[1] = cls_gain; p.gain[2] = dfl_gain; p.gain[3] = angle_gain;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  QUAN_CUDA(cudaMemsetAsync(scratch, 0, 5 * sizeof(double), st));
  const int64_t n = (int64_t)B * A * nc;
  QUAN_TIMED(st);
  QUAN_LAUNCH((obb_tss_kernel), (unsigned)((n + 256 * 8 - 1) / (256 * 8) < 1184 ? (n + 256 * 8 - 1) / (256 * 8) : 1184), 256, 0, st, target_scores, n,
              scratch, scratch + 1);
  QUAN_CHECK_LAUNCH("obb_loss_tss");
  const unsigned blocks = (unsigned)(((int64_t)B * A + OL_THREADS - 1) / OL_THREADS);
  size_t smem = 0;
  unsigned sblocks = 0;
  QUAN_TIMED(st);
  if (dtype == QUAN_BF16) {
    if (ol_stage_plan<__nv_bfloat16>(p, true, smem, sblocks)) {
      static thread_local DeviceOnce attr;
      if (attr.first()) QUAN_CUDA(cudaFuncSetAttribute(obb_loss_staged_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
      QUAN_LAUNCH((obb_loss_staged_kernel<__nv_bfloat16>), sblocks, OL_THREADS, smem, st, p);
    } else QUAN_LAUNCH((obb_loss_kernel<__nv_bfloat16>), blocks, OL_THREADS, 0, st, p);
  } else {
    if (ol_stage_plan<float>(p, true, smem, sblocks)) {
      static thread_local DeviceOnce attr;
      if (attr.first()) QUAN_CUDA(cudaFuncSetAttribute(obb_loss_staged_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
      QUAN_LAUNCH((obb_loss_staged_kernel<float>), sblocks, OL_THREADS, smem, st, p);
    } else QUAN_LAUNCH((obb_loss_kernel<float>), blocks, OL_THREADS, 0, st, p);
  }
  QUAN_CHECK_LAUNCH("obb_loss");
  QUAN_TIMED(st);
  QUAN_LAUNCH((obb_loss_tail_kernel), 1, 32, 0, st, (const double*)(scratch + 1), (const double*)scratch, box_gain, cls_gain, dfl_gain, angle_gain,
              (int)B, items, total);
  QUAN_CHECK_LAUNCH("obb_loss_tail");
  return QUAN_OK;
}

}  // extern "C"
