// tal.cu — rotated task-aligned assigner of the OBB loss as three kernels with static shapes and no host
// synchronisation (so the whole loss can live in a CUDA graph).
//
// Reference: ultralytics/utils/tal.py:298-330 RotatedTaskAlignedAssigner (+ the base class :14-296), called from
// utils/loss.py:985-993 under no_grad: for every image b, ground-truth box j and anchor a
//   in_gts   = anchor centre inside the rotated box                     (tal.py:306-330, corners by ops.py:572-600)
//   overlap  = probiou(gt, pred).clamp(0) where in_gts, else 0           (metrics.py:198-241, tal.py:301-303)
//   align    = score[b,a,label_j]^alpha * overlap^beta                   (tal.py:147)
//   mask_pos = top-`topk` anchors of align per (b,j)  AND in_gts AND mask_gt   (tal.py:118-129, :155-186)
//   anchors claimed by several boxes go to the box of highest overlap    (tal.py:271-296)
//   target_scores = one_hot(label) * align * max_a(overlap) / (max_a(align) + eps)   (tal.py:108-113)
// The reference materialises ~10 [B, n, A] tensors, indexes with boolean masks (host syncs) and builds index tensors on the CPU
// (tal.py:138-140): 9-14 ms of host-bound time per YOLO11n step.  Here: one block per (b, j) computes its row of metrics, keeps the
// row in shared memory and selects the top-k by k block-wide arg-max passes (ties -> lowest anchor index); one thread per (b, a)
// resolves multiply-claimed anchors and scatters the per-box maxima with atomicMax; one thread per (b, a) writes the targets.
#include "common.cuh"
#include <math.h>

namespace quan {

constexpr int TAL_THREADS = 256;

struct RBox {
  float x, y, w, h, r;
};
__device__ __forceinline__ RBox ld_box(const float* p) { return RBox{p[0], p[1], p[2], p[3], p[4]}; }

__device__ __forceinline__ void covariance(const RBox& b, float& a, float& bb, float& c) {      // metrics.py:178-195
  const float A = b.w * b.w / 12.f, B = b.h * b.h / 12.f;
  const float cs = cosf(b.r), sn = sinf(b.r);
  const float c2 = cs * cs, s2 = sn * sn;
  a = A * c2 + B * s2;
  bb = A * s2 + B * c2;
  c = (A - B) * cs * sn;
}

__device__ __forceinline__ float probiou_dev(const RBox& g, float a1, float b1, float c1, const RBox& p) {   // metrics.py:217-233
  const float eps = 1e-7f;
  float a2, b2, c2;
  covariance(p, a2, b2, c2);
  const float sa = a1 + a2, sb = b1 + b2, sc = c1 + c2;
  const float den = sa * sb - sc * sc + eps;
  const float dx = g.x - p.x, dy = g.y - p.y;
  const float t1 = ((sa * dy * dy + sb * dx * dx) / den) * 0.25f;
  const float t2 = ((sc * (p.x - g.x) * dy) / den) * 0.5f;
  const float d1 = fmaxf(a1 * b1 - c1 * c1, 0.f), d2 = fmaxf(a2 * b2 - c2 * c2, 0.f);
  const float t3 = logf((sa * sb - sc * sc) / (4.f * sqrtf(d1 * d2) + eps) + eps) * 0.5f;
  const float bd = fminf(fmaxf(t1 + t2 + t3, eps), 100.f);
  const float hd = sqrtf(1.f - expf(-bd) + eps);
  return 1.f - hd;
}

struct InBox {                       // tal.py:317-330 with the corners of ops.py:589-600: a = pt1, b = pt2, d = pt4
  float ax, ay, abx, aby, adx, ady, nab, nad;
};
__device__ __forceinline__ InBox make_inbox(const RBox& g) {
  const float cs = cosf(g.r), sn = sinf(g.r);
  const float v1x = g.w * 0.5f * cs, v1y = g.w * 0.5f * sn, v2x = -g.h * 0.5f * sn, v2y = g.h * 0.5f * cs;
  InBox t;
  t.ax = g.x + v1x + v2x; t.ay = g.y + v1y + v2y;                    // pt1
  const float bx = g.x + v1x - v2x, by = g.y + v1y - v2y;            // pt2
  const float dx = g.x - v1x + v2x, dy = g.y - v1y + v2y;            // pt4
  t.abx = bx - t.ax; t.aby = by - t.ay; t.adx = dx - t.ax; t.ady = dy - t.ay;
  t.nab = t.abx * t.abx + t.aby * t.aby;
  t.nad = t.adx * t.adx + t.ady * t.ady;
  return t;
}
__device__ __forceinline__ bool inside(const InBox& t, float px, float py) {
  const float apx = px - t.ax, apy = py - t.ay;
  const float d1 = apx * t.abx + apy * t.aby, d2 = apx * t.adx + apy * t.ady;
  return d1 >= 0.f && d1 <= t.nab && d2 >= 0.f && d2 <= t.nad;
}

struct TalArgs {
  const float* scores; const float* boxes; const float* anc; const float* labels; const float* gts; const float* mask;
  int B, A, n, nc, topk;
  float alpha, beta, eps;
  float* overlaps; float* align; int* cnt; int* cand; unsigned* pos_align; unsigned* pos_over;
  float* t_boxes; float* t_scores; unsigned char* fg; long long* tgi;
  int full_scan;                   // QUAN_TAL_FULLSCAN=1: always the full-row scans (tests)
};

__global__ void tal_clear_kernel(TalArgs t) {
  pdl_prologue();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < (int64_t)t.B * t.A) t.cnt[i] = 0;
  if (i < (int64_t)t.B * t.n) { t.pos_align[i] = 0u; t.pos_over[i] = 0u; }
}

// grid (n, B), one block per ground-truth box.  Top-k with the reference's tie rule (torch.topk on the metric row; equal values ->
// lowest anchor index, which is what matters for the zeros): the anchors with a POSITIVE metric — a few hundred of the 21 504 of a
// 1024^2 image, those inside the box — are compacted into shared memory while the row is computed, the k arg-max passes run over that
// list, and if fewer than k are positive the remaining picks are the lowest-index anchors whose metric is exactly zero (they only count
// if they lie inside the box, tested below as before).  The first version kept the whole row in shared memory (84 KB per block) and
// scanned all of it k times: 120 us per step.  A box with more positive anchors than the list holds falls back to the full scans.
constexpr int TAL_CAP = 4096;

__device__ __forceinline__ void tal_argmax_block(float& bv, int& bi, int& bp, float* red_v, int* red_i, int* red_p) {
  // block-wide arg-max of (value desc, anchor index asc); every thread passes its candidate, thread 0 returns the winner
  for (int o = 16; o > 0; o >>= 1) {
    const float ov2 = __shfl_xor_sync(0xffffffffu, bv, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    const int op = __shfl_xor_sync(0xffffffffu, bp, o);
    if (ov2 > bv || (ov2 == bv && oi < bi)) { bv = ov2; bi = oi; bp = op; }
  }
  if ((threadIdx.x & 31) == 0) { red_v[threadIdx.x >> 5] = bv; red_i[threadIdx.x >> 5] = bi; red_p[threadIdx.x >> 5] = bp; }
  __syncthreads();
  if (threadIdx.x < 32) {
    bv = threadIdx.x < TAL_THREADS / 32 ? red_v[threadIdx.x] : -2.f;
    bi = threadIdx.x < TAL_THREADS / 32 ? red_i[threadIdx.x] : 0x7fffffff;
    bp = threadIdx.x < TAL_THREADS / 32 ? red_p[threadIdx.x] : 0;
    for (int o = 16; o > 0; o >>= 1) {
      const float ov2 = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      const int op = __shfl_xor_sync(0xffffffffu, bp, o);
      if (ov2 > bv || (ov2 == bv && oi < bi)) { bv = ov2; bi = oi; bp = op; }
    }
  }
}

__global__ void __launch_bounds__(TAL_THREADS) tal_metrics_topk_kernel(TalArgs t) {
  pdl_prologue();
  __shared__ float cval[TAL_CAP];
  __shared__ int cidx[TAL_CAP];
  __shared__ float red_v[TAL_THREADS / 32];
  __shared__ int red_i[TAL_THREADS / 32], red_p[TAL_THREADS / 32];
  __shared__ int sel[64];
  __shared__ int s_cnt;
  const int j = blockIdx.x, b = blockIdx.y;
  const int64_t rowoff = ((int64_t)b * t.n + j) * t.A;
  float* ov = t.overlaps + rowoff;
  float* al = t.align + rowoff;
  const bool valid = t.mask[b * t.n + j] != 0.f;
  if (!valid) {                                    // tal.py:121 (mask_in_gts * mask_gt): the row is all zero and selects nothing
    for (int a = threadIdx.x; a < t.A; a += TAL_THREADS) { ov[a] = 0.f; al[a] = 0.f; }
    return;
  }
  if (threadIdx.x == 0) s_cnt = 0;
  if (threadIdx.x < 64) sel[threadIdx.x] = 0x7fffffff;
  __syncthreads();
  const RBox g = ld_box(t.gts + ((int64_t)b * t.n + j) * 5);
  const int label = (int)t.labels[b * t.n + j];
  float a1, b1, c1;
  covariance(g, a1, b1, c1);
  const InBox ib = make_inbox(g);
  for (int a = threadIdx.x; a < t.A; a += TAL_THREADS) {
    float o = 0.f, m = 0.f;
    if (inside(ib, t.anc[2 * a], t.anc[2 * a + 1])) {
      const RBox p = ld_box(t.boxes + ((int64_t)b * t.A + a) * 5);
      o = fmaxf(probiou_dev(g, a1, b1, c1, p), 0.f);
      const float s = t.scores[((int64_t)b * t.A + a) * t.nc + label];
      m = powf(s, t.alpha) * powf(o, t.beta);
      if (m > 0.f) {
        const int k = atomicAdd(&s_cnt, 1);
        if (k < TAL_CAP) { cval[k] = m; cidx[k] = a; }
      }
    }
    ov[a] = o;
    al[a] = m;
  }
  __syncthreads();
  const int npos = s_cnt;
  const bool compact = npos <= TAL_CAP && !t.full_scan;
  if (compact) {
    const int nsel = npos < t.topk ? npos : t.topk;
    for (int k = 0; k < nsel; ++k) {
      float bv = -2.f;
      int bi = 0x7fffffff, bp = 0;
      for (int e = threadIdx.x; e < npos; e += TAL_THREADS) {
        const float v = cval[e];
        const int a = cidx[e];
        if (v > bv || (v == bv && a < bi)) { bv = v; bi = a; bp = e; }
      }
      tal_argmax_block(bv, bi, bp, red_v, red_i, red_p);
      if (threadIdx.x == 0) { sel[k] = bi; cval[bp] = -1.f; }
      __syncthreads();
    }
    if (threadIdx.x == 0) {                        // fewer positives than k: the lowest-index zeros (al[] of this block is visible after the barrier)
      int k = nsel;
      for (int a = 0; a < t.A && k < t.topk; ++a)
        if (al[a] == 0.f) sel[k++] = a;
    }
    __syncthreads();
  } else {
    // full scans of the row in global memory (ties -> lowest index); selected entries are hidden with -1 and restored afterwards
    for (int k = 0; k < t.topk; ++k) {
      float bv = -2.f;
      int bi = 0x7fffffff, bp = 0;
      for (int a = threadIdx.x; a < t.A; a += TAL_THREADS) {
        const float v = al[a];
        if (v > bv) { bv = v; bi = a; }            // strided ascending scan: first maximum of this thread
      }
      tal_argmax_block(bv, bi, bp, red_v, red_i, red_p);
      if (threadIdx.x == 0) {
        sel[k] = bi;
        if (bi < t.A) al[bi] = -1.f;
      }
      __syncthreads();
    }
  }
  if (threadIdx.x < t.topk) {
    const int a = sel[threadIdx.x];
    if (a < t.A) {
      if (!compact) al[a] = powf(t.scores[((int64_t)b * t.A + a) * t.nc + label], t.alpha) * powf(ov[a], t.beta);   // restore
      if (inside(ib, t.anc[2 * a], t.anc[2 * a + 1])) {            // mask_topk * mask_in_gts (tal.py:127)
        atomicAdd(t.cnt + (int64_t)b * t.A + a, 1);
        t.cand[(int64_t)b * t.A + a] = j;
      }
    }
  }
}

// one thread per (b, a): tal.py:271-296 select_highest_overlaps, then the per-box maxima of tal.py:109-110
__global__ void tal_resolve_kernel(TalArgs t) {
  pdl_prologue();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)t.B * t.A) return;
  const int b = (int)(i / t.A), a = (int)(i % t.A);
  const int c = t.cnt[i];
  int j = 0;
  if (c == 1) {
    j = t.cand[i];
  } else if (c > 1) {
    float best = -1.f;
    for (int q = 0; q < t.n; ++q) {
      const float o = t.overlaps[((int64_t)b * t.n + q) * t.A + a];
      if (o > best) { best = o; j = q; }           // first maximum (torch.argmax)
    }
  }
  t.fg[i] = c > 0;
  t.tgi[i] = c > 0 ? j : 0;
  if (c > 0) {
    const int64_t e = ((int64_t)b * t.n + j) * t.A + a;
    atomicMax(t.pos_align + b * t.n + j, __float_as_uint(t.align[e]));      // non-negative floats order like their bit patterns
    atomicMax(t.pos_over + b * t.n + j, __float_as_uint(t.overlaps[e]));
  }
}

// one thread per (b, a): tal.py:188-236 get_targets and the normalisation :108-113.  The [B][A][nc] score rows and [B][A][5] boxes of a
// block's 256 consecutive anchors are contiguous: every thread computes its anchor's (label, normalised metric, box) into shared
// memory and the block writes the rows with consecutive lanes on consecutive floats (the per-thread row writes — 20 floats 60 / 20
// bytes apart across lanes — made this 36 us for 27 MB).
__global__ void __launch_bounds__(256) tal_targets_kernel(TalArgs t) {
  pdl_prologue();
  __shared__ float s_norm[256];
  __shared__ int s_label[256];
  __shared__ float s_box[256 * 5];
  const int64_t i0 = (int64_t)blockIdx.x * 256;
  const int64_t i = i0 + threadIdx.x;
  const int64_t total = (int64_t)t.B * t.A;
  if (i < total) {
    const int b = (int)(i / t.A), a = (int)(i % t.A);
    const int j = (int)t.tgi[i];
    const bool fg = t.fg[i] != 0;
    const float* g = t.gts + ((int64_t)b * t.n + j) * 5;
#pragma unroll
    for (int k = 0; k < 5; ++k) s_box[threadIdx.x * 5 + k] = g[k];
    int label = (int)t.labels[b * t.n + j];
    label = label < 0 ? 0 : label;
    float norm = 0.f;
    if (fg) {
      const int64_t e = ((int64_t)b * t.n + j) * t.A + a;
      norm = t.align[e] * __uint_as_float(t.pos_over[b * t.n + j]) / (__uint_as_float(t.pos_align[b * t.n + j]) + t.eps);
    }
    s_norm[threadIdx.x] = norm;
    s_label[threadIdx.x] = fg ? label : -1;
  }
  __syncthreads();
  const int n = (int)(total - i0 < 256 ? total - i0 : 256);
  for (int e = threadIdx.x; e < n * t.nc; e += 256) {
    const int r = e / t.nc, c = e - r * t.nc;
    t.t_scores[i0 * t.nc + e] = c == s_label[r] ? s_norm[r] : 0.f;
  }
  for (int e = threadIdx.x; e < n * 5; e += 256) t.t_boxes[i0 * 5 + e] = s_box[e];
}

}  // namespace quan

extern "C" {

size_t quan_rotated_tal_workspace_bytes(int32_t B, int32_t A, int32_t n) {
  if (B <= 0 || A <= 0 || n <= 0) return 0;
  return ((size_t)2 * B * n * A + (size_t)2 * B * A + (size_t)2 * B * n) * 4 + 256;
}

int quan_rotated_tal_assign(const float* pd_scores, const float* pd_bboxes, const float* anc_points, const float* gt_labels,
                            const float* gt_bboxes, const float* mask_gt, int32_t B, int32_t A, int32_t n, int32_t nc, int32_t topk,
                            float alpha, float beta, float eps, float* target_bboxes, float* target_scores, uint8_t* fg_mask,
                            int64_t* target_gt_idx, void* workspace, size_t ws_bytes, void* stream) {
  using namespace quan;
  QUAN_REQUIRE(B > 0 && A > 0 && n > 0 && nc > 0 && topk > 0 && topk <= 64, QUAN_E_ARG, "rotated_tal_assign: bad sizes B=%d A=%d n=%d nc=%d topk=%d",
               B, A, n, nc, topk);
  QUAN_REQUIRE(pd_scores && pd_bboxes && anc_points && gt_labels && gt_bboxes && mask_gt && target_bboxes && target_scores && fg_mask &&
                   target_gt_idx && workspace, QUAN_E_ARG, "rotated_tal_assign: null pointer");
  QUAN_REQUIRE(ws_bytes >= quan_rotated_tal_workspace_bytes(B, A, n), QUAN_E_WORKSPACE, "rotated_tal_assign: workspace needs %zu bytes, got %zu",
               quan_rotated_tal_workspace_bytes(B, A, n), ws_bytes);
  QUAN_REQUIRE(n <= 65535 && B <= 65535, QUAN_E_UNSUPPORTED, "rotated_tal_assign: grid limit");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  TalArgs t;
  t.scores = pd_scores; t.boxes = pd_bboxes; t.anc = anc_points; t.labels = gt_labels; t.gts = gt_bboxes; t.mask = mask_gt;
  t.B = B; t.A = A; t.n = n; t.nc = nc; t.topk = topk; t.alpha = alpha; t.beta = beta; t.eps = eps;
  float* w = reinterpret_cast<float*>(workspace);
  t.overlaps = w; w += (size_t)B * n * A;
  t.align = w; w += (size_t)B * n * A;
  t.cnt = reinterpret_cast<int*>(w); w += (size_t)B * A;
  t.cand = reinterpret_cast<int*>(w); w += (size_t)B * A;
  t.pos_align = reinterpret_cast<unsigned*>(w); w += (size_t)B * n;
  t.pos_over = reinterpret_cast<unsigned*>(w);
  t.t_boxes = target_bboxes; t.t_scores = target_scores; t.fg = fg_mask; t.tgi = reinterpret_cast<long long*>(target_gt_idx);
  static const int env_full = [] { const char* e = getenv("QUAN_TAL_FULLSCAN"); return e ? atoi(e) : 0; }();
  t.full_scan = env_full;
  auto kern = tal_metrics_topk_kernel;
  const int64_t na = (int64_t)B * A;
  const int blocks = (int)((na + 255) / 256);
  QUAN_TIMED(st);
  QUAN_LAUNCH((tal_clear_kernel), blocks, 256, 0, st, t);
  QUAN_CHECK_LAUNCH("tal_clear");
  QUAN_TIMED(st);
  QUAN_LAUNCH((kern), dim3(n, B), TAL_THREADS, 0, st, t);
  QUAN_CHECK_LAUNCH("tal_metrics_topk");
  QUAN_TIMED(st);
  QUAN_LAUNCH((tal_resolve_kernel), blocks, 256, 0, st, t);
  QUAN_CHECK_LAUNCH("tal_resolve");
  QUAN_TIMED(st);
  QUAN_LAUNCH((tal_targets_kernel), blocks, 256, 0, st, t);
  QUAN_CHECK_LAUNCH("tal_targets");
  return QUAN_OK;
}

}  // extern "C"
