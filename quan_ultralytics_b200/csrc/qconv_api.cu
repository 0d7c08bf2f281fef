// qconv_api.cu — C-ABI entry points for QConv2D (argument validation, engine selection, workspace carving).
// Replaces the reference extension's host wrappers qconv_forward_cuda / qconv_backward_cuda
// (ultralytics/nn/cuda/quaternion_ops.cu:735-799, :532-679); unlike those, every launch goes to the caller's
// stream, outputs/workspace are caller-allocated, and launch errors are reported.
#include "qconv_internal.cuh"
#include <mutex>

namespace quan {

static int validate(const quan_conv_dims* d, int dtype, int layout, const float* mix) {
  QUAN_REQUIRE(d != nullptr, QUAN_E_ARG, "qconv2d: null dims");
  QUAN_REQUIRE(dtype == QUAN_F32 || dtype == QUAN_BF16, QUAN_E_ARG, "qconv2d: bad dtype %d", dtype);
  QUAN_REQUIRE(layout == QUAN_LAYOUT_BCHWQ || layout == QUAN_LAYOUT_BHWQC, QUAN_E_ARG, "qconv2d: bad layout %d", layout);
  QUAN_REQUIRE(mix != nullptr, QUAN_E_ARG, "qconv2d: null mixing matrix");
  QUAN_REQUIRE(d->B > 0 && d->Ci > 0 && d->Co > 0 && d->H > 0 && d->W > 0, QUAN_E_ARG, "qconv2d: non-positive tensor dims");
  QUAN_REQUIRE(d->kH > 0 && d->kW > 0 && d->sH > 0 && d->sW > 0 && d->dH > 0 && d->dW > 0 && d->pH >= 0 && d->pW >= 0,
               QUAN_E_ARG, "qconv2d: bad kernel/stride/padding/dilation");
  QUAN_REQUIRE(d->groups > 0 && d->Ci % d->groups == 0 && d->Co % d->groups == 0, QUAN_E_SHAPE,
               "qconv2d: Ci=%d / Co=%d not divisible by groups=%d", d->Ci, d->Co, d->groups);
  QUAN_REQUIRE(conv_out(d->H, d->kH, d->sH, d->pH, d->dH) > 0 && conv_out(d->W, d->kW, d->sW, d->pW, d->dW) > 0,
               QUAN_E_SHAPE, "qconv2d: empty output (input %dx%d, kernel %dx%d)", d->H, d->W, d->kH, d->kW);
  return QUAN_OK;
}

// engines whose backward kernels read the raw output gradient (M^T applied on load)
static bool raw_dy(int algo) { return algo == QUAN_ALGO_DEPTHWISE || algo == QUAN_ALGO_SMALLC; }

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static size_t g_bytes(const quan_conv_dims& d, int dtype) {
  const int Ho = conv_out(d.H, d.kH, d.sH, d.pH, d.dH), Wo = conv_out(d.W, d.kW, d.sW, d.pW, d.dW);
  return align_up((size_t)d.B * d.Co * Ho * Wo * 4 * (dtype == QUAN_F32 ? 4 : 2), 1024);
}

static int resolve_algo(const quan_conv_dims& d, int dtype, int layout, int pass, int algo) {
  if (algo == QUAN_ALGO_DIRECT) return QUAN_ALGO_DIRECT;
  if (algo == QUAN_ALGO_DEPTHWISE) return qconv_dw_supported(d, dtype, layout, pass) ? QUAN_ALGO_DEPTHWISE : -1;
  if (algo == QUAN_ALGO_SMALLC) return qconv_small_supported(d, dtype, layout, pass) ? QUAN_ALGO_SMALLC : -1;
  const bool ok = qconv_tc_supported(d, dtype, layout, pass);
  if (algo == QUAN_ALGO_TCGEN05) return ok ? QUAN_ALGO_TCGEN05 : -1;
  if (ok) return QUAN_ALGO_TCGEN05;
  if (qconv_dw_supported(d, dtype, layout, pass)) return QUAN_ALGO_DEPTHWISE;
  return qconv_small_supported(d, dtype, layout, pass) ? QUAN_ALGO_SMALLC : QUAN_ALGO_DIRECT;
}

}  // namespace quan

using namespace quan;

extern "C" {

int quan_qconv2d_pick_algo(const quan_conv_dims* d, int dtype, int layout, int pass) {
  if (d == nullptr || pass < 0 || pass > 2) return QUAN_E_ARG;
  return resolve_algo(*d, dtype, layout, pass, QUAN_ALGO_AUTO);
}

// workspace layout: [ G = M^T dY | packed weights (fwd / dgrad) | wgrad split-K partials ] — dgrad and wgrad get disjoint
// regions because the backward of a narrow layer runs them concurrently (fork / join on a side stream)
static size_t pack_region_bytes(const quan_conv_dims& d, int dtype, int layout) {
  size_t b = 0;
  for (int pass = PASS_FWD; pass <= PASS_DGRAD; ++pass)
    if (qconv_tc_supported(d, dtype, layout, pass)) {
      const size_t v = qconv_tc_workspace_bytes(d, dtype, layout, pass);
      if (v > b) b = v;
    }
  return align_up(b, 1024);
}

size_t quan_qconv2d_workspace_bytes(const quan_conv_dims* d, int dtype, int layout, int algo) {
  if (d == nullptr) return 0;
  size_t tc = 0;
  if (algo != QUAN_ALGO_DIRECT && layout == QUAN_LAYOUT_BHWQC) {
    tc = pack_region_bytes(*d, dtype, layout);
    if (qconv_tc_supported(*d, dtype, layout, PASS_WGRAD)) tc += align_up(qconv_tc_workspace_bytes(*d, dtype, layout, PASS_WGRAD), 1024);
  }
  return g_bytes(*d, dtype) + tc;
}

// side stream + events for the dgrad || wgrad fork (one set per host thread and device)
struct ForkJoin {
  cudaStream_t side = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr;
};
// Deferred join (quan_bwd_side_stream_set): the caller lends a side stream and a workspace of its own for the wgrad chains; the
// wgrad of a narrow layer is then forked onto that stream and NOT joined by the call — the weight gradient is only needed by the
// optimizer, so the ~80 wgrad launches of a step leave the dX critical path of the backward pass.  quan_bwd_side_stream_join makes a
// stream wait for everything forked so far.  Per device, process-wide.
struct DeferredSide {
  cudaStream_t side = nullptr;
  void* ws = nullptr;
  size_t ws_bytes = 0;
  cudaEvent_t fork = nullptr, join = nullptr;
  bool pending = false;
};
// process-wide, one per device: the setter runs on the host thread that drives the step, the backward kernels are launched from the
// autograd engine's device thread
static std::mutex g_deferred_mu;
static DeferredSide& deferred_side() {
  static DeferredSide d[16];
  int dev = 0;
  cudaGetDevice(&dev);
  return d[dev & 15];
}

static int get_fork_join(ForkJoin** out) {
  static thread_local ForkJoin fj[16];
  int dev = 0;
  QUAN_CUDA(cudaGetDevice(&dev));
  QUAN_REQUIRE(dev >= 0 && dev < 16, QUAN_E_UNSUPPORTED, "qconv2d_bwd: device index %d out of range", dev);
  ForkJoin& f = fj[dev];
  if (f.side == nullptr) {
    QUAN_CUDA(cudaStreamCreateWithFlags(&f.side, cudaStreamNonBlocking));
    QUAN_CUDA(cudaEventCreateWithFlags(&f.fork, cudaEventDisableTiming));
    QUAN_CUDA(cudaEventCreateWithFlags(&f.join, cudaEventDisableTiming));
  }
  *out = &f;
  return QUAN_OK;
}

// Algorithmic work of one conv pass for the kernel-timing table (bench.py roofline): the separable FLOP count of SURVEY §8(d),
// 4*2*B*Ho*Wo*Co*(Ci/g)*kH*kW, and the bytes a pass must move — both activation tensors it touches plus the four weight sets.
static void announce_conv_work(const quan_conv_dims& d, int dtype, int pass) {
  const double Ho = conv_out(d.H, d.kH, d.sH, d.pH, d.dH), Wo = conv_out(d.W, d.kW, d.sW, d.pW, d.dW);
  const double esz = dtype == QUAN_BF16 ? 2.0 : 4.0;
  const double flops = 8.0 * d.B * Ho * Wo * d.Co * (double)(d.Ci / d.groups) * d.kH * d.kW;
  const double sx = 4.0 * d.B * d.Ci * (double)d.H * d.W * esz, sy = 4.0 * d.B * d.Co * Ho * Wo * esz;
  const double sw = 4.0 * d.Co * (double)(d.Ci / d.groups) * d.kH * d.kW * 4.0;
  (void)pass;
  timing_work("qconv_", "qconv_bias_grad", sx + sy + sw, flops);
}

static int qconv2d_fwd_impl(const void* x, const float* const w[4], const float* bias_r, void* y, const quan_conv_dims* d,
                            int dtype, int layout, const float* mix, int algo, void* workspace, size_t ws_bytes,
                            void* stream, double* stat_part, int* stat_nparts) {
  if (stat_nparts != nullptr) *stat_nparts = *stat_nparts < 0 ? -1 : 0;      // < 0 on entry: the caller accepts accumulated statistics
  int rc = validate(d, dtype, layout, mix);
  if (rc) return rc;
  QUAN_REQUIRE(x != nullptr && y != nullptr && w != nullptr && w[0] && w[1] && w[2] && w[3], QUAN_E_ARG,
               "qconv2d_fwd: null tensor pointer");
  cudaStream_t st = (cudaStream_t)stream;
  announce_conv_work(*d, dtype, PASS_FWD);
  const int a = resolve_algo(*d, dtype, layout, PASS_FWD, algo);
  QUAN_REQUIRE(a > 0, QUAN_E_UNSUPPORTED, "qconv2d_fwd: requested engine does not serve this shape/layout");
  if (stat_nparts != nullptr && a != QUAN_ALGO_TCGEN05) *stat_nparts = 0;     // only the tensor-core epilogue emits statistics
  if (a == QUAN_ALGO_DEPTHWISE) return qconv_dw_fwd(x, w, bias_r, y, *d, dtype, mix, st);
  if (a == QUAN_ALGO_SMALLC) return qconv_small_fwd(x, w, bias_r, y, *d, dtype, mix, st);
  if (a == QUAN_ALGO_TCGEN05) {
    const size_t need = qconv_tc_workspace_bytes(*d, dtype, layout, PASS_FWD);
    QUAN_REQUIRE(workspace != nullptr && ws_bytes >= need, QUAN_E_WORKSPACE,
                 "qconv2d_fwd: tcgen05 engine needs %zu workspace bytes, got %zu", need, ws_bytes);
    return qconv_tc_fwd(x, w, bias_r, y, *d, dtype, qconv_tc_mode(*d, dtype, layout, PASS_FWD), mix, workspace, ws_bytes, st,
                        stat_part, stat_nparts);
  }
  return qconv_fwd_direct_launch(x, w, bias_r, y, *d, dtype, layout, mix, st);
}

int quan_qconv2d_fwd(const void* x, const float* const w[4], const float* bias_r, void* y, const quan_conv_dims* d,
                     int dtype, int layout, const float* mix, int algo, void* workspace, size_t ws_bytes,
                     void* stream) {
  return qconv2d_fwd_impl(x, w, bias_r, y, d, dtype, layout, mix, algo, workspace, ws_bytes, stream, nullptr, nullptr);
}

int quan_qconv2d_fwd_stats(const void* x, const float* const w[4], const float* bias_r, void* y, const quan_conv_dims* d,
                           int dtype, int layout, const float* mix, int algo, void* workspace, size_t ws_bytes,
                           void* iqbn_workspace, size_t iqbn_ws_bytes, int* nparts, void* stream) {
  QUAN_REQUIRE(nparts != nullptr && d != nullptr, QUAN_E_ARG, "qconv2d_fwd_stats: null pointer");
  const bool room = iqbn_workspace != nullptr && iqbn_ws_bytes >= quan_iqbn_workspace_bytes(d->Co);
  return qconv2d_fwd_impl(x, w, bias_r, y, d, dtype, layout, mix, algo, workspace, ws_bytes, stream,
                          room ? (double*)iqbn_workspace : nullptr, nparts);
}

}  // extern "C"

// premixed: `dy` already holds G = M^T dY (emitted by quan_iqbn_bwd_apply with mix_t); only legal when no selected pass
// wants the raw gradient (quan_qconv2d_bwd_wants_mixed)
static int qconv2d_bwd_impl(const void* dy, const void* x, const float* const w[4], void* dx, float* const dw[4],
                            float* dbias_r, const quan_conv_dims* d, int dtype, int layout, const float* mix, int algo,
                            void* workspace, size_t ws_bytes, void* stream, bool premixed) {
  int rc = validate(d, dtype, layout, mix);
  if (rc) return rc;
  QUAN_REQUIRE(dy != nullptr && w != nullptr && w[0] && w[1] && w[2] && w[3], QUAN_E_ARG, "qconv2d_bwd: null tensor pointer");
  QUAN_REQUIRE(dw == nullptr || (dw[0] && dw[1] && dw[2] && dw[3]), QUAN_E_ARG, "qconv2d_bwd: dw must hold 4 pointers");
  QUAN_REQUIRE(dw == nullptr || x != nullptr, QUAN_E_ARG, "qconv2d_bwd: weight gradient needs the input x");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t gb = g_bytes(*d, dtype);
  QUAN_REQUIRE(workspace != nullptr && ws_bytes >= gb, QUAN_E_WORKSPACE,
               "qconv2d_bwd: workspace needs >= %zu bytes (see quan_qconv2d_workspace_bytes), got %zu", gb, ws_bytes);
  const void* gq = premixed ? dy : workspace;
  // [ G | packed weights | wgrad partials ]
  const size_t pack_bytes = layout == QUAN_LAYOUT_BHWQC ? pack_region_bytes(*d, dtype, layout) : 0;
  void* tc_ws = (char*)workspace + gb;
  const size_t tc_ws_bytes = ws_bytes - gb;
  void* wg_ws = ws_bytes >= gb + pack_bytes ? (char*)workspace + gb + pack_bytes : nullptr;
  size_t wg_ws_bytes = ws_bytes >= gb + pack_bytes ? ws_bytes - gb - pack_bytes : 0;

  // engine per pass; the dense tensor-core form consumes dY directly (M is folded into its weights / reduce step)
  int a_dx = 0, a_dw = 0, m_dx = TC_NONE, m_dw = TC_NONE;
  if (dx != nullptr) {
    a_dx = resolve_algo(*d, dtype, layout, PASS_DGRAD, algo);
    QUAN_REQUIRE(a_dx > 0, QUAN_E_UNSUPPORTED, "qconv2d_bwd: requested engine does not serve this dgrad shape/layout");
    if (a_dx == QUAN_ALGO_TCGEN05) m_dx = qconv_tc_mode(*d, dtype, layout, PASS_DGRAD);
  }
  if (dw != nullptr) {
    a_dw = resolve_algo(*d, dtype, layout, PASS_WGRAD, algo);
    QUAN_REQUIRE(a_dw > 0, QUAN_E_UNSUPPORTED, "qconv2d_bwd: requested engine does not serve this wgrad shape/layout");
    if (a_dw == QUAN_ALGO_TCGEN05) m_dw = qconv_tc_mode(*d, dtype, layout, PASS_WGRAD);
  }
  // G = M^T dY, once, shared by every consumer that needs it (separable dgrad / wgrad, direct engine, bias grad)
  const bool need_g = (dx != nullptr && m_dx != TC_DENSE && !raw_dy(a_dx)) ||
                      (dw != nullptr && m_dw != TC_DENSE && !raw_dy(a_dw)) || dbias_r != nullptr;
  QUAN_REQUIRE(!premixed || ((dx == nullptr || (m_dx != TC_DENSE && !raw_dy(a_dx))) &&
                             (dw == nullptr || (m_dw != TC_DENSE && !raw_dy(a_dw)))),
               QUAN_E_UNSUPPORTED, "qconv2d_bwd_premixed: this shape's dgrad/wgrad consume the raw gradient (dense / depthwise form)");
  if (need_g && !premixed) {
    float mix_t[16];
    for (int p = 0; p < 4; ++p)
      for (int q = 0; q < 4; ++q) mix_t[q * 4 + p] = mix[p * 4 + q];
    const int Ho = conv_out(d->H, d->kH, d->sH, d->pH, d->dH), Wo = conv_out(d->W, d->kW, d->sW, d->pW, d->dW);
    rc = quan_mix(dy, workspace, d->B, d->Co, Ho, Wo, dtype, layout, mix_t, stream);
    if (rc) return rc;
  }

  // Narrow layers (dense tensor-core form, depthwise, small-channel kernels) do not fill the GPU with one kernel: their
  // dgrad and wgrad chains are independent and run concurrently — wgrad forks onto a side stream and joins at the end
  // (captured as parallel branches under CUDA graphs).  QUAN_BWD_CONCURRENT=0 disables.
  static const int env_conc = [] { const char* e = getenv("QUAN_BWD_CONCURRENT"); return e ? atoi(e) : 1; }();
  const bool narrow = (m_dx == TC_DENSE || raw_dy(a_dx)) && (m_dw == TC_DENSE || raw_dy(a_dw));
  ForkJoin* fj = nullptr;
  cudaStream_t st_w = st;
  std::unique_lock<std::mutex> ds_lock(g_deferred_mu);
  DeferredSide& ds = deferred_side();
  const bool dw_narrow = m_dw == TC_DENSE || raw_dy(a_dw);          // the wgrad reads dY itself (no G pre-pass on the main stream)
  // ... or G handed in by the caller (premixed: a tensor of its own, not this call's workspace)
  if (env_conc && dw != nullptr && (premixed ? a_dw == QUAN_ALGO_TCGEN05 : dw_narrow) && ds.side != nullptr &&
      (a_dw != QUAN_ALGO_TCGEN05 || ds.ws_bytes >= qconv_tc_workspace_bytes(*d, dtype, layout, PASS_WGRAD))) {
    // deferred: fork only (the caller joins once, before anything reads the weight gradients); partials live in the lender's workspace
    QUAN_CUDA(cudaEventRecord(ds.fork, st));
    QUAN_CUDA(cudaStreamWaitEvent(ds.side, ds.fork, 0));
    st_w = ds.side;
    wg_ws = ds.ws;
    wg_ws_bytes = ds.ws_bytes;
    ds.pending = true;
    ds_lock.unlock();
  } else if (env_conc && dx != nullptr && dw != nullptr && narrow) {
    rc = get_fork_join(&fj);
    if (rc) return rc;
    QUAN_CUDA(cudaEventRecord(fj->fork, st));
    QUAN_CUDA(cudaStreamWaitEvent(fj->side, fj->fork, 0));
    st_w = fj->side;
  }
  if (ds_lock.owns_lock()) ds_lock.unlock();

  if (dx != nullptr) {
    announce_conv_work(*d, dtype, PASS_DGRAD);
    if (a_dx == QUAN_ALGO_TCGEN05) {
      const size_t need = qconv_tc_workspace_bytes(*d, dtype, layout, PASS_DGRAD);
      QUAN_REQUIRE(tc_ws_bytes >= need, QUAN_E_WORKSPACE, "qconv2d_bwd: dgrad needs %zu more workspace bytes", need);
      rc = qconv_tc_dgrad(m_dx == TC_DENSE ? dy : gq, w, dx, *d, dtype, m_dx, mix, tc_ws, tc_ws_bytes, st);
    } else if (a_dx == QUAN_ALGO_DEPTHWISE) {
      rc = qconv_dw_dgrad(dy, w, dx, *d, dtype, mix, st);
    } else if (a_dx == QUAN_ALGO_SMALLC) {
      rc = qconv_small_dgrad(dy, w, dx, *d, dtype, mix, st);
    } else {
      rc = qconv_dgrad_direct_launch(gq, w, dx, *d, dtype, layout, st);
    }
    if (rc) {
      if (fj != nullptr) {                     // nothing was launched on the side stream yet: close the fork
        cudaEventRecord(fj->join, fj->side);
        cudaStreamWaitEvent(st, fj->join, 0);
      }
      return rc;
    }
  }
  if (dw != nullptr) {
    announce_conv_work(*d, dtype, PASS_WGRAD);
    if (a_dw == QUAN_ALGO_TCGEN05) {
      const size_t need = qconv_tc_workspace_bytes(*d, dtype, layout, PASS_WGRAD);
      QUAN_REQUIRE(wg_ws != nullptr && wg_ws_bytes >= need, QUAN_E_WORKSPACE, "qconv2d_bwd: wgrad needs %zu more workspace bytes", need);
      rc = qconv_tc_wgrad(m_dw == TC_DENSE ? dy : gq, x, dw, *d, dtype, m_dw, mix, wg_ws, wg_ws_bytes, st_w);
    } else if (a_dw == QUAN_ALGO_DEPTHWISE) {
      rc = qconv_dw_wgrad(dy, x, dw, *d, dtype, mix, st_w);
    } else if (a_dw == QUAN_ALGO_SMALLC) {
      rc = qconv_small_wgrad(dy, x, dw, *d, dtype, mix, st_w);
    } else {
      rc = qconv_wgrad_direct_launch(gq, x, dw, *d, dtype, layout, st_w);
    }
    if (fj != nullptr) {                       // join even when the wgrad launch failed: the side stream must not dangle
      cudaEventRecord(fj->join, fj->side);
      cudaStreamWaitEvent(st, fj->join, 0);
    }
    if (rc) return rc;
  }
  if (dbias_r != nullptr) {
    rc = qconv_bias_grad_launch(gq, dbias_r, *d, dtype, layout, st);
    if (rc) return rc;
  }
  return QUAN_OK;
}

extern "C" {

int quan_qconv2d_bwd(const void* dy, const void* x, const float* const w[4], void* dx, float* const dw[4],
                     float* dbias_r, const quan_conv_dims* d, int dtype, int layout, const float* mix, int algo,
                     void* workspace, size_t ws_bytes, void* stream) {
  return qconv2d_bwd_impl(dy, x, w, dx, dw, dbias_r, d, dtype, layout, mix, algo, workspace, ws_bytes, stream, false);
}

int quan_qconv2d_bwd_premixed(const void* g, const void* x, const float* const w[4], void* dx, float* const dw[4],
                              float* dbias_r, const quan_conv_dims* d, int dtype, int layout, const float* mix, int algo,
                              void* workspace, size_t ws_bytes, void* stream) {
  return qconv2d_bwd_impl(g, x, w, dx, dw, dbias_r, d, dtype, layout, mix, algo, workspace, ws_bytes, stream, true);
}

int quan_bwd_side_stream_set(void* side_stream, void* wgrad_workspace, size_t ws_bytes) {
  std::lock_guard<std::mutex> lock(g_deferred_mu);
  DeferredSide& ds = deferred_side();
  if (side_stream == nullptr) {                // off: back to fork + join inside every call
    QUAN_REQUIRE(!ds.pending, QUAN_E_ARG, "bwd_side_stream_set: join the forked work first (quan_bwd_side_stream_join)");
    ds.side = nullptr; ds.ws = nullptr; ds.ws_bytes = 0;
    return QUAN_OK;
  }
  QUAN_REQUIRE(wgrad_workspace != nullptr && ws_bytes > 0, QUAN_E_ARG, "bwd_side_stream_set: a side stream needs its own workspace");
  if (ds.fork == nullptr) {
    QUAN_CUDA(cudaEventCreateWithFlags(&ds.fork, cudaEventDisableTiming));
    QUAN_CUDA(cudaEventCreateWithFlags(&ds.join, cudaEventDisableTiming));
  }
  ds.side = (cudaStream_t)side_stream; ds.ws = wgrad_workspace; ds.ws_bytes = ws_bytes;
  return QUAN_OK;
}

int quan_bwd_side_stream_join(void* stream) {
  std::lock_guard<std::mutex> lock(g_deferred_mu);
  DeferredSide& ds = deferred_side();
  if (ds.side == nullptr) return QUAN_OK;
  // under stream capture the side stream only belongs to the capture once something was forked onto it; joining an uncaptured
  // stream into a capturing one is an error (and there is nothing to wait for)
  cudaStreamCaptureStatus cs_main = cudaStreamCaptureStatusNone, cs_side = cudaStreamCaptureStatusNone;
  QUAN_CUDA(cudaStreamIsCapturing((cudaStream_t)stream, &cs_main));
  QUAN_CUDA(cudaStreamIsCapturing(ds.side, &cs_side));
  if (cs_main != cudaStreamCaptureStatusNone && cs_side == cudaStreamCaptureStatusNone) { ds.pending = false; return QUAN_OK; }
  QUAN_CUDA(cudaEventRecord(ds.join, ds.side));
  QUAN_CUDA(cudaStreamWaitEvent((cudaStream_t)stream, ds.join, 0));
  ds.pending = false;
  return QUAN_OK;
}

int quan_qconv2d_bwd_wants_mixed(const quan_conv_dims* d, int dtype, int layout, int algo, int need_dx, int need_dw) {
  if (d == nullptr) return QUAN_E_ARG;
  for (int pass = PASS_DGRAD; pass <= PASS_WGRAD; ++pass) {
    if (pass == PASS_DGRAD ? !need_dx : !need_dw) continue;
    const int a = resolve_algo(*d, dtype, layout, pass, algo);
    if (a <= 0) return 0;
    if (a == QUAN_ALGO_DEPTHWISE || a == QUAN_ALGO_SMALLC) return 0;
    if (a == QUAN_ALGO_TCGEN05 && qconv_tc_mode(*d, dtype, layout, pass) == TC_DENSE) return 0;
  }
  return 1;
}

}  // extern "C"
