// api.cu — library-level entry points of libquan_sm100.so (version, error string).
#include "common.cuh"
#include <string.h>
#include <atomic>
#include <map>
#include <mutex>
#include <string>
#include <vector>

namespace quan {
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
static std::atomic<uint64_t> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

// ---- optional per-kernel device timing (bench.py's kernel table): a CUDA-event pair around every launch that is
// bracketed by timing_begin() ... QUAN_CHECK_LAUNCH(name), on the stream the kernel is launched on.  Off by default.
struct TimedLaunch {
  cudaEvent_t e0, e1;
  double bytes, flops;          // algorithmic work of this launch (timing_work), 0 when the launch site did not declare any
};
static std::atomic<int> g_timing{0};
static std::mutex g_timing_mu;
static std::map<std::string, std::vector<TimedLaunch>> g_timed;
static thread_local cudaEvent_t t_begin = nullptr;
static thread_local cudaStream_t t_stream = nullptr;
// algorithmic work announced by an API entry for the NEXT launch whose name starts with `t_work_prefix` (auxiliary kernels —
// weight packing, folds, the mix pre-pass — carry other names and pass it on)
static thread_local double t_work_bytes = 0.0, t_work_flops = 0.0;
static thread_local const char* t_work_prefix = nullptr;
static thread_local const char* t_work_skip = nullptr;

void timing_work(const char* prefix, const char* skip, double bytes, double flops) {
  if (!g_timing.load(std::memory_order_relaxed)) return;
  t_work_prefix = prefix;
  t_work_skip = skip;
  t_work_bytes = bytes;
  t_work_flops = flops;
}

void timing_begin(cudaStream_t st) {
  if (!g_timing.load(std::memory_order_relaxed)) return;
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) return;
  cudaEvent_t e;
  if (cudaEventCreate(&e) != cudaSuccess) return;
  cudaEventRecord(e, st);
  t_begin = e;
  t_stream = st;
}
void timing_end(const char* name) {
  if (t_begin == nullptr) return;
  cudaEvent_t e1;
  if (cudaEventCreate(&e1) == cudaSuccess) {
    cudaEventRecord(e1, t_stream);
    double wb = 0.0, wf = 0.0;
    if (t_work_prefix != nullptr && strncmp(name, t_work_prefix, strlen(t_work_prefix)) == 0 &&
        (t_work_skip == nullptr || strncmp(name, t_work_skip, strlen(t_work_skip)) != 0)) {
      wb = t_work_bytes;
      wf = t_work_flops;
      t_work_prefix = nullptr;
    }
    std::lock_guard<std::mutex> lk(g_timing_mu);
    g_timed[name].push_back({t_begin, e1, wb, wf});
  } else {
    cudaEventDestroy(t_begin);
  }
  t_begin = nullptr;
}
static void timing_clear() {
  std::lock_guard<std::mutex> lk(g_timing_mu);
  for (auto& kv : g_timed)
    for (auto& t : kv.second) { cudaEventDestroy(t.e0); cudaEventDestroy(t.e1); }
  g_timed.clear();
}
}  // namespace quan

extern "C" {

int quan_version(void) { return QUAN_ABI_VERSION; }

const char* quan_last_error(void) { return quan::g_err; }

uint64_t quan_launch_count(void) { return quan::g_launches.load(std::memory_order_relaxed); }

int quan_kernel_timing_enable(int on) {
  if (on) quan::timing_clear();
  quan::g_timing.store(on ? 1 : 0);
  return QUAN_OK;
}

size_t quan_kernel_timing_report(char* buf, size_t cap) {
  std::lock_guard<std::mutex> lk(quan::g_timing_mu);
  std::string out;
  for (auto& kv : quan::g_timed) {
    double total = 0.0, bytes = 0.0, flops = 0.0;
    for (auto& t : kv.second) {
      float ms = 0.f;
      if (cudaEventSynchronize(t.e1) == cudaSuccess && cudaEventElapsedTime(&ms, t.e0, t.e1) == cudaSuccess) total += ms;
      bytes += t.bytes;
      flops += t.flops;
    }
    char line[320];
    snprintf(line, sizeof(line), "%s %zu %.6f %.6e %.6e\n", kv.first.c_str(), kv.second.size(), total, bytes, flops);
    out += line;
  }
  if (buf != nullptr && cap > 0) {
    const size_t n = out.size() < cap - 1 ? out.size() : cap - 1;
    memcpy(buf, out.data(), n);
    buf[n] = 0;
  }
  return out.size() + 1;
}

const char* quan_build_info(void) {
  return "libquan_sm100 abi=1 arch=sm_100a nvcc=" QUAN_STR(__CUDACC_VER_MAJOR__) "." QUAN_STR(__CUDACC_VER_MINOR__)
         " engines=direct,tcgen05,depthwise,smallc";
}

}  // extern "C"
