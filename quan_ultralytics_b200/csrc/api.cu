// api.cu — library-level entry points of libquan_sm100.so (version, error string).
#include "common.cuh"
#include <string.h>
#include <atomic>

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
}  // namespace quan

extern "C" {

int quan_version(void) { return QUAN_ABI_VERSION; }

const char* quan_last_error(void) { return quan::g_err; }

uint64_t quan_launch_count(void) { return quan::g_launches.load(std::memory_order_relaxed); }

const char* quan_build_info(void) {
  return "libquan_sm100 abi=1 arch=sm_100a nvcc=" QUAN_STR(__CUDACC_VER_MAJOR__) "." QUAN_STR(__CUDACC_VER_MINOR__)
         " engines=direct,tcgen05";
}

}  // extern "C"
