// Library-level entry points: version and the thread-local error message of the C ABI.
#include <stdarg.h>
#include <string.h>

#include <atomic>

#include <stdlib.h>

#include "common.cuh"

namespace rb {

static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

static std::atomic<uint64_t> g_launches{0};

void count_launches(int n) { g_launches.fetch_add(static_cast<uint64_t>(n), std::memory_order_relaxed); }

// programmatic dependent launch (common.cuh): -1 = the RB_PDL environment variable decides (default on), 0 / 1 = set by rb_set_pdl
static std::atomic<int> g_pdl{-1};
bool pdl_enabled() {
  const int v = g_pdl.load(std::memory_order_relaxed);
  if (v >= 0) return v != 0;
  static const bool env_on = [] {
    const char* e = getenv("RB_PDL");
    return e == nullptr || e[0] != '0';
  }();
  return env_on;
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) in %s", static_cast<int>(e), cudaGetErrorString(e), what);
  return RB_ERR_CUDA;
}

}  // namespace rb

extern "C" int rb_version(void) { return RB_VERSION; }

extern "C" const char* rb_last_error(void) { return rb::g_error; }

extern "C" uint64_t rb_kernel_launches(void) { return rb::g_launches.load(std::memory_order_relaxed); }
extern "C" void rb_set_pdl(int32_t mode) { rb::g_pdl.store(mode < 0 ? -1 : (mode != 0 ? 1 : 0), std::memory_order_relaxed); }
