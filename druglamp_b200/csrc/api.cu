// Library-wide plumbing of the C ABI: version, thread-local error message, launch counter.
#include <atomic>
#include <stdlib.h>
#include <stdarg.h>
#include <string.h>

#include "../../include/druglamp_sm100.h"
#include "common.cuh"

namespace dl {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int sm_count() {
  static int n = [] {
    int dev = 0, v = 148;
    if (cudaGetDevice(&dev) == cudaSuccess)
      cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    return v > 0 ? v : 148;
  }();
  return n;
}

bool pdl_enabled() {
  static const bool on = [] {
    const char* e = getenv("DL_NO_PDL");
    return !(e && atoi(e) != 0);
  }();
  return on;
}

}  // namespace dl

extern "C" int dl_version(void) { return 100; }
extern "C" const char* dl_last_error(void) { return dl::g_err; }
extern "C" int64_t dl_launch_count(void) { return (int64_t)dl::g_launches.load(); }
