#include "common.h"

#include <atomic>
#include <stdarg.h>
#include <string.h>

#include "../../include/dexbotic_b200.h"

namespace b200 {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int num_sms() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

}  // namespace b200

extern "C" {
int b200_version(void) { return 100; }
const char* b200_last_error(void) { return b200::g_err; }
int64_t b200_launch_count(void) { return b200::g_launches.load(); }
}
