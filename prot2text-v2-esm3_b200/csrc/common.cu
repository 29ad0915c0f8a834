#include "common.h"

#include <atomic>
#include <cstdarg>
#include <cstdio>

namespace p2t {

static thread_local char g_err[512] = "";
static std::atomic<unsigned long long> g_launches{0};

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error((int)e, "%s: %s", what, cudaGetErrorString(e));
  count_launch();
  return 0;
}
const char* last_error() { return g_err; }
unsigned long long launches() { return g_launches.load(std::memory_order_relaxed); }
void reset_launches() { g_launches.store(0, std::memory_order_relaxed); }

}  // namespace p2t
