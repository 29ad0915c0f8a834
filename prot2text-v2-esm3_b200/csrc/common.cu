#include "common.h"

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

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
// ---- launch stamps: one event after every launch, resolved on demand ----
static std::atomic<bool> g_stamping{false};
static std::mutex g_stamp_mu;  // forward runs on the caller's thread, backward on autograd's worker
static std::vector<cudaEvent_t> g_stamp_ev;
static std::vector<const char*> g_stamp_name;  // string literals
static size_t g_stamp_used = 0;
void stamp_launch(const char* name, cudaStream_t st) {
  if (!g_stamping.load(std::memory_order_relaxed)) return;
  std::lock_guard<std::mutex> lock(g_stamp_mu);
  if (g_stamp_used == g_stamp_ev.size()) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    g_stamp_ev.push_back(e);
    g_stamp_name.push_back(name);
  }
  g_stamp_name[g_stamp_used] = name;
  cudaEventRecord(g_stamp_ev[g_stamp_used++], st);
}
void launch_timing_enable(bool on) {
  std::lock_guard<std::mutex> lock(g_stamp_mu);
  g_stamp_used = 0;
  g_stamping.store(on, std::memory_order_relaxed);
}
// names (newline separated, in launch order) and the time from the previous stamp to each stamp, in ms; a "mark"
// stamp has time 0 and starts a new step.  Caller must have synchronised.  Returns the number of stamps.
int launch_timing_collect(double* ms, int cap, int* n_out, char* names, int names_cap) {
  std::lock_guard<std::mutex> lock(g_stamp_mu);
  std::string joined;
  int n = 0;
  for (size_t i = 0; i < g_stamp_used; ++i) {
    float t = 0.f;
    const bool is_mark = std::strcmp(g_stamp_name[i], "mark") == 0;
    if (i > 0 && !is_mark) {
      cudaError_t e = cudaEventElapsedTime(&t, g_stamp_ev[i - 1], g_stamp_ev[i]);
      if (e != cudaSuccess) return set_error((int)e, "cudaEventElapsedTime: %s", cudaGetErrorString(e));
    }
    if (ms != nullptr && n < cap) ms[n] = t;
    joined += g_stamp_name[i];
    joined += '\n';
    ++n;
  }
  if (names != nullptr && names_cap > 0) {
    std::strncpy(names, joined.c_str(), (size_t)names_cap - 1);
    names[names_cap - 1] = 0;
  }
  *n_out = n;
  g_stamp_used = 0;
  return 0;
}

int check_launch(const char* what, cudaStream_t st) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error((int)e, "%s: %s", what, cudaGetErrorString(e));
  count_launch();
  stamp_launch(what, st);
  return 0;
}
const char* last_error() { return g_err; }
unsigned long long launches() { return g_launches.load(std::memory_order_relaxed); }
void reset_launches() { g_launches.store(0, std::memory_order_relaxed); }


// ---- GEMM launch timing: event pairs recorded around each GEMM launch, resolved on demand ----
static bool g_timing = false;
static std::vector<cudaEvent_t> g_ev;  // begin/end pairs
static size_t g_ev_used = 0;
bool gemm_timing_enabled() { return g_timing; }
void gemm_timing_record(cudaStream_t st, bool begin) {
  (void)begin;
  if (g_ev_used == g_ev.size()) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    g_ev.push_back(e);
  }
  cudaEventRecord(g_ev[g_ev_used++], st);
}
void gemm_timing_enable(bool on) { g_timing = on; g_ev_used = 0; }
// sum of (end - begin) over recorded pairs, in ms; caller must have synchronised the stream
int gemm_timing_collect(double* total_ms, int* pairs, double* each_ms, int each_cap) {
  double t = 0.0;
  int n = 0;
  for (size_t i = 0; i + 1 < g_ev_used; i += 2) {
    float ms = 0.f;
    cudaError_t e = cudaEventElapsedTime(&ms, g_ev[i], g_ev[i + 1]);
    if (e != cudaSuccess) return set_error((int)e, "cudaEventElapsedTime: %s", cudaGetErrorString(e));
    t += ms;
    if (each_ms != nullptr && n < each_cap) each_ms[n] = ms;
    ++n;
  }
  *total_ms = t;
  *pairs = n;
  g_ev_used = 0;
  return 0;
}

}  // namespace p2t
