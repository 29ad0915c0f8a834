// Device-side pieces of the peer-memory exchange protocol (see peer.cu for the protocol itself), shared by the
// channel kernels in peer.cu and by kernels that wait for a gathered block themselves (loss_fused.cu): the table of
// mapped peer buffers, the control-block layout, system-scope loads/stores and the flag wait.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace p2t {

constexpr int kPeerMaxWorld = 16;
constexpr unsigned long long kPeerTimeoutNs = 20ull * 1000ull * 1000ull * 1000ull;  // 20 s

struct PeerTable {
  void* base[kPeerMaxWorld];  // base[r] = this process's mapping of rank r's peer buffer (base[rank] = own)
};

// ---- control block at the start of every channel buffer ----
//   u32 [0]            epoch of the last completed round
//   u32 [1], [2], [3]  grid counters of the channel's kernels
//   u32 [4]            status (0 = ok, 1 + r = timed out waiting for rank r); sticky until p2t_peer_reset
//   u32 [64 + ph*16 + r]  arrival flag of phase ph (0, 1) from rank r
constexpr size_t kPeerCtrlBytes = 1024;
__host__ __device__ constexpr size_t peer_flag_row_off(int phase) { return (64 + phase * 16) * sizeof(unsigned); }

__device__ __forceinline__ uint4 ld_sys_v4(const uint4* p) {  // relaxed system-scope load: never served from a stale L1 line
  uint4 v;
  asm volatile("ld.relaxed.sys.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ unsigned ld_relaxed_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_sys(unsigned* p, unsigned v) {
  asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Threads [0, world) of the CTA wait for flags[t] to reach `epoch` (relaxed polls, then one system fence = acquire);
// the CTA barrier that follows extends the ordering to every thread of the CTA.  System-scope fences cost a few
// microseconds each (they wait for the thread's outstanding NVLink traffic), so they are issued by the polling /
// publishing threads only, never by all threads.  Returns false (to every thread of the CTA) when a wait gave up:
// the caller must then POISON what it produces (NaN) instead of consuming a stale or partial block.
__device__ __forceinline__ bool wait_flags(const unsigned* flags, int world, unsigned epoch, unsigned* status) {
  __shared__ int s_timed_out;
  if (threadIdx.x == 0) s_timed_out = 0;
  __syncthreads();
  if ((int)threadIdx.x < world) {
    const unsigned long long t0 = global_ns();
    while ((int)(ld_relaxed_sys(flags + threadIdx.x) - epoch) < 0) {
      if (global_ns() - t0 > kPeerTimeoutNs) {
        atomicExch(status, 1u + threadIdx.x);
        s_timed_out = 1;
        break;
      }
      __nanosleep(32);
    }
    __threadfence_system();
  }
  __syncthreads();
  return s_timed_out == 0 && ld_relaxed_sys(status) == 0u;
}

}  // namespace p2t
