// Device-side pieces of the peer-memory exchange protocol (see peer.cu for the protocol itself), shared by the
// channel kernels in peer.cu and by kernels that wait for a gathered block themselves (loss_fused.cu): the table of
// mapped peer buffers, the control-block layout, system-scope loads/stores and the flag wait.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "mathfn.cuh"

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
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {  // LDG.STRONG.SYS + CCTL.IVALL: no MEMBAR
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// Release side of the flag protocol: acq_rel is all the pattern "data stores ; fence ; relaxed flag store" needs (the PTX
// model's release pattern); __threadfence_system() is the sequentially consistent fence (MEMBAR.SC.SYS).
__device__ __forceinline__ void fence_release_sys() { asm volatile("fence.acq_rel.sys;" ::: "memory"); }
__device__ __forceinline__ void st_relaxed_sys(unsigned* p, unsigned v) {
  asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Threads [0, world) of the CTA wait for flags[t] to reach `epoch`: relaxed polls, then ONE acquire load of the flag
// (it reads the releasing store or a later one; an acquire load costs an L1 invalidate, not a MEMBAR — ncu showed the
// sequentially consistent system fence that used to stand here, and the two on the publishing side, as most of a
// 56 us reduce launch); the CTA barrier that follows extends the ordering to every thread of the CTA.  Returns false
// (to every thread of the CTA) when a wait gave up: the caller must then POISON what it produces (NaN) instead of
// consuming a stale or partial block.
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
    (void)ld_acquire_sys(flags + threadIdx.x);
  }
  __syncthreads();
  return s_timed_out == 0 && ld_relaxed_sys(status) == 0u;
}

constexpr unsigned kPeerPoison = 0x7fffffffu;  // NaN as fp32 and as a pair of bf16

// `n_ctas` CTAs cooperate on a channel phase: the CTA's stores are done (barrier), thread 0 fences them at system scope
// and counts the CTA in; the last of the `n_ctas` to get here fences once more and publishes `epoch` in slot `rank` of
// the flag row at byte offset `flag_off` of every peer buffer (and, optionally, advances the local epoch).
__device__ __forceinline__ void peer_publish_when_done(const PeerTable& peers, int world, int rank, size_t flag_off,
                                                       unsigned epoch, unsigned* counter, unsigned n_ctas,
                                                       unsigned* epoch_word) {
  __syncthreads();
  if (threadIdx.x == 0) {
    const bool publish = flag_off != (size_t)-1;  // closing a round touches local words only: no system-scope ordering needed
    if (publish) fence_release_sys(); else __threadfence();
    const unsigned done = atomicAdd(counter, 1u);
    if (done == n_ctas - 1) {
      if (publish) {
        fence_release_sys();  // acquire of the other CTAs' counts + release of everything before the flags
        for (int r = 0; r < world; ++r)
          st_relaxed_sys(reinterpret_cast<unsigned*>(static_cast<char*>(peers.base[(rank + r) % world]) + flag_off) + rank, epoch);
      } else {
        __threadfence();
      }
      *counter = 0;
      if (epoch_word) *epoch_word = epoch;
    }
  }
}

// One rank's share of the two-shot mean all-reduce (see peer.cu): sum slice `rank` of every peer's contribution area in
// fp32, in rank order (bit-identical on all ranks), and store the mean into every peer's result area.  Executed by
// `n_ctas` CTAs (this one is `cta`); vectors below f32_begin hold 8 bf16 values, the others 4 fp32 values.
template <int RB, int U>
__device__ __forceinline__ void peer_reduce_slice(const PeerTable& peers, int world, int rank, long long n_vec,
                                                  long long f32_begin, float scale, bool ok, int cta, int n_ctas) {
  const long long per = (n_vec + world - 1) / world;
  const long long lo = per * rank, hi = min(n_vec, lo + per);
  const size_t in_off = kPeerCtrlBytes, out_off = kPeerCtrlBytes + (size_t)n_vec * sizeof(uint4);
  const long long stride = (long long)n_ctas * blockDim.x;
  for (long long i = lo + (long long)cta * blockDim.x + threadIdx.x; i < hi; i += U * stride) {
    float acc[U][8];
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[u][k] = 0.f;
    for (int r0 = 0; r0 < world; r0 += RB) {
      uint4 v[U][RB];
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int j = 0; j < RB; ++j)
          if (r0 + j < world && i + u * stride < hi)
            v[u][j] = ld_sys_v4(reinterpret_cast<const uint4*>(static_cast<char*>(peers.base[r0 + j]) + in_off) + i + u * stride);
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int j = 0; j < RB; ++j)
          if (r0 + j < world && i + u * stride < hi) {
            const uint4& q = v[u][j];
            if (i + u * stride >= f32_begin) {
              acc[u][0] += __uint_as_float(q.x); acc[u][1] += __uint_as_float(q.y);
              acc[u][2] += __uint_as_float(q.z); acc[u][3] += __uint_as_float(q.w);
            } else {
              const float2 a = unpack_bf16x2(q.x), b = unpack_bf16x2(q.y), c = unpack_bf16x2(q.z), d = unpack_bf16x2(q.w);
              acc[u][0] += a.x; acc[u][1] += a.y; acc[u][2] += b.x; acc[u][3] += b.y;
              acc[u][4] += c.x; acc[u][5] += c.y; acc[u][6] += d.x; acc[u][7] += d.y;
            }
          }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (i + u * stride >= hi) continue;
      uint4 o;
      if (!ok)
        o = make_uint4(kPeerPoison, kPeerPoison, kPeerPoison, kPeerPoison);  // a contribution never arrived: NaN, not a partial mean
      else if (i + u * stride >= f32_begin)
        o = make_uint4(__float_as_uint(acc[u][0] * scale), __float_as_uint(acc[u][1] * scale), __float_as_uint(acc[u][2] * scale),
                       __float_as_uint(acc[u][3] * scale));
      else
        o = make_uint4(pack_bf16x2(acc[u][0] * scale, acc[u][1] * scale), pack_bf16x2(acc[u][2] * scale, acc[u][3] * scale),
                       pack_bf16x2(acc[u][4] * scale, acc[u][5] * scale), pack_bf16x2(acc[u][6] * scale, acc[u][7] * scale));
      for (int r = 0; r < world; ++r) {
        const int dst = (rank + r) % world;
        reinterpret_cast<uint4*>(static_cast<char*>(peers.base[dst]) + out_off)[i + u * stride] = o;
      }
    }
  }
}

// ---- the same protocol pieces for a GROUP of warps inside a CTA that is busy with something else ----
// (the epilogue warps of a GEMM CTA while its first accumulator is being computed: gemm_sm100.cuh).  `tid` in
// [0, nthr) is the thread's index in the group, `bar_id` a named barrier reserved for the group (nthr % 32 == 0).
__device__ __forceinline__ void group_sync(int bar_id, int nthr) {
  asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(nthr) : "memory");
}
__device__ __forceinline__ bool wait_flags_group(const unsigned* flags, int world, unsigned epoch, unsigned* status, int tid,
                                                 int nthr, int bar_id) {
  __shared__ int s_group_timed_out;
  if (tid == 0) s_group_timed_out = 0;
  group_sync(bar_id, nthr);
  if (tid < world) {
    const unsigned long long t0 = global_ns();
    while ((int)(ld_relaxed_sys(flags + tid) - epoch) < 0) {
      if (global_ns() - t0 > kPeerTimeoutNs) {
        atomicExch(status, 1u + tid);
        s_group_timed_out = 1;
        break;
      }
      __nanosleep(64);
    }
    (void)ld_acquire_sys(flags + tid);
  }
  group_sync(bar_id, nthr);
  return s_group_timed_out == 0 && ld_relaxed_sys(status) == 0u;
}

// One rank's share of the two-shot mean all-reduce for a channel that carries fp32 only (f32_begin == 0), light enough
// on registers to run inside a kernel compiled for another job (96 registers per thread in the GEMM): RB * U = 8
// 16-byte loads in flight per thread, 4 accumulators per vector.  `part` of `n_parts` groups of `nthr` threads.
template <int RB, int U>
__device__ __forceinline__ void peer_reduce_slice_f32(const PeerTable& peers, int world, int rank, long long n_vec, float scale,
                                                      bool ok, int part, int n_parts, int tid, int nthr) {
  const long long per = (n_vec + world - 1) / world;
  const long long lo = per * rank, hi = min(n_vec, lo + per);
  const size_t in_off = kPeerCtrlBytes, out_off = kPeerCtrlBytes + (size_t)n_vec * sizeof(uint4);
  const long long stride = (long long)n_parts * nthr;
  for (long long i = lo + (long long)part * nthr + tid; i < hi; i += U * stride) {
    float4 acc[U];
#pragma unroll
    for (int u = 0; u < U; ++u) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r0 = 0; r0 < world; r0 += RB) {
      uint4 v[U][RB];
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int j = 0; j < RB; ++j)
          if (r0 + j < world && i + u * stride < hi)
            v[u][j] = ld_sys_v4(reinterpret_cast<const uint4*>(static_cast<char*>(peers.base[r0 + j]) + in_off) + i + u * stride);
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int j = 0; j < RB; ++j)
          if (r0 + j < world && i + u * stride < hi) {
            acc[u].x += __uint_as_float(v[u][j].x); acc[u].y += __uint_as_float(v[u][j].y);
            acc[u].z += __uint_as_float(v[u][j].z); acc[u].w += __uint_as_float(v[u][j].w);
          }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (i + u * stride >= hi) continue;
      const uint4 o = ok ? make_uint4(__float_as_uint(acc[u].x * scale), __float_as_uint(acc[u].y * scale),
                                      __float_as_uint(acc[u].z * scale), __float_as_uint(acc[u].w * scale))
                         : make_uint4(kPeerPoison, kPeerPoison, kPeerPoison, kPeerPoison);
      for (int r = 0; r < world; ++r)
        reinterpret_cast<uint4*>(static_cast<char*>(peers.base[(rank + r) % world]) + out_off)[i + u * stride] = o;
    }
  }
}

// The contribution phase + reduce phase of one channel round, run by the otherwise idle epilogue warps of EVERY CTA of
// a GEMM launch before their first accumulator is ready (gemm_sm100.cuh): no SM is taken from the GEMM and no second
// kernel competes with the persistent one.  This rank's contribution is complete before the launch (stream order),
// so CTA 0 announces it right away; every group then waits for all ranks' announcements, reduces its part of this
// rank's slice, and the last group to finish publishes the phase-1 flags.  The channel carries fp32 only.
struct GemmCommReduce {
  PeerTable peers;
  int world, rank;
  long long n_vec;
  float scale;
};
static __device__ __noinline__ void comm_reduce_role(const GemmCommReduce& c, int cta, int n_ctas, int tid, int nthr, int bar_id) {
  unsigned* ctrl = static_cast<unsigned*>(c.peers.base[c.rank]);
  const unsigned epoch = ctrl[0] + 1;
  if (cta == 0 && tid == 0) {
    fence_release_sys();
    for (int r = 0; r < c.world; ++r)
      st_relaxed_sys(reinterpret_cast<unsigned*>(static_cast<char*>(c.peers.base[(c.rank + r) % c.world]) + peer_flag_row_off(0)) + c.rank, epoch);
  }
  const bool ok = wait_flags_group(reinterpret_cast<const unsigned*>(static_cast<char*>(c.peers.base[c.rank]) + peer_flag_row_off(0)),
                                   c.world, epoch, ctrl + 4, tid, nthr, bar_id);
  if (c.world <= 2) peer_reduce_slice_f32<2, 4>(c.peers, c.world, c.rank, c.n_vec, c.scale, ok, cta, n_ctas, tid, nthr);
  else if (c.world <= 4) peer_reduce_slice_f32<4, 2>(c.peers, c.world, c.rank, c.n_vec, c.scale, ok, cta, n_ctas, tid, nthr);
  else peer_reduce_slice_f32<8, 1>(c.peers, c.world, c.rank, c.n_vec, c.scale, ok, cta, n_ctas, tid, nthr);
  // this group's stores are done -> count the CTA in; the last one publishes phase 1 to every peer
  group_sync(bar_id, nthr);
  if (tid == 0) {
    fence_release_sys();
    const unsigned done = atomicAdd(ctrl + 2, 1u);
    if (done == (unsigned)n_ctas - 1) {
      fence_release_sys();
      for (int r = 0; r < c.world; ++r)
        st_relaxed_sys(reinterpret_cast<unsigned*>(static_cast<char*>(c.peers.base[(c.rank + r) % c.world]) + peer_flag_row_off(1)) + c.rank, epoch);
      ctrl[2] = 0;
    }
  }
}

}  // namespace p2t
