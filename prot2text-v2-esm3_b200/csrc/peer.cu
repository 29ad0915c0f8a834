// Exchange steps of the sharded contrastive step over NVLink peer memory (no NCCL on the data path):
//   * all-gather of the unit-norm text embeddings (global negatives), push model: every rank stores
//     its block straight into every peer's gather buffer and raises a per-source flag there;
//   * mean all-reduce of the adapter weight gradients (what DDP does in the reference, scripts/train_contrast.py:611-614),
//     two-shot: rank k sums slice k out of every peer's buffer (P2P loads, fixed rank order = identical bits on all
//     ranks) and stores the mean back into every peer's result buffer.
// Everything is plain kernels on the caller's stream with device-resident epochs, so the whole sharded step
// (exchange included) can be captured in one CUDA graph.  Peer buffers are cudaMalloc'ed here (they must be
// exportable with cudaIpcGetMemHandle, which torch's caching allocator does not guarantee) and mapped by the peers
// with cudaIpcOpenMemHandle; the handles travel through torch.distributed on the host side (peer.py).
//
// Flag protocol.  A channel owns, in its peer buffer, `world` arrival words per phase; a writer publishes with
// data stores -> CTA barrier -> thread 0: fence.acq_rel.sys, count the CTA in -> (last CTA) fence, flag := epoch
// (relaxed system-scope stores); a reader polls the flag with relaxed system-scope loads until flag - epoch >= 0,
// re-reads it once with ld.acquire.sys (an L1 invalidate, no MEMBAR), and then reads the data with L1-bypassing
// loads.  (The sequentially consistent __threadfence_system() this started with was most of the channel kernels'
// time: profiles/r02_summary.md.)  The epoch lives in device
// memory and is advanced by the channel's last kernel of a round, so graph replays need no host-side argument.
// A reader that waits longer than `timeout_ns` gives up, records it in `status` and lets the stream drain (the host
// raises on the next status check) instead of hanging the GPU.
#include "common.h"
#include "mathfn.cuh"
#include "rows.h"
#include "peer_dev.cuh"

namespace p2t {
namespace {

__device__ __forceinline__ void publish_when_grid_done(const PeerTable& peers, int world, int rank, size_t flag_off,
                                                       unsigned epoch, unsigned* grid_counter, unsigned* epoch_word) {
  peer_publish_when_done(peers, world, rank, flag_off, epoch, grid_counter, gridDim.x, epoch_word);
}

constexpr size_t kCtrlBytes = kPeerCtrlBytes;
__host__ __device__ constexpr size_t flag_row_off(int phase) { return peer_flag_row_off(phase); }
constexpr unsigned kPoison = kPeerPoison;

// ------------------------------------------------------------------------------------------------
// all-gather, push half: slot (epoch parity, rank) of every peer's buffer := src
// ------------------------------------------------------------------------------------------------
// U independent 16-byte loads per thread and trip, each fanned out to every peer: (grid x 256 x U x 16) bytes per
// destination are in flight against an NVLink round trip of ~2 us (one load per thread and 74 CTAs moved 67 MB at
// 115 GB/s: latency-bound, not link-bound).
template <int U>
__global__ void __launch_bounds__(256)
peer_allgather_push_kernel(PeerTable peers, int world, int rank, const uint4* __restrict__ src, long long vecs_per_rank) {
  unsigned* ctrl = static_cast<unsigned*>(peers.base[rank]);
  const unsigned epoch = ctrl[0] + 1;
  const size_t slot_off = kCtrlBytes + ((size_t)(epoch & 1) * world + rank) * (size_t)vecs_per_rank * sizeof(uint4);
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < vecs_per_rank; i += U * stride) {
    uint4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (i + u * stride < vecs_per_rank) v[u] = __ldg(src + i + u * stride);
    for (int r = 0; r < world; ++r) {
      const int dst = (rank + r) % world;  // ranks start on different links
      uint4* out = reinterpret_cast<uint4*>(static_cast<char*>(peers.base[dst]) + slot_off);
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (i + u * stride < vecs_per_rank) out[i + u * stride] = v[u];
    }
  }
  publish_when_grid_done(peers, world, rank, flag_row_off(0), epoch, ctrl + 1, nullptr);
}

// all-gather, arrival half: wait for every rank's block of this epoch, copy the gathered rows out, close the round.
// A wait that gave up (or a channel whose status word is already set) writes NaN instead of the rows: the loss of
// the step becomes NaN, which no caller can miss, instead of a silently stale set of negatives.
__global__ void __launch_bounds__(256)
peer_allgather_wait_kernel(PeerTable peers, int world, int rank, uint4* __restrict__ dst, long long vecs_per_rank) {
  unsigned* ctrl = static_cast<unsigned*>(peers.base[rank]);
  const unsigned epoch = ctrl[0] + 1;
  const bool ok = wait_flags(reinterpret_cast<const unsigned*>(static_cast<char*>(peers.base[rank]) + flag_row_off(0)), world, epoch, ctrl + 4);
  const uint4* slots = reinterpret_cast<const uint4*>(static_cast<char*>(peers.base[rank]) + kCtrlBytes +
                                                      (size_t)(epoch & 1) * world * (size_t)vecs_per_rank * sizeof(uint4));
  const long long total = vecs_per_rank * world;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
    dst[i] = ok ? ld_sys_v4(slots + i) : make_uint4(kPoison, kPoison, kPoison, kPoison);
  publish_when_grid_done(peers, world, rank, (size_t)-1, epoch, ctrl + 2, ctrl);
}

// ------------------------------------------------------------------------------------------------
// mean all-reduce of a flat bf16 vector (adapter weight gradients), two-shot over peer memory
//   buffer layout after the control block: in[n_vec] (this rank's contribution), out[n_vec] (reduced result)
// ------------------------------------------------------------------------------------------------
// phase 0: this rank's contribution is in place (stream order) -> tell everybody
__global__ void peer_allreduce_ready_kernel(PeerTable peers, int world, int rank) {
  unsigned* ctrl = static_cast<unsigned*>(peers.base[rank]);
  const unsigned epoch = ctrl[0] + 1;
  publish_when_grid_done(peers, world, rank, flag_row_off(0), epoch, ctrl + 1, nullptr);
}

// phase 1: reduce my slice out of every peer's `in`, store the mean into every peer's `out`.
// RB ranks x U vectors = 8 independent 16-byte loads in flight per thread whatever the world size, two CTAs per SM
// (an NVLink round trip is ~2 us); ranks are always added in rank order, so every rank computes bit-identical sums.
// Vectors below `f32_begin` hold 8 bf16 values (mean stored as bf16), vectors from `f32_begin` on hold 4 fp32 values
// (mean stored as fp32: the bias gradients travel unrounded and are rounded to bf16 once, after the mean).
template <int RB, int U>
__global__ void __launch_bounds__(256, 2)
peer_allreduce_reduce_kernel(PeerTable peers, int world, int rank, long long n_vec, long long f32_begin, float scale, int announce) {
  unsigned* ctrl = static_cast<unsigned*>(peers.base[rank]);
  const unsigned epoch = ctrl[0] + 1;
  if (announce && blockIdx.x == 0 && threadIdx.x == 0) {
    // phase 0 folded into this launch: the contribution was complete before the launch (stream order)
    fence_release_sys();
    for (int r = 0; r < world; ++r)
      st_relaxed_sys(reinterpret_cast<unsigned*>(static_cast<char*>(peers.base[(rank + r) % world]) + flag_row_off(0)) + rank, epoch);
  }
  const bool ok = wait_flags(reinterpret_cast<const unsigned*>(static_cast<char*>(peers.base[rank]) + flag_row_off(0)), world, epoch, ctrl + 4);
  peer_reduce_slice<RB, U>(peers, world, rank, n_vec, f32_begin, scale, ok, blockIdx.x, gridDim.x);
  publish_when_grid_done(peers, world, rank, flag_row_off(1), epoch, ctrl + 2, nullptr);
}

// The same phase for a channel that carries fp32 only (the adapter's gradient reducer): 4 accumulators per vector
// instead of 8, so two CTAs fit an SM and the whole slice is one resident wave (the generic kernel needs 164 registers
// at world 2: one CTA per SM, two waves — 89 us against 42 us for the adapter's 55 MB, ncu, world of one).
template <int RB, int U>
__global__ void __launch_bounds__(256, 2)
peer_allreduce_reduce_f32_kernel(PeerTable peers, int world, int rank, long long n_vec, float scale, int announce) {
  unsigned* ctrl = static_cast<unsigned*>(peers.base[rank]);
  const unsigned epoch = ctrl[0] + 1;
  if (announce && blockIdx.x == 0 && threadIdx.x == 0) {
    fence_release_sys();
    for (int r = 0; r < world; ++r)
      st_relaxed_sys(reinterpret_cast<unsigned*>(static_cast<char*>(peers.base[(rank + r) % world]) + flag_row_off(0)) + rank, epoch);
  }
  const bool ok = wait_flags(reinterpret_cast<const unsigned*>(static_cast<char*>(peers.base[rank]) + flag_row_off(0)), world, epoch, ctrl + 4);
  peer_reduce_slice_f32<RB, U>(peers, world, rank, n_vec, scale, ok, blockIdx.x, gridDim.x, threadIdx.x, blockDim.x);
  publish_when_grid_done(peers, world, rank, flag_row_off(1), epoch, ctrl + 2, nullptr);
}

// phase 2: every slice of `out` has arrived -> (optionally) copy it to the caller's flat buffer, close the round.
// After a timed-out wait the result area is overwritten with NaN (see peer_allgather_wait_kernel).
__global__ void __launch_bounds__(256)
peer_allreduce_wait_kernel(PeerTable peers, int world, int rank, long long n_vec, uint4* __restrict__ dst) {
  unsigned* ctrl = static_cast<unsigned*>(peers.base[rank]);
  const unsigned epoch = ctrl[0] + 1;
  const bool ok = wait_flags(reinterpret_cast<const unsigned*>(static_cast<char*>(peers.base[rank]) + flag_row_off(1)), world, epoch, ctrl + 4);
  uint4* out = reinterpret_cast<uint4*>(static_cast<char*>(peers.base[rank]) + kCtrlBytes) + n_vec;
  if (dst != nullptr || !ok) {
    const uint4 poison = make_uint4(kPoison, kPoison, kPoison, kPoison);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += (long long)gridDim.x * blockDim.x) {
      if (!ok) out[i] = poison;
      if (dst != nullptr) dst[i] = ok ? ld_sys_v4(out + i) : poison;
    }
  }
  publish_when_grid_done(peers, world, rank, (size_t)-1, epoch, ctrl + 3, ctrl);
}

int fill_table(PeerTable& t, void* const* peers, int world, int rank, const char* what) {
  if (!peers) return set_error(-1, "%s: null peer table", what);
  if (world < 1 || world > kPeerMaxWorld) return set_error(-1, "%s: world %d outside [1, %d]", what, world, kPeerMaxWorld);
  if (rank < 0 || rank >= world) return set_error(-1, "%s: rank %d outside [0, %d)", what, rank, world);
  for (int r = 0; r < world; ++r) {
    if (!peers[r]) return set_error(-1, "%s: peer %d not mapped", what, r);
    t.base[r] = peers[r];
  }
  return 0;
}

// Waiting CTAs depend only on flags raised by kernels that precede this one in some rank's stream, never on each other,
// so the grid need not be co-resident; `per_sm` bounds how much of the GPU a channel kernel may take.
int grid_for(long long vecs, int per_sm_x2 = 1) {
  const long long want = (vecs + 255) / 256;
  const long long cap = sm_count() * per_sm_x2 / 2 > 0 ? sm_count() * per_sm_x2 / 2 : 1;
  return (int)(want < 1 ? 1 : (want > cap ? cap : want));
}

}  // namespace

size_t peer_ctrl_bytes() { return kCtrlBytes; }

int peer_alloc(size_t bytes, void** dptr, unsigned char* handle64) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  if (!dptr || !handle64 || bytes == 0) return set_error(-1, "p2t_peer_alloc: bad argument");
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e != cudaSuccess) return set_error((int)e, "p2t_peer_alloc: cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e));
  e = cudaMemset(p, 0, bytes);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    return set_error((int)e, "p2t_peer_alloc: %s", cudaGetErrorString(e));
  }
  memcpy(handle64, &h, 64);
  *dptr = p;
  return 0;
}

int peer_open(const unsigned char* handle64, void** dptr) {
  if (!dptr || !handle64) return set_error(-1, "p2t_peer_open: bad argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  void* p = nullptr;
  cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) return set_error((int)e, "p2t_peer_open: cudaIpcOpenMemHandle: %s", cudaGetErrorString(e));
  *dptr = p;
  return 0;
}

int peer_close(void* dptr) {
  cudaError_t e = cudaIpcCloseMemHandle(dptr);
  if (e != cudaSuccess) return set_error((int)e, "p2t_peer_close: %s", cudaGetErrorString(e));
  return 0;
}

int peer_free(void* dptr) {
  cudaError_t e = cudaFree(dptr);
  if (e != cudaSuccess) return set_error((int)e, "p2t_peer_free: %s", cudaGetErrorString(e));
  return 0;
}

// Collective re-initialisation of a channel after a timed-out round (every rank, between two host-side barriers):
// epochs, grid counters, status word and arrival flags go back to zero.
int peer_reset(void* base, cudaStream_t st) {
  if (!base) return set_error(-1, "p2t_peer_reset: null pointer");
  cudaError_t e = cudaMemsetAsync(base, 0, kCtrlBytes, st);
  if (e != cudaSuccess) return set_error((int)e, "p2t_peer_reset: %s", cudaGetErrorString(e));
  return 0;
}

int peer_allgather(void* const* peers, int world, int rank, const void* src, long long bytes_per_rank, void* dst,
                   int phases, cudaStream_t st) {
  PeerTable t{};
  if (int r = fill_table(t, peers, world, rank, "p2t_peer_allgather")) return r;
  if (bytes_per_rank <= 0 || bytes_per_rank % 16) return set_error(-1, "p2t_peer_allgather: bytes_per_rank must be a positive multiple of 16");
  const long long vecs = bytes_per_rank / 16;
  if (phases & 1) {
    if (!src) return set_error(-1, "p2t_peer_allgather: null src");
    if (vecs >= (1 << 18))  // >= 4 MB per rank: bandwidth matters, take every SM twice over
      peer_allgather_push_kernel<4><<<grid_for(vecs / 4, 4), 256, 0, st>>>(t, world, rank, static_cast<const uint4*>(src), vecs);
    else
      peer_allgather_push_kernel<1><<<grid_for(vecs), 256, 0, st>>>(t, world, rank, static_cast<const uint4*>(src), vecs);
    if (int r = check_launch("peer_allgather_push_kernel", st)) return r;
  }
  if (phases & 2) {
    if (!dst) return set_error(-1, "p2t_peer_allgather: null dst");
    peer_allgather_wait_kernel<<<grid_for(vecs * world), 256, 0, st>>>(t, world, rank, static_cast<uint4*>(dst), vecs);
    if (int r = check_launch("peer_allgather_wait_kernel", st)) return r;
  }
  return 0;
}

int peer_allreduce_mean(void* const* peers, int world, int rank, long long n_bytes, long long f32_from_byte, void* dst, int phases,
                        cudaStream_t st) {
  PeerTable t{};
  if (int r = fill_table(t, peers, world, rank, "p2t_peer_allreduce_mean_bf16")) return r;
  if (n_bytes <= 0 || n_bytes % 16) return set_error(-1, "p2t_peer_allreduce_mean_bf16: n_bytes must be a positive multiple of 16");
  const long long n_vec = n_bytes / 16;
  if (f32_from_byte < 0 || f32_from_byte > n_bytes) f32_from_byte = n_bytes;
  if (f32_from_byte % 16) return set_error(-1, "p2t_peer_allreduce_mean: the fp32 part must start on a 16-byte boundary");
  const long long f32_begin = f32_from_byte / 16;
  const int announce_in_reduce = (phases & 3) == 3;  // announce + reduce in one launch (CTA 0 raises the flags first)
  if ((phases & 1) && !announce_in_reduce) {
    peer_allreduce_ready_kernel<<<1, 32, 0, st>>>(t, world, rank);
    if (int r = check_launch("peer_allreduce_ready_kernel", st)) return r;
  }
  if (phases & 2) {
    const float scale = 1.f / (float)world;
    const long long slice = (n_vec + world - 1) / world;
    if (f32_begin == 0) {
      // two resident CTAs per SM, one wave
      if (world <= 2) peer_allreduce_reduce_f32_kernel<2, 8><<<grid_for(slice / 8, 4), 256, 0, st>>>(t, world, rank, n_vec, scale, announce_in_reduce);
      else if (world <= 4) peer_allreduce_reduce_f32_kernel<4, 4><<<grid_for(slice / 4, 4), 256, 0, st>>>(t, world, rank, n_vec, scale, announce_in_reduce);
      else peer_allreduce_reduce_f32_kernel<8, 2><<<grid_for(slice / 2, 4), 256, 0, st>>>(t, world, rank, n_vec, scale, announce_in_reduce);
    } else if (world <= 2) {
      peer_allreduce_reduce_kernel<2, 4><<<grid_for(slice / 4, 4), 256, 0, st>>>(t, world, rank, n_vec, f32_begin, scale, announce_in_reduce);
    } else if (world <= 4) {
      peer_allreduce_reduce_kernel<4, 2><<<grid_for(slice / 2, 4), 256, 0, st>>>(t, world, rank, n_vec, f32_begin, scale, announce_in_reduce);
    } else {
      peer_allreduce_reduce_kernel<8, 1><<<grid_for(slice, 4), 256, 0, st>>>(t, world, rank, n_vec, f32_begin, scale, announce_in_reduce);
    }
    if (int r = check_launch("peer_allreduce_reduce_kernel", st)) return r;
  }
  if (phases & 4) {
    // zero-copy callers (dst == NULL) only wait and close the round: one CTA (every CTA pays two system fences)
    peer_allreduce_wait_kernel<<<dst ? grid_for(n_vec, 4) : 1, 256, 0, st>>>(t, world, rank, n_vec, static_cast<uint4*>(dst));
    if (int r = check_launch("peer_allreduce_wait_kernel", st)) return r;
  }
  return 0;
}

}  // namespace p2t
