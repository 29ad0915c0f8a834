// HBM-bound kernels over ragged residue rows: mask -> row plan, row gather (packing), masked
// mean/std ("mix") pooling forward with per-row L2 normalisation folded in, its backward, L2
// normalisation of the pooled embeddings, and column sums for the bias gradients.
//
// Reference sites: scripts/train_contrast.py:198-248 (readout_embeddings), :354/:365 (F.normalize),
// models/modeling_esm2llama_instruct.py:67 (per-residue F.normalize) and their autograd.
//
// Layout: residue rows of all sequences are PACKED back to back ([sum L_b][D], bf16); `seq_off[b]`
// is the first packed row of sequence b (B+1 entries).  Every streaming kernel moves 16-byte
// (8 x bf16) vectors per thread with the column index fastest across the warp, accumulates in fp32
// and reduces in a fixed order (no floating-point atomics), so results are run-to-run identical.
#include "common.h"
#include "mathfn.cuh"
#include "ptx.cuh"
#include "rows.h"
#include <cuda.h>
#include <algorithm>

namespace p2t {

constexpr float kEpsNorm = 1e-12f;  // F.normalize eps

__device__ __forceinline__ bool mask_at(const void* mask, int mask_bytes, long long i) {
  switch (mask_bytes) {
    case 1: return reinterpret_cast<const uint8_t*>(mask)[i] != 0;
    case 4: return reinterpret_cast<const int32_t*>(mask)[i] != 0;
    default: return reinterpret_cast<const long long*>(mask)[i] != 0;
  }
}

// number of non-zero mask elements of one row, counted by a warp: 16-byte loads, four in flight per lane
__device__ __forceinline__ int warp_count_row(const void* mask, int mask_bytes, long long first, int L, int lane) {
  int c = 0;
  const char* base = reinterpret_cast<const char*>(mask) + first * mask_bytes;
  const long long bytes = (long long)L * mask_bytes;
  if (((reinterpret_cast<uintptr_t>(base) | (uintptr_t)bytes) & 15) == 0) {
    const uint4* vp = reinterpret_cast<const uint4*>(base);
    const int nvec = (int)(bytes >> 4);
    auto nz = [&](const uint4& u) {
      if (mask_bytes == 8) return ((u.x | u.y) != 0) + ((u.z | u.w) != 0);
      if (mask_bytes == 4) return (u.x != 0) + (u.y != 0) + (u.z != 0) + (u.w != 0);
      return __popc(__vcmpne4(u.x, 0u) & 0x01010101u) + __popc(__vcmpne4(u.y, 0u) & 0x01010101u) +
             __popc(__vcmpne4(u.z, 0u) & 0x01010101u) + __popc(__vcmpne4(u.w, 0u) & 0x01010101u);
    };
    int v = lane;
    for (; v + 96 < nvec; v += 128) {
      const uint4 a = __ldg(vp + v), b = __ldg(vp + v + 32), d = __ldg(vp + v + 64), e = __ldg(vp + v + 96);
      c += nz(a) + nz(b) + nz(d) + nz(e);
    }
    for (; v < nvec; v += 32) c += nz(__ldg(vp + v));
  } else {
    for (int r = lane; r < L; r += 32) c += mask_at(mask, mask_bytes, first + r) ? 1 : 0;
  }
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  return c;
}

template <int THREADS>
__device__ __forceinline__ float block_sum(float v, float* smem /*[THREADS/32]*/) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) smem[w] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int i = 0; i < THREADS / 32; ++i) t += smem[i];
  return t;
}

// ------------------------------------------------------------------------------------------------
// row plan: counts -> offsets -> source row list
// ------------------------------------------------------------------------------------------------
__global__ void plan_count_kernel(const void* mask, int mask_bytes, int B, int L, int* counts) {
  const int b = blockIdx.x;
  int c = 0;
  for (int r = threadIdx.x; r < L; r += blockDim.x) c += mask_at(mask, mask_bytes, (long long)b * L + r) ? 1 : 0;
  __shared__ int sh[8];
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += sh[i];
    counts[b] = t;
  }
}

// grid B, block 256.  seq_off[B+1], chunk_off[B+1] (chunks of `rc` rows), n_rows[0] = total,
// row_src[i] = flat source row (b*L + r) of packed row i.
template <bool COUNT_HERE>
__global__ void plan_fill_kernel(const void* mask, int mask_bytes, int B, int L, int* counts, int rc,
                                 int* seq_off, int* chunk_off, int* n_rows, int* row_src, int* chunk_seq) {
  const int b = blockIdx.x;
  __shared__ int sh[2][8];
  __shared__ int s_base, s_cbase, s_cnt;
  int off = 0, coff = 0;
  if constexpr (COUNT_HERE) {
    // one launch instead of two (small batches): this block counts sequences 0..b itself, one warp per sequence
    // (block b reads (b + 1) * L mask words: <= 256 KB at config 2, L2-resident after the first block touched them)
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    for (int i = w; i <= b; i += (int)(blockDim.x >> 5)) {
      const int c = warp_count_row(mask, mask_bytes, (long long)i * L, L, lane);
      if (lane == 0) {
        if (i < b) { off += c; coff += (c + rc - 1) / rc; }
        else { s_cnt = c; counts[b] = c; }
      }
    }
  } else {
    for (int i = threadIdx.x; i < b; i += blockDim.x) {
      const int c = counts[i];
      off += c;
      coff += (c + rc - 1) / rc;
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    off += __shfl_xor_sync(0xffffffffu, off, o);
    coff += __shfl_xor_sync(0xffffffffu, coff, o);
  }
  if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = off; sh[1][threadIdx.x >> 5] = coff; }
  __syncthreads();
  const int cnt_b = COUNT_HERE ? s_cnt : counts[b];
  if (threadIdx.x == 0) {
    int t0 = 0, t1 = 0;
    for (int i = 0; i < 8; ++i) { t0 += sh[0][i]; t1 += sh[1][i]; }
    seq_off[b] = t0;
    chunk_off[b] = t1;
    if (b == B - 1) {
      const int c = cnt_b;
      seq_off[B] = t0 + c;
      chunk_off[B] = t1 + (c + rc - 1) / rc;
      n_rows[0] = t0 + c;
    }
    s_base = t0;
    s_cbase = t1;
  }
  __syncthreads();
  if (chunk_seq != nullptr) {  // one descriptor per pooling chunk: {first packed row, end row, sequence, 0}
    const int cnt = cnt_b;
    const int nchunks = (cnt + rc - 1) / rc;
    int4* desc = reinterpret_cast<int4*>(chunk_seq);
    for (int j = threadIdx.x; j < nchunks; j += blockDim.x)
      desc[s_cbase + j] = make_int4(s_base + j * rc, s_base + min(cnt, (j + 1) * rc), b, 0);
  }
  if (row_src == nullptr) return;
  // ordered compaction of the valid positions of this sequence
  int base = s_base;
  __shared__ int wsum[8];
  for (int r0 = 0; r0 < L; r0 += blockDim.x) {
    const int r = r0 + threadIdx.x;
    const bool v = (r < L) && mask_at(mask, mask_bytes, (long long)b * L + r);
    const unsigned ball = __ballot_sync(0xffffffffu, v);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) wsum[w] = __popc(ball);
    __syncthreads();
    int before = 0, total = 0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) {
      if (i < w) before += wsum[i];
      total += wsum[i];
    }
    if (v) row_src[base + before + __popc(ball & ((1u << lane) - 1u))] = b * L + r;
    base += total;
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// gather: out[i] = src[row_src[i]] for i < n; zeros for n <= i < min(cap, roundup(n, 256))
// ------------------------------------------------------------------------------------------------
__global__ void gather_rows_kernel(const __nv_bfloat16* __restrict__ src, long long ld_src,
                                   const int* __restrict__ row_src, const int* __restrict__ n_rows, int cap, int D,
                                   __nv_bfloat16* __restrict__ out) {
  const int n = min(*n_rows, cap);
  const int n_pad = min(cap, (n + 255) & ~255);
  const int vec_per_row = D >> 3;
  const long long total = (long long)n_pad * vec_per_row;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int row = (int)(i / vec_per_row), v = (int)(i % vec_per_row);
    uint4 val = make_uint4(0, 0, 0, 0);
    if (row < n) val = __ldg(reinterpret_cast<const uint4*>(src + (long long)row_src[row] * ld_src) + v);
    reinterpret_cast<uint4*>(out + (long long)row * D)[v] = val;
  }
}

// ------------------------------------------------------------------------------------------------
// pooling forward, phase 1: per (row chunk, column) shifted partial moments
//   value of element = src[row][col] * inv_norm[row]   (inv_norm from rowsq partials; 1 if rowsq == null)
//   partial[chunk][col] = (mean_c, M2_c) over the chunk's rows
// grid (max_chunks, ceil(D / (128*8))), block 128
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int find_segment(const int* __restrict__ off, int n, int x) {
  // largest b in [0, n) with off[b] <= x   (off non-decreasing, off[0] == 0)
  int lo = 0, hi = n - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (off[mid] <= x) lo = mid; else hi = mid - 1;
  }
  return lo;
}

// inverse L2 norm of each adapter output row from the fc2 epilogue's partial sums of squares
// rowsq is [nblk][cap] (partial index major).  Block = 32 rows (lanes) x 8 groups (warps): group g sums partials
// g, g+8, ... of its row with all loads in flight, the 8 group sums meet in shared memory in a fixed order — two
// dependent memory trips per row instead of nblk/4.
__global__ void __launch_bounds__(256)
row_inv_norm_kernel(const float* __restrict__ rowsq, int nblk, const int* __restrict__ n_rows, int cap,
                    float* __restrict__ inv_norm) {
  __shared__ float part[8][33];
  const int n = min(*n_rows, cap);
  const int lane = threadIdx.x & 31, g = threadIdx.x >> 5;
  for (int r0 = blockIdx.x * 32; r0 < n; r0 += gridDim.x * 32) {
    const int r = r0 + lane;
    float s = 0.f;
    if (r < n) {
      float v[8];
      for (int j0 = g; j0 < nblk; j0 += 64) {
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = (j0 + 8 * u < nblk) ? rowsq[(long long)(j0 + 8 * u) * cap + r] : 0.f;
#pragma unroll
        for (int u = 0; u < 8; ++u) s += v[u];
      }
    }
    part[g][lane] = s;
    __syncthreads();
    if (g == 0 && r < n) {
      float tot = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) tot += part[k][lane];
      inv_norm[r] = 1.f / fmaxf(sqrtf(tot), kEpsNorm);
    }
    __syncthreads();
  }
}

template <bool F16>
__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  const float2 a = unpack_h2<F16>(u.x), b = unpack_h2<F16>(u.y), c = unpack_h2<F16>(u.z), d = unpack_h2<F16>(u.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}

// Persistent TMA-fed streaming kernel.  grid (ctas_per_slab, ceil(D/256)), at most one CTA per SM; block = 8 consumer
// warps + 1 producer warp.  A CTA owns one 256-column slab and a CONTIGUOUS range of the pooling chunks (<= 64 rows of
// ONE sequence each, described by an int4 {first row, end row, sequence, 0}); consecutive chunks of the same sequence
// accumulate in registers and meet in shared memory only when the sequence changes, so a CTA merges its row subsets
// and writes a partial record once per (CTA, sequence) segment (~5 times at config 2) instead of once per chunk (~33).
// partial[c] = (mean, M2) and seg_rows[c] = row count of the segment that STARTS at chunk c (0 for the other chunks).  The rows of a chunk are contiguous in the
// source (packed activations; padded tensors with contiguous masks), so the producer brings a chunk in with one 2D TMA
// box (64 rows x 256 columns = 32 KB) into a 5-deep shared-memory ring, together with the rows' scale factors:
// ~128 KB of loads stay in flight per SM independent of register pressure, which is what an HBM latency of ~2 us under
// load needs.  Chunks whose source rows are not contiguous (masks with holes) are read straight from global memory.
// Consumer thread = 8 columns (one 16-byte shared-memory load per row) x the rows {sub, sub+8, ...}; it accumulates
// sums shifted by the chunk's first row, s1 = sum(v - K), s2 = sum((v - K)^2), which merge across the 8 row subsets
// by plain addition (fixed order) and stay well conditioned; thread (sub, group) then finishes column 8*group + sub.
constexpr int POOL_NSUB = 8, POOL_RPT = 8, POOL_CONSUMERS = 32 * POOL_NSUB, POOL_THREADS = POOL_CONSUMERS + 32;
constexpr int POOL_STAGES = 5, POOL_BOX_ROWS = 64, POOL_BOX_COLS = 256, POOL_STAGE_BYTES = POOL_BOX_ROWS * POOL_BOX_COLS * 2;
struct PoolSmem {
  uint8_t stage[POOL_STAGES][POOL_STAGE_BYTES];  // 1024-byte aligned
  float comb[2][POOL_NSUB][32][17];              // [buffer][row subset][column group][s1[8] | s2[8]] (+1 pad)
  float sc[POOL_STAGES][POOL_BOX_ROWS];          // per-row scale of the chunk held by each stage (0 beyond its rows)
  uint64_t full[POOL_STAGES], empty[POOL_STAGES];
  int4 desc[POOL_STAGES];   // descriptor of the chunk held by each stage
  int src_row[POOL_STAGES]; // first source row of that chunk, or -1: not contiguous, read from global memory
};
template <bool F16>
__global__ void __launch_bounds__(POOL_THREADS, 1)
pool_partial_kernel(const __grid_constant__ CUtensorMap tmap, const __nv_bfloat16* __restrict__ src, long long ld_src,
                    const int* __restrict__ row_src, const float* __restrict__ inv_norm,
                    const int* __restrict__ chunk_off, const int4* __restrict__ chunk_desc, int B, int D,
                    float2* __restrict__ partial, int* __restrict__ seg_rows) {
  extern __shared__ uint8_t pool_smem_raw[];
  PoolSmem& sm = *reinterpret_cast<PoolSmem*>(pool_smem_raw + ((1024u - (smem_u32(pool_smem_raw) & 1023u)) & 1023u));
  const int total = chunk_off[B];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int per = (total + (int)gridDim.x - 1) / (int)gridDim.x;
  const int c_lo = min(total, (int)blockIdx.x * per), c_hi = min(total, c_lo + per);
  const int col0 = blockIdx.y * POOL_BOX_COLS;
  if (tid == 0) {
    tma_prefetch_desc(&tmap);
    for (int s = 0; s < POOL_STAGES; ++s) {
      mbar_init(&sm.full[s], 1);
      mbar_init(&sm.empty[s], POOL_NSUB);  // one arrive per consumer warp
    }
    fence_barrier_init();
  }
  __syncthreads();
  if (warp == POOL_NSUB) {
    // ------------------------------- producer warp -------------------------------
    // Software-pipelined: the chunk descriptor is loaded two iterations ahead, the chunk's per-row scale factors
    // (and, for padded sources, its first source row) one iteration ahead, so that no iteration waits for a chain
    // of dependent global loads before it can issue its TMA box.  (Round 2, first version: descriptor -> scale
    // factors -> TMA issue cost ~1 us of exposed latency per 32 KB chunk, which capped a CTA at ~30 GB/s, the kernel
    // at 0.67 of the HBM peak — ncu: the producer's STS of the scale factors held 70 % of its samples in long_scoreboard.)
    int stage = 0;
    uint32_t phase = 0;
    auto load_desc = [&](int c) { return (c < c_hi) ? __ldg(chunk_desc + c) : make_int4(0, 0, -1, 0); };
    auto load_aux = [&](const int4& d, float (&scv)[POOL_BOX_ROWS / 32], int& first) {
      first = d.x;
      if (row_src != nullptr && d.z >= 0) {
        first = __ldg(row_src + d.x);
        if (__ldg(row_src + d.y - 1) - first != d.y - 1 - d.x) first = -1;  // row_src is increasing: holes inside
      }
#pragma unroll
      for (int h = 0; h < POOL_BOX_ROWS / 32; ++h) {
        const int r = d.x + lane + 32 * h;
        scv[h] = (d.z >= 0 && r < d.y) ? (inv_norm ? __ldg(inv_norm + r) : 1.f) : 0.f;
      }
    };
    int4 d_cur = load_desc(c_lo), d_nxt = load_desc(c_lo + 1);
    float sc_cur[POOL_BOX_ROWS / 32];
    int first_cur;
    load_aux(d_cur, sc_cur, first_cur);
    for (int c = c_lo; c < c_hi; ++c) {
      const int4 d_nn = load_desc(c + 2);
      float sc_nxt[POOL_BOX_ROWS / 32];
      int first_nxt;
      load_aux(d_nxt, sc_nxt, first_nxt);  // in flight while this iteration waits for its ring slot
      mbar_wait(&sm.empty[stage], phase ^ 1);
#pragma unroll
      for (int h = 0; h < POOL_BOX_ROWS / 32; ++h) sm.sc[stage][lane + 32 * h] = sc_cur[h];
      if (lane == 0) {
        sm.desc[stage] = d_cur;
        sm.src_row[stage] = first_cur;
      }
      __syncwarp();
      if (lane == 0) {
        if (first_cur >= 0) {
          mbar_arrive_expect_tx(&sm.full[stage], POOL_STAGE_BYTES);
          tma_load_2d(sm.stage[stage], &tmap, &sm.full[stage], col0, first_cur);
        } else {
          mbar_arrive(&sm.full[stage]);
        }
      }
      d_cur = d_nxt;
      d_nxt = d_nn;
      first_cur = first_nxt;
#pragma unroll
      for (int h = 0; h < POOL_BOX_ROWS / 32; ++h) sc_cur[h] = sc_nxt[h];
      if (++stage == POOL_STAGES) { stage = 0; phase ^= 1; }
    }
    return;
  }
  // ------------------------------- consumer warps -------------------------------
  const int grp = lane, sub = warp;                 // 8-column group inside the slab, row subset
  const int colb = col0 + grp * 8;                  // first of this thread's 8 columns
  const bool active = colb < D;
  int stage = 0, buf = 0;
  uint32_t phase = 0;
  int cur_seq = -1, seg_first = 0, seg_n = 0;
  float s1[8], s2[8], K[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { s1[i] = 0.f; s2[i] = 0.f; K[i] = 0.f; }
  // close the running segment: the 8 row subsets meet in shared memory (fixed order), thread (sub, grp) finishes
  // column 8*grp + sub.  Called by all consumer threads together.
  auto flush = [&]() {
#pragma unroll
    for (int i = 0; i < 8; ++i) { sm.comb[buf][sub][grp][i] = s1[i]; sm.comb[buf][sub][grp][8 + i] = s2[i]; }
    asm volatile("bar.sync 1, %0;" ::"n"(POOL_CONSUMERS) : "memory");  // consumers only; `buf` alternates
    float t1 = 0.f, t2 = 0.f;
#pragma unroll
    for (int k = 0; k < POOL_NSUB; ++k) {
      t1 += sm.comb[buf][k][grp][sub];
      t2 += sm.comb[buf][k][grp][8 + sub];
    }
    float kk = K[0];
#pragma unroll
    for (int i = 1; i < 8; ++i) kk = (sub == i) ? K[i] : kk;
    const float n = (float)seg_n;
    if (colb + sub < D) partial[(long long)seg_first * D + colb + sub] = make_float2(kk + t1 / n, fmaxf(t2 - t1 * t1 / n, 0.f));
    if (blockIdx.y == 0 && tid == 0) seg_rows[seg_first] = seg_n;
    buf ^= 1;
  };
  for (int c = c_lo; c < c_hi; ++c) {
    mbar_wait(&sm.full[stage], phase);  // also orders the producer's desc / sc / src_row writes before these reads
    const int4 d = sm.desc[stage];
    const int first = sm.src_row[stage];
    const bool new_seg = d.z != cur_seq;
    if (new_seg) {
      if (cur_seq >= 0) flush();
      cur_seq = d.z; seg_first = c; seg_n = 0;
#pragma unroll
      for (int i = 0; i < 8; ++i) { s1[i] = 0.f; s2[i] = 0.f; }
    } else if (blockIdx.y == 0 && tid == 0) {
      seg_rows[c] = 0;  // this chunk continues a segment
    }
    seg_n += d.y - d.x;
    if (first >= 0) {
      const uint8_t* base = sm.stage[stage] + grp * 16;
      if (new_seg) {  // shift of the segment: its first row
        float f[8];
        unpack8<F16>(*reinterpret_cast<const uint4*>(base), f);
        const float sc0 = sm.sc[stage][0];
#pragma unroll
        for (int i = 0; i < 8; ++i) K[i] = f[i] * sc0;
      }
      uint4 u[POOL_RPT];
      float sc[POOL_RPT];
#pragma unroll
      for (int k = 0; k < POOL_RPT; ++k) {
        const int rr = sub + POOL_NSUB * k;
        u[k] = *reinterpret_cast<const uint4*>(base + rr * (POOL_BOX_COLS * 2));  // rows past the chunk: ignored below
        sc[k] = sm.sc[stage][rr];
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.empty[stage]);  // this warp holds its part of the stage in registers
      const int nrows = d.y - d.x;
#pragma unroll
      for (int k = 0; k < POOL_RPT; ++k) {
        if (sub + POOL_NSUB * k < nrows) {  // warp-uniform
          float f[8];
          unpack8<F16>(u[k], f);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float v = fmaf(f[i], sc[k], -K[i]);
            s1[i] += v;
            s2[i] = fmaf(v, v, s2[i]);
          }
        }
      }
    } else {
      // rows scattered in the source: straight from global memory (rare; correctness path)
      if (active) {
        float f[8];
        if (new_seg) {
          unpack8<F16>(__ldg(reinterpret_cast<const uint4*>(src + (long long)__ldg(row_src + d.x) * ld_src + colb)), f);
          const float sc0 = sm.sc[stage][0];
#pragma unroll
          for (int i = 0; i < 8; ++i) K[i] = f[i] * sc0;
        }
        for (int r = d.x + sub; r < d.y; r += POOL_NSUB) {
          unpack8<F16>(__ldg(reinterpret_cast<const uint4*>(src + (long long)__ldg(row_src + r) * ld_src + colb)), f);
          const float scr = sm.sc[stage][r - d.x];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float v = fmaf(f[i], scr, -K[i]);
            s1[i] += v;
            s2[i] = fmaf(v, v, s2[i]);
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.empty[stage]);
    }
    if (++stage == POOL_STAGES) { stage = 0; phase ^= 1; }
  }
  if (cur_seq >= 0) flush();
}

// phase 2: combine the chunks of each sequence in order (Chan), emit mean / std / mix.
// grid (B, ceil(D/256)), block 256.  mode: 1 mean, 2 std, 3 mix.  out [B][ld_out] fp32.
// stats[b] = (mean[D] | std[D]) is always written when `stats` != null (needed by backward).
__global__ void pool_finalize_kernel(const float2* __restrict__ partial, const int* __restrict__ seg_rows,
                                     const int* __restrict__ seq_off, const int* __restrict__ chunk_off, int B, int D,
                                     int mode, float* __restrict__ out, long long ld_out) {
  const int b = blockIdx.x;
  const int col = blockIdx.y * blockDim.x + threadIdx.x;
  if (col >= D) return;
  const int c0 = chunk_off[b], c1 = chunk_off[b + 1];
  const int n_total = seq_off[b + 1] - seq_off[b];
  float n = 0.f, mean = 0.f, m2 = 0.f;
  for (int c = c0; c < c1; ++c) {
    const int rows_c = seg_rows[c];
    if (rows_c == 0) continue;  // the chunk continues a segment: its rows are inside an earlier record
    const float nc = (float)rows_c;
    const float2 pc = partial[(long long)c * D + col];
    const float tot = n + nc;
    const float delta = pc.x - mean;
    mean += delta * (nc / tot);
    m2 += pc.y + delta * delta * (n * nc / tot);
    n = tot;
  }
  float mu, sd;
  if (n_total > 0) { mu = mean; sd = sqrtf(m2 / n); }
  else { mu = __int_as_float(0x7fc00000); sd = mu; }  // 0/0 in the reference
  float* o = out + (long long)b * ld_out;
  if (mode == 1) o[col] = mu;
  else if (mode == 2) o[col] = sd;
  else { o[col] = mu; o[D + col] = sd; }
}

// phase 2 of the fused step ('mix' readout followed by F.normalize, scripts/train_contrast.py:277-281 + :354/:365):
// combine the chunks, write stats = (mean | std), then p = stats / max(|stats|, eps).
// grid (FIN_CLUSTER, B) launched as thread-block clusters of FIN_CLUSTER CTAs along x: the CTAs of a cluster share
// one sequence, each finalises a column slice, and the slices' sums of squares meet through distributed shared
// memory (fixed order), so the whole embedding is normalised without a second kernel and with 8x more CTAs in flight
// than one CTA per sequence would give.
constexpr int FIN_CLUSTER = 8, FIN_THREADS = 256;
__global__ void __launch_bounds__(FIN_THREADS)
pool_finalize_normalize_kernel(const float2* __restrict__ partial, const int* __restrict__ seg_rows,
                               const int* __restrict__ seq_off, const int* __restrict__ chunk_off, int D, float* __restrict__ stats,
                               __nv_bfloat16* __restrict__ p_bf16, float* __restrict__ p_f32, float* __restrict__ norm_out) {
  const int b = blockIdx.y;
  const uint32_t rank = cluster_ctarank();
  const int c0 = chunk_off[b], c1 = chunk_off[b + 1];
  const int n_total = seq_off[b + 1] - seq_off[b];
  float* st = stats + (long long)b * 2 * D;
  __shared__ float sh[FIN_THREADS / 32];
  __shared__ float cluster_sq[FIN_CLUSTER];  // slot r of EVERY CTA receives CTA r's partial sum of squares
  float sq = 0.f;
  for (int col = rank * FIN_THREADS + threadIdx.x; col < D; col += FIN_CLUSTER * FIN_THREADS) {
    float n = 0.f, mean = 0.f, m2 = 0.f;
    for (int c = c0; c < c1; ++c) {
      const int rows_c = seg_rows[c];
      if (rows_c == 0) continue;
      const float nc = (float)rows_c;
      const float2 pc = partial[(long long)c * D + col];
      const float tot = n + nc;
      const float delta = pc.x - mean;
      mean += delta * (nc / tot);
      m2 += pc.y + delta * delta * (n * nc / tot);
      n = tot;
    }
    float mu, sd;
    if (n_total > 0) { mu = mean; sd = sqrtf(m2 / n); }
    else { mu = __int_as_float(0x7fc00000); sd = mu; }  // 0/0 in the reference
    st[col] = mu;
    st[D + col] = sd;
    sq = fmaf(mu, mu, fmaf(sd, sd, sq));
  }
  sq = block_sum<FIN_THREADS>(sq, sh);
  if (threadIdx.x < FIN_CLUSTER) {  // broadcast this CTA's partial into slot `rank` of every CTA of the cluster
    const uint32_t remote = mapa_shared(smem_u32(&cluster_sq[rank]), threadIdx.x);
    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(remote), "f"(sq) : "memory");
  }
  cluster_sync_all();  // release/acquire: the remote stores above are visible after the barrier
  float total = 0.f;
#pragma unroll
  for (int r = 0; r < FIN_CLUSTER; ++r) total += cluster_sq[r];  // same order in every CTA: identical norm
  const float nrm = sqrtf(total);
  const float inv = 1.f / fmaxf(nrm, kEpsNorm);
  if (rank == 0 && threadIdx.x == 0 && norm_out) norm_out[b] = nrm;
  for (int col = rank * FIN_THREADS + threadIdx.x; col < D; col += FIN_CLUSTER * FIN_THREADS) {
    const float mu = st[col] * inv, sd = st[D + col] * inv;  // written by this very thread above
    const long long o = (long long)b * 2 * D;
    if (p_f32) { p_f32[o + col] = mu; p_f32[o + D + col] = sd; }
    if (p_bf16) { p_bf16[o + col] = __float2bfloat16_rn(mu); p_bf16[o + D + col] = __float2bfloat16_rn(sd); }
  }
}

// ------------------------------------------------------------------------------------------------
// L2 normalisation of pooled embeddings: p = e / max(|e|, eps).  grid B, block 256.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
l2norm_fwd_kernel(const float* __restrict__ e, int E, __nv_bfloat16* __restrict__ p_bf16, float* __restrict__ p_f32,
                  float* __restrict__ norm_out) {
  const int b = blockIdx.x;
  __shared__ float sh[8];
  const float* row = e + (long long)b * E;
  float s = 0.f;
  for (int i = threadIdx.x; i < E; i += 256) { const float v = row[i]; s = fmaf(v, v, s); }
  s = block_sum<256>(s, sh);
  const float nrm = sqrtf(s);
  const float inv = 1.f / fmaxf(nrm, kEpsNorm);
  if (threadIdx.x == 0 && norm_out) norm_out[b] = nrm;
  for (int i = threadIdx.x; i < E; i += 256) {
    const float v = row[i] * inv;
    if (p_f32) p_f32[(long long)b * E + i] = v;
    if (p_bf16) p_bf16[(long long)b * E + i] = __float2bfloat16_rn(v);
  }
}

// de = (dp - p (p.dp)) / |e|   (dp/eps where the clamp is active).  grid B, block 256.
__global__ void __launch_bounds__(256)
l2norm_bwd_kernel(const float* __restrict__ dp, const float* __restrict__ p, const float* __restrict__ norm, int E,
                  float* __restrict__ de) {
  const int b = blockIdx.x;
  __shared__ float sh[8];
  const float* dpr = dp + (long long)b * E;
  const float* pr = p + (long long)b * E;
  float s = 0.f;
  for (int i = threadIdx.x; i < E; i += 256) s = fmaf(dpr[i], pr[i], s);
  s = block_sum<256>(s, sh);
  const float nrm = norm[b];
  const bool clamped = nrm < kEpsNorm;
  const float inv = 1.f / fmaxf(nrm, kEpsNorm);
  for (int i = threadIdx.x; i < E; i += 256)
    de[(long long)b * E + i] = clamped ? dpr[i] * inv : (dpr[i] - pr[i] * s) * inv;
}

// ------------------------------------------------------------------------------------------------
// pooling backward, phase 1: per-sequence coefficient vectors so that  dy_r = c1 + c2 * y_r
//   mean: c1 = dmu/n                   std: c2 = dsd/(n sd), c1 = -c2*mu         mix: both
// grid (B, ceil(D/256)), block 256.  de [B][ld_de] fp32 holds (dmu | dsd) per mode; stats (mu | sd).
// ------------------------------------------------------------------------------------------------
__global__ void pool_bwd_coef_kernel(const float* __restrict__ de, long long ld_de, const float* __restrict__ stats,
                                     long long ld_stats, const int* __restrict__ seq_off, int D, int mode,
                                     float* __restrict__ c1, float* __restrict__ c2) {
  const int b = blockIdx.x;
  const int col = blockIdx.y * blockDim.x + threadIdx.x;
  if (col >= D) return;
  const float n = (float)(seq_off[b + 1] - seq_off[b]);
  const float* d = de + (long long)b * ld_de;
  const float* st = stats + (long long)b * ld_stats;
  float dmu = 0.f, dsd = 0.f;
  if (mode == 1) dmu = d[col];
  else if (mode == 2) dsd = d[col];
  else { dmu = d[col]; dsd = d[D + col]; }
  float k1 = dmu / n, k2 = 0.f;
  if (mode != 1) {
    const float mu = st[col], sd = st[D + col];
    k2 = dsd / (n * sd);  // inf/NaN when sd == 0, as autograd on the reference
    k1 = k1 - k2 * mu;
  }
  c1[(long long)b * D + col] = k1;
  c2[(long long)b * D + col] = k2;
}

// ------------------------------------------------------------------------------------------------
// Backward from dLogits to the pooling coefficients (small / medium similarity blocks, fp32 embeddings), two kernels:
//   (1) loss_bwd_dp_kernel, grid E/64, block 256: dp[:, slice] = (dloss / tau) * dS t[:, slice] for ALL rows
//       (t's 64-column slab is read once by one CTA and shared by every row: the work scales with R*C*E / #SM,
//       not with C per row) + per-slice partial dot products  dotp[i][slice] = sum_e dp_ie p_ie
//   (2) loss_bwd_coef_kernel, grid B, block 1024: de_i = (dp_i - p_i (p_i . dp_i)) / |e_i| (autograd of
//       F.normalize, eps clamp as in l2norm_bwd_kernel), then the 'mix' coefficients
//       c2 = dsd / (n sd),  c1 = dmu / n - c2 mu   (de = (dmu | dsd)), so that dy_r = c1 + c2 * y_r
// Rows i >= R were dropped by the segment split: dp = 0.  dS fp32 [R][C] already carries the 1/R of the mean.
// ------------------------------------------------------------------------------------------------
constexpr int LBW_SLICE = 64;
constexpr int LBW_JC = 64;  // columns of dS (= rows of t) staged per trip
__global__ void __launch_bounds__(256)
loss_bwd_dp_kernel(const float* __restrict__ dS, const float* __restrict__ t, const float* __restrict__ p,
                   const float* __restrict__ dloss, int R, int B, int C, int E, float inv_tau, float* __restrict__ dp,
                   float* __restrict__ dotp) {
  // thread = (row group rg of 16 -> rows rg, rg+16, ...; column quad cq -> 4 columns), 16 column quads x 16 row groups.
  // t's 64-column slab and the matching block of dS are staged in shared memory 64 logit-columns at a time by all
  // 256 threads (every load of a trip in flight together): the cost per CTA is C/64 trips, not C dependent loads.
  __shared__ __align__(16) float ts[LBW_JC][LBW_SLICE];
  __shared__ float ws[64][LBW_JC + 1];
  const int e0 = blockIdx.x * LBW_SLICE;
  const int cq = threadIdx.x & 15, rg = threadIdx.x >> 4;
  const int col = e0 + cq * 4;
  const float scale = inv_tau * (dloss ? dloss[0] : 1.f);
  const int nslice = gridDim.x;
  for (int i0 = 0; i0 < B; i0 += 64) {  // 64 rows per pass: 4 rows per thread
    float4 acc[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) acc[q] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int j0 = 0; j0 < C; j0 += LBW_JC) {
      __syncthreads();
#pragma unroll
      for (int m = 0; m < (LBW_JC * LBW_SLICE / 4) / 256; ++m) {
        const int idx = threadIdx.x + 256 * m;
        const int jr = idx >> 4, ev = (idx & 15) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (j0 + jr < C && e0 + ev < E) v = __ldg(reinterpret_cast<const float4*>(t + (long long)(j0 + jr) * E + e0 + ev));
        *reinterpret_cast<float4*>(&ts[jr][ev]) = v;
      }
#pragma unroll
      for (int m = 0; m < (64 * LBW_JC) / 256; ++m) {
        const int idx = threadIdx.x + 256 * m;
        const int r = idx >> 6, c = idx & 63;
        const int i = i0 + r, j = j0 + c;
        ws[r][c] = (i < R && j < C) ? __ldg(dS + (long long)i * C + j) : 0.f;
      }
      __syncthreads();
      const int jn = min(LBW_JC, C - j0);
#pragma unroll 8
      for (int jj = 0; jj < jn; ++jj) {
        const float4 tv = *reinterpret_cast<const float4*>(&ts[jj][cq * 4]);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float w = ws[rg + 16 * q][jj];
          acc[q].x = fmaf(w, tv.x, acc[q].x); acc[q].y = fmaf(w, tv.y, acc[q].y);
          acc[q].z = fmaf(w, tv.z, acc[q].z); acc[q].w = fmaf(w, tv.w, acc[q].w);
        }
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int i = i0 + rg + 16 * q;
      float d = 0.f;
      if (i < B && col < E) {
        const float4 v = make_float4(acc[q].x * scale, acc[q].y * scale, acc[q].z * scale, acc[q].w * scale);
        *reinterpret_cast<float4*>(dp + (long long)i * E + col) = v;
        const float4 pv = *reinterpret_cast<const float4*>(p + (long long)i * E + col);
        d = fmaf(v.x, pv.x, fmaf(v.y, pv.y, fmaf(v.z, pv.z, v.w * pv.w)));
      }
      // the 16 column quads of a row sit in 16 consecutive lanes: fixed-order shuffle reduction
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
      if (cq == 0 && i < B) dotp[(long long)i * nslice + blockIdx.x] = d;
    }
  }
}

__global__ void __launch_bounds__(1024)
loss_bwd_coef_kernel(const float* __restrict__ dp, const float* __restrict__ dotp, int nslice, const float* __restrict__ p,
                     const float* __restrict__ pnorm, const float* __restrict__ stats, const int* __restrict__ seq_off,
                     int D, float* __restrict__ c1, float* __restrict__ c2) {
  __shared__ float sh[32];
  const int i = blockIdx.x;
  const int E = 2 * D;
  float dot = 0.f;
  for (int k = threadIdx.x; k < nslice; k += 1024) dot += dotp[(long long)i * nslice + k];
  dot = block_sum<1024>(dot, sh);
  const float nrm = pnorm[i];
  const bool clamped = nrm < kEpsNorm;
  const float inv = 1.f / fmaxf(nrm, kEpsNorm);
  const float n = (float)(seq_off[i + 1] - seq_off[i]);
  const float* st = stats + (long long)i * E;
  const float* pr = p + (long long)i * E;
  const float* dr = dp + (long long)i * E;
  for (int col = threadIdx.x; col < D; col += 1024) {
    const float dmu = clamped ? dr[col] * inv : (dr[col] - pr[col] * dot) * inv;
    const float dsd = clamped ? dr[D + col] * inv : (dr[D + col] - pr[D + col] * dot) * inv;
    const float mu = st[col], sd = st[D + col];
    const float k2 = dsd / (n * sd);  // inf/NaN when sd == 0, as autograd on the reference
    c1[(long long)i * D + col] = dmu / n - k2 * mu;
    c2[(long long)i * D + col] = k2;
  }
}

// ------------------------------------------------------------------------------------------------
// adapter tail backward over packed rows, persistent: CTA k owns the contiguous row range [k*per, (k+1)*per) of the
// n valid rows (per = ceil(n / grid) rounded to the batch size), so every SM streams the same number of rows whatever
// the sequence lengths are (one CTA per 64-row pooling chunk left 541 CTAs for 148 SMs at config 2: 3.65 waves).
//   y = a * inv;  dy = c1[b] + c2[b] * y;  da = (dy - y (y.dy)) * inv;  dz2 = da * g
// thread = 8 columns; inside one sequence the coefficient vectors c1[b], c2[b] live in registers and rows go through
// in batches of 2, double-buffered so that 8 independent 16-byte loads are in flight per thread; the per-row dot
// product y.dy is a block reduction (one barrier per batch).  A range that crosses a sequence boundary restarts the
// pipeline there with the next sequence's coefficients.
// The column sums of dz2 (the fc2 bias gradient) accumulate in registers over the CTA's whole range and leave as ONE
// partial row per CTA (fixed row -> CTA assignment: deterministic).  a, g fp16 [rows][D]; dz2 bf16.  Rows in
// [n, roundup(n,256)) are zeroed so the weight-gradient GEMMs can run their K loop over whole 64-row blocks.
// ------------------------------------------------------------------------------------------------
constexpr int TAIL_R = 2;
template <int THREADS>
__global__ void __launch_bounds__(THREADS, 1)
adapter_tail_bwd_kernel(const __half* __restrict__ a, const __half* __restrict__ g, const float* __restrict__ inv_norm,
                        const int* __restrict__ seq_off, int B, const float* __restrict__ c1,
                        const float* __restrict__ c2, const int* __restrict__ n_rows, int cap, int D,
                        __nv_bfloat16* __restrict__ dz2, float* __restrict__ colsum_partial, int* __restrict__ nparts_out) {
  const int n = min(*n_rows, cap);
  const int n_pad = min(cap, (n + 255) & ~255);
  const int nvec = D >> 3;
  const int tid = threadIdx.x;
  const int G = gridDim.x;
  const bool active = tid < nvec;
  const int warp = tid >> 5, lane = tid & 31;
  constexpr int NWARPS = THREADS / 32;
  __shared__ float red[2][NWARPS][TAIL_R];
  if (blockIdx.x == 0 && tid == 0 && nparts_out != nullptr) nparts_out[0] = G;
  // pad rows
  for (int row = n + blockIdx.x; row < n_pad; row += G)
    if (active) reinterpret_cast<uint4*>(dz2 + (long long)row * D)[tid] = make_uint4(0, 0, 0, 0);
  const int per = (((n + G - 1) / G) + 2 * TAIL_R - 1) / (2 * TAIL_R) * (2 * TAIL_R);
  const int lo = (int)min((long long)n, (long long)blockIdx.x * per), hi = min(n, lo + per);
  float k1[8], k2[8], csum[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { k1[i] = 0.f; k2[i] = 0.f; csum[i] = 0.f; }
  struct Batch {
    uint4 av[TAIL_R], gv[TAIL_R];
    float inv[TAIL_R];
  };
  int b = (lo < hi) ? find_segment(seq_off, B, lo) : 0;
  for (int r0 = lo; r0 < hi;) {
    while (b + 1 < B && r0 >= __ldg(seq_off + b + 1)) ++b;  // (sequences without rows are stepped over)
    const int r1 = min(hi, __ldg(seq_off + b + 1));
    if (active) {
      const float4* p1 = reinterpret_cast<const float4*>(c1 + (long long)b * D) + 2 * tid;
      const float4* p2 = reinterpret_cast<const float4*>(c2 + (long long)b * D) + 2 * tid;
      const float4 x0 = __ldg(p1), x1 = __ldg(p1 + 1), y0 = __ldg(p2), y1 = __ldg(p2 + 1);
      k1[0] = x0.x; k1[1] = x0.y; k1[2] = x0.z; k1[3] = x0.w; k1[4] = x1.x; k1[5] = x1.y; k1[6] = x1.z; k1[7] = x1.w;
      k2[0] = y0.x; k2[1] = y0.y; k2[2] = y0.z; k2[3] = y0.w; k2[4] = y1.x; k2[5] = y1.y; k2[6] = y1.z; k2[7] = y1.w;
    }
    auto load = [&](Batch& bt, int rb) {
#pragma unroll
      for (int q = 0; q < TAIL_R; ++q) {
        const int r = min(rb + q, r1 - 1);
        bt.av[q] = active ? __ldg(reinterpret_cast<const uint4*>(a + (long long)r * D) + tid) : make_uint4(0, 0, 0, 0);
        bt.gv[q] = active ? __ldg(reinterpret_cast<const uint4*>(g + (long long)r * D) + tid) : make_uint4(0, 0, 0, 0);
        bt.inv[q] = inv_norm[r];
      }
    };
    auto process = [&](const Batch& bt, int rb, int buf) {
      float dot[TAIL_R];
#pragma unroll
      for (int q = 0; q < TAIL_R; ++q) {
        float f[8];
        unpack8<true>(bt.av[q], f);
        float d = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float y = f[i] * bt.inv[q];
          d = fmaf(y, fmaf(k2[i], y, k1[i]), d);
        }
        dot[q] = warp_sum(d);
      }
      if (lane == 0) {
#pragma unroll
        for (int q = 0; q < TAIL_R; ++q) red[buf][warp][q] = dot[q];
      }
      __syncthreads();  // the only barrier per batch: every thread then adds the NWARPS partials itself (smem broadcasts)
      float totq[TAIL_R];
#pragma unroll
      for (int q = 0; q < TAIL_R; ++q) totq[q] = 0.f;
#pragma unroll
      for (int w = 0; w < NWARPS; ++w) {
#pragma unroll
        for (int q = 0; q < TAIL_R; ++q) totq[q] += red[buf][w][q];
      }
#pragma unroll
      for (int q = 0; q < TAIL_R; ++q) {
        if (active && rb + q < r1) {
          const float dt = totq[q];
          float f[8], gg[8], o[8];
          unpack8<true>(bt.av[q], f);
          unpack8<true>(bt.gv[q], gg);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float y = f[i] * bt.inv[q];
            const float dy = fmaf(k2[i], y, k1[i]);
            o[i] = (dy - y * dt) * bt.inv[q] * gg[i];
            csum[i] += o[i];
          }
          reinterpret_cast<uint4*>(dz2 + (long long)(rb + q) * D)[tid] =
              make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]), pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7]));
        }
      }
    };
    // double-buffered: the loads of batch i+1 are in flight while batch i is reduced and written
    __syncthreads();  // the previous sequence's last batch may still be reading red[0]
    Batch b0, b1;
    load(b0, r0);
    for (int rb = r0; rb < r1; rb += 2 * TAIL_R) {
      if (rb + TAIL_R < r1) load(b1, rb + TAIL_R);
      process(b0, rb, 0);
      if (rb + TAIL_R < r1) {
        if (rb + 2 * TAIL_R < r1) load(b0, rb + 2 * TAIL_R);
        process(b1, rb + TAIL_R, 1);
      }
    }
    r0 = r1;
  }
  if (active && colsum_partial != nullptr) {
    float4* o = reinterpret_cast<float4*>(colsum_partial + (long long)blockIdx.x * D) + 2 * tid;
    o[0] = make_float4(csum[0], csum[1], csum[2], csum[3]);
    o[1] = make_float4(csum[4], csum[5], csum[6], csum[7]);
  }
}

// sum `nparts` partial rows: out[col] = sum_k partial[k][col].  grid ceil(D/32), block (32, 32): thread (x, y)
// adds parts y, y+32, ... of column 32*blockIdx.x + x, then the 32 rows are merged in shared memory (fixed order).
constexpr int PARTS_Y = 32;
__device__ __forceinline__ float sum_parts(const float* __restrict__ partial, int nparts, int D, int col) {
  // four independent chains: the loads of a thread are latency-bound, not bandwidth-bound (fixed order: deterministic)
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  int k = threadIdx.y;
  for (; k + 3 * PARTS_Y < nparts; k += 4 * PARTS_Y) {
    s0 += partial[(long long)k * D + col];
    s1 += partial[(long long)(k + PARTS_Y) * D + col];
    s2 += partial[(long long)(k + 2 * PARTS_Y) * D + col];
    s3 += partial[(long long)(k + 3 * PARTS_Y) * D + col];
  }
  for (; k < nparts; k += PARTS_Y) s0 += partial[(long long)k * D + col];
  return (s0 + s1) + (s2 + s3);
}
__global__ void __launch_bounds__(32 * PARTS_Y)
parts_colsum_final_kernel(const float* __restrict__ partial, const int* __restrict__ chunk_off, int B,
                          const int* __restrict__ n_rows, int n_static, int D, __nv_bfloat16* __restrict__ out_bf16,
                          float* __restrict__ out_f32) {
  // number of partial rows: chunk_off[B] (pooling chunks) or ceil(n/64) (plain 64-row blocks)
  int nparts;
  if (chunk_off != nullptr) nparts = chunk_off[B];
  else nparts = ((n_rows ? min(*n_rows, n_static) : n_static) + 63) / 64;
  const int col = blockIdx.x * 32 + threadIdx.x;
  __shared__ float sm[PARTS_Y][33];
  sm[threadIdx.y][threadIdx.x] = (col < D) ? sum_parts(partial, nparts, D, col) : 0.f;
  __syncthreads();
  if (threadIdx.y == 0 && col < D) {
    float s = 0.f;
#pragma unroll
    for (int y = 0; y < PARTS_Y; ++y) s += sm[y][threadIdx.x];
    if (out_bf16) out_bf16[col] = __float2bfloat16_rn(s);
    if (out_f32) out_f32[col] = s;
  }
}

// Both bias gradients of the step in one launch: db = sum of the partial rows of a job, in a fixed order.
//   job 0 (fc1.bias): partials written by the dgrad GEMM's epilogue, one row per 32 residue rows
//   job 1 (fc2.bias): partials written by adapter_tail_bwd_kernel, one row per CTA
// Outputs: fp32 (what the gradient all-reduce carries: a single rounding to bf16 AFTER the mean over ranks) and/or
// bf16 (what the parameter's .grad holds).  `accumulate`: add to the fp32 output first (micro-batch accumulation).
__global__ void __launch_bounds__(32 * PARTS_Y)
bias_grads_final_kernel(BiasJob j0, BiasJob j1, int blocks0) {
  const bool first = (int)blockIdx.x < blocks0;
  const BiasJob& j = first ? j0 : j1;
  const int blk = first ? blockIdx.x : blockIdx.x - blocks0;
  int nparts;
  if (j.nparts_dev != nullptr) nparts = min(*j.nparts_dev, j.nparts_max);
  else nparts = min(((j.n_rows ? min(*j.n_rows, j.n_static) : j.n_static) + j.block_rows - 1) / j.block_rows, j.nparts_max);
  const int col = blk * 32 + threadIdx.x;
  __shared__ float sm[PARTS_Y][33];
  sm[threadIdx.y][threadIdx.x] = (col < j.D) ? sum_parts(j.partial, nparts, j.D, col) : 0.f;
  __syncthreads();
  if (threadIdx.y == 0 && col < j.D) {
    float s = 0.f;
#pragma unroll
    for (int y = 0; y < PARTS_Y; ++y) s += sm[y][threadIdx.x];
    if (j.accumulate && j.out_f32) s += j.out_f32[col];
    if (j.out_f32) j.out_f32[col] = s;
    if (j.out_bf16) reinterpret_cast<__nv_bfloat16*>(j.out_bf16)[col] = __float2bfloat16_rn(s);
  }
}

// same, but the upstream gradient dy is given per row (module API: y = adapter(x) was returned)
template <int NV>
__global__ void __launch_bounds__(256)
adapter_tail_bwd_dy_kernel(const __half* __restrict__ a, const __half* __restrict__ g,
                           const float* __restrict__ inv_norm, const __nv_bfloat16* __restrict__ dy, int n,
                           const int* __restrict__ n_dev, int cap, int D, __nv_bfloat16* __restrict__ dz2) {
  if (n_dev) n = min(n, *n_dev);
  const int n_pad = min(cap, (n + 255) & ~255);
  const int lane = threadIdx.x & 31;
  const int warps_total = (gridDim.x * blockDim.x) >> 5;
  const int nvec = D >> 3;
  for (int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < n_pad; row += warps_total) {
    uint4* out = reinterpret_cast<uint4*>(dz2 + (long long)row * D);
    if (row >= n) {
      for (int v = lane; v < nvec; v += 32) out[v] = make_uint4(0, 0, 0, 0);
      continue;
    }
    const float inv = inv_norm[row];
    const uint4* ar = reinterpret_cast<const uint4*>(a + (long long)row * D);
    const uint4* gr = reinterpret_cast<const uint4*>(g + (long long)row * D);
    const uint4* dr = reinterpret_cast<const uint4*>(dy + (long long)row * D);
    uint4 av[NV], dv[NV];
    float dot = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int v = lane + 32 * k;
      av[k] = (v < nvec) ? __ldg(ar + v) : make_uint4(0, 0, 0, 0);
      dv[k] = (v < nvec) ? __ldg(dr + v) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const uint32_t w[4] = {av[k].x, av[k].y, av[k].z, av[k].w};
      const uint32_t dw[4] = {dv[k].x, dv[k].y, dv[k].z, dv[k].w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 f = unpack_f16x2(w[i]);
        const float2 d = unpack_bf16x2(dw[i]);
        dot = fmaf(f.x * inv, d.x, dot);
        dot = fmaf(f.y * inv, d.y, dot);
      }
    }
    dot = warp_sum(dot);
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int v = lane + 32 * k;
      if (v < nvec) {
        const uint4 gu = __ldg(gr + v);
        const uint32_t w[4] = {av[k].x, av[k].y, av[k].z, av[k].w};
        const uint32_t dw[4] = {dv[k].x, dv[k].y, dv[k].z, dv[k].w};
        const uint32_t gw[4] = {gu.x, gu.y, gu.z, gu.w};
        uint32_t o[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 f = unpack_f16x2(w[i]);
          const float2 d = unpack_bf16x2(dw[i]);
          const float2 gg = unpack_f16x2(gw[i]);
          o[i] = pack_bf16x2((d.x - f.x * inv * dot) * inv * gg.x, (d.y - f.y * inv * dot) * inv * gg.y);
        }
        out[v] = make_uint4(o[0], o[1], o[2], o[3]);
      }
    }
  }
}

// y[dst(row)] = a[row] * inv_norm[row]  (module API forward tail).  inv from rowsq partials.  dst(row) = row, or
// row_dst[row] when the rows go straight into another tensor's slots (Stage-2 placeholder replacement,
// models/esmc_qwen_arc.py:142: inputs_embeds[placeholder_mask] = encoder_hidden_states[encoder_mask]).
// n_dev / n_dst_dev (optional, device) bound the row count from both sides without a host sync.
__global__ void __launch_bounds__(256)
scale_rows_kernel(const __half* __restrict__ a, const float* __restrict__ rowsq, int nblk, int cap, int n, int D,
                  __nv_bfloat16* __restrict__ y, long long ld_y, const int* __restrict__ row_dst,
                  const int* __restrict__ n_dev, const int* __restrict__ n_dst_dev, float* __restrict__ inv_norm_out) {
  const int lane = threadIdx.x & 31;
  const int warps_total = (gridDim.x * blockDim.x) >> 5;
  const int nvec = D >> 3;
  if (n_dev) n = min(n, *n_dev);
  if (n_dst_dev) n = min(n, *n_dst_dev);
  for (int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < n; row += warps_total) {
    float s = 0.f;
    for (int j = lane; j < nblk; j += 32) s += rowsq[(long long)j * cap + row];  // (order differs from row_inv_norm: module API only)
    s = warp_sum(s);
    const float inv = 1.f / fmaxf(sqrtf(s), kEpsNorm);
    if (lane == 0 && inv_norm_out) inv_norm_out[row] = inv;
    const uint4* ar = reinterpret_cast<const uint4*>(a + (long long)row * D);
    uint4* yr = reinterpret_cast<uint4*>(y + (long long)(row_dst ? row_dst[row] : row) * ld_y);
    for (int v = lane; v < nvec; v += 32) {
      const uint4 u = __ldg(ar + v);
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
      uint32_t o[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 f = unpack_f16x2(w[i]);
        o[i] = pack_bf16x2(f.x * inv, f.y * inv);
      }
      yr[v] = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// generic readout backward on a padded (B, S, D) tensor: dx[b,r] = m[b,r] (c1[b] + c2[b]*x[b,r])
// grid-stride over 16-byte vectors; output bf16, same layout as x.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
readout_bwd_kernel(const __nv_bfloat16* __restrict__ x, const void* __restrict__ mask, int mask_bytes, int B, int S,
                   int D, const float* __restrict__ c1, const float* __restrict__ c2, __nv_bfloat16* __restrict__ dx) {
  const int nvec = D >> 3;
  const long long total = (long long)B * S * nvec;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / nvec;
    const int v = (int)(i % nvec);
    const int b = (int)(row / S);
    uint4 o = make_uint4(0, 0, 0, 0);
    if (mask_at(mask, mask_bytes, row)) {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(x + row * D) + v);
      const float4 p0 = __ldg(reinterpret_cast<const float4*>(c1 + (long long)b * D) + 2 * v);
      const float4 p1 = __ldg(reinterpret_cast<const float4*>(c1 + (long long)b * D) + 2 * v + 1);
      const float4 q0 = __ldg(reinterpret_cast<const float4*>(c2 + (long long)b * D) + 2 * v);
      const float4 q1 = __ldg(reinterpret_cast<const float4*>(c2 + (long long)b * D) + 2 * v + 1);
      const float k1[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
      const float k2[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
      uint32_t r[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = unpack_bf16x2(w[j]);
        r[j] = pack_bf16x2(fmaf(k2[2 * j], f.x, k1[2 * j]), fmaf(k2[2 * j + 1], f.y, k1[2 * j + 1]));
      }
      o = make_uint4(r[0], r[1], r[2], r[3]);
    }
    reinterpret_cast<uint4*>(dx + row * D)[v] = o;
  }
}

// ------------------------------------------------------------------------------------------------
// column sums of a packed bf16 matrix (bias gradients): two deterministic phases.
// phase 1 grid (ceil(cap/64), ceil(D/512)), block 256 = 64 column groups x 4 row subsets over a 64-row
// block -> partial[block][col]; phase 2 sums the row blocks below ceil(n/64) in a fixed order.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
colsum_partial_kernel(const __nv_bfloat16* __restrict__ x, const int* __restrict__ n_rows, int n_static, int D,
                      float* __restrict__ partial) {
  const int n = n_rows ? min(*n_rows, n_static) : n_static;
  const int r0 = blockIdx.x * 64, r1 = min(n, r0 + 64);
  if (r0 >= n) return;
  const int cgl = threadIdx.x & 63, sub = threadIdx.x >> 6;
  const int g = blockIdx.y * 64 + cgl;
  const bool active = g * 8 < D;
  __shared__ float comb[3][64][9];
  float s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (active) {
    uint4 u[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const int r = r0 + sub + 4 * k;
      u[k] = (r < r1) ? __ldg(reinterpret_cast<const uint4*>(x + (long long)r * D) + g) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      float f[8];
      unpack8<false>(u[k], f);
#pragma unroll
      for (int i = 0; i < 8; ++i) s[i] += f[i];
    }
  }
  if (sub > 0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) comb[sub - 1][cgl][i] = s[i];
  }
  __syncthreads();
  if (sub == 0 && active) {
    float* o = partial + (long long)blockIdx.x * D + g * 8;
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = s[i] + comb[0][cgl][i] + comb[1][cgl][i] + comb[2][cgl][i];
  }
}
// 'last' readout: out[b] = x[b, sum(mask[b]) - 1]   (scripts/train_contrast.py:207-215)
__global__ void readout_last_kernel(const __nv_bfloat16* __restrict__ x, const int* __restrict__ counts, int S, int D,
                                    float* __restrict__ out) {
  const int b = blockIdx.x;
  int idx = counts[b] - 1;
  if (idx < 0) idx += S;  // python negative indexing, as the reference's advanced indexing does
  const __nv_bfloat16* row = x + ((long long)b * S + idx) * D;
  for (int i = threadIdx.x; i < D; i += blockDim.x) out[(long long)b * D + i] = __bfloat162float(row[i]);
}

// its backward: dx = 0 except dx[b, sum(mask[b]) - 1] = dout[b]; every element of dx is written here (no memset)
__global__ void __launch_bounds__(256)
readout_last_bwd_kernel(const __nv_bfloat16* __restrict__ dout, const int* __restrict__ counts, int S, int D,
                        __nv_bfloat16* __restrict__ dx) {
  const int b = blockIdx.y;
  int idx = counts[b] - 1;
  if (idx < 0) idx += S;
  const long long total = (long long)S * D;
  __nv_bfloat16* base = dx + (long long)b * total;
  const __nv_bfloat16 zero = __float2bfloat16_rn(0.f);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int row = (int)(i / D);
    base[i] = row == idx ? dout[(long long)b * D + (i - (long long)row * D)] : zero;
  }
}

// ------------------------------------------------------------------------------------------------
// host -> device staging by PULL: a few CTAs read the valid rows of a padded batch straight out of pinned (mapped)
// host memory and write them packed.  One launch for the whole batch: the copy-engine form pays ~3.5 us of set-up per
// sequence (64 copies of 0.25-5 MB reach 49 GB/s where one packed copy reaches 54, tools/h2d_probe.py).
//   table[3*s + 0..2] = byte offset of segment s in the host batch, in the packed destination, and its length
//   (all multiples of 16); piece_prefix[s] = number of 32 KB pieces before segment s, piece_prefix[n_seg] = total.
// Each CTA takes pieces round-robin; a piece is one sweep of 256 threads x 8 x 16 B = 32 KB, all loads of a thread
// in flight before its first store (a PCIe read round trip is ~1.5 us).  Loads are system-scope (never a stale line).
// ------------------------------------------------------------------------------------------------
constexpr int kPullPieceBytes = 256 * 8 * 16;
__global__ void __launch_bounds__(256)
stage_rows_pull_kernel(const char* __restrict__ host_base, const long long* __restrict__ table,
                       const int* __restrict__ piece_prefix, int n_seg, char* __restrict__ dst) {
  const int n_pieces = piece_prefix[n_seg];
  for (int piece = blockIdx.x; piece < n_pieces; piece += gridDim.x) {
    const int s = find_segment(piece_prefix, n_seg, piece);
    const long long off = (long long)(piece - piece_prefix[s]) * kPullPieceBytes;
    const long long len = min((long long)kPullPieceBytes, table[3 * s + 2] - off);
    const uint4* src = reinterpret_cast<const uint4*>(host_base + table[3 * s] + off);
    uint4* out = reinterpret_cast<uint4*>(dst + table[3 * s + 1] + off);
    const int n_vec = (int)(len >> 4);
    uint4 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int i = u * 256 + threadIdx.x;
      if (i < n_vec)
        asm volatile("ld.relaxed.sys.global.v4.u32 {%0,%1,%2,%3}, [%4];"
                     : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w) : "l"(src + i) : "memory");
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int i = u * 256 + threadIdx.x;
      if (i < n_vec) out[i] = v[u];
    }
  }
}

// ================================================================================================
// host launchers
// ================================================================================================
int rows_plan(const void* mask, int mask_bytes, int B, int L, int rc, int* counts, int* seq_off, int* chunk_off,
              int* n_rows, int* row_src, int* chunk_seq, cudaStream_t st) {
  if (B <= 0 || L <= 0) return set_error(-1, "rows_plan: empty batch");
  // (count-and-fill in one launch exists — plan_fill_kernel<true> — but measured SLOWER at config 2: block b has to
  //  count b + 1 sequences behind one another, 26 us against 6 + 8 us for the two launches; it serves batches of a few
  //  sequences only)
  if (B <= 4 && (long long)B * L <= (1 << 14)) {
    stamp_begin(st);
    plan_fill_kernel<true><<<B, 256, 0, st>>>(mask, mask_bytes, B, L, counts, rc, seq_off, chunk_off, n_rows, row_src, chunk_seq);
    return check_launch("plan_fill_kernel", st);
  }
  plan_count_kernel<<<B, 256, 0, st>>>(mask, mask_bytes, B, L, counts);
  if (int rc_ = check_launch("plan_count_kernel", st)) return rc_;
  plan_fill_kernel<false><<<B, 256, 0, st>>>(mask, mask_bytes, B, L, counts, rc, seq_off, chunk_off, n_rows, row_src, chunk_seq);
  return check_launch("plan_fill_kernel", st);
}

// plan for rows that are ALREADY packed: only the per-sequence counts are given
int rows_plan_counts(const int* counts, int B, int rc, int* seq_off, int* chunk_off, int* n_rows, int* chunk_seq,
                     cudaStream_t st) {
  if (B <= 0) return set_error(-1, "rows_plan_counts: empty batch");
  plan_fill_kernel<false><<<B, 256, 0, st>>>(nullptr, 0, B, 0, const_cast<int*>(counts), rc, seq_off, chunk_off, n_rows, nullptr, chunk_seq);
  return check_launch("plan_fill_kernel", st);
}

int gather_rows(const void* src, long long ld_src, const int* row_src, const int* n_rows, int cap, int D, void* out,
                cudaStream_t st) {
  if (D % 8) return set_error(-1, "gather_rows: D must be a multiple of 8");
  const long long total = (long long)cap * (D / 8);
  const int blocks = (int)std::min<long long>((total + 255) / 256, (long long)sm_count() * 16);
  stamp_begin(st);
  gather_rows_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(src), ld_src, row_src, n_rows, cap, D,
                                             reinterpret_cast<__nv_bfloat16*>(out));
  return check_launch("gather_rows_kernel", st);
}

int row_inv_norm(const float* rowsq, int nblk, const int* n_rows, int cap, float* inv_norm, cudaStream_t st) {
  const int blocks = std::min((cap + 31) / 32, sm_count() * 8);
  stamp_begin(st);
  row_inv_norm_kernel<<<blocks, 256, 0, st>>>(rowsq, nblk, n_rows, cap, inv_norm);
  return check_launch("row_inv_norm_kernel", st);
}

int pool_forward(const void* src, bool src_is_f16, long long ld_src, int src_rows, const int* row_src, const float* inv_norm,
                 const int* seq_off, const int* chunk_off, const int* chunk_seq, int B, int D, int rc, int max_chunks,
                 int mode, float2* partial, float* out, long long ld_out, void* norm_p_bf16, float* norm_p_f32,
                 float* norm_out, cudaStream_t st) {
  if (D % 8) return set_error(-1, "pool_forward: D must be a multiple of 8");
  if (rc != POOL_BOX_ROWS) return set_error(-1, "pool_forward: chunk_rows must be 64");
  if (src_rows <= 0) return set_error(-1, "pool_forward: source row count required");
  const int slabs = (D + POOL_BOX_COLS - 1) / POOL_BOX_COLS;
  const int per_slab = std::max(1, std::min(max_chunks, sm_count() / slabs));  // at most one CTA per SM: a single wave
  dim3 grid(per_slab, slabs);
  const __nv_bfloat16* sp = reinterpret_cast<const __nv_bfloat16*>(src);
  const int4* desc = reinterpret_cast<const int4*>(chunk_seq);
  CUtensorMap tmap;
  if (int r = make_tmap_16bit(&tmap, src, D, src_rows, ld_src, POOL_BOX_COLS, POOL_BOX_ROWS, CU_TENSOR_MAP_SWIZZLE_NONE)) return r;
  const int smem = (int)sizeof(PoolSmem) + 1024;
  static bool configured_dev[64] = {false};  // per-device setting
  int cur_dev = 0;
  cudaGetDevice(&cur_dev);
  if (cur_dev < 0 || cur_dev >= 64) cur_dev = 0;
  bool& configured = configured_dev[cur_dev];
  if (!configured) {
    cudaError_t e1 = cudaFuncSetAttribute(pool_partial_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaError_t e2 = cudaFuncSetAttribute(pool_partial_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e1 != cudaSuccess || e2 != cudaSuccess) return set_error((int)(e1 != cudaSuccess ? e1 : e2), "pool_forward: cannot reserve %d bytes of shared memory", smem);
    configured = true;
  }
  stamp_begin(st);
  int* seg_rows = reinterpret_cast<int*>(partial + (size_t)max_chunks * D);  // behind the partial records
  if (src_is_f16)
    pool_partial_kernel<true><<<grid, POOL_THREADS, smem, st>>>(tmap, sp, ld_src, row_src, inv_norm, chunk_off, desc, B, D, partial, seg_rows);
  else
    pool_partial_kernel<false><<<grid, POOL_THREADS, smem, st>>>(tmap, sp, ld_src, row_src, inv_norm, chunk_off, desc, B, D, partial, seg_rows);
  if (int r = check_launch("pool_partial_kernel", st)) return r;
  if (norm_p_bf16 != nullptr || norm_p_f32 != nullptr) {
    if (mode != 3 || ld_out != 2LL * D) return set_error(-1, "pool_forward: fused normalise needs the dense 'mix' layout");
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(FIN_CLUSTER, B);
    cfg.blockDim = dim3(FIN_THREADS);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = FIN_CLUSTER;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    stamp_begin(st);
    cudaError_t le = cudaLaunchKernelEx(&cfg, pool_finalize_normalize_kernel, (const float2*)partial, (const int*)seg_rows, seq_off,
                                        chunk_off, D, out, reinterpret_cast<__nv_bfloat16*>(norm_p_bf16), norm_p_f32, norm_out);
    if (le != cudaSuccess) return set_error((int)le, "pool_finalize_normalize_kernel: %s", cudaGetErrorString(le));
    count_launch();
    stamp_launch("pool_finalize_normalize_kernel", st);
    return 0;
  }
  dim3 g2(B, (D + 255) / 256);
  pool_finalize_kernel<<<g2, 256, 0, st>>>(partial, seg_rows, seq_off, chunk_off, B, D, mode, out, ld_out);
  return check_launch("pool_finalize_kernel", st);
}

int l2norm_forward(const float* e, int B, int E, void* p_bf16, float* p_f32, float* norm, cudaStream_t st) {
  l2norm_fwd_kernel<<<B, 256, 0, st>>>(e, E, reinterpret_cast<__nv_bfloat16*>(p_bf16), p_f32, norm);
  return check_launch("l2norm_fwd_kernel", st);
}
int l2norm_backward(const float* dp, const float* p, const float* norm, int B, int E, float* de, cudaStream_t st) {
  l2norm_bwd_kernel<<<B, 256, 0, st>>>(dp, p, norm, E, de);
  return check_launch("l2norm_bwd_kernel", st);
}
int pool_bwd_coef(const float* de, long long ld_de, const float* stats, long long ld_stats, const int* seq_off, int B,
                  int D, int mode, float* c1, float* c2, cudaStream_t st) {
  dim3 g(B, (D + 255) / 256);
  pool_bwd_coef_kernel<<<g, 256, 0, st>>>(de, ld_de, stats, ld_stats, seq_off, D, mode, c1, c2);
  return check_launch("pool_bwd_coef_kernel", st);
}

int loss_bwd_coef(const float* dS, const float* t, const float* p, const float* pnorm, const float* stats,
                  const int* seq_off, const float* dloss, int R, int B, int C, int D, float tau, float* dp_ws,
                  float* c1, float* c2, cudaStream_t st) {
  if (D % 4) return set_error(-1, "loss_bwd_coef: D must be a multiple of 4");
  const int E = 2 * D;
  const int nslice = (E + LBW_SLICE - 1) / LBW_SLICE;
  float* dp = dp_ws;                         // [B][E]
  float* dotp = dp_ws + (size_t)B * E;       // [B][nslice]
  loss_bwd_dp_kernel<<<nslice, 256, 0, st>>>(dS, t, p, dloss, R, B, C, E, 1.f / tau, dp, dotp);
  if (int r = check_launch("loss_bwd_dp_kernel", st)) return r;
  loss_bwd_coef_kernel<<<B, 1024, 0, st>>>(dp, dotp, nslice, p, pnorm, stats, seq_off, D, c1, c2);
  return check_launch("loss_bwd_coef_kernel", st);
}

int adapter_tail_backward(const void* a, const void* g, const float* inv_norm, const int* seq_off, int B, const float* c1,
                          const float* c2, const int* n_rows, int cap, int D, void* dz2, float* colsum_partial, int ws_rows,
                          int* nparts_dev, void* db2, cudaStream_t st) {
  if (D % 8 || D > 8192) return set_error(-1, "adapter_tail_backward: D must be a multiple of 8 and <= 8192");
  const __half* ap = reinterpret_cast<const __half*>(a);
  const __half* gp = reinterpret_cast<const __half*>(g);
  __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(dz2);
  const int threads = D <= 2048 ? 256 : (D <= 4096 ? 512 : 1024);
  int grid = sm_count() * std::max(1, 512 / threads);
  if (colsum_partial != nullptr) {
    if (ws_rows < 1 || nparts_dev == nullptr) return set_error(-1, "adapter_tail_backward: the column-sum workspace needs rows and a count word");
    grid = std::min(grid, ws_rows);
  }
  grid = std::max(1, std::min(grid, (cap + 2 * TAIL_R - 1) / (2 * TAIL_R)));
  stamp_begin(st);
  if (threads == 256)
    adapter_tail_bwd_kernel<256><<<grid, 256, 0, st>>>(ap, gp, inv_norm, seq_off, B, c1, c2, n_rows, cap, D, op, colsum_partial, nparts_dev);
  else if (threads == 512)
    adapter_tail_bwd_kernel<512><<<grid, 512, 0, st>>>(ap, gp, inv_norm, seq_off, B, c1, c2, n_rows, cap, D, op, colsum_partial, nparts_dev);
  else
    adapter_tail_bwd_kernel<1024><<<grid, 1024, 0, st>>>(ap, gp, inv_norm, seq_off, B, c1, c2, n_rows, cap, D, op, colsum_partial, nparts_dev);
  if (int r = check_launch("adapter_tail_bwd_kernel", st)) return r;
  if (db2 != nullptr) {
    if (!colsum_partial) return set_error(-1, "adapter_tail_backward: db2 needs the partial workspace");
    BiasJob none{};
    BiasJob j{};
    j.partial = colsum_partial; j.D = D; j.nparts_dev = nparts_dev; j.nparts_max = ws_rows; j.out_bf16 = db2;
    return bias_grads_final(none, j, st);
  }
  return 0;
}

int bias_grads_final(const BiasJob& j0, const BiasJob& j1, cudaStream_t st) {
  const int blocks0 = j0.partial ? (j0.D + 31) / 32 : 0, blocks1 = j1.partial ? (j1.D + 31) / 32 : 0;
  if (blocks0 + blocks1 == 0) return 0;
  stamp_begin(st);
  bias_grads_final_kernel<<<blocks0 + blocks1, dim3(32, PARTS_Y), 0, st>>>(j0, j1, blocks0);
  return check_launch("bias_grads_final_kernel", st);
}

template <int NV>
static int tail_bwd_dy_launch(const void* a, const void* g, const float* inv_norm, const void* dy, int n,
                              const int* n_dev, int cap, int D, void* dz2, cudaStream_t st) {
  const int blocks = min((cap + 7) / 8, sm_count() * 8);
  adapter_tail_bwd_dy_kernel<NV><<<blocks, 256, 0, st>>>(
      reinterpret_cast<const __half*>(a), reinterpret_cast<const __half*>(g), inv_norm,
      reinterpret_cast<const __nv_bfloat16*>(dy), n, n_dev, cap, D, reinterpret_cast<__nv_bfloat16*>(dz2));
  return check_launch("adapter_tail_bwd_dy_kernel", st);
}
int adapter_tail_backward_dy(const void* a, const void* g, const float* inv_norm, const void* dy, int n,
                             const int* n_dev, int cap, int D, void* dz2, cudaStream_t st) {
  if (D % 8 || D > 4096) return set_error(-1, "adapter_tail_backward_dy: D must be a multiple of 8 and <= 4096");
  if (D <= 2048) return tail_bwd_dy_launch<8>(a, g, inv_norm, dy, n, n_dev, cap, D, dz2, st);
  return tail_bwd_dy_launch<16>(a, g, inv_norm, dy, n, n_dev, cap, D, dz2, st);
}

int scale_rows(const void* a, const float* rowsq, int nblk, int cap, int n, int D, void* y, float* inv_norm_out,
               cudaStream_t st) {
  return scatter_scaled_rows(a, rowsq, nblk, cap, n, D, y, D, nullptr, nullptr, nullptr, inv_norm_out, st);
}

int scatter_scaled_rows(const void* a, const float* rowsq, int nblk, int cap, int n, int D, void* y, long long ld_y,
                        const int* row_dst, const int* n_dev, const int* n_dst_dev, float* inv_norm_out, cudaStream_t st) {
  if (n <= 0) return 0;
  if (D % 8 || ld_y % 8) return set_error(-1, "scale_rows: D and the destination row stride must be multiples of 8");
  const int blocks = min((n + 7) / 8, sm_count() * 8);
  scale_rows_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<const __half*>(a), rowsq, nblk, cap, n, D,
                                            reinterpret_cast<__nv_bfloat16*>(y), ld_y, row_dst, n_dev, n_dst_dev, inv_norm_out);
  return check_launch("scale_rows_kernel", st);
}

int readout_backward(const void* x, const void* mask, int mask_bytes, int B, int S, int D, const float* c1,
                     const float* c2, void* dx, cudaStream_t st) {
  const long long total = (long long)B * S * (D / 8);
  const int blocks = (int)std::min<long long>((total + 255) / 256, (long long)sm_count() * 16);
  readout_bwd_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(x), mask, mask_bytes, B, S, D, c1, c2,
                                             reinterpret_cast<__nv_bfloat16*>(dx));
  return check_launch("readout_bwd_kernel", st);
}

int colsum(const void* x, const int* n_rows, int n_static, int D, float* partial, void* out_bf16, float* out_f32,
           cudaStream_t st) {
  if (D % 8) return set_error(-1, "colsum: D must be a multiple of 8");
  if (n_static <= 0) return set_error(-1, "colsum: empty matrix");
  dim3 g((n_static + 63) / 64, (D / 8 + 63) / 64);
  colsum_partial_kernel<<<g, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(x), n_rows, n_static, D, partial);
  if (int r = check_launch("colsum_partial_kernel", st)) return r;
  parts_colsum_final_kernel<<<(D + 31) / 32, dim3(32, PARTS_Y), 0, st>>>(partial, nullptr, 0, n_rows, n_static, D,
                                                                   reinterpret_cast<__nv_bfloat16*>(out_bf16), out_f32);
  return check_launch("parts_colsum_final_kernel", st);
}

int readout_last(const void* x, const int* counts, int B, int S, int D, float* out, cudaStream_t st) {
  readout_last_kernel<<<B, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(x), counts, S, D, out);
  return check_launch("readout_last_kernel", st);
}

int stage_rows_pull(const void* host_base, const long long* table, const int* piece_prefix, int n_seg, void* dst, int ctas,
                    cudaStream_t st) {
  if (n_seg <= 0) return 0;
  if (ctas <= 0) ctas = 32;
  stage_rows_pull_kernel<<<ctas, 256, 0, st>>>(static_cast<const char*>(host_base), table, piece_prefix, n_seg,
                                              static_cast<char*>(dst));
  return check_launch("stage_rows_pull_kernel", st);
}

int readout_last_bwd(const void* dout, const int* counts, int B, int S, int D, void* dx, cudaStream_t st) {
  const long long per = (long long)S * D;
  const int gx = (int)std::min<long long>((per + 255) / 256, 64);
  readout_last_bwd_kernel<<<dim3(gx, B), 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(dout), counts, S, D,
                                                       reinterpret_cast<__nv_bfloat16*>(dx));
  return check_launch("readout_last_bwd_kernel", st);
}

}  // namespace p2t
