// The loss block of the fused step for small similarity blocks (R*C <= 16384 logits: the single-GPU step and the
// B x B_global block of the sharded step), ONE cooperative kernel from the unit-norm embeddings to the pooling
// coefficients of the backward pass:
//
//   phase 0  (sharded step only) wait for the gathered text embeddings in the exchange channel (csrc/peer.cu flags)
//   phase 1  S = p t^T / tau on CUDA cores with fp32 embeddings (4x8 or 2x4 register tiles, E split over the CTA)
//   -------- grid barrier --------
//   phase 2  every CTA takes S into shared memory: online-softmax row statistics, column statistics (symmetric term,
//            text->protein retrieval), loss, argmax, dLogits in place of the logits (probabilities never exist in
//            memory); then its 64-column slice of dp = dLogits t / tau and of the row dot products p.dp
//   -------- grid barrier --------
//   phase 3  de = (dp - p (p.dp)) / |e| (autograd of F.normalize) -> 'mix' pooling coefficients c1, c2
//
// Reference: scripts/train_contrast.py:100-114 (SegmentedBatchInfoNCELoss: S, exp, sum, log, mean), :354/:365
// (F.normalize), :237-248 (mix readout) and their autograd.  Replaces six launches of the round-1 step (similarity,
// column statistics, row cross-entropy, loss mean, dLogits -> dp, dp -> coefficients), whose cost at 32 x 32 .. 32 x 256
// logits was launch latency, not work.  Every reduction runs in a fixed order: results are run-to-run identical.
#include "common.h"
#include "mathfn.cuh"
#include "peer_dev.cuh"
#include "rows.h"

#include <algorithm>

namespace p2t {
namespace {

constexpr int LF_THREADS = 256;
constexpr int LF_SLICE = 64;  // embedding columns per CTA trip of the dp phase
constexpr int LF_JC = 64;     // logit columns (= rows of t) staged per trip
constexpr float kEpsNormLF = 1e-12f;
constexpr unsigned long long kBarrierTimeoutNs = 4ull * 1000ull * 1000ull * 1000ull;  // 4 s

struct LossFusedParams {
  const float* p;      // [B][E] unit-norm protein embeddings; rows [0, R) enter the loss
  const float* t;      // [C][E] unit-norm text embeddings, or nullptr: taken from the exchange channel
  PeerTable peers;     // exchange channel of the gathered text embeddings (t == nullptr)
  int world, rank;
  long long vecs_per_rank;
  const int* labels;   // [R] column of each row's positive
  int R, B, C, E;
  float inv_tau, w_row, w_col, scale;  // scale: 1/R or the caller's normaliser
  int all_cols_labelled, want_col, need_grad;
  const float* dloss;  // optional device scalar multiplying the gradient
  const float* pnorm;  // [B] |e| before normalisation
  const float* stats;  // [B][E] (mean | std) of the pooled rows
  const int* seq_off;  // [B+1]
  float* S;            // workspace [R][C]
  float* dp;           // workspace [B][E]
  float* dotp;         // workspace [B][nslice]
  unsigned* bar;       // workspace, 4 words, zero at first use: [0] barrier count, [1] exit count, [2] error
  float* loss;
  float* row_lse;      // [R] or null
  int* argmax_row;     // [R] or null
  int* argmax_col;     // [C] or null
  float* col_max;      // [C] or null
  float* col_sum;      // [C] or null
  float* c1;           // [B][D]
  float* c2;           // [B][D]
};

__device__ __forceinline__ unsigned ld_acquire_gpu_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// All CTAs of the (cooperative, hence co-resident) grid meet; `target` = arrivals expected in total so far.
// A CTA that waits longer than 4 s records it and goes on: the kernel then ends with a NaN loss instead of hanging.
__device__ __forceinline__ void grid_barrier(unsigned* bar, unsigned target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(bar, 1u);
    const unsigned long long t0 = global_ns();
    while (ld_acquire_gpu_u32(bar) < target) {
      if (global_ns() - t0 > kBarrierTimeoutNs) {
        atomicExch(bar + 2, 1u);
        break;
      }
      __nanosleep(20);
    }
    __threadfence();
  }
  __syncthreads();
}

template <int RT, int CT>
__device__ __forceinline__ void similarity_phase(const LossFusedParams& q, const float* __restrict__ t, float* red /*[8][32]*/) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int tiles_r = (q.R + RT - 1) / RT, tiles_c = (q.C + CT - 1) / CT;
  const int nvec = q.E >> 2;
  for (int tile = blockIdx.x; tile < tiles_r * tiles_c; tile += gridDim.x) {
    const int i0 = (tile / tiles_c) * RT, j0 = (tile % tiles_c) * CT;
    const float4* pr[RT];
    const float4* tr[CT];
#pragma unroll
    for (int a = 0; a < RT; ++a) pr[a] = reinterpret_cast<const float4*>(q.p + (long long)min(i0 + a, q.R - 1) * q.E);
#pragma unroll
    for (int b = 0; b < CT; ++b) tr[b] = reinterpret_cast<const float4*>(t + (long long)min(j0 + b, q.C - 1) * q.E);
    float acc[RT][CT];
#pragma unroll
    for (int a = 0; a < RT; ++a)
#pragma unroll
      for (int b = 0; b < CT; ++b) acc[a][b] = 0.f;
    for (int v = threadIdx.x; v < nvec; v += LF_THREADS) {
      float4 pv[RT], tv[CT];
#pragma unroll
      for (int a = 0; a < RT; ++a) pv[a] = __ldg(pr[a] + v);
#pragma unroll
      for (int b = 0; b < CT; ++b) tv[b] = __ldcg(tr[b] + v);  // t may sit in a peer-written channel buffer: L2, never L1
#pragma unroll
      for (int a = 0; a < RT; ++a)
#pragma unroll
        for (int b = 0; b < CT; ++b) {
          acc[a][b] = fmaf(pv[a].x, tv[b].x, acc[a][b]);
          acc[a][b] = fmaf(pv[a].y, tv[b].y, acc[a][b]);
          acc[a][b] = fmaf(pv[a].z, tv[b].z, acc[a][b]);
          acc[a][b] = fmaf(pv[a].w, tv[b].w, acc[a][b]);
        }
    }
#pragma unroll
    for (int a = 0; a < RT; ++a)
#pragma unroll
      for (int b = 0; b < CT; ++b) {
        const float s = warp_sum(acc[a][b]);
        if (lane == 0) red[warp * 32 + a * CT + b] = s;
      }
    __syncthreads();
    if (threadIdx.x < RT * CT) {
      const int i = i0 + threadIdx.x / CT, j = j0 + threadIdx.x % CT;
      if (i < q.R && j < q.C) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < LF_THREADS / 32; ++w) s += red[w * 32 + threadIdx.x];  // fixed order
        q.S[(long long)i * q.C + j] = s * q.inv_tau;
      }
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(LF_THREADS, 1) loss_fused_kernel(const LossFusedParams q) {
  extern __shared__ __align__(16) float lf_smem[];
  // shared-memory carve-up (floats): dS [R*C] | ts [64][64] | red [256] | row_lse [R] | row_pos [R] | col_lse [C] |
  //                                  labels [R] (int) | marks [C] (bytes)
  float* Ssm = lf_smem;
  float* ts = Ssm + (((size_t)q.R * q.C + 3) & ~(size_t)3);
  float* red = ts + LF_JC * LF_SLICE;
  float* row_lse = red + LF_THREADS;
  float* row_pos = row_lse + q.R;
  float* col_lse = row_pos + q.R;
  int* lab_s = reinterpret_cast<int*>(col_lse + q.C);
  unsigned char* marks = reinterpret_cast<unsigned char*>(lab_s + q.R);
  __shared__ float s_bad;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int R = q.R, B = q.B, C = q.C, E = q.E;
  const int D = E >> 1;

  // ---------------- phase 0: the gathered text embeddings ----------------
  const float* t = q.t;
  unsigned* ctrl = nullptr;
  unsigned epoch = 0;
  bool gathered_ok = true;
  if (t == nullptr) {
    ctrl = static_cast<unsigned*>(q.peers.base[q.rank]);
    epoch = ctrl[0] + 1;
    gathered_ok = wait_flags(reinterpret_cast<const unsigned*>(static_cast<char*>(q.peers.base[q.rank]) + peer_flag_row_off(0)),
                             q.world, epoch, ctrl + 4);
    t = reinterpret_cast<const float*>(static_cast<char*>(q.peers.base[q.rank]) + kPeerCtrlBytes +
                                       (size_t)(epoch & 1) * q.world * (size_t)q.vecs_per_rank * sizeof(uint4));
  }

  // ---------------- phase 1: S = p t^T / tau ----------------
  if ((long long)((R + 3) / 4) * ((C + 7) / 8) >= (long long)gridDim.x) similarity_phase<4, 8>(q, t, red);
  else similarity_phase<2, 4>(q, t, red);
  grid_barrier(q.bar, gridDim.x);

  // ---------------- phase 2: statistics, loss, dLogits ----------------
  const bool lead = blockIdx.x == 0;
  if (q.need_grad || lead) {
    for (int idx = tid; idx < R * C; idx += LF_THREADS) Ssm[idx] = __ldcg(q.S + idx);
    for (int i = tid; i < R; i += LF_THREADS) lab_s[i] = q.labels[i];
    if (tid == 0) s_bad = 0.f;
    __syncthreads();
    // rows: one warp per row, online (max, sum-exp, argmax with lowest-index ties)
    for (int i = warp; i < R; i += LF_THREADS / 32) {
      const float* row = Ssm + (size_t)i * C;
      float m = -INFINITY, s = 0.f;
      int am = 0x7fffffff;
      for (int j = lane; j < C; j += 32) {
        const float v = row[j];
        if (v > m) { s = s * __expf(m - v) + 1.f; m = v; am = j; }
        else s += __expf(v - m);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float m2 = __shfl_xor_sync(0xffffffffu, m, o);
        const float s2 = __shfl_xor_sync(0xffffffffu, s, o);
        const int a2 = __shfl_xor_sync(0xffffffffu, am, o);
        const float M = fmaxf(m, m2);
        const float sa = (m > -INFINITY) ? s * __expf(m - M) : 0.f;
        const float sb = (m2 > -INFINITY) ? s2 * __expf(m2 - M) : 0.f;
        if (m2 > m || (m2 == m && a2 < am)) am = a2;
        m = M;
        s = sa + sb;
      }
      if (lane == 0) {
        const float lse = m + __logf(s);
        const int lab = lab_s[i];
        const bool lab_ok = lab >= 0 && lab < C;  // a label outside the block poisons the loss instead of reading out of bounds
        row_lse[i] = lse;
        row_pos[i] = lab_ok ? row[lab] : 0.f;
        if (!lab_ok) s_bad = 1.f;
        if (lead) {
          if (q.row_lse) q.row_lse[i] = lse;
          if (q.argmax_row) q.argmax_row[i] = am;
        }
      }
    }
    // columns: one thread per column
    const bool cols = q.w_col != 0.f || (lead && q.want_col);
    if (cols) {
      for (int j = tid; j < C; j += LF_THREADS) {
        float m = -INFINITY, s = 0.f;
        int am = -1;
        for (int i = 0; i < R; ++i) {
          const float v = Ssm[(size_t)i * C + j];
          if (v > m) { s = s * __expf(m - v) + 1.f; m = v; am = i; }
          else s += __expf(v - m);
        }
        col_lse[j] = m + __logf(s);
        if (lead) {
          if (q.col_max) q.col_max[j] = m;
          if (q.col_sum) q.col_sum[j] = s;
          if (q.argmax_col) q.argmax_col[j] = am;
        }
      }
    }
    if (q.w_col != 0.f) {
      for (int j = tid; j < C; j += LF_THREADS) marks[j] = q.all_cols_labelled ? 1 : 0;
      __syncthreads();
      if (!q.all_cols_labelled)
        for (int i = tid; i < R; i += LF_THREADS) {
          const int lab = lab_s[i];
          if (lab >= 0 && lab < C) marks[lab] = 1;
        }
    }
    __syncthreads();
    if (lead) {
      // loss = scale * sum_i [ w_row (lse_i - S_i,lab) + w_col (lse_col[lab] - S_i,lab) ], fixed-order tree
      float s = 0.f;
      for (int i = tid; i < R; i += LF_THREADS) {
        const int lab = min(max(lab_s[i], 0), C - 1);
        float l = q.w_row * (row_lse[i] - row_pos[i]);
        if (q.w_col != 0.f) l += q.w_col * (col_lse[lab] - row_pos[i]);
        s += l;
      }
      red[tid] = s;
      __syncthreads();
      for (int o = LF_THREADS / 2; o > 0; o >>= 1) {
        if (tid < o) red[tid] += red[tid + o];
        __syncthreads();
      }
      if (tid == 0) {
        const bool bad = s_bad != 0.f || !gathered_ok || ld_acquire_gpu_u32(q.bar + 2) != 0u;
        q.loss[0] = bad ? __int_as_float(0x7fc00000) : red[0] * q.scale;
      }
      __syncthreads();
    }
  }

  if (q.need_grad) {
    // dLogits in place of the logits (shared memory only)
    const float wr = q.w_row * q.scale, wc = q.w_col * q.scale;
    for (int idx = tid; idx < R * C; idx += LF_THREADS) {
      const int i = idx / C, j = idx - i * C;
      const float v = Ssm[idx];
      float d = wr * __expf(v - row_lse[i]);
      if (q.w_col != 0.f && marks[j]) d += wc * __expf(v - col_lse[j]);
      if (j == lab_s[i]) d -= (wr + wc);
      Ssm[idx] = d;
    }
    __syncthreads();
    // dp[:, slice] = (dloss / tau) dS t[:, slice] for all B rows (rows >= R were dropped by the segment split: 0)
    const int nslice = (E + LF_SLICE - 1) / LF_SLICE;
    const int cq = tid & 15, rg = tid >> 4;
    const float gscale = q.inv_tau * (q.dloss ? q.dloss[0] : 1.f);
    for (int slice = blockIdx.x; slice < nslice; slice += gridDim.x) {
      const int e0 = slice * LF_SLICE;
      const int col = e0 + cq * 4;
      for (int i0 = 0; i0 < B; i0 += 64) {
        float4 acc[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int j0 = 0; j0 < C; j0 += LF_JC) {
          __syncthreads();
#pragma unroll
          for (int m = 0; m < (LF_JC * LF_SLICE / 4) / LF_THREADS; ++m) {
            const int idx = tid + LF_THREADS * m;
            const int jr = idx >> 4, ev = (idx & 15) * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (j0 + jr < C && e0 + ev < E) v = __ldcg(reinterpret_cast<const float4*>(t + (long long)(j0 + jr) * E + e0 + ev));
            *reinterpret_cast<float4*>(&ts[jr * LF_SLICE + ev]) = v;
          }
          __syncthreads();
          const int jn = min(LF_JC, C - j0);
          const float* wrow[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) wrow[k] = Ssm + (size_t)min(i0 + rg + 16 * k, R - 1) * C + j0;
#pragma unroll 4
          for (int jj = 0; jj < jn; ++jj) {
            const float4 tv = *reinterpret_cast<const float4*>(&ts[jj * LF_SLICE + cq * 4]);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float w = (i0 + rg + 16 * k < R) ? wrow[k][jj] : 0.f;
              acc[k].x = fmaf(w, tv.x, acc[k].x); acc[k].y = fmaf(w, tv.y, acc[k].y);
              acc[k].z = fmaf(w, tv.z, acc[k].z); acc[k].w = fmaf(w, tv.w, acc[k].w);
            }
          }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int i = i0 + rg + 16 * k;
          float d = 0.f;
          if (i < B && col < E) {
            const float4 v = make_float4(acc[k].x * gscale, acc[k].y * gscale, acc[k].z * gscale, acc[k].w * gscale);
            *reinterpret_cast<float4*>(q.dp + (long long)i * E + col) = v;
            const float4 pv = *reinterpret_cast<const float4*>(q.p + (long long)i * E + col);
            d = fmaf(v.x, pv.x, fmaf(v.y, pv.y, fmaf(v.z, pv.z, v.w * pv.w)));
          }
#pragma unroll
          for (int o = 8; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
          if (cq == 0 && i < B) q.dotp[(long long)i * nslice + slice] = d;
        }
      }
    }
    grid_barrier(q.bar, 2u * gridDim.x);

    // ---------------- phase 3: pooling coefficients ----------------
    const int nchunk = (D + 4 * LF_THREADS - 1) / (4 * LF_THREADS);
    for (int u = blockIdx.x; u < B * nchunk; u += gridDim.x) {
      const int i = u / nchunk, chunk = u - i * nchunk;
      float dot = 0.f;
      for (int k = tid; k < nslice; k += LF_THREADS) dot += __ldcg(q.dotp + (long long)i * nslice + k);
      __syncthreads();
      dot = warp_sum(dot);
      if (lane == 0) red[warp] = dot;
      __syncthreads();
      dot = 0.f;
#pragma unroll
      for (int w = 0; w < LF_THREADS / 32; ++w) dot += red[w];
      const float nrm = q.pnorm[i];
      const bool clamped = nrm < kEpsNormLF;
      const float inv = 1.f / fmaxf(nrm, kEpsNormLF);
      const float n = (float)(q.seq_off[i + 1] - q.seq_off[i]);
      const float* st = q.stats + (long long)i * E;
      const float* pr = q.p + (long long)i * E;
      const float* dr = q.dp + (long long)i * E;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int col = chunk * 4 * LF_THREADS + k * LF_THREADS + tid;
        if (col < D) {
          const float d0 = __ldcg(dr + col), d1 = __ldcg(dr + D + col);
          const float dmu = clamped ? d0 * inv : (d0 - pr[col] * dot) * inv;
          const float dsd = clamped ? d1 * inv : (d1 - pr[D + col] * dot) * inv;
          const float mu = st[col], sd = st[D + col];
          const float k2 = dsd / (n * sd);  // inf/NaN when sd == 0, as autograd on the reference
          q.c1[(long long)i * D + col] = dmu / n - k2 * mu;
          q.c2[(long long)i * D + col] = k2;
        }
      }
    }
  }

  // ---------------- exit: the last CTA re-arms the barrier words and closes the exchange round ----------------
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    const unsigned done = atomicAdd(q.bar + 1, 1u);
    if (done == gridDim.x - 1) {
      q.bar[0] = 0u;
      q.bar[1] = 0u;
      q.bar[2] = 0u;
      if (ctrl != nullptr) ctrl[0] = epoch;
      __threadfence();
    }
  }
}

int lf_max_grid = 0;
int lf_max_smem = 0;
int lf_configured_device = -1;

}  // namespace

size_t loss_fused_smem_bytes(int R, int C) {
  const size_t floats = (((size_t)R * C + 3) & ~(size_t)3) + (size_t)LF_JC * LF_SLICE + LF_THREADS + 2 * (size_t)R + (size_t)C + (size_t)R;
  return floats * sizeof(float) + (size_t)C + 16;
}

// eligible: the logits block fits the shared memory of one SM next to the staging tiles
bool loss_fused_eligible(int R, int B, int C, int E) {
  return R >= 1 && R <= B && C >= 1 && (long long)R * C <= 16384 && E % 8 == 0 && E >= 8;
}

int loss_fused(const float* p, const float* t, void* const* peers, int world, int rank, long long bytes_per_rank,
               const int* labels, int R, int B, int C, int E, float tau, float w_row, float w_col, float scale,
               int all_cols_labelled, int want_col, int need_grad, const float* dloss, const float* pnorm,
               const float* stats, const int* seq_off, float* S_ws, float* dp_ws, unsigned* bar_ws, float* loss,
               float* row_lse, int* argmax_row, int* argmax_col, float* col_max, float* col_sum, float* c1, float* c2,
               cudaStream_t st) {
  if (!loss_fused_eligible(R, B, C, E)) return set_error(-1, "loss_fused: block of %d x %d logits (E = %d) is not eligible", R, C, E);
  if (tau <= 0.f) return set_error(-1, "loss_fused: temperature must be positive");
  LossFusedParams q{};
  q.p = p; q.t = t;
  if (t == nullptr) {
    if (!peers || world < 1 || world > kPeerMaxWorld || rank < 0 || rank >= world)
      return set_error(-1, "loss_fused: neither text embeddings nor a valid exchange channel given");
    if (bytes_per_rank <= 0 || bytes_per_rank % 16 || (long long)world * bytes_per_rank != (long long)C * E * 4)
      return set_error(-1, "loss_fused: the channel holds %lld bytes per rank, expected C*E*4/world", bytes_per_rank);
    for (int r = 0; r < world; ++r) {
      if (!peers[r]) return set_error(-1, "loss_fused: peer %d not mapped", r);
      q.peers.base[r] = peers[r];
    }
    q.world = world; q.rank = rank; q.vecs_per_rank = bytes_per_rank / 16;
  }
  q.labels = labels; q.R = R; q.B = B; q.C = C; q.E = E;
  q.inv_tau = 1.f / tau; q.w_row = w_row; q.w_col = w_col; q.scale = scale;
  q.all_cols_labelled = all_cols_labelled; q.want_col = want_col; q.need_grad = need_grad;
  q.dloss = dloss; q.pnorm = pnorm; q.stats = stats; q.seq_off = seq_off;
  q.S = S_ws; q.dp = dp_ws; q.dotp = dp_ws ? dp_ws + (size_t)B * E : nullptr; q.bar = bar_ws;
  q.loss = loss; q.row_lse = row_lse; q.argmax_row = argmax_row; q.argmax_col = argmax_col;
  q.col_max = col_max; q.col_sum = col_sum; q.c1 = c1; q.c2 = c2;

  int dev = 0;
  cudaGetDevice(&dev);
  if (lf_configured_device != dev) {
    int optin = 0;
    cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    cudaError_t e = cudaFuncSetAttribute(loss_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - 1024);
    if (e != cudaSuccess) return set_error((int)e, "loss_fused: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    lf_max_smem = optin - 1024;
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    lf_max_grid = sms;  // one CTA per SM: always co-resident under a cooperative launch
    lf_configured_device = dev;
  }
  const size_t smem = loss_fused_smem_bytes(R, C);
  if ((int)smem > lf_max_smem) return set_error(-1, "loss_fused: %zu bytes of shared memory needed, %d available", smem, lf_max_smem);
  const int nslice = (E + LF_SLICE - 1) / LF_SLICE;
  const int tiles = ((R + 1) / 2) * ((C + 3) / 4);
  int grid = std::min(lf_max_grid, std::max(nslice, std::min(tiles, lf_max_grid)));
  if (grid < 1) grid = 1;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(LF_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  stamp_begin(st);
  cudaError_t e = cudaLaunchKernelEx(&cfg, loss_fused_kernel, q);
  if (e != cudaSuccess) return set_error((int)e, "loss_fused_kernel: %s", cudaGetErrorString(e));
  count_launch();
  stamp_launch("loss_fused_kernel", st);
  return 0;
}

}  // namespace p2t
