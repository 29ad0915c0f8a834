// InfoNCE on a materialised, temperature-scaled similarity block S [R][C] (fp32):
//   loss = w_row * mean_i( lse_j S_ij - S_i,lab(i) ) + w_col * mean_i( lse_i' S_i',lab(i) - S_i,lab(i) )
// Reference: scripts/train_contrast.py:86-91 (BatchInfoNCELoss), :100-114 (SegmentedBatchInfoNCELoss);
// the column term is the same class with its arguments swapped (north_star's symmetric form).
//
// The cross-entropy is ONE pass over each row with an online (max, sum-exp) softmax that never
// materialises probabilities: the row is read once for the statistics and once to emit dS in place
// of the logits; the row/column argmax (retrieval indices) falls out of the same pass.
// Column statistics are produced as (max, sumexp) pairs so that ranks can combine them (multi-GPU).
#include "common.h"
#include "mathfn.cuh"
#include "rows.h"
#include <algorithm>

namespace p2t {

// ------------------------------------------------------------------------------------------------
// small-problem similarity on CUDA cores: S[i][j] = alpha * sum_e p[i][e] t[j][e]
// one warp per (i, 4 consecutive j); 16-byte loads.  (large problems use the tcgen05 GEMM)
// ------------------------------------------------------------------------------------------------
// fp32-embedding flavour: random-init / early-training embeddings are nearly parallel (loss ~ ln B), so the
// informative part of S and of dp = dS t is a small difference of large common components; rounding p and t
// to bf16 costs ~1 % of that difference.  The fused step therefore keeps p and t in fp32 on this path.
// one CTA per (row i, 4 consecutive columns j): its 8 warps split E, partial dot products meet in shared memory
template <bool F32>
__global__ void __launch_bounds__(256)
sim_small_kernel(const void* __restrict__ pv, const void* __restrict__ tv, int R, int C, int E, float alpha,
                 float* __restrict__ S) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int i = blockIdx.x, j0 = blockIdx.y * 4;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  if constexpr (F32) {
    const float* p = reinterpret_cast<const float*>(pv);
    const float* t = reinterpret_cast<const float*>(tv);
    const float4* pr = reinterpret_cast<const float4*>(p + (long long)i * E);
    const float4* tr[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) tr[q] = reinterpret_cast<const float4*>(t + (long long)min(j0 + q, C - 1) * E);
    const int nvec = E >> 2;
    for (int v = threadIdx.x; v < nvec; v += 256) {
      const float4 a = __ldg(pr + v);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 b = __ldg(tr[q] + v);
        acc[q] = fmaf(a.x, b.x, acc[q]);
        acc[q] = fmaf(a.y, b.y, acc[q]);
        acc[q] = fmaf(a.z, b.z, acc[q]);
        acc[q] = fmaf(a.w, b.w, acc[q]);
      }
    }
  } else {
    const __nv_bfloat16* p = reinterpret_cast<const __nv_bfloat16*>(pv);
    const __nv_bfloat16* t = reinterpret_cast<const __nv_bfloat16*>(tv);
    const uint4* pr = reinterpret_cast<const uint4*>(p + (long long)i * E);
    const uint4* tr[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) tr[q] = reinterpret_cast<const uint4*>(t + (long long)min(j0 + q, C - 1) * E);
    const int nvec = E >> 3;
    for (int v = threadIdx.x; v < nvec; v += 256) {
      const uint4 pu = __ldg(pr + v);
      const uint32_t pw[4] = {pu.x, pu.y, pu.z, pu.w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint4 tu = __ldg(tr[q] + v);
        const uint32_t tw[4] = {tu.x, tu.y, tu.z, tu.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 a = unpack_bf16x2(pw[k]), b = unpack_bf16x2(tw[k]);
          acc[q] = fmaf(a.x, b.x, acc[q]);
          acc[q] = fmaf(a.y, b.y, acc[q]);
        }
      }
    }
  }
  __shared__ float red[8][4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float s = warp_sum(acc[q]);
    if (lane == 0) red[warp][q] = s;
  }
  __syncthreads();
  if (threadIdx.x < 4 && j0 + threadIdx.x < C) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w][threadIdx.x];  // fixed order: deterministic
    S[(long long)i * C + j0 + threadIdx.x] = s * alpha;
  }
}

// Register-tiled form for blocks with more than a few hundred logits (the sharded step's B x B_global block):
// one CTA per 4 rows x 8 columns, its 256 threads split E (each thread: 12 independent 16-byte loads per trip feed
// 128 FMAs), so p and t are streamed from L2 (R/4)(C/8) x 12 row-lengths in total instead of R (C/4) x 5.
// Reduction in a fixed order (xor-shuffle tree, then warps 0..7): deterministic.
// Two shapes: 4 x 8 tiles with 256 threads from 4096 logits up; 2 x 4 tiles with 1024 threads below (4x the CTAs and
// a quarter of the dependent trips over E when there are too few tiles to fill the SMs).
template <int RT, int CT, int THREADS>
__global__ void __launch_bounds__(THREADS)
sim_tile_f32_kernel(const float* __restrict__ p, const float* __restrict__ t, int R, int C, int E, float alpha,
                    float* __restrict__ S) {
  constexpr int NW = THREADS / 32;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int i0 = blockIdx.x * RT, j0 = blockIdx.y * CT;
  const float4* pr[RT];
  const float4* tr[CT];
#pragma unroll
  for (int a = 0; a < RT; ++a) pr[a] = reinterpret_cast<const float4*>(p + (long long)min(i0 + a, R - 1) * E);
#pragma unroll
  for (int b = 0; b < CT; ++b) tr[b] = reinterpret_cast<const float4*>(t + (long long)min(j0 + b, C - 1) * E);
  float acc[RT][CT];
#pragma unroll
  for (int a = 0; a < RT; ++a)
#pragma unroll
    for (int b = 0; b < CT; ++b) acc[a][b] = 0.f;
  const int nvec = E >> 2;
  for (int v = threadIdx.x; v < nvec; v += THREADS) {
    float4 pv[RT], tv[CT];
#pragma unroll
    for (int a = 0; a < RT; ++a) pv[a] = __ldg(pr[a] + v);
#pragma unroll
    for (int b = 0; b < CT; ++b) tv[b] = __ldg(tr[b] + v);
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
  __shared__ float red[NW][RT * CT];
#pragma unroll
  for (int a = 0; a < RT; ++a)
#pragma unroll
    for (int b = 0; b < CT; ++b) {
      const float s = warp_sum(acc[a][b]);
      if (lane == 0) red[warp][a * CT + b] = s;
    }
  __syncthreads();
  if (threadIdx.x < RT * CT) {
    const int i = i0 + threadIdx.x / CT, j = j0 + threadIdx.x % CT;
    if (i < R && j < C) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < NW; ++w) s += red[w][threadIdx.x];
      S[(long long)i * C + j] = s * alpha;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// column statistics: for each column j, online (max, sumexp, argmax-row) over the R rows.
// grid ceil(C/32), block (32, 8): thread (x, y) walks rows y, y+8, ... of column 32*blockIdx.x + x.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
col_stats_kernel(const float* __restrict__ S, int R, int C, float* __restrict__ col_max, float* __restrict__ col_sum,
                 int* __restrict__ col_argmax, int row_index_base) {
  const int j = blockIdx.x * 32 + threadIdx.x;
  float m = -INFINITY, s = 0.f;
  int am = 0x7fffffff;
  if (j < C) {
    for (int i = threadIdx.y; i < R; i += 8) {
      const float v = S[(long long)i * C + j];
      if (v > m) { s = s * __expf(m - v) + 1.f; m = v; am = i; }
      else s += __expf(v - m);
    }
  }
  __shared__ float sm[8][33], ss[8][33];
  __shared__ int sa[8][33];
  sm[threadIdx.y][threadIdx.x] = m;
  ss[threadIdx.y][threadIdx.x] = s;
  sa[threadIdx.y][threadIdx.x] = am;
  __syncthreads();
  if (threadIdx.y == 0 && j < C) {
    float M = sm[0][threadIdx.x], Ssum = ss[0][threadIdx.x];
    int A = sa[0][threadIdx.x];
    for (int y = 1; y < 8; ++y) {
      const float m2 = sm[y][threadIdx.x], s2 = ss[y][threadIdx.x];
      const int a2 = sa[y][threadIdx.x];
      if (m2 > M || (m2 == M && a2 < A)) {
        if (m2 > M) { Ssum = Ssum * __expf(M - m2) + s2; M = m2; } else Ssum += s2;
        A = a2;
      } else if (m2 > -INFINITY) {
        Ssum += s2 * __expf(m2 - M);
      }
    }
    col_max[j] = M;
    col_sum[j] = Ssum;
    if (col_argmax) col_argmax[j] = (A == 0x7fffffff) ? -1 : A + row_index_base;
  }
}

// ------------------------------------------------------------------------------------------------
// fused cross-entropy: one warp per row.
//   in : S [R][C] fp32, labels[R], optional column (max,sum) for the column term
//   out: row_loss[R] (both terms), row_lse[R], argmax_row[R]; dS (fp32 in place and/or bf16 copy)
// dS_ij = w_row/Rn (softmax_row_ij - d_ij) + w_col/Rn (exp(S_ij - lse_col_j) [j labelled here] - d_ij)
// `col_labelled[j]` != 0 marks columns whose positive lives in this row block; Rn = normaliser.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
ce_rows_kernel(float* __restrict__ S, const int* __restrict__ labels, int R, int C, float w_row, float w_col,
               float inv_rn, const float* __restrict__ col_max, const float* __restrict__ col_sum,
               const unsigned char* __restrict__ col_labelled, float* __restrict__ row_loss,
               float* __restrict__ row_lse, int* __restrict__ argmax_row, __nv_bfloat16* __restrict__ dS_bf16,
               int write_ds) {
  const int lane = threadIdx.x & 31;
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (i >= R) return;
  float* row = S + (long long)i * C;
  float m = -INFINITY, s = 0.f;
  int am = 0x7fffffff;
  for (int j = lane; j < C; j += 32) {
    const float v = row[j];
    if (v > m) { s = s * __expf(m - v) + 1.f; m = v; am = j; }
    else s += __expf(v - m);
  }
  // combine lanes: (max, sum, argmax with lowest-index tie-break)
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
  const float lse = m + __logf(s);
  const int lab_raw = labels[i];
  const bool lab_ok = lab_raw >= 0 && lab_raw < C;  // a label outside the block poisons the loss, never reads out of bounds
  const int lab = lab_ok ? lab_raw : 0;
  const float pos = row[lab];
  float loss = w_row * (lse - pos);
  if (w_col != 0.f) loss += w_col * (col_max[lab] + __logf(col_sum[lab]) - pos);
  if (!lab_ok) loss = __int_as_float(0x7fc00000);
  if (lane == 0) {
    row_loss[i] = loss;
    if (row_lse) row_lse[i] = lse;
    if (argmax_row) argmax_row[i] = am;
  }
  if (!write_ds) return;
  const float wr = w_row * inv_rn, wc = w_col * inv_rn;
  for (int j = lane; j < C; j += 32) {
    const float v = row[j];
    float d = wr * __expf(v - lse);
    if (w_col != 0.f && col_labelled[j]) d += wc * __expf(v - col_max[j]) / col_sum[j];
    if (j == lab) d -= (wr + wc);
    row[j] = d;
    if (dS_bf16) dS_bf16[(long long)i * C + j] = __float2bfloat16_rn(d);
  }
}

// deterministic mean of row_loss -> loss[0] (fp32); single block
__global__ void __launch_bounds__(256) loss_mean_kernel(const float* __restrict__ row_loss, int R, float scale,
                                                        float* __restrict__ loss, int accumulate) {
  __shared__ float sh[256];
  float s = 0.f;
  for (int i = threadIdx.x; i < R; i += 256) s += row_loss[i];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) loss[0] = (accumulate ? loss[0] : 0.f) + sh[0] * scale;
}

__global__ void mark_labelled_kernel(const int* __restrict__ labels, int R, int C, unsigned char* __restrict__ marks) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < R) {
    const int l = labels[i];
    if (l >= 0 && l < C) marks[l] = 1;
  }
}

// ------------------------------------------------------------------------------------------------
// small-problem gradient contraction on CUDA cores: out[i][e] = alpha * sum_j W[i][j] * X[j][e]
//   (dp = dS t / tau with W = dS [R][C], X = t [C][E];  dt = dS^T p / tau with W read transposed)
// thread = (row i, 8 columns e); grid (ceil(E/8/256), R)
// ------------------------------------------------------------------------------------------------
template <bool TRANSPOSE_W, bool X_F32>
__global__ void __launch_bounds__(256)
contract_small_kernel(const float* __restrict__ W, int ldw, const void* __restrict__ Xv, int n_out, int n_red,
                      int E, float alpha, float* __restrict__ out) {
  const int i = blockIdx.y;
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v * 8 >= E) return;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int j = 0; j < n_red; ++j) {
    const float w = TRANSPOSE_W ? W[(long long)j * ldw + i] : W[(long long)i * ldw + j];
    if constexpr (X_F32) {
      const float4* xr = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(Xv) + (long long)j * E) + 2 * v;
      const float4 a = __ldg(xr), b = __ldg(xr + 1);
      acc[0] = fmaf(w, a.x, acc[0]); acc[1] = fmaf(w, a.y, acc[1]); acc[2] = fmaf(w, a.z, acc[2]); acc[3] = fmaf(w, a.w, acc[3]);
      acc[4] = fmaf(w, b.x, acc[4]); acc[5] = fmaf(w, b.y, acc[5]); acc[6] = fmaf(w, b.z, acc[6]); acc[7] = fmaf(w, b.w, acc[7]);
    } else {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(Xv) + (long long)j * E) + v);
      const uint32_t x[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 f = unpack_bf16x2(x[k]);
        acc[2 * k] = fmaf(w, f.x, acc[2 * k]);
        acc[2 * k + 1] = fmaf(w, f.y, acc[2 * k + 1]);
      }
    }
  }
  float4* o = reinterpret_cast<float4*>(out + (long long)i * E) + 2 * v;
  o[0] = make_float4(acc[0] * alpha, acc[1] * alpha, acc[2] * alpha, acc[3] * alpha);
  o[1] = make_float4(acc[4] * alpha, acc[5] * alpha, acc[6] * alpha, acc[7] * alpha);
  (void)n_out;
}

// fp32 -> bf16 conversion with optional transpose-free copy (for feeding the tcgen05 GEMM)
__global__ void f32_to_bf16_kernel(const float* __restrict__ in, long long n, __nv_bfloat16* __restrict__ out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = __float2bfloat16_rn(in[i]);
}
__global__ void bf16_to_f32_kernel(const __nv_bfloat16* __restrict__ in, long long n, float* __restrict__ out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = __bfloat162float(in[i]);
}

// ------------------------------------------------------------------------------------------------
// Large blocks: the similarity GEMM's EPI_SIM_STATS epilogue (gemm_sm100.cuh) leaves per-tile partials instead of
// logits; these two kernels merge them in a fixed order.
//   columns: col_part [nrb][C] (max, sum exp, arg-max row) over 32-row blocks -> col_max, col_sum, argmax_col
//   rows   : row_part [npart][ld] (max, sum exp, arg-max column) over 64-column pieces -> row_lse, argmax_row,
//            row_loss = w_row (lse - S_i,lab) + w_col (lse_col[lab] - S_i,lab); the same launch writes
//            col_lse[j] = col_max[j] + log col_sum[j] for the dLogits pass
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
sim_cols_combine_kernel(const float4* __restrict__ col_part, int nrb, int C, float* __restrict__ col_max,
                        float* __restrict__ col_sum, int* __restrict__ col_argmax) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= C) return;
  float M = -INFINITY, S = 0.f;
  int A = -1;
  for (int rb = 0; rb < nrb; ++rb) {
    const float4 q = col_part[(long long)rb * C + j];
    const float m2 = q.x, s2 = q.y;
    if (!(m2 > -INFINITY)) continue;
    if (m2 > M) { S = S * __expf(M - m2) + s2; M = m2; A = __float_as_int(q.z); }  // ties keep the lower row block
    else S += s2 * __expf(m2 - M);
  }
  col_max[j] = M;
  col_sum[j] = S;
  if (col_argmax) col_argmax[j] = A;
}

__global__ void __launch_bounds__(256)
sim_finish_kernel(const float4* __restrict__ row_part, int npart, int ld, const float* __restrict__ pos,
                  const int* __restrict__ labels, int R, int C, float w_row, float w_col,
                  const float* __restrict__ col_max, const float* __restrict__ col_sum, float* __restrict__ col_lse,
                  float* __restrict__ row_loss, float* __restrict__ row_lse, int* __restrict__ argmax_row) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < C && col_lse != nullptr && col_max != nullptr) col_lse[i] = col_max[i] + __logf(col_sum[i]);
  if (i >= R) return;
  float M = -INFINITY, S = 0.f;
  int A = 0x7fffffff;
  for (int k = 0; k < npart; ++k) {
    const float4 q = row_part[(long long)k * ld + i];
    const float m2 = q.x, s2 = q.y;
    if (!(m2 > -INFINITY)) continue;
    if (m2 > M) { S = S * __expf(M - m2) + s2; M = m2; A = __float_as_int(q.z); }  // ties keep the lower column
    else S += s2 * __expf(m2 - M);
  }
  const float lse = M + __logf(S);
  const int lab = labels[i];
  const bool lab_ok = lab >= 0 && lab < C;
  float loss = __int_as_float(0x7fc00000);  // a label outside the block poisons the loss
  if (lab_ok) {
    const float ps = pos[i];
    loss = w_row * (lse - ps);
    if (w_col != 0.f) loss += w_col * (col_max[lab] + __logf(col_sum[lab]) - ps);
  }
  row_loss[i] = loss;
  if (row_lse) row_lse[i] = lse;
  if (argmax_row) argmax_row[i] = A;
}

// ================================================================================================
int sim_small(const void* p, const void* t, bool in_f32, int R, int C, int E, float alpha, float* S, cudaStream_t st) {
  if (E % 8) return set_error(-1, "similarity: E must be a multiple of 8");
  if (in_f32 && (long long)R * C >= 4096) {
    const dim3 tiles((R + 3) / 4, (C + 7) / 8);
    sim_tile_f32_kernel<4, 8, 256><<<tiles, 256, 0, st>>>(static_cast<const float*>(p), static_cast<const float*>(t), R, C, E, alpha, S);
    return check_launch("sim_tile_f32_kernel", st);
  }
  if (in_f32 && (long long)R * C >= 512) {
    const dim3 tiles((R + 1) / 2, (C + 3) / 4);
    sim_tile_f32_kernel<2, 4, 1024><<<tiles, 1024, 0, st>>>(static_cast<const float*>(p), static_cast<const float*>(t), R, C, E, alpha, S);
    return check_launch("sim_tile_f32_kernel", st);
  }
  const dim3 grid(R, (C + 3) / 4);
  if (in_f32) sim_small_kernel<true><<<grid, 256, 0, st>>>(p, t, R, C, E, alpha, S);
  else sim_small_kernel<false><<<grid, 256, 0, st>>>(p, t, R, C, E, alpha, S);
  return check_launch("sim_small_kernel", st);
}

int col_stats(const float* S, int R, int C, float* col_max, float* col_sum, int* col_argmax, int row_index_base,
              cudaStream_t st) {
  col_stats_kernel<<<(C + 31) / 32, dim3(32, 8), 0, st>>>(S, R, C, col_max, col_sum, col_argmax, row_index_base);
  return check_launch("col_stats_kernel", st);
}

int mark_labelled(const int* labels, int R, int C, unsigned char* marks, cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(marks, 0, C, st);
  if (e != cudaSuccess) return set_error((int)e, "memset: %s", cudaGetErrorString(e));
  mark_labelled_kernel<<<(R + 255) / 256, 256, 0, st>>>(labels, R, C, marks);
  return check_launch("mark_labelled_kernel", st);
}

int ce_rows(float* S, const int* labels, int R, int C, float w_row, float w_col, float inv_rn, const float* col_max,
            const float* col_sum, const unsigned char* col_labelled, float* row_loss, float* row_lse, int* argmax_row,
            void* dS_bf16, int write_ds, cudaStream_t st) {
  ce_rows_kernel<<<(R + 7) / 8, 256, 0, st>>>(S, labels, R, C, w_row, w_col, inv_rn, col_max, col_sum, col_labelled,
                                              row_loss, row_lse, argmax_row, reinterpret_cast<__nv_bfloat16*>(dS_bf16),
                                              write_ds);
  return check_launch("ce_rows_kernel", st);
}

int loss_mean(const float* row_loss, int R, float scale, float* loss, int accumulate, cudaStream_t st) {
  loss_mean_kernel<<<1, 256, 0, st>>>(row_loss, R, scale, loss, accumulate);
  return check_launch("loss_mean_kernel", st);
}

int contract_small(const float* W, int ldw, bool transpose_w, const void* X, bool x_f32, int n_out, int n_red, int E,
                   float alpha, float* out, cudaStream_t st) {
  if (E % 8) return set_error(-1, "contract: E must be a multiple of 8");
  dim3 g((E / 8 + 255) / 256, n_out);
  if (transpose_w) {
    if (x_f32) contract_small_kernel<true, true><<<g, 256, 0, st>>>(W, ldw, X, n_out, n_red, E, alpha, out);
    else contract_small_kernel<true, false><<<g, 256, 0, st>>>(W, ldw, X, n_out, n_red, E, alpha, out);
  } else {
    if (x_f32) contract_small_kernel<false, true><<<g, 256, 0, st>>>(W, ldw, X, n_out, n_red, E, alpha, out);
    else contract_small_kernel<false, false><<<g, 256, 0, st>>>(W, ldw, X, n_out, n_red, E, alpha, out);
  }
  return check_launch("contract_small_kernel", st);
}

int convert_f32_to_bf16(const float* in, long long n, void* out, cudaStream_t st) {
  const int blocks = (int)std::min<long long>((n + 255) / 256, (long long)sm_count() * 8);
  f32_to_bf16_kernel<<<blocks, 256, 0, st>>>(in, n, reinterpret_cast<__nv_bfloat16*>(out));
  return check_launch("f32_to_bf16_kernel", st);
}
int sim_cols_combine(const float4* col_part, int nrb, int C, float* col_max, float* col_sum, int* col_argmax, cudaStream_t st) {
  sim_cols_combine_kernel<<<(C + 255) / 256, 256, 0, st>>>(col_part, nrb, C, col_max, col_sum, col_argmax);
  return check_launch("sim_cols_combine_kernel", st);
}
int sim_finish(const float4* row_part, int npart, int ld, const float* pos, const int* labels, int R, int C, float w_row,
               float w_col, const float* col_max, const float* col_sum, float* col_lse, float* row_loss, float* row_lse,
               int* argmax_row, cudaStream_t st) {
  const int n = R > C ? R : C;
  sim_finish_kernel<<<(n + 255) / 256, 256, 0, st>>>(row_part, npart, ld, pos, labels, R, C, w_row, w_col, col_max, col_sum,
                                                     col_lse, row_loss, row_lse, argmax_row);
  return check_launch("sim_finish_kernel", st);
}

int convert_bf16_to_f32(const void* in, long long n, float* out, cudaStream_t st) {
  const int blocks = (int)std::min<long long>((n + 255) / 256, (long long)sm_count() * 8);
  bf16_to_f32_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(in), n, out);
  return check_launch("bf16_to_f32_kernel", st);
}

}  // namespace p2t
