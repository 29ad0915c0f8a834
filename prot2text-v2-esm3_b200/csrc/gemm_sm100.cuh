// Persistent, warp-specialised tcgen05 GEMM for sm_100a with fused epilogues.
//
//   D[M,N] = epilogue( sum_k A[m,k] * B[n,k] )      bf16 operands, fp32 accumulation in TMEM
//
// * operands arrive by TMA (128-byte swizzle) into a multi-stage shared-memory ring; either
//   operand may be K-major (memory [rows][K]) or MN-major (memory [K][rows]) — the latter is what
//   the weight-gradient GEMMs need (dW = dZ^T H with the long residue axis as K)
// * one elected thread issues tcgen05.mma (UMMA 128x256x16, or 256x256x16 across a CTA pair with
//   cta_group::2, each CTA loading half of B)
// * two 256-column fp32 accumulators live in TMEM so the epilogue of one piece of work overlaps the
//   MMAs of the next; 16 epilogue warps (four per SM sub-partition) read TMEM with tcgen05.ld, one
//   row per thread
// * M and/or K may be read from device memory (ragged batches: the residue-row count is produced
//   on the device by the packing kernel, no host sync)
// * work distribution: whole output tiles round-robin over the persistent workers.  When the last wave
//   would leave SMs idle (the weight gradients: 80 resp. 128 tiles of 256x256 for 74 CTA pairs) the
//   launcher asks for a split-K tail: the tiles of the incomplete last wave are cut into S K-ranges
//   each, so that R*S pieces fill the machine.  Pieces of one tile run side by side on neighbouring
//   workers over the same K range as their row/column neighbours (L2 sharing of operand slabs is kept,
//   which a free-running stream-K cut destroyed: measured slower).  A piece that does not hold a
//   tile's first K range dumps its raw fp32 accumulator to a workspace slot and bumps the tile's flag;
//   the piece that holds the first range owns the tile, adds the dumps in a fixed order
//   (deterministic) and runs the epilogue.  Every worker executes ALL its dump pieces first (they never
//   wait) and its owning pieces and whole tiles afterwards, so a wait can only target work that
//   needs nothing but an SM to run.
//
// Epilogues (reference sites in models/modeling_esm2llama_instruct.py:60-68 and its autograd):
//   EPI_STORE_BF16 / EPI_STORE_F32 : D = alpha * acc
//   EPI_FC1    : z1 = acc + b1;  D0 = h1 = keep*GELU(z1) (bf16);  D1 = g1 = keep*GELU'(z1) (fp16)   (:62-63)
//   EPI_FC2    : z2 = acc + b2;  D0 = a  = keep*GELU(z2) (fp16);  D1 = g2 = keep*GELU'(z2) (fp16);
//                rowsq[4*n_blk + part][row] = sum over 64 columns of a^2                         (:65-67)
//   (fp16 for tensors only our own streaming kernels read: bf16's 8-bit mantissa adds rounding noise
//    of 2^-9*|a| to every residue, which inflates the pooled std and its 1/std backward)
//   EPI_MUL_AUX: D0 = alpha * acc * aux   (fc2 dgrad chained into GELU'(z1): aux = g1); optionally the column sums of
//                D0 over each 32-row block (db1 = sum over rows of dz1, finished by bias_grads_final_kernel)
//   EPI_SIM_STATS: S = alpha * acc (the temperature-scaled similarity, scripts/train_contrast.py:108) is NEVER stored:
//                every warp reduces its 32 x 64 piece of the tile to per-row online-softmax partials (max, sum exp,
//                argmax), per-column partials over its 32 rows (max, sum exp, arg-max row) and the label logit
//                S[i][lab(i)] — the forward of the InfoNCE cross-entropy for blocks too large for one SM (:109-113)
//   EPI_SIM_DS : S recomputed tile by tile (recomputation is not counted) and turned into dLogits
//                = w_row (softmax_row - onehot) + w_col (softmax_col - onehot), written as the bf16 A operand of
//                dp = dLogits t.  Probabilities and fp32 logits never exist in memory.
// Storing (value, derivative) pairs instead of the pre-activation keeps every later HBM-bound pass
// free of erf/exp and of Philox re-generation: the dropout multiplier is folded into both.
#pragma once
#include <cuda.h>
#include <type_traits>
#include "ptx.cuh"
#include "mathfn.cuh"
#include "peer_dev.cuh"

namespace p2t {

enum GemmEpilogue : int { EPI_STORE_BF16 = 0, EPI_STORE_F32 = 1, EPI_FC1 = 2, EPI_FC2 = 3, EPI_MUL_AUX = 4,
                          EPI_SIM_STATS = 5, EPI_SIM_DS = 6 };

constexpr int GEMM_BLOCK_M = 128;  // rows per CTA
constexpr int GEMM_BLOCK_N = 256;  // UMMA N
constexpr int GEMM_BLOCK_K = 64;   // 64 bf16 = one 128-byte swizzle row
constexpr int GEMM_UMMA_K = 16;
constexpr int GEMM_EPI_WARPS = 16;
constexpr int GEMM_THREADS = 64 + 32 * GEMM_EPI_WARPS;  // warp0 TMA, warp1 MMA + TMEM alloc, warps 2.. epilogue
constexpr int GEMM_TMEM_COLS = 512;
constexpr int GEMM_SK_SLOT_FLOATS = GEMM_BLOCK_M * GEMM_BLOCK_N;  // one CTA's raw accumulator (128 KB)
constexpr int GEMM_SK_FLAG_BYTES = 4096;                           // flags live in front of the slots
constexpr int GEMM_SK_SLOTS_PER_SM = 3;                            // R*S <= 3 * workers
constexpr int GEMM_SK_MAX_SPLITS = 16;

struct GemmParams {
  int m, n, k;          // problem size (upper bounds when dyn_* are set)
  const int* dyn_m;     // optional device scalar: actual M (<= m)
  const int* dyn_k;     // optional device scalar: actual K (<= k)
  int rows_cap;         // rows physically present in D0/D1/aux (>= m rounded as allocated)
  int a_extent;         // rows physically present in A's storage (0 -> m, or k when MN-major); TMA zero-fills beyond
  int b_extent;         // same for B (0 -> n, or k when MN-major)
  void* d0;
  long long ldd0;
  void* d1;
  long long ldd1;
  const __nv_bfloat16* bias;  // [n]
  const __half* aux;          // EPI_MUL_AUX: fp16 multiplier [rows][ldaux]
  long long ldaux;
  float* rowsq;               // EPI_FC2: [4 * n_blocks][ld_rowsq = rows_cap] partial sums of squares (four per 256-column N block)
  int ld_rowsq;
  float* colsum;              // EPI_MUL_AUX (optional): [ceil(rows_cap / 32)][n] column sums of D0 over each 32-row block
                              // (fixed-order shuffle tree; the bias gradient db1 without a second pass over dz1)
  // EPI_SIM_STATS / EPI_SIM_DS (InfoNCE on large blocks; M = rows R, N = columns C)
  const int* sim_labels;      // [M] column of each row's positive
  float4* sim_row_part;       // STATS out: [4 * n_blocks][ld_rowsq] (max, sum exp, argmax as int bits, -)
  float4* sim_col_part;       // STATS out: [ceil(M / 32)][n] (max, sum exp, arg-max row as int bits, -), or nullptr
  float* sim_pos;             // STATS out: [M] S[i][lab(i)]
  const float* sim_row_lse;   // DS in: [M]
  const float* sim_col_lse;   // DS in: [n] (column term) or nullptr
  const unsigned char* sim_col_marks;  // DS in: [n] 1 = the column's positive is among the rows of the (global) batch
  float sim_wr, sim_wc;       // DS: w_row * scale, w_col * scale
  int accumulate;             // EPI_STORE_BF16 / EPI_STORE_F32: D0 += alpha * acc (read-modify-write: gradient accumulation
                              // over micro-batches, scripts/train_contrast.py:448 without optimizer.zero_grad in between)
  float alpha;
  DropoutParams drop;         // p == 0 -> disabled
  // split-K tail (see header comment); sk_ws == nullptr -> whole tiles only
  void* sk_ws;                // [GEMM_SK_FLAG_BYTES of int flags][GEMM_SK_SLOTS_PER_SM * #SM slots of 128x256 fp32]
  int sk_splits;              // S, set by the launcher (1 = whole tiles only; -1 = chosen on the device from the actual
                              // tile count, for GEMMs whose M is a device scalar)
  int tma_store;              // FC epilogues: tmap_d0/tmap_d1 are valid, write outputs through shared memory + TMA
  int debug_nostore;          // timing aid (P2T_DEBUG_NOSTORE): 1 = run the epilogue math, skip its global stores; 2 / 3: see the
                              // staged-store path (results are wrong in all three modes)
};

// The FC epilogues stage their two 16-bit outputs in shared memory (one 32-row x 32-column box per output and
// epilogue warp, 64-byte swizzle) and write them with TMA stores; they give up pipeline stages for it.
template <int CTA_GROUP, int EPI>
struct GemmSmem {
  static constexpr bool TMA_STORE = (EPI == EPI_FC1 || EPI == EPI_FC2);
  static constexpr int A_BYTES = GEMM_BLOCK_M * GEMM_BLOCK_K * 2;                 // 16 KB
  static constexpr int B_ROWS = GEMM_BLOCK_N / CTA_GROUP;                         // rows of B this CTA loads
  static constexpr int B_BYTES = B_ROWS * GEMM_BLOCK_K * 2;                       // 32 / 16 KB
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  // (5 x 32 KB of operands in flight per SM is what keeps the MMA fed from L2: with 4 stages ncu showed the epilogue
  //  warps waiting for the accumulator 29 % of the time while the tensor pipe was 63 % active)
#ifndef P2T_FC_STAGES
#define P2T_FC_STAGES 5
#endif
#ifndef P2T_ST_STAGES
#define P2T_ST_STAGES 6
#endif
  static constexpr int STAGES = TMA_STORE ? ((CTA_GROUP == 1) ? 3 : P2T_FC_STAGES) : ((CTA_GROUP == 1) ? 4 : P2T_ST_STAGES);
  static constexpr int BAR_BYTES = 1024;
  static constexpr int OUT_BOX_BYTES = 32 * 32 * 2;        // one staged output box
  static constexpr int OUT_BYTES = TMA_STORE ? GEMM_EPI_WARPS * 2 * OUT_BOX_BYTES : 0;
  static constexpr int TOTAL = STAGES * STAGE_BYTES + BAR_BYTES + OUT_BYTES + 1024;  // + slack for 1024-B alignment
};

// ------------------------------------------------------------------------------------------------
// work distribution: identical sequence of pieces (tile, kb0, kb1) in every role of a worker.
//   tiles [0, t_full)           whole, round-robin (t_full = complete waves; = num_tiles when S == 1)
//   tiles [t_full, num_tiles)   cut into S K-ranges each: tail piece idx = r*S + s on worker idx % W
// order per worker: its dump pieces (s > 0), then its owning pieces (s == 0), then its whole tiles.
// ------------------------------------------------------------------------------------------------
struct PieceIter {
  int num_tiles, num_kb, worker, W, S, t_full, n_tail;
  int phase, idx;
  __device__ __forceinline__ PieceIter(int tiles, int kb, int w, int nw, int splits)
      : num_tiles(tiles), num_kb(kb), worker(w), W(nw), S(splits) {
    t_full = (S > 1) ? (tiles / nw) * nw : tiles;
    n_tail = (tiles - t_full) * S;
    phase = (S > 1) ? 0 : 2;
    idx = (S > 1) ? w : first_whole();
  }
  __device__ __forceinline__ int first_whole() const { return ((worker - n_tail) % W + W) % W; }
  // tail_r: index of the tile among the tail tiles (flag / slot addressing), -1 for a whole tile; s: K-range index
  __device__ __forceinline__ bool next(int& tile, int& kb0, int& kb1, int& tail_r, int& s) {
    while (phase < 2) {
      while (idx < n_tail) {
        const int r = idx / S, ss = idx - r * S;
        idx += W;
        if ((ss != 0) == (phase == 0)) {
          tile = t_full + r; tail_r = r; s = ss;
          kb0 = (int)((long long)num_kb * ss / S);
          kb1 = (int)((long long)num_kb * (ss + 1) / S);
          return true;
        }
      }
      ++phase;
      idx = (phase == 1) ? worker : first_whole();
    }
    if (idx >= t_full) return false;
    tile = idx; tail_r = -1; s = 0; kb0 = 0; kb1 = num_kb;
    idx += W;
    return true;
  }
};

// The launcher's cost model for the split-K tail, evaluated on the device when M (hence the tile count) is only known
// there: the tiles of the incomplete last wave are cut into the S that minimises rounds(S) / S + a per-dump charge.
__host__ __device__ inline int choose_sk_splits(int tiles, int kb, int workers) {
  const int tail = tiles % workers;
  if (tail == 0) return 1;
  float best = 1.f;
  int best_s = 1;
  for (int sp = 2; sp <= GEMM_SK_MAX_SPLITS; ++sp) {
    if ((long long)tail * sp > (long long)GEMM_SK_SLOTS_PER_SM * workers || kb / sp < 4) break;
    const int rounds = (tail * sp + workers - 1) / workers;
    const float cost = (float)rounds / sp + 0.004f * (sp - 1);
    if (cost < best - 1e-6f) { best = cost; best_s = sp; }
  }
  return (best_s > 1 && best <= 0.92f) ? best_s : 1;
}

__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// ------------------------------------------------------------------------------------------------
// Epilogue math, 8-column groups (= one 16-byte store per output, one Philox block of dropout bits)
// ------------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ void store8_vec(T* dst, const float (&f)[8]) {
  if constexpr (std::is_same<T, __half>::value) {
    *reinterpret_cast<uint4*>(dst) =
        make_uint4(pack_f16x2(f[0], f[1]), pack_f16x2(f[2], f[3]), pack_f16x2(f[4], f[5]), pack_f16x2(f[6], f[7]));
  } else if constexpr (std::is_same<T, __nv_bfloat16>::value) {
    *reinterpret_cast<uint4*>(dst) =
        make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
  } else {
    reinterpret_cast<float4*>(dst)[0] = make_float4(f[0], f[1], f[2], f[3]);
    reinterpret_cast<float4*>(dst)[1] = make_float4(f[4], f[5], f[6], f[7]);
  }
}
template <typename T>
__device__ __forceinline__ void store8(T* dst, const float (&f)[8], int n_ok, bool vec_ok) {
  if (n_ok == 8 && vec_ok) {  // vec_ok: every row of the tensor starts 16-byte aligned
    store8_vec(dst, f);
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (i < n_ok) {
        if constexpr (std::is_same<T, __half>::value) dst[i] = __float2half_rn(f[i]);
        else if constexpr (std::is_same<T, __nv_bfloat16>::value) dst[i] = __float2bfloat16_rn(f[i]);
        else dst[i] = f[i];
      }
  }
}
template <typename T>
__device__ __forceinline__ void load8(const T* src, float (&f)[8], int n_ok) {
  if (n_ok == 8 && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(src));
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 x = unpack_h2<std::is_same<T, __half>::value>(w[i]);
      f[2 * i] = x.x;
      f[2 * i + 1] = x.y;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if constexpr (std::is_same<T, __half>::value) f[i] = (i < n_ok) ? __half2float(src[i]) : 0.f;
      else f[i] = (i < n_ok) ? __bfloat162float(src[i]) : 0.f;
    }
  }
}

// read-modify-write partner of store8 (accumulating epilogue): ordinary loads, never the read-only path
__device__ __forceinline__ void load8_plain(const float* src, float (&f)[8], int n_ok, bool vec_ok) {
  if (n_ok == 8 && vec_ok) {
    const float4 a = reinterpret_cast<const float4*>(src)[0], b = reinterpret_cast<const float4*>(src)[1];
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = (i < n_ok) ? src[i] : 0.f;
  }
}
__device__ __forceinline__ void load8_plain(const __nv_bfloat16* src, float (&f)[8], int n_ok, bool vec_ok) {
  if (n_ok == 8 && vec_ok) {
    const uint4 u = *reinterpret_cast<const uint4*>(src);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 x = unpack_bf16x2(w[i]);
      f[2 * i] = x.x;
      f[2 * i + 1] = x.y;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = (i < n_ok) ? __bfloat162float(src[i]) : 0.f;
  }
}

// Column sums of a 32 x 32 block held one row per lane (v[c] = element (lane, c)): a butterfly that halves the
// number of live columns per step (16 + 8 + 4 + 2 + 1 shuffles).  Returns, in lane l, the sum of column l over the
// 32 lanes; the pairing is fixed, so the result is deterministic.
__device__ __forceinline__ float warp_colsum32(float (&v)[32], int lane) {
#pragma unroll
  for (int half = 16; half >= 1; half >>= 1) {
    const bool upper = (lane & half) != 0;
#pragma unroll
    for (int i = 0; i < half; ++i) {
      const float send = upper ? v[i] : v[i + half];
      const float keep = upper ? v[i + half] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, half);
    }
  }
  return v[0];
}

// same butterfly for the column maxima (returns, in lane l, max over the 32 lanes of column l) ...
__device__ __forceinline__ float warp_colmax32(float (&v)[32], int lane) {
#pragma unroll
  for (int half = 16; half >= 1; half >>= 1) {
    const bool upper = (lane & half) != 0;
#pragma unroll
    for (int i = 0; i < half; ++i) {
      const float send = upper ? v[i] : v[i + half];
      const float keep = upper ? v[i + half] : v[i];
      v[i] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, half));
    }
  }
  return v[0];
}
// ... and for the column minima of integers (lowest row index among the rows that attain the column maximum)
__device__ __forceinline__ int warp_colmin32(int (&v)[32], int lane) {
#pragma unroll
  for (int half = 16; half >= 1; half >>= 1) {
    const bool upper = (lane & half) != 0;
#pragma unroll
    for (int i = 0; i < half; ++i) {
      const int send = upper ? v[i] : v[i + half];
      const int keep = upper ? v[i + half] : v[i];
      v[i] = min(keep, __shfl_xor_sync(0xffffffffu, send, half));
    }
  }
  return v[0];
}

// bias[col .. col+8) as fp32: one 16-byte load when the eight bf16 values are in range and aligned (every lane of
// the warp reads the same address: a broadcast served by L1), element-wise at the ragged end of N
__device__ __forceinline__ void load_bias8(const __nv_bfloat16* __restrict__ bias, int col, int n, float (&b)[8]) {
  if (bias == nullptr) {
#pragma unroll
    for (int i = 0; i < 8; ++i) b[i] = 0.f;
  } else if (col + 8 <= n && ((reinterpret_cast<uintptr_t>(bias + col) & 15) == 0)) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(bias + col));
    const float2 a = unpack_bf16x2(u.x), c = unpack_bf16x2(u.y), d = unpack_bf16x2(u.z), e = unpack_bf16x2(u.w);
    b[0] = a.x; b[1] = a.y; b[2] = c.x; b[3] = c.y; b[4] = d.x; b[5] = d.y; b[6] = e.x; b[7] = e.y;
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) b[i] = (col + i < n) ? __bfloat162float(bias[col + i]) : 0.f;
  }
}

// what the epilogue warps need from GemmParams, read once into registers (the parameter bank costs
// an LDC round trip per use inside the loop otherwise)
struct EpiArgs {
  const __nv_bfloat16* bias;
  int n;
  char* d0;
  char* d1;
  long long ldd0, ldd1, ldaux;
  const __half* aux;
  float alpha, scale;
  float gscale;  // dropout multiplier folded into the GELU pair (1 when dropout is off)
  uint32_t threshold, layer;
  uint2 key;
  int nostore;
  int accumulate;
};

// One 32-column chunk of one accumulator row.  `ncols` < 32 only in the last chunk of a ragged N; columns
// beyond N hold zero accumulators (TMA zero fill) and zero bias, so the math may run on them and only the
// stores are clipped.  Rows in [M, rows_cap) are the zero padding of the packed buffers.
template <int EPI>
__device__ __forceinline__ void epilogue_chunk(const uint32_t (&v)[32], const EpiArgs& e, const uint4 (&rnd)[4],
                                               int row, int col, int ncols, bool row_valid,
                                               bool row_in_buf, bool vec_ok, float& sumsq, float (&res)[32]) {
  using T0 = typename std::conditional<EPI == EPI_STORE_F32, float,
             typename std::conditional<EPI == EPI_FC2, __half, __nv_bfloat16>::type>::type;
  T0* dst0 = reinterpret_cast<T0*>(e.d0) + (long long)row * e.ldd0 + col;
  __half* dst1 = reinterpret_cast<__half*>(e.d1) + (long long)row * e.ldd1 + col;
  const bool want_der = (EPI == EPI_FC1 || EPI == EPI_FC2) && e.d1 != nullptr;
  if constexpr (EPI == EPI_MUL_AUX) {
#pragma unroll
    for (int i = 0; i < 32; ++i) res[i] = 0.f;
  }
  if (!row_valid) {
    if constexpr (EPI == EPI_FC1 || EPI == EPI_FC2 || EPI == EPI_MUL_AUX) {
      if (row_in_buf) {
        const float zero[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int n_ok = min(8, ncols - 8 * j);
          if (n_ok <= 0) break;
          store8(dst0 + 8 * j, zero, n_ok, vec_ok);
          if (want_der) store8(dst1 + 8 * j, zero, n_ok, vec_ok);
        }
      }
    }
    return;
  }
  uint4 g4[4];
  if constexpr (EPI == EPI_MUL_AUX) {
    // all 16-byte loads of the multiplier first (the only long-latency operation of this epilogue)
    const __half* ap = e.aux + (long long)row * e.ldaux + col;
    if (vec_ok && ncols == 32) {
#pragma unroll
      for (int j = 0; j < 4; ++j) g4[j] = __ldg(reinterpret_cast<const uint4*>(ap) + j);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float g[8];
        load8(ap + 8 * j, g, max(0, min(8, ncols - 8 * j)));
        g4[j] = make_uint4(pack_f16x2(g[0], g[1]), pack_f16x2(g[2], g[3]), pack_f16x2(g[4], g[5]), pack_f16x2(g[6], g[7]));
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int n_ok = min(8, ncols - 8 * j);
    if (n_ok <= 0) break;
    float val[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) val[i] = __uint_as_float(v[8 * j + i]);
    if constexpr (EPI == EPI_STORE_BF16 || EPI == EPI_STORE_F32) {
#pragma unroll
      for (int i = 0; i < 8; ++i) val[i] *= e.alpha;
      if (e.accumulate) {  // plain loads: this kernel writes the same addresses
        float old[8];
        load8_plain(dst0 + 8 * j, old, n_ok, vec_ok);
#pragma unroll
        for (int i = 0; i < 8; ++i) val[i] += old[i];
      }
      store8(dst0 + 8 * j, val, n_ok, vec_ok);
    } else if constexpr (EPI == EPI_SIM_DS) {
      // (handled by epilogue_sim_ds: this generic body is never instantiated for it)
    } else if constexpr (EPI == EPI_MUL_AUX) {
      const uint32_t w[4] = {g4[j].x, g4[j].y, g4[j].z, g4[j].w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 g = unpack_f16x2(w[i]);
        val[2 * i] *= e.alpha * g.x;
        val[2 * i + 1] *= e.alpha * g.y;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) res[8 * j + i] = (i < n_ok) ? val[i] : 0.f;
      store8(dst0 + 8 * j, val, n_ok, vec_ok);
    } else {
      float bias[8], der[8];
      load_bias8(e.bias, col + 8 * j, e.n, bias);
#pragma unroll
      for (int i = 0; i < 8; ++i) gelu_erf_both_scaled(val[i] + bias[i], e.gscale, val[i], der[i]);
      if (e.threshold != 0) {  // the multiplier 1/(1-p) is already inside val/der: dropped elements are zeroed
        const uint32_t w[4] = {rnd[j].x, rnd[j].y, rnd[j].z, rnd[j].w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const bool k0 = (w[i] & 0xFFFFu) >= e.threshold, k1 = (w[i] >> 16) >= e.threshold;
          val[2 * i] = k0 ? val[2 * i] : 0.f; der[2 * i] = k0 ? der[2 * i] : 0.f;
          val[2 * i + 1] = k1 ? val[2 * i + 1] : 0.f; der[2 * i + 1] = k1 ? der[2 * i + 1] : 0.f;
        }
      }
      if constexpr (EPI == EPI_FC2) {
#pragma unroll
        for (int i = 0; i < 8; ++i) sumsq = fmaf(val[i], val[i], sumsq);
      }
      if (e.nostore && val[0] != 12345.678f) continue;
      store8(dst0 + 8 * j, val, n_ok, vec_ok);
      if (want_der) store8(dst1 + 8 * j, der, n_ok, vec_ok);
    }
  }
}

// FC epilogues, TMA-store flavour: the chunk's 32 rows x 32 columns of each output go to this warp's staging boxes
// (row = lane, 64-byte rows, 64-byte swizzle: the 16-byte unit u of row r sits at unit u ^ ((r >> 1) & 3), which is
// also what makes the 32 lanes' 16-byte stores bank-conflict free).  d0 box at `stage_addr`, d1 box 2 KB behind it.
// Rows past M are staged as zeros (the zero padding of the packed buffers); TMA clips rows/columns outside the tensor.
template <int EPI>
__device__ __forceinline__ void epilogue_chunk_staged(const uint32_t (&v)[32], const EpiArgs& e, const uint4 (&rnd)[4],
                                                      int col, uint32_t stage_addr, int lane, bool row_valid,
                                                      bool want_der, float& sumsq) {
  const uint32_t row_addr = stage_addr + lane * 64;
  const uint32_t sw = (lane >> 1) & 3;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float val[8], der[8];
    if (row_valid) {
      float bias[8];
      load_bias8(e.bias, col + 8 * j, e.n, bias);
#pragma unroll
      for (int i = 0; i < 8; ++i) gelu_erf_both_scaled(__uint_as_float(v[8 * j + i]) + bias[i], e.gscale, val[i], der[i]);
      if (e.threshold != 0) {  // the multiplier 1/(1-p) is already inside val/der: dropped elements are zeroed
        const uint32_t w[4] = {rnd[j].x, rnd[j].y, rnd[j].z, rnd[j].w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const bool k0 = (w[i] & 0xFFFFu) >= e.threshold, k1 = (w[i] >> 16) >= e.threshold;
          val[2 * i] = k0 ? val[2 * i] : 0.f; der[2 * i] = k0 ? der[2 * i] : 0.f;
          val[2 * i + 1] = k1 ? val[2 * i + 1] : 0.f; der[2 * i + 1] = k1 ? der[2 * i + 1] : 0.f;
        }
      }
      if constexpr (EPI == EPI_FC2) {
#pragma unroll
        for (int i = 0; i < 8; ++i) sumsq = fmaf(val[i], val[i], sumsq);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) { val[i] = 0.f; der[i] = 0.f; }
    }
    const uint32_t unit = ((static_cast<uint32_t>(j) ^ sw) << 4);
    if constexpr (EPI == EPI_FC1)  // h1 feeds the next GEMM: bf16
      sts_b32x4(row_addr + unit, pack_bf16x2(val[0], val[1]), pack_bf16x2(val[2], val[3]), pack_bf16x2(val[4], val[5]),
                pack_bf16x2(val[6], val[7]));
    else                           // a is read only by our own streaming kernels: fp16
      sts_b32x4(row_addr + unit, pack_f16x2(val[0], val[1]), pack_f16x2(val[2], val[3]), pack_f16x2(val[4], val[5]),
                pack_f16x2(val[6], val[7]));
    if (want_der)
      sts_b32x4(row_addr + 2048 + unit, pack_f16x2(der[0], der[1]), pack_f16x2(der[2], der[3]), pack_f16x2(der[4], der[5]),
                pack_f16x2(der[6], der[7]));
  }
}

// ------------------------------------------------------------------------------------------------
// COMM: the launch also services one round of a peer-memory mean all-reduce channel (announce, wait for all ranks,
// reduce this rank's slice, publish).  The work is done by the epilogue warps of every CTA BEFORE their first
// accumulator is ready: in a weight-gradient GEMM (K = all residue rows of the batch) they would otherwise sleep
// on the TMEM barrier for ~100 us, so the collective costs the GEMM neither SMs nor issue slots it needs, and no
// second kernel has to fight the persistent one for a place to run.  Used for the dW1 GEMM of the sharded training
// step: the mean of dW2 / db2 over the ranks crosses NVLink behind it.
template <int CTA_GROUP, bool A_MN, bool B_MN, int EPI, bool COMM = false>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                         const __grid_constant__ CUtensorMap tmap_d0, const __grid_constant__ CUtensorMap tmap_d1,
                         const GemmParams p, const typename std::conditional<COMM, GemmCommReduce, int>::type comm) {
  const int grid_ctas = gridDim.x;
  using S = GemmSmem<CTA_GROUP, EPI>;
  extern __shared__ uint8_t smem_raw[];
  // align inside the shared window with pointer arithmetic (keeps the address space known to the compiler)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + S::STAGES * S::A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::STAGES * S::STAGE_BYTES);
  uint64_t* full_bar = bars;                     // [STAGES]
  uint64_t* empty_bar = bars + S::STAGES;        // [STAGES]
  uint64_t* tmem_full_bar = bars + 2 * S::STAGES;      // [2]
  uint64_t* tmem_empty_bar = bars + 2 * S::STAGES + 2; // [2]
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(bars + 2 * S::STAGES + 4);
  uint8_t* out_smem = smem + S::STAGES * S::STAGE_BYTES + S::BAR_BYTES;  // 1024-byte aligned

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = (CTA_GROUP == 2) ? cluster_ctarank() : 0u;
  const bool is_leader = (cta_rank == 0);

  const int M = p.dyn_m ? min(*p.dyn_m, p.m) : p.m;
  const int K = p.dyn_k ? min(*p.dyn_k, p.k) : p.k;
  const int N = p.n;
  // K == 0 (an empty batch) gives pieces with an empty K range: no MMA is issued and the epilogue runs on a
  // zero accumulator
  const int num_kb = (max(K, 0) + GEMM_BLOCK_K - 1) / GEMM_BLOCK_K;
  const int tile_m = GEMM_BLOCK_M * CTA_GROUP;
  const int num_m_blk = (M + tile_m - 1) / tile_m;
  const int num_n_blk = (N + GEMM_BLOCK_N - 1) / GEMM_BLOCK_N;
  const int num_tiles = num_m_blk * num_n_blk;
  const int worker = blockIdx.x / CTA_GROUP;
  const int num_workers = grid_ctas / CTA_GROUP;
  const int sk_splits = (p.sk_splits < 0) ? choose_sk_splits(num_tiles, num_kb, num_workers) : p.sk_splits;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < S::STAGES; ++s) {
        mbar_init(&full_bar[s], 1);          // the leader's expect_tx arrive; both CTAs' TMA bytes complete_tx on it
        mbar_init(&empty_bar[s], 1);         // one tcgen05.commit
      }
      for (int a = 0; a < 2; ++a) {
        mbar_init(&tmem_full_bar[a], 1);               // one tcgen05.commit
        mbar_init(&tmem_empty_bar[a], GEMM_EPI_WARPS * CTA_GROUP);  // one arrive per epilogue warp of every CTA in the group
      }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc<CTA_GROUP>(tmem_base_slot, GEMM_TMEM_COLS);
  }
  tcgen05_fence_before();
  if constexpr (CTA_GROUP == 2) cluster_sync_all(); else __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  if (warp == 0) {
    // =============================== TMA producer ===============================
    // (whole warp converged, one elected lane issues: see the MMA warp below for why)
    {
      const bool issuer = elect_one();
      int stage = 0;
      uint32_t phase = 0;
      PieceIter pieces(num_tiles, num_kb, worker, num_workers, sk_splits);
      int t, kb0, kb1, tail_r, split;
      while (pieces.next(t, kb0, kb1, tail_r, split)) {
        const int m_blk = t / num_n_blk, n_blk = t % num_n_blk;
        const int m_base = (m_blk * CTA_GROUP + (int)cta_rank) * GEMM_BLOCK_M;
        const int n_base = n_blk * GEMM_BLOCK_N + (int)cta_rank * S::B_ROWS;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (issuer) {
            uint8_t* sa = smem_a + stage * S::A_BYTES;
            uint8_t* sb = smem_b + stage * S::B_BYTES;
            const int k0 = kb * GEMM_BLOCK_K;
            if (is_leader) mbar_arrive_expect_tx(&full_bar[stage], S::STAGE_BYTES * CTA_GROUP);
            auto load = [&](void* dst, const CUtensorMap* tm, int c0, int c1) {
              if constexpr (CTA_GROUP == 2) tma_load_2d_pair(dst, tm, &full_bar[stage], c0, c1);
              else tma_load_2d(dst, tm, &full_bar[stage], c0, c1);
            };
            if constexpr (!A_MN) {
              load(sa, &tmap_a, k0, m_base);  // box {64 k, 128 rows}
            } else {
#pragma unroll
              for (int j = 0; j < GEMM_BLOCK_M / 64; ++j)  // box {64 m, 64 k}: one 128-B-wide slab each
                load(sa + j * (GEMM_BLOCK_K * 128), &tmap_a, m_base + j * 64, k0);
            }
            if constexpr (!B_MN) {
              load(sb, &tmap_b, k0, n_base);  // box {64 k, B_ROWS rows}
            } else {
#pragma unroll
              for (int j = 0; j < S::B_ROWS / 64; ++j)
                load(sb + j * (GEMM_BLOCK_K * 128), &tmap_b, n_base + j * 64, k0);
            }
          }
          __syncwarp();
          if (++stage == S::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer (leader CTA) ===============================
    // The whole warp runs the loop converged and only the tcgen05 instructions are issued by one elected lane:
    // all addresses and descriptors are then warp-uniform values (uniform datapath, no per-MMA vector->uniform
    // register moves).  That matters: with the epilogue warps busy, a ~110-instruction issue sequence per K block
    // could not keep up with the 512 cycles its four MMAs take (ncu: tensor pipe 63 % active while the epilogue
    // warps waited 45 % of their time for the accumulator and the MMA warp never waited for operands).
    if (is_leader) {
      // (fp16 x bf16 mixed operands are rejected by the hardware: 'illegal instruction' on sm_100a, tried in round 1)
      constexpr uint32_t idesc = make_idesc_bf16(GEMM_BLOCK_M * CTA_GROUP, GEMM_BLOCK_N, A_MN, B_MN);
      // shared-memory descriptors: high word constant, low word = (address >> 4) | (LBO >> 4) << 16
      //   K-major SW128: rows are 128 B, 8-row groups 1024 B apart (SBO); k-step = 32 B inside the row (LBO unused: 1).
      //   MN-major SW128: each 64-element slab is [64 k][128 B]; k-step = 16 rows = 2048 B,
      //                   8-k groups 1024 B apart (SBO), slabs BLOCK_K*128 B apart (LBO).
      constexpr uint32_t desc_hi = ((1024u >> 4) & 0x3FFFu) | (1u << 14) | (2u << 29);  // SBO | version 1 | SW128
      constexpr uint32_t a_lbo = A_MN ? (uint32_t)(GEMM_BLOCK_K * 128) : 16u;
      constexpr uint32_t b_lbo = B_MN ? (uint32_t)(GEMM_BLOCK_K * 128) : 16u;
      constexpr uint32_t a_kstep = (A_MN ? GEMM_UMMA_K * 128 : GEMM_UMMA_K * 2) >> 4;
      constexpr uint32_t b_kstep = (B_MN ? GEMM_UMMA_K * 128 : GEMM_UMMA_K * 2) >> 4;
      const uint32_t a_lo0 = ((smem_u32(smem_a) & 0x3FFFFu) >> 4) | (((a_lbo >> 4) & 0x3FFFu) << 16);
      const uint32_t b_lo0 = ((smem_u32(smem_b) & 0x3FFFFu) >> 4) | (((b_lbo >> 4) & 0x3FFFu) << 16);
      const bool issuer = elect_one();
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      PieceIter pieces(num_tiles, num_kb, worker, num_workers, sk_splits);
      int t, kb0, kb1, tail_r, split;
      while (pieces.next(t, kb0, kb1, tail_r, split)) {
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + acc * GEMM_BLOCK_N;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
          const uint32_t a_lo = a_lo0 + stage * (S::A_BYTES >> 4);
          const uint32_t b_lo = b_lo0 + stage * (S::B_BYTES >> 4);
          if (issuer) {
#pragma unroll
            for (int k = 0; k < GEMM_BLOCK_K / GEMM_UMMA_K; ++k) {
              const uint64_t adesc = (static_cast<uint64_t>(desc_hi) << 32) | (a_lo + k * a_kstep);
              const uint64_t bdesc = (static_cast<uint64_t>(desc_hi) << 32) | (b_lo + k * b_kstep);
              umma_bf16<CTA_GROUP>(tmem_d, adesc, bdesc, idesc, ((kb - kb0) | k) != 0 ? 1u : 0u);
            }
            if constexpr (CTA_GROUP == 2) umma_commit_pair(&empty_bar[stage], 0x3);
            else umma_commit_1cta(&empty_bar[stage]);
          }
          __syncwarp();
          if (++stage == S::STAGES) { stage = 0; phase ^= 1; }
        }
        if (issuer) {
          if constexpr (CTA_GROUP == 2) umma_commit_pair(&tmem_full_bar[acc], 0x3);
          else umma_commit_1cta(&tmem_full_bar[acc]);
        }
        __syncwarp();
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // =============================== epilogue warps ===============================
    // 16 warps (four per SM sub-partition): warp w may read TMEM lanes [32*(w%4), +32); the four warps of a
    // lane quarter split the tile's 256 columns into 64-column parts.  Per 32-column chunk: issue the TMEM
    // load, generate the chunk's dropout bits (4 Philox blocks, independent of the data) while it is in
    // flight, then wait and run the math.  The chunk loop is NOT unrolled: the body is ~1000 instructions
    // and must stay inside the instruction cache.
    if constexpr (COMM) comm_reduce_role(comm, (int)blockIdx.x, grid_ctas, (int)threadIdx.x - 64, 32 * GEMM_EPI_WARPS, 1);
    const int quarter = warp & 3;
    const int part = (warp - 2) >> 2;
    constexpr int COLS_PER_WARP = GEMM_BLOCK_N / (GEMM_EPI_WARPS / 4);
    constexpr int CHUNKS = COLS_PER_WARP / 32;
    uint8_t* out_stage = out_smem + (warp - 2) * 2 * S::OUT_BOX_BYTES;  // this warp's two staging boxes (FC epilogues)
    const uint32_t stage_addr = smem_u32(out_stage);
    EpiArgs e;
    e.bias = p.bias; e.n = N;
    e.d0 = reinterpret_cast<char*>(p.d0); e.d1 = reinterpret_cast<char*>(p.d1);
    e.ldd0 = p.ldd0; e.ldd1 = p.ldd1; e.ldaux = p.ldaux; e.aux = p.aux;
    e.alpha = p.alpha; e.scale = p.drop.scale; e.threshold = p.drop.threshold; e.layer = p.drop.layer;
    e.gscale = p.drop.threshold != 0 ? p.drop.scale : 1.f;
    const unsigned long long seed = p.drop.seed + (p.drop.seed_dev ? *p.drop.seed_dev : 0ull);
    e.key = make_uint2(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
    e.nostore = p.debug_nostore == 1;
    e.accumulate = p.accumulate;
    const int rows_cap = p.rows_cap;
    // the interior path moves 16-byte vectors: every row of every output/aux tensor must start 16-byte aligned
    auto rows_aligned = [](const void* ptr, long long ld, int elem_bytes) {
      return ptr == nullptr || (((reinterpret_cast<uintptr_t>(ptr) | (uintptr_t)(ld * elem_bytes)) & 15) == 0);
    };
    const bool vec_ok = rows_aligned(p.d0, p.ldd0, EPI == EPI_STORE_F32 ? 4 : 2) && rows_aligned(p.d1, p.ldd1, 2) &&
                        rows_aligned(p.aux, p.ldaux, 2);
    int* sk_flags = reinterpret_cast<int*>(p.sk_ws);
    float* sk_slots = reinterpret_cast<float*>(reinterpret_cast<char*>(p.sk_ws) + GEMM_SK_FLAG_BYTES);
    const int row_in_cta = quarter * 32 + lane;
    const int S_ = sk_splits;
    int acc = 0;
    uint32_t acc_phase = 0;
    PieceIter pieces(num_tiles, num_kb, worker, num_workers, S_);
    int t, kb0, kb1, tail_r, split;
    while (pieces.next(t, kb0, kb1, tail_r, split)) {
      const int m_blk = t / num_n_blk, n_blk = t % num_n_blk;
      const int row = (m_blk * CTA_GROUP + (int)cta_rank) * GEMM_BLOCK_M + row_in_cta;
      const int col0 = n_blk * GEMM_BLOCK_N + part * COLS_PER_WARP;
      const bool is_head = split == 0;             // owns the tile: runs the epilogue
      const bool has_peers = is_head && tail_r >= 0;  // ... after adding the S-1 dumped K ranges
      const bool empty_k = kb1 <= kb0;
      if (has_peers) {
        if (lane == 0)
          while (ld_acquire_gpu(sk_flags + tail_r) < (S_ - 1) * GEMM_EPI_WARPS * CTA_GROUP) __nanosleep(100);
        __syncwarp();
      }
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tcgen05_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * GEMM_BLOCK_N +
                             part * COLS_PER_WARP;
      const bool row_valid = row < M;
      const bool row_in_buf = row < rows_cap;
      float sumsq = 0.f;
      // EPI_SIM_*: this row's label and running online-softmax state over the warp's 64 columns of the tile
      int sim_lab = -1;
      float sim_m = -INFINITY, sim_s = 0.f, sim_lse = 0.f;
      int sim_a = 0x7fffffff;
      if constexpr (EPI == EPI_SIM_STATS || EPI == EPI_SIM_DS) {
        if (row_valid) {
          sim_lab = __ldg(p.sim_labels + row);
          if constexpr (EPI == EPI_SIM_DS) sim_lse = __ldg(p.sim_row_lse + row);
        }
      }
#pragma unroll 1
      for (int c = 0; c < CHUNKS; ++c) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(taddr + c * 32, v);
        const int col = col0 + c * 32;
        uint4 rnd[4];
        if constexpr (EPI == EPI_FC1 || EPI == EPI_FC2) {
          if (e.threshold != 0 && is_head) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              rnd[j] = philox4x32<kDropoutPhiloxRounds>(make_uint4(static_cast<uint32_t>(row), static_cast<uint32_t>((col >> 3) + j), e.layer, 0u), e.key);
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) rnd[j] = make_uint4(0u, 0u, 0u, 0u);
          }
        }
        tmem_ld_wait();
        if (empty_k) {  // no MMA wrote this accumulator
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = 0u;
        }
        if (!is_head) {
          // raw partial accumulator -> the piece's slot, column-major so that the 32 rows of a warp coalesce
          float* slot = sk_slots + ((size_t)((tail_r * S_ + split) * CTA_GROUP + (int)cta_rank)) * GEMM_SK_SLOT_FLOATS +
                        (size_t)(part * COLS_PER_WARP + c * 32) * GEMM_BLOCK_M + row_in_cta;
#pragma unroll
          for (int i = 0; i < 32; ++i) slot[(size_t)i * GEMM_BLOCK_M] = __uint_as_float(v[i]);
          continue;
        }
        if (has_peers) {
          for (int s2 = 1; s2 < S_; ++s2) {
            const float* slot = sk_slots + ((size_t)((tail_r * S_ + s2) * CTA_GROUP + (int)cta_rank)) * GEMM_SK_SLOT_FLOATS +
                                (size_t)(part * COLS_PER_WARP + c * 32) * GEMM_BLOCK_M + row_in_cta;
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) + __ldcg(slot + (size_t)i * GEMM_BLOCK_M));
          }
        }
        if constexpr (EPI == EPI_SIM_STATS) {
          if (col < N) {  // warp-uniform
            // logits of this thread's row; -inf outside the block so that they drop out of every max / sum
            float sv[32];
#pragma unroll
            for (int i = 0; i < 32; ++i)
              sv[i] = (row_valid && col + i < N) ? __uint_as_float(v[i]) * e.alpha : -INFINITY;
            // row side: online (max, sum exp, first arg-max) over the chunk
            float cm = -INFINITY;
            int ca = 0;
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (sv[i] > cm) { cm = sv[i]; ca = i; }
            if (cm > -INFINITY) {
              const float nm = fmaxf(sim_m, cm);
              float add = 0.f;
#pragma unroll
              for (int i = 0; i < 32; ++i) add += __expf(sv[i] - nm);
              sim_s = sim_s * __expf(sim_m - nm) + add;
              if (cm > sim_m) sim_a = col + ca;  // strictly greater: ties keep the lower column
              sim_m = nm;
            }
            if (sim_lab >= col && sim_lab < col + 32) {
              float pos = 0.f;
#pragma unroll
              for (int i = 0; i < 32; ++i) pos = (sim_lab - col == i) ? sv[i] : pos;
              p.sim_pos[row] = pos;
            }
            // column side over the warp's 32 rows: max -> exp against it -> sum; arg-max = lowest row attaining the max
            if (p.sim_col_part != nullptr) {
              float tmpf[32];
#pragma unroll
              for (int i = 0; i < 32; ++i) tmpf[i] = sv[i];
              const float cmax_l = warp_colmax32(tmpf, lane);  // lane l: column col + l
              int cand[32];
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                const float cs = __shfl_sync(0xffffffffu, cmax_l, i);
                tmpf[i] = (sv[i] > -INFINITY) ? __expf(sv[i] - cs) : 0.f;
                cand[i] = (sv[i] > -INFINITY && sv[i] == cs) ? row : 0x7fffffff;
              }
              const float csum_l = warp_colsum32(tmpf, lane);
              const int carg_l = warp_colmin32(cand, lane);
              const int row0 = row - lane;
              if (col + lane < N && row0 < M)
                p.sim_col_part[(long long)(row0 >> 5) * N + col + lane] = make_float4(cmax_l, csum_l, __int_as_float(carg_l), 0.f);
            }
          }
          continue;
        }
        if constexpr (EPI == EPI_SIM_DS) {
          if (col < N && row_valid) {
            __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(e.d0) + (long long)row * e.ldd0 + col;
            const int ncols = min(32, N - col);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int n_ok = min(8, ncols - 8 * j);
              if (n_ok <= 0) break;
              float d[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const int cc = col + 8 * j + i;
                const float sl = __uint_as_float(v[8 * j + i]) * e.alpha;
                float g = p.sim_wr * __expf(sl - sim_lse);
                if (p.sim_col_lse != nullptr && cc < N) {
                  if (__ldg(p.sim_col_marks + cc)) g += p.sim_wc * __expf(sl - __ldg(p.sim_col_lse + cc));
                }
                if (cc == sim_lab) g -= (p.sim_wr + p.sim_wc);
                d[i] = g;
              }
              store8(dst + 8 * j, d, n_ok, vec_ok);
            }
          }
          continue;
        }
        if (col < N) {
          if constexpr (S::TMA_STORE) {
            if (p.tma_store) {
              const bool want_der = p.d1 != nullptr;
              if (lane == 0) tma_store_wait_read();  // the previous chunk's boxes have been read out
              __syncwarp();
              epilogue_chunk_staged<EPI>(v, e, rnd, col, stage_addr, lane, row_valid, want_der, sumsq);
              fence_proxy_async_smem();
              __syncwarp();
              if (lane == 0 && p.debug_nostore != 3) {   // (3: stage only, no TMA store — timing experiments)
                // (2: every tile stores into the same 256 rows — the stores stay in L2, no DRAM write traffic)
                const int row0 = p.debug_nostore == 2 ? ((row - lane) & 255) : row - lane;
                tma_store_2d(&tmap_d0, out_stage, col, row0);
                if (want_der) tma_store_2d(&tmap_d1, out_stage + S::OUT_BOX_BYTES, col, row0);
                tma_store_commit();
              }
              continue;
            }
          }
          float res[32];
          epilogue_chunk<EPI>(v, e, rnd, row, col, min(32, N - col), row_valid, row_in_buf, vec_ok, sumsq, res);
          if constexpr (EPI == EPI_MUL_AUX) {
            if (p.colsum != nullptr) {  // warp-uniform: every lane takes part in the shuffles, invalid rows hold zeros
              const float cs = warp_colsum32(res, lane);
              const int row0 = row - lane;
              if (row0 < rows_cap && col + lane < N) p.colsum[(long long)(row0 >> 5) * N + col + lane] = cs;
            }
          }
        }
      }
      if constexpr (EPI == EPI_FC2) {
        if (row_in_buf && is_head) p.rowsq[(long long)(n_blk * (GEMM_EPI_WARPS / 4) + part) * p.ld_rowsq + row] = sumsq;
      }
      if constexpr (EPI == EPI_SIM_STATS) {
        if (row_valid && is_head)
          p.sim_row_part[(long long)(n_blk * (GEMM_EPI_WARPS / 4) + part) * p.ld_rowsq + row] =
              make_float4(sim_m, sim_s, __int_as_float(sim_a), 0.f);
      }
      tcgen05_fence_before();
      if (!is_head) __threadfence();  // the dumped piece must be visible before the flag
      __syncwarp();
      if (lane == 0) {
        if constexpr (CTA_GROUP == 2) mbar_arrive_cluster(&tmem_empty_bar[acc], 0);
        else mbar_arrive(&tmem_empty_bar[acc]);
        if (!is_head) atomicAdd(sk_flags + tail_r, 1);
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if constexpr (S::TMA_STORE) {
      if (lane == 0) tma_store_wait_all();  // staged outputs are in global memory before the CTA retires
    }
  }

  tcgen05_fence_before();
  if constexpr (CTA_GROUP == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc<CTA_GROUP>(tmem_base, GEMM_TMEM_COLS);
  }
}

}  // namespace p2t
