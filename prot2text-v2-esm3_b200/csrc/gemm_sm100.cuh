// Persistent, warp-specialised tcgen05 GEMM for sm_100a with fused epilogues.
//
//   D[M,N] = epilogue( sum_k A[m,k] * B[n,k] )      bf16 operands, fp32 accumulation in TMEM
//
// * operands arrive by TMA (128-byte swizzle) into a multi-stage shared-memory ring; either
//   operand may be K-major (memory [rows][K]) or MN-major (memory [K][rows]) — the latter is what
//   the weight-gradient GEMMs need (dW = dZ^T H with the long residue axis as K)
// * one elected thread issues tcgen05.mma (UMMA 128x256x16, or 256x256x16 across a CTA pair with
//   cta_group::2, each CTA loading half of B)
// * two 256-column fp32 accumulators live in TMEM so the epilogue of tile i overlaps the MMAs of
//   tile i+1; 16 epilogue warps (four per SM sub-partition) read TMEM with tcgen05.ld, one row per thread
// * M and/or K may be read from device memory (ragged batches: the residue-row count is produced
//   on the device by the packing kernel, no host sync)
//
// Epilogues (reference sites in models/modeling_esm2llama_instruct.py:60-68 and its autograd):
//   EPI_STORE_BF16 / EPI_STORE_F32 : D = alpha * acc
//   EPI_FC1    : z1 = acc + b1;  D0 = h1 = keep*GELU(z1) (bf16);  D1 = g1 = keep*GELU'(z1) (fp16)   (:62-63)
//   EPI_FC2    : z2 = acc + b2;  D0 = a  = keep*GELU(z2) (fp16);  D1 = g2 = keep*GELU'(z2) (fp16);
//                rowsq[row][4*n_blk + part] = sum over 64 columns of a^2                         (:65-67)
//   (fp16 for tensors only our own streaming kernels read: bf16's 8-bit mantissa adds rounding noise
//    of 2^-9*|a| to every residue, which inflates the pooled std and its 1/std backward)
//   EPI_MUL_AUX: D0 = alpha * acc * aux   (fc2 dgrad chained into GELU'(z1): aux = g1)
// Storing (value, derivative) pairs instead of the pre-activation keeps every later HBM-bound pass
// free of erf/exp and of Philox re-generation: the dropout multiplier is folded into both.
#pragma once
#include <cuda.h>
#include <type_traits>
#include "ptx.cuh"
#include "mathfn.cuh"

namespace p2t {

enum GemmEpilogue : int { EPI_STORE_BF16 = 0, EPI_STORE_F32 = 1, EPI_FC1 = 2, EPI_FC2 = 3, EPI_MUL_AUX = 4 };

constexpr int GEMM_BLOCK_M = 128;  // rows per CTA
constexpr int GEMM_BLOCK_N = 256;  // UMMA N
constexpr int GEMM_BLOCK_K = 64;   // 64 bf16 = one 128-byte swizzle row
constexpr int GEMM_UMMA_K = 16;
constexpr int GEMM_EPI_WARPS = 16;
constexpr int GEMM_THREADS = 64 + 32 * GEMM_EPI_WARPS;  // warp0 TMA, warp1 MMA + TMEM alloc, warps 2.. epilogue
constexpr int GEMM_TMEM_COLS = 512;

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
  float* rowsq;               // EPI_FC2: [rows][ld_rowsq] partial sums of squares, four per 256-column N block
  int ld_rowsq;
  float alpha;
  DropoutParams drop;         // p == 0 -> disabled
};

template <int CTA_GROUP>
struct GemmSmem {
  static constexpr int A_BYTES = GEMM_BLOCK_M * GEMM_BLOCK_K * 2;                 // 16 KB
  static constexpr int B_ROWS = GEMM_BLOCK_N / CTA_GROUP;                         // rows of B this CTA loads
  static constexpr int B_BYTES = B_ROWS * GEMM_BLOCK_K * 2;                       // 32 / 16 KB
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (CTA_GROUP == 1) ? 4 : 6;
  static constexpr int BAR_BYTES = 1024;
  static constexpr int BIAS_BYTES = GEMM_BLOCK_N * 4 * 4;  // per-epilogue-warp bias slices (16 warps x 64 floats)
  static constexpr int TOTAL = STAGES * STAGE_BYTES + BAR_BYTES + BIAS_BYTES + 1024;  // + slack for 1024-B alignment
};

// ------------------------------------------------------------------------------------------------
// Epilogue math for 32 consecutive columns of one accumulator row (one thread), in 8-column groups
// (= one 16-byte store per output, one Philox block of dropout decisions).
// ------------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ void store8(T* dst, const float (&f)[8], int n_ok) {
  if (n_ok == 8 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
    uint4 u;
    if constexpr (sizeof(T) == 2 && std::is_same<T, __half>::value) {
      u = make_uint4(pack_f16x2(f[0], f[1]), pack_f16x2(f[2], f[3]), pack_f16x2(f[4], f[5]), pack_f16x2(f[6], f[7]));
      *reinterpret_cast<uint4*>(dst) = u;
    } else if constexpr (std::is_same<T, __nv_bfloat16>::value) {
      u = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
      *reinterpret_cast<uint4*>(dst) = u;
    } else {
      reinterpret_cast<float4*>(dst)[0] = make_float4(f[0], f[1], f[2], f[3]);
      reinterpret_cast<float4*>(dst)[1] = make_float4(f[4], f[5], f[6], f[7]);
    }
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

template <int EPI>
__device__ __forceinline__ void epilogue_chunk(const uint32_t (&v)[32], const GemmParams& p, const float* bias_s, int row,
                                               int col, int ncols, bool row_valid, bool row_in_buf, float& sumsq) {
  // dropout decisions for the whole 32-column chunk first: four independent Philox blocks in flight
  uint4 rnd[4];
  if constexpr (EPI == EPI_FC1 || EPI == EPI_FC2) {
    if (p.drop.threshold != 0) {
      const uint2 key = make_uint2(static_cast<uint32_t>(p.drop.seed), static_cast<uint32_t>(p.drop.seed >> 32));
#pragma unroll
      for (int j = 0; j < 4; ++j)
        rnd[j] = philox4x32_10(make_uint4(static_cast<uint32_t>(row), static_cast<uint32_t>((col >> 3) + j), p.drop.layer, 0u), key);
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c8 = col + 8 * j;
    const int n_ok = min(8, ncols - 8 * j);
    if (n_ok <= 0) break;
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = __uint_as_float(v[8 * j + i]);
    if constexpr (EPI == EPI_STORE_BF16) {
      if (row_valid) {
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] *= p.alpha;
        store8(reinterpret_cast<__nv_bfloat16*>(p.d0) + (long long)row * p.ldd0 + c8, acc, n_ok);
      }
    } else if constexpr (EPI == EPI_STORE_F32) {
      if (row_valid) {
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] *= p.alpha;
        store8(reinterpret_cast<float*>(p.d0) + (long long)row * p.ldd0 + c8, acc, n_ok);
      }
    } else if constexpr (EPI == EPI_FC1 || EPI == EPI_FC2) {
      if (row_in_buf) {
        float keep[8], val[8], der[8];
        if (p.drop.threshold != 0) dropout_keep8_from(p.drop, rnd[j], keep);
        else {
#pragma unroll
          for (int i = 0; i < 8; ++i) keep[i] = 1.f;
        }
        const float4 bz0 = *reinterpret_cast<const float4*>(bias_s + 8 * j);
        const float4 bz1 = *reinterpret_cast<const float4*>(bias_s + 8 * j + 4);
        const float bias[8] = {bz0.x, bz0.y, bz0.z, bz0.w, bz1.x, bz1.y, bz1.z, bz1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float gl, gd;
          gelu_erf_both_fast(acc[i] + bias[i], gl, gd);
          const float kp = (row_valid && i < n_ok) ? keep[i] : 0.f;
          val[i] = gl * kp;
          der[i] = gd * kp;
          if constexpr (EPI == EPI_FC2) sumsq = fmaf(val[i], val[i], sumsq);
        }
        if constexpr (EPI == EPI_FC1)  // h1 feeds the next GEMM: bf16
          store8(reinterpret_cast<__nv_bfloat16*>(p.d0) + (long long)row * p.ldd0 + c8, val, n_ok);
        else                           // a is read only by our own streaming kernels: fp16
          store8(reinterpret_cast<__half*>(p.d0) + (long long)row * p.ldd0 + c8, val, n_ok);
        if (p.d1 != nullptr) store8(reinterpret_cast<__half*>(p.d1) + (long long)row * p.ldd1 + c8, der, n_ok);
      }
    } else if constexpr (EPI == EPI_MUL_AUX) {
      if (row_in_buf) {
        float g[8];
        if (row_valid) load8(p.aux + (long long)row * p.ldaux + c8, g, n_ok);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = row_valid ? acc[i] * p.alpha * g[i] : 0.f;
        store8(reinterpret_cast<__nv_bfloat16*>(p.d0) + (long long)row * p.ldd0 + c8, acc, n_ok);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
template <int CTA_GROUP, bool A_MN, bool B_MN, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                         const GemmParams p) {
  using S = GemmSmem<CTA_GROUP>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + S::STAGES * S::A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::STAGES * S::STAGE_BYTES);
  uint64_t* full_bar = bars;                     // [STAGES]
  uint64_t* empty_bar = bars + S::STAGES;        // [STAGES]
  uint64_t* tmem_full_bar = bars + 2 * S::STAGES;      // [2]
  uint64_t* tmem_empty_bar = bars + 2 * S::STAGES + 2; // [2]
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(bars + 2 * S::STAGES + 4);
  float* bias_smem = reinterpret_cast<float*>(smem + S::STAGES * S::STAGE_BYTES + S::BAR_BYTES);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = (CTA_GROUP == 2) ? cluster_ctarank() : 0u;
  const bool is_leader = (cta_rank == 0);

  const int M = p.dyn_m ? min(*p.dyn_m, p.m) : p.m;
  const int K = p.dyn_k ? min(*p.dyn_k, p.k) : p.k;
  const int N = p.n;
  const int num_kb = (K + GEMM_BLOCK_K - 1) / GEMM_BLOCK_K;
  const int tile_m = GEMM_BLOCK_M * CTA_GROUP;
  const int num_m_blk = (M + tile_m - 1) / tile_m;
  const int num_n_blk = (N + GEMM_BLOCK_N - 1) / GEMM_BLOCK_N;
  const int num_tiles = num_m_blk * num_n_blk;
  const int worker = blockIdx.x / CTA_GROUP;
  const int num_workers = gridDim.x / CTA_GROUP;

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
    if (lane == 0 && num_kb > 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = worker; t < num_tiles; t += num_workers) {
        const int m_blk = t / num_n_blk, n_blk = t % num_n_blk;
        const int m_base = (m_blk * CTA_GROUP + (int)cta_rank) * GEMM_BLOCK_M;
        const int n_base = n_blk * GEMM_BLOCK_N + (int)cta_rank * S::B_ROWS;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
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
          if (++stage == S::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer (leader CTA) ===============================
    if (is_leader && lane == 0 && num_kb > 0) {
      // (fp16 x bf16 mixed operands are rejected by the hardware: 'illegal instruction' on sm_100a, tried in round 1)
      constexpr uint32_t idesc = make_idesc_bf16(GEMM_BLOCK_M * CTA_GROUP, GEMM_BLOCK_N, A_MN, B_MN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int t = worker; t < num_tiles; t += num_workers) {
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + acc * GEMM_BLOCK_N;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
          const uint32_t sa = smem_u32(smem_a + stage * S::A_BYTES);
          const uint32_t sb = smem_u32(smem_b + stage * S::B_BYTES);
#pragma unroll
          for (int k = 0; k < GEMM_BLOCK_K / GEMM_UMMA_K; ++k) {
            // K-major SW128: rows are 128 B, 8-row groups 1024 B apart (SBO); k-step = 32 B inside the row.
            // MN-major SW128: each 64-element slab is [64 k][128 B]; k-step = 16 rows = 2048 B,
            //                 8-k groups 1024 B apart (SBO), slabs BLOCK_K*128 B apart (LBO).
            const uint64_t adesc = A_MN ? make_smem_desc_sw128(sa + k * (GEMM_UMMA_K * 128), GEMM_BLOCK_K * 128, 1024)
                                        : make_smem_desc_sw128(sa + k * (GEMM_UMMA_K * 2), 16, 1024);
            const uint64_t bdesc = B_MN ? make_smem_desc_sw128(sb + k * (GEMM_UMMA_K * 128), GEMM_BLOCK_K * 128, 1024)
                                        : make_smem_desc_sw128(sb + k * (GEMM_UMMA_K * 2), 16, 1024);
            umma_bf16<CTA_GROUP>(tmem_d, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          if constexpr (CTA_GROUP == 2) umma_commit_pair(&empty_bar[stage], 0x3);
          else umma_commit_1cta(&empty_bar[stage]);
          if (++stage == S::STAGES) { stage = 0; phase ^= 1; }
        }
        if constexpr (CTA_GROUP == 2) umma_commit_pair(&tmem_full_bar[acc], 0x3);
        else umma_commit_1cta(&tmem_full_bar[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // =============================== epilogue warps ===============================
    // 16 warps (four per SM sub-partition, enough thread-level parallelism to hide the MUFU / Philox
    // dependency chains): warp w may read TMEM lanes [32*(w%4), +32); the four warps of a lane quarter
    // split the tile's 256 columns into 64-column parts.  The chunk loop is NOT unrolled: the epilogue
    // body is ~1000 instructions and must stay inside the instruction cache.
    const int quarter = warp & 3;
    const int part = (warp - 2) >> 2;
    constexpr int COLS_PER_WARP = GEMM_BLOCK_N / (GEMM_EPI_WARPS / 4);
    constexpr int CHUNKS = COLS_PER_WARP / 32;
    float* bias_s = bias_smem + (warp - 2) * COLS_PER_WARP;  // this warp's private slice
    int acc = 0;
    uint32_t acc_phase = 0;
    if (num_kb > 0) {
      for (int t = worker; t < num_tiles; t += num_workers) {
        const int m_blk = t / num_n_blk, n_blk = t % num_n_blk;
        const int row = (m_blk * CTA_GROUP + (int)cta_rank) * GEMM_BLOCK_M + quarter * 32 + lane;
        const int col0 = n_blk * GEMM_BLOCK_N + part * COLS_PER_WARP;
        if constexpr (EPI == EPI_FC1 || EPI == EPI_FC2) {
          // stage this warp's bias slice once per tile (read back as shared-memory broadcasts)
          __syncwarp();
#pragma unroll
          for (int i = lane; i < COLS_PER_WARP; i += 32)
            bias_s[i] = (p.bias != nullptr && col0 + i < N) ? __bfloat162float(p.bias[col0 + i]) : 0.f;
          __syncwarp();
        }
        mbar_wait(&tmem_full_bar[acc], acc_phase);
        tcgen05_fence_after();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * GEMM_BLOCK_N +
                               part * COLS_PER_WARP;
        const bool row_valid = row < M;
        const bool row_in_buf = row < p.rows_cap;
        float sumsq = 0.f;
#pragma unroll 1
        for (int c = 0; c < CHUNKS; ++c) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(taddr + c * 32, v);
          tmem_ld_wait();
          const int col = col0 + c * 32;
          if (col < N) epilogue_chunk<EPI>(v, p, bias_s + c * 32, row, col, min(32, N - col), row_valid, row_in_buf, sumsq);
        }
        if constexpr (EPI == EPI_FC2) {
          if (row_in_buf) p.rowsq[(long long)row * p.ld_rowsq + n_blk * (GEMM_EPI_WARPS / 4) + part] = sumsq;
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) {
          if constexpr (CTA_GROUP == 2) mbar_arrive_cluster(&tmem_empty_bar[acc], 0);
          else mbar_arrive(&tmem_empty_bar[acc]);
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
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
