// Host side of the tcgen05 GEMM: TMA tensor-map construction (driver entry point resolved at run
// time, so the library links without libcuda) and template dispatch.
#include "common.h"
#include "gemm_sm100.cuh"
#include <cstdlib>
#include <cstring>

namespace p2t {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !sym) return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  return fn;
}

// 16-bit row-major matrix [outer][inner] with leading dimension `ld` elements; box = {box_inner, box_outer}
// (operands: 64-element = 128-byte inner box, 128-byte swizzle; staged outputs: 32-element inner box, 64-byte swizzle)
int make_tmap_16bit(CUtensorMap* map, const void* ptr, long long inner, long long outer, long long ld,
                    int box_inner, int box_outer, CUtensorMapSwizzle swizzle) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return set_error(-10, "cuTensorMapEncodeTiled entry point not available");
  // the encoder is a driver-API call and needs the primary context current on THIS host thread; a thread whose
  // first CUDA call is this one (e.g. autograd's backward worker) has none yet: one runtime call binds it
  static thread_local bool ctx_bound = false;
  if (!ctx_bound) {
    cudaFree(nullptr);
    ctx_bound = true;
  }
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (ld * 2) % 16)
    return set_error(-11, "GEMM operand must be 16-byte aligned with a leading dimension multiple of 8");
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(inner), static_cast<cuuint64_t>(outer)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(box_inner), static_cast<cuuint32_t>(box_outer)};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(-12, "cuTensorMapEncodeTiled failed (code %d)", (int)r);
  return 0;
}
static int make_tmap_bf16(CUtensorMap* map, const void* ptr, long long inner, long long outer, long long ld,
                          int box_outer) {
  return make_tmap_16bit(map, ptr, inner, outer, ld, 64, box_outer, CU_TENSOR_MAP_SWIZZLE_128B);
}

size_t gemm_streamk_workspace_bytes() {
  return (size_t)GEMM_SK_FLAG_BYTES + (size_t)GEMM_SK_SLOTS_PER_SM * sm_count() * GEMM_SK_SLOT_FLOATS * sizeof(float);
}

int sm_count() {
  // per device: a process may drive more than one GPU (single-process multi-GPU tests, a changed current device)
  static int n[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (!n[dev]) cudaDeviceGetAttribute(&n[dev], cudaDevAttrMultiProcessorCount, dev);
  return n[dev];
}

template <int CTA_GROUP, bool A_MN, bool B_MN, int EPI, bool COMM = false>
static int launch_inst(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, cudaStream_t stream,
                       const GemmCommReduce* comm = nullptr) {
  using S = GemmSmem<CTA_GROUP, EPI>;
  auto kern = gemm_bf16_tcgen05_kernel<CTA_GROUP, A_MN, B_MN, EPI, COMM>;
  static bool configured[64] = {false};  // cudaFuncSetAttribute is a per-device setting
  int cur_dev = 0;
  cudaGetDevice(&cur_dev);
  if (cur_dev < 0 || cur_dev >= 64) cur_dev = 0;
  if (!configured[cur_dev]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL);
    if (e != cudaSuccess) return set_error((int)e, "cudaFuncSetAttribute(smem=%d): %s", S::TOTAL, cudaGetErrorString(e));
    configured[cur_dev] = true;
  }
  const int tile_m = GEMM_BLOCK_M * CTA_GROUP;
  const long long tiles = (long long)((p.m + tile_m - 1) / tile_m) * ((p.n + GEMM_BLOCK_N - 1) / GEMM_BLOCK_N);
  if (tiles == 0) return 0;
  const int all_workers = sm_count() / CTA_GROUP;
  int workers = all_workers;
  if (tiles < workers) workers = (int)tiles;
  GemmParams q = p;
  q.sk_splits = 1;
  static const int nostore = getenv("P2T_DEBUG_NOSTORE") ? atoi(getenv("P2T_DEBUG_NOSTORE")) : 0;
  q.debug_nostore = nostore;
  const int tail = (int)(tiles % all_workers);  // tiles of the incomplete last wave
  const int kb = (p.k + GEMM_BLOCK_K - 1) / GEMM_BLOCK_K;
  bool reset_flags = false;
  if (p.sk_ws != nullptr && p.dyn_m != nullptr && p.dyn_k == nullptr) {
    // M is a device scalar (ragged row count): the kernel derives the tile count and picks S itself (same cost model)
    q.sk_splits = -1;
    workers = all_workers;
    reset_flags = true;
  } else if (p.sk_ws != nullptr && p.dyn_m == nullptr && tail != 0) {
    // split-K tail with the tile count known on the host: cut the last wave's tiles into S K ranges so that tail*S
    // pieces fill the machine; cost model = rounds of pieces / S + a small per-dump charge
    const int best_s = choose_sk_splits((int)tiles, kb, all_workers);
    if (best_s > 1) {
      q.sk_splits = best_s;
      workers = all_workers;
      reset_flags = true;
    }
  }
  if (reset_flags) {
    cudaError_t me = cudaMemsetAsync(p.sk_ws, 0, GEMM_SK_FLAG_BYTES, stream);
    if (me != cudaSuccess) return set_error((int)me, "split-K flag reset failed: %s", cudaGetErrorString(me));
  }
  cudaLaunchConfig_t cfg{};
  if (COMM) workers = all_workers;  // every CTA's epilogue warps take part in the channel round, tiles or not
  cfg.gridDim = dim3(workers * CTA_GROUP);
  cfg.blockDim = dim3(GEMM_THREADS);
  cfg.dynamicSmemBytes = S::TOTAL;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CTA_GROUP;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  const bool timed = gemm_timing_enabled();
  if (timed) gemm_timing_record(stream, true);
  // FC epilogues write their 16-bit outputs with TMA stores when the output rows are 16-byte aligned
  CUtensorMap td0, td1;
  memset(&td0, 0, sizeof(td0));
  memset(&td1, 0, sizeof(td1));
  q.tma_store = 0;
  static const int tma_store_on = getenv("P2T_TMA_STORE") ? atoi(getenv("P2T_TMA_STORE")) : 1;  // 0: direct global stores (experiments)
  if (S::TMA_STORE && nostore != 1 && tma_store_on) {
    const bool ok0 = !((reinterpret_cast<uintptr_t>(p.d0) & 15) || (p.ldd0 * 2) % 16);
    const bool ok1 = p.d1 == nullptr || !((reinterpret_cast<uintptr_t>(p.d1) & 15) || (p.ldd1 * 2) % 16);
    if (ok0 && ok1) {
      if (int rc = make_tmap_16bit(&td0, p.d0, p.n, p.rows_cap, p.ldd0, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B)) return rc;
      if (p.d1)
        if (int rc = make_tmap_16bit(&td1, p.d1, p.n, p.rows_cap, p.ldd1, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B)) return rc;
      q.tma_store = 1;
    }
  }
  stamp_begin(stream);
  cudaError_t e;
  if constexpr (COMM) e = cudaLaunchKernelEx(&cfg, kern, ta, tb, td0, td1, q, *comm);
  else e = cudaLaunchKernelEx(&cfg, kern, ta, tb, td0, td1, q, 0);
  if (timed) gemm_timing_record(stream, false);
  if (e != cudaSuccess) return set_error((int)e, "GEMM launch failed: %s", cudaGetErrorString(e));
  count_launch();
  stamp_launch(EPI == EPI_FC1 ? "gemm_fc1" : EPI == EPI_FC2 ? "gemm_fc2" : EPI == EPI_MUL_AUX ? "gemm_dgrad" : EPI == EPI_SIM_STATS ? "gemm_sim_stats" : EPI == EPI_SIM_DS ? "gemm_sim_dlogits"
               : (A_MN && B_MN) ? "gemm_wgrad" : "gemm_plain", stream);
  return 0;
}

template <int CTA_GROUP, bool A_MN, bool B_MN>
static int launch_epi(int epi, const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, cudaStream_t s) {
  switch (epi) {
    case EPI_STORE_BF16: return launch_inst<CTA_GROUP, A_MN, B_MN, EPI_STORE_BF16>(ta, tb, p, s);
    case EPI_STORE_F32: return launch_inst<CTA_GROUP, A_MN, B_MN, EPI_STORE_F32>(ta, tb, p, s);
    default: break;
  }
  if constexpr (!A_MN && !B_MN) {
    if (epi == EPI_SIM_STATS) return launch_inst<CTA_GROUP, false, false, EPI_SIM_STATS>(ta, tb, p, s);
    if (epi == EPI_SIM_DS) return launch_inst<CTA_GROUP, false, false, EPI_SIM_DS>(ta, tb, p, s);
  }
  if constexpr (!A_MN && !B_MN) {
    if (epi == EPI_FC1) return launch_inst<CTA_GROUP, false, false, EPI_FC1>(ta, tb, p, s);
    if (epi == EPI_FC2) return launch_inst<CTA_GROUP, false, false, EPI_FC2>(ta, tb, p, s);
  }
  if constexpr (!A_MN) {
    if (epi == EPI_MUL_AUX) return launch_inst<CTA_GROUP, false, B_MN, EPI_MUL_AUX>(ta, tb, p, s);
  }
  return set_error(-13, "unsupported GEMM epilogue/layout combination (epi=%d)", epi);
}

// A: logical [M][K]; memory [M][lda] (K-major) or [K][lda] (MN-major).  Same for B with N.
// `a_rows_cap`/`b_rows_cap`: extent of the operand along its row (M/N resp. K for MN-major) axis that is
// physically present, used for TMA bounds (zero fill beyond).
int launch_gemm(const void* a, long long lda, bool a_mn, const void* b, long long ldb, bool b_mn, int epi,
                GemmParams p, int cta_group, cudaStream_t stream, const GemmCommReduce* comm) {
  if (p.m <= 0 || p.n <= 0 || p.k <= 0) return 0;
  if (comm != nullptr && !(a_mn && b_mn && epi == EPI_STORE_F32 && cta_group == 2))
    return set_error(-16, "the comm role rides on the fp32 weight-gradient GEMM (both operands MN-major, cta_group 2) only");
  CUtensorMap ta, tb;
  int rc;
  if (!a_mn) rc = make_tmap_bf16(&ta, a, p.k, p.a_extent > 0 ? p.a_extent : p.m, lda, GEMM_BLOCK_M);
  else rc = make_tmap_bf16(&ta, a, p.m, p.a_extent > 0 ? p.a_extent : p.k, lda, GEMM_BLOCK_K);
  if (rc) return rc;
  const int b_rows = GEMM_BLOCK_N / cta_group;
  if (!b_mn) rc = make_tmap_bf16(&tb, b, p.k, p.b_extent > 0 ? p.b_extent : p.n, ldb, b_rows);
  else rc = make_tmap_bf16(&tb, b, p.n, p.b_extent > 0 ? p.b_extent : p.k, ldb, GEMM_BLOCK_K);
  if (rc) return rc;
  if (p.rows_cap <= 0) p.rows_cap = p.m;
  if (comm != nullptr) return launch_inst<2, true, true, EPI_STORE_F32, true>(ta, tb, p, stream, comm);
  if (cta_group == 2) {
    if (a_mn && b_mn) return launch_epi<2, true, true>(epi, ta, tb, p, stream);
    if (a_mn) return launch_epi<2, true, false>(epi, ta, tb, p, stream);
    if (b_mn) return launch_epi<2, false, true>(epi, ta, tb, p, stream);
    return launch_epi<2, false, false>(epi, ta, tb, p, stream);
  } else if (cta_group == 1) {
    if (a_mn && b_mn) return launch_epi<1, true, true>(epi, ta, tb, p, stream);
    if (a_mn) return launch_epi<1, true, false>(epi, ta, tb, p, stream);
    if (b_mn) return launch_epi<1, false, true>(epi, ta, tb, p, stream);
    return launch_epi<1, false, false>(epi, ta, tb, p, stream);
  }
  return set_error(-14, "cta_group must be 1 or 2");
}

}  // namespace p2t
