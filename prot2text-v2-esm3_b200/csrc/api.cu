// C ABI (include/p2t_b200.h): argument checking and composition of the kernels.
#include "../../include/p2t_b200.h"

#include "common.h"
#include "gemm_sm100.cuh"
#include "rows.h"

using namespace p2t;

namespace {

inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }

DropoutParams make_dropout(float p, unsigned long long seed, int layer, const unsigned long long* seed_dev = nullptr) {
  DropoutParams d{};
  d.seed = seed;
  d.seed_dev = seed_dev;
  d.layer = (uint32_t)layer;
  if (p > 0.f) {
    long t = lroundf(p * 65536.f);
    if (t < 1) t = 1;
    if (t > 65535) t = 65535;
    d.threshold = (uint32_t)t;
    d.scale = 1.f / (1.f - p);
  } else {
    d.threshold = 0;
    d.scale = 1.f;
  }
  return d;
}

GemmParams base_params(int m, int n, int k) {
  GemmParams p{};
  p.m = m; p.n = n; p.k = k;
  p.alpha = 1.f;
  p.rows_cap = m;
  p.drop = make_dropout(0.f, 0, 0);
  return p;
}

__global__ void dropout_mask_kernel(int rows, int cols, DropoutParams d, float* out) {
  const long long groups = (long long)rows * ((cols + 7) / 8);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < groups; i += (long long)gridDim.x * blockDim.x) {
    const int row = (int)(i / ((cols + 7) / 8)), g = (int)(i % ((cols + 7) / 8));
    float k[8];
    dropout_keep8(d, (uint32_t)row, (uint32_t)g, k);
    for (int j = 0; j < 8; ++j)
      if (g * 8 + j < cols) out[(long long)row * cols + g * 8 + j] = k[j];
  }
}

}  // namespace

#pragma GCC visibility push(default)
extern "C" {

int p2t_abi_version(void) { return P2T_ABI_VERSION; }
const char* p2t_last_error(void) { return last_error(); }
unsigned long long p2t_launch_count(void) { return launches(); }
void p2t_reset_launch_count(void) { reset_launches(); }
void p2t_gemm_timing_enable(int on) { gemm_timing_enable(on != 0); }
int p2t_gemm_timing_collect(double* total_ms, int* launches, double* each_ms, int each_cap) {
  if (!total_ms || !launches) return set_error(-1, "p2t_gemm_timing_collect: null pointer");
  return gemm_timing_collect(total_ms, launches, each_ms, each_cap);
}

void p2t_launch_timing_enable(int on) { launch_timing_enable(on != 0); }
int p2t_launch_timing_mark(void* stream) {
  stamp_launch("mark", S(stream));
  return 0;
}
int p2t_launch_timing_collect(double* ms, int cap, int* n, char* names, int names_cap) {
  if (!n) return set_error(-1, "p2t_launch_timing_collect: null pointer");
  return launch_timing_collect(ms, cap, n, names, names_cap);
}

unsigned long long p2t_gemm_workspace_bytes(void) { return (unsigned long long)gemm_streamk_workspace_bytes(); }

int p2t_gemm_bf16(const void* a, long long lda, int a_mn_major, const void* b, long long ldb, int b_mn_major,
                  void* d, long long ldd, int d_is_f32, int m, int n, int k, float alpha, const int* dyn_m,
                  const int* dyn_k, void* gemm_ws, int cta_group, void* stream) {
  if (!a || !b || !d) return set_error(-1, "p2t_gemm_bf16: null pointer");
  GemmParams p = base_params(m, n, k);
  p.dyn_m = dyn_m; p.dyn_k = dyn_k; p.sk_ws = gemm_ws;
  p.d0 = d; p.ldd0 = ldd; p.alpha = alpha;
  return launch_gemm(a, lda, a_mn_major != 0, b, ldb, b_mn_major != 0, d_is_f32 ? EPI_STORE_F32 : EPI_STORE_BF16, p,
                     cta_group, S(stream));
}

int p2t_rows_plan(const void* mask, int mask_bytes, int B, int L, int chunk_rows, int* counts, int* seq_off,
                  int* chunk_off, int* n_rows_dev, int* row_src, int* chunk_seq, void* stream) {
  if (!mask || !counts || !seq_off || !chunk_off || !n_rows_dev) return set_error(-1, "p2t_rows_plan: null pointer");
  if (mask_bytes != 1 && mask_bytes != 4 && mask_bytes != 8) return set_error(-1, "p2t_rows_plan: mask_bytes must be 1, 4 or 8");
  if (chunk_rows <= 0) return set_error(-1, "p2t_rows_plan: chunk_rows must be positive");
  return rows_plan(mask, mask_bytes, B, L, chunk_rows, counts, seq_off, chunk_off, n_rows_dev, row_src, chunk_seq, S(stream));
}

int p2t_rows_plan_counts(const int* counts, int B, int chunk_rows, int* seq_off, int* chunk_off, int* n_rows_dev,
                         int* chunk_seq, void* stream) {
  if (!counts || !seq_off || !chunk_off || !n_rows_dev) return set_error(-1, "p2t_rows_plan_counts: null pointer");
  if (chunk_rows <= 0) return set_error(-1, "p2t_rows_plan_counts: chunk_rows must be positive");
  return rows_plan_counts(counts, B, chunk_rows, seq_off, chunk_off, n_rows_dev, chunk_seq, S(stream));
}

int p2t_stage_rows_h2d(const void* host_src, long long seq_stride_bytes, long long row_bytes, const int* starts,
                       const int* counts, int B, void* dev_dst, void* const* streams, int n_streams) {
  if (!host_src || !starts || !counts || !dev_dst || !streams) return set_error(-1, "p2t_stage_rows_h2d: null pointer");
  if (row_bytes <= 0 || seq_stride_bytes < 0 || n_streams < 1) return set_error(-1, "p2t_stage_rows_h2d: bad strides / stream count");
  // one plain cudaMemcpyAsync per sequence (the batched-memcpy entry points are not used on purpose), dealt round-robin
  // to the caller's copy streams so that the DMA set-up of one copy overlaps the transfer of its neighbours
  char* dst = reinterpret_cast<char*>(dev_dst);
  const char* src = reinterpret_cast<const char*>(host_src);
  for (int b = 0; b < B; ++b) {
    if (counts[b] < 0 || starts[b] < 0) return set_error(-1, "p2t_stage_rows_h2d: negative start/count");
    const size_t bytes = (size_t)counts[b] * (size_t)row_bytes;
    if (bytes) {
      cudaError_t e = cudaMemcpyAsync(dst, src + (size_t)b * seq_stride_bytes + (size_t)starts[b] * row_bytes, bytes,
                                      cudaMemcpyHostToDevice, S(streams[b % n_streams]));
      if (e != cudaSuccess) return set_error((int)e, "cudaMemcpyAsync: %s", cudaGetErrorString(e));
    }
    dst += bytes;
  }
  return 0;
}

int p2t_stage_rows_pull(const void* host_mapped, const long long* table_dev, const int* piece_prefix_dev, int n_seg,
                        void* dev_dst, int ctas, void* stream) {
  if (!host_mapped || !table_dev || !piece_prefix_dev || !dev_dst) return set_error(-1, "p2t_stage_rows_pull: null pointer");
  if (n_seg < 0 || ctas < 0) return set_error(-1, "p2t_stage_rows_pull: bad segment / CTA count");
  return stage_rows_pull(host_mapped, table_dev, piece_prefix_dev, n_seg, dev_dst, ctas, S(stream));
}

int p2t_gather_rows(const void* src, long long ld_src, const int* row_src, const int* n_rows_dev, int rows_cap,
                    int D, void* out, void* stream) {
  if (!src || !row_src || !n_rows_dev || !out) return set_error(-1, "p2t_gather_rows: null pointer");
  return gather_rows(src, ld_src, row_src, n_rows_dev, rows_cap, D, out, S(stream));
}

int p2t_adapter_fwd(const void* x, int x_rows, const void* w1, const void* b1, const void* w2, const void* b2, int d_in,
                    int d_mid, int d_out, int rows_cap, const int* n_rows_dev, void* h1, void* g1, void* a,
                    void* g2, float* rowsq, float dropout_p, unsigned long long seed, const unsigned long long* seed_dev,
                    void* gemm_ws, int cta_group, void* stream) {
  if (!x || !w1 || !w2 || !h1 || !a || !rowsq) return set_error(-1, "p2t_adapter_fwd: null pointer");
  if (d_in % 8 || d_mid % 8 || d_out % 8) return set_error(-1, "p2t_adapter_fwd: dims must be multiples of 8");
  if (dropout_p < 0.f || dropout_p >= 1.f) return set_error(-1, "p2t_adapter_fwd: dropout_p must be in [0, 1)");
  GemmParams p = base_params(rows_cap, d_mid, d_in);
  p.dyn_m = n_rows_dev;
  p.d0 = h1; p.ldd0 = d_mid; p.d1 = g1; p.ldd1 = d_mid;
  p.bias = reinterpret_cast<const __nv_bfloat16*>(b1);
  p.drop = make_dropout(dropout_p, seed, 1, seed_dev);
  p.a_extent = x_rows;
  p.sk_ws = gemm_ws;  // split-K tail for the incomplete last wave of tiles (S chosen on the device: M is ragged)
  if (int rc = launch_gemm(x, d_in, false, w1, d_in, false, EPI_FC1, p, cta_group, S(stream))) return rc;
  GemmParams q = base_params(rows_cap, d_out, d_mid);
  q.dyn_m = n_rows_dev;
  q.d0 = a; q.ldd0 = d_out; q.d1 = g2; q.ldd1 = d_out;
  q.bias = reinterpret_cast<const __nv_bfloat16*>(b2);
  q.rowsq = rowsq; q.ld_rowsq = rows_cap;
  q.drop = make_dropout(dropout_p, seed, 2, seed_dev);
  q.sk_ws = gemm_ws;
  return launch_gemm(h1, d_mid, false, w2, d_mid, false, EPI_FC2, q, cta_group, S(stream));
}

int p2t_adapter_scale_rows(const void* a, const float* rowsq, int nblk, int rows_cap, int n, int d_out, void* y,
                           float* inv_norm, void* stream) {
  if (!a || !rowsq || !y) return set_error(-1, "p2t_adapter_scale_rows: null pointer");
  if (d_out % 8) return set_error(-1, "p2t_adapter_scale_rows: d_out must be a multiple of 8");
  return scale_rows(a, rowsq, nblk, rows_cap, n, d_out, y, inv_norm, S(stream));
}

int p2t_row_inv_norm(const float* rowsq, int nblk, const int* n_rows_dev, int rows_cap, float* inv_norm, void* stream) {
  if (!rowsq || !n_rows_dev || !inv_norm) return set_error(-1, "p2t_row_inv_norm: null pointer");
  return row_inv_norm(rowsq, nblk, n_rows_dev, rows_cap, inv_norm, S(stream));
}

int p2t_pool_fwd(const void* src, int src_is_f16, long long ld_src, int src_rows, const int* row_src, const float* inv_norm,
                 const int* seq_off, const int* chunk_off, const int* chunk_seq, int B, int D, int chunk_rows,
                 int max_chunks, int mode, void* partial_ws, float* out, long long ld_out, void* p_bf16, float* p_f32,
                 float* norm, void* stream) {
  if (!src || !seq_off || !chunk_off || !chunk_seq || !partial_ws || !out) return set_error(-1, "p2t_pool_fwd: null pointer");
  if (mode < 1 || mode > 3) return set_error(-1, "p2t_pool_fwd: mode must be MEAN, STD or MIX");
  if (max_chunks <= 0) return 0;
  return pool_forward(src, src_is_f16 != 0, ld_src, src_rows, row_src, inv_norm, seq_off, chunk_off, chunk_seq, B, D, chunk_rows,
                      max_chunks, mode, reinterpret_cast<float2*>(partial_ws), out, ld_out, p_bf16, p_f32, norm, S(stream));
}

int p2t_readout_last(const void* x, const int* counts, int B, int S_, int D, float* out, void* stream) {
  if (!x || !counts || !out) return set_error(-1, "p2t_readout_last: null pointer");
  return readout_last(x, counts, B, S_, D, out, S(stream));
}
int p2t_readout_last_bwd(const void* dout, const int* counts, int B, int S_, int D, void* dx, void* stream) {
  if (!dout || !counts || !dx) return set_error(-1, "p2t_readout_last_bwd: null pointer");
  if (B <= 0 || S_ <= 0 || D <= 0) return set_error(-1, "p2t_readout_last_bwd: empty shape");
  return readout_last_bwd(dout, counts, B, S_, D, dx, S(stream));
}

int p2t_l2norm_fwd(const float* e, int B, int E, void* p_bf16, float* p_f32, float* norm, void* stream) {
  if (!e) return set_error(-1, "p2t_l2norm_fwd: null pointer");
  return l2norm_forward(e, B, E, p_bf16, p_f32, norm, S(stream));
}
int p2t_l2norm_bwd(const float* dp, const float* p_f32, const float* norm, int B, int E, float* de, void* stream) {
  if (!dp || !p_f32 || !norm || !de) return set_error(-1, "p2t_l2norm_bwd: null pointer");
  return l2norm_backward(dp, p_f32, norm, B, E, de, S(stream));
}

int p2t_pool_bwd_coef(const float* de, long long ld_de, const float* stats, long long ld_stats, const int* seq_off,
                      int B, int D, int mode, float* c1, float* c2, void* stream) {
  if (!de || !seq_off || !c1 || !c2) return set_error(-1, "p2t_pool_bwd_coef: null pointer");
  if (mode != P2T_READOUT_MEAN && !stats) return set_error(-1, "p2t_pool_bwd_coef: stats required for std/mix");
  return pool_bwd_coef(de, ld_de, stats, ld_stats, seq_off, B, D, mode, c1, c2, S(stream));
}
int p2t_loss_bwd_coef(const float* dS, const float* t_f32, const float* p_f32, const float* pnorm, const float* stats,
                      const int* seq_off, const float* dloss, int R, int B, int C, int D, float tau, float* dp_ws,
                      float* c1, float* c2, void* stream) {
  if (!dS || !t_f32 || !p_f32 || !pnorm || !stats || !seq_off || !dp_ws || !c1 || !c2) return set_error(-1, "p2t_loss_bwd_coef: null pointer");
  if (R > B || tau <= 0.f) return set_error(-1, "p2t_loss_bwd_coef: bad sizes");
  return loss_bwd_coef(dS, t_f32, p_f32, pnorm, stats, seq_off, dloss, R, B, C, D, tau, dp_ws, c1, c2, S(stream));
}

int p2t_readout_bwd(const void* x, const void* mask, int mask_bytes, int B, int S_, int D, const float* c1,
                    const float* c2, void* dx, void* stream) {
  if (!x || !mask || !c1 || !c2 || !dx) return set_error(-1, "p2t_readout_bwd: null pointer");
  if (D % 8) return set_error(-1, "p2t_readout_bwd: D must be a multiple of 8");
  return readout_backward(x, mask, mask_bytes, B, S_, D, c1, c2, dx, S(stream));
}

int p2t_adapter_tail_bwd(const void* a, const void* g2, const float* inv_norm, const int* seq_off, int B, const float* c1,
                         const float* c2, const int* n_rows_dev, int rows_cap, int d_out, void* dz2, float* colsum_ws,
                         int ws_rows, int* nparts_dev, void* db2, void* stream) {
  if (!a || !g2 || !inv_norm || !seq_off || !c1 || !c2 || !n_rows_dev || !dz2)
    return set_error(-1, "p2t_adapter_tail_bwd: null pointer");
  if (B < 1) return set_error(-1, "p2t_adapter_tail_bwd: empty batch");
  return adapter_tail_backward(a, g2, inv_norm, seq_off, B, c1, c2, n_rows_dev, rows_cap, d_out, dz2, colsum_ws, ws_rows,
                               nparts_dev, db2, S(stream));
}
int p2t_adapter_tail_bwd_dy(const void* a, const void* g2, const float* inv_norm, const void* dy, int n,
                            const int* n_rows_dev, int rows_cap, int d_out, void* dz2, void* stream) {
  if (!a || !g2 || !inv_norm || !dy || !dz2) return set_error(-1, "p2t_adapter_tail_bwd_dy: null pointer");
  return adapter_tail_backward_dy(a, g2, inv_norm, dy, n, n_rows_dev, rows_cap, d_out, dz2, S(stream));
}

int p2t_adapter_bwd(const void* x, int x_rows, const void* w1, const void* w2, const void* h1, const void* g1, const void* dz2,
                    int d_in, int d_mid, int d_out, int rows_cap, const int* n_rows_dev, void* dz1, void* dw1,
                    void* db1, void* dw2, void* db2, void* dx, float* colsum_ws, void* gemm_ws, int accumulate,
                    int dw_is_f32, int phases, const p2t_overlap_reduce_t* overlap, int cta_group, void* stream) {
  if (!x || !w2 || !h1 || !g1 || !dz2 || !dz1 || !dw1 || !dw2 || !colsum_ws)
    return set_error(-1, "p2t_adapter_bwd: null pointer");
  if (phases == 0) phases = P2T_BWD_DGRAD | P2T_BWD_DW2 | P2T_BWD_DW1;
  cudaStream_t st = S(stream);
  float* db1_partial = colsum_ws;                                               // [ceil(rows_cap/32)][d_mid]
  float* db2_ws = colsum_ws + (size_t)((rows_cap + 31) / 32) * (size_t)d_mid;   // [ceil(rows_cap/64)][d_out], only with db2
  // dz1 = (dz2 W2) * g1      A = dz2 [rows][d_out] (K-major), B[n][k] = W2[k][n] (MN-major, ld d_mid)
  // + per-32-row column sums of dz1 from the epilogue (db1 without a second pass over dz1)
  if (phases & P2T_BWD_DGRAD) {
    GemmParams p = base_params(rows_cap, d_mid, d_out);
    p.dyn_m = n_rows_dev;
    p.d0 = dz1; p.ldd0 = d_mid;
    p.aux = reinterpret_cast<const __half*>(g1); p.ldaux = d_mid;
    p.colsum = db1_partial;
    if (int rc = launch_gemm(dz2, d_out, false, w2, d_mid, true, EPI_MUL_AUX, p, cta_group, st)) return rc;
  }
  // dW2 = dz2^T h1           both operands MN-major, K = residue rows
  if (phases & P2T_BWD_DW2) {
    GemmParams p = base_params(d_out, d_mid, rows_cap);
    p.dyn_k = n_rows_dev;
    p.sk_ws = gemm_ws;
    p.d0 = dw2; p.ldd0 = d_mid;
    p.accumulate = accumulate;
    if (int rc = launch_gemm(dz2, d_out, true, h1, d_mid, true, dw_is_f32 ? EPI_STORE_F32 : EPI_STORE_BF16, p, cta_group, st)) return rc;
    if (db2) if (int rc = colsum(dz2, n_rows_dev, rows_cap, d_out, db2_ws, db2, nullptr, st)) return rc;
  }
  if (!(phases & P2T_BWD_DW1)) return 0;
  // dW1 = dz1^T x  — optionally with its idle epilogue warps servicing a gradient-mean channel behind it (the mean of
  // dW2 / db2 over the ranks travels over NVLink while this GEMM runs: one launch, no SM taken from the GEMM)
  {
    GemmParams p = base_params(d_mid, d_in, rows_cap);
    p.dyn_k = n_rows_dev;
    p.sk_ws = gemm_ws;
    p.d0 = dw1; p.ldd0 = d_in;
    p.b_extent = x_rows;
    p.accumulate = accumulate;
    GemmCommReduce comm{};
    const GemmCommReduce* cp = nullptr;
    if (overlap != nullptr) {
      if (!dw_is_f32 || cta_group != 2) return set_error(-1, "p2t_adapter_bwd: the overlapped reduce rides on the fp32 dW1 GEMM with cta_group 2");
      if (!overlap->peers || overlap->world < 1 || overlap->world > kPeerMaxWorld || overlap->rank < 0 || overlap->rank >= overlap->world)
        return set_error(-1, "p2t_adapter_bwd: bad overlap channel");
      if (overlap->n_bytes <= 0 || overlap->n_bytes % 16 || overlap->f32_from_byte % 16)
        return set_error(-1, "p2t_adapter_bwd: overlap channel sizes must be multiples of 16 bytes");
      if (overlap->f32_from_byte != 0) return set_error(-1, "p2t_adapter_bwd: the overlapped channel carries fp32 only (f32_from_byte = 0)");
      for (int r = 0; r < overlap->world; ++r) {
        if (!overlap->peers[r]) return set_error(-1, "p2t_adapter_bwd: overlap peer %d not mapped", r);
        comm.peers.base[r] = overlap->peers[r];
      }
      comm.world = overlap->world; comm.rank = overlap->rank;
      comm.n_vec = overlap->n_bytes / 16;
      comm.scale = 1.f / (float)overlap->world;
      cp = &comm;
    }
    if (int rc = launch_gemm(dz1, d_mid, true, x, d_in, true, dw_is_f32 ? EPI_STORE_F32 : EPI_STORE_BF16, p, cta_group, st, cp)) return rc;
  }
  if (db1) {
    BiasJob j{}, none{};
    j.partial = db1_partial; j.D = d_mid; j.nparts_max = (rows_cap + 31) / 32; j.n_rows = n_rows_dev; j.n_static = rows_cap;
    j.block_rows = 32; j.out_bf16 = db1;
    if (int rc = bias_grads_final(j, none, st)) return rc;
  }
  if (dx) {
    if (!w1) return set_error(-1, "p2t_adapter_bwd: w1 required for dx");
    GemmParams p = base_params(rows_cap, d_in, d_mid);
    p.dyn_m = n_rows_dev;
    p.d0 = dx; p.ldd0 = d_in;
    if (int rc = launch_gemm(dz1, d_mid, false, w1, d_in, true, EPI_STORE_BF16, p, cta_group, st)) return rc;
  }
  return 0;
}

int p2t_bias_grads(const float* db1_partial, int rows_cap, const int* n_rows_dev, int d_mid, void* db1_bf16, float* db1_f32,
                   const float* db2_partial, const int* nparts2_dev, int ws_rows2, int d_out, void* db2_bf16,
                   float* db2_f32, int accumulate, void* stream) {
  if (!db1_partial && !db2_partial) return set_error(-1, "p2t_bias_grads: null pointer");
  if (accumulate && ((db1_partial && !db1_f32) || (db2_partial && !db2_f32)))
    return set_error(-1, "p2t_bias_grads: accumulation runs in the fp32 outputs");
  BiasJob j0{}, j1{};
  if (db1_partial) {
    if (!n_rows_dev) return set_error(-1, "p2t_bias_grads: n_rows_dev required");
    j0.partial = db1_partial; j0.D = d_mid; j0.nparts_max = (rows_cap + 31) / 32; j0.n_rows = n_rows_dev; j0.n_static = rows_cap;
    j0.block_rows = 32; j0.out_bf16 = db1_bf16; j0.out_f32 = db1_f32; j0.accumulate = accumulate;
  }
  if (db2_partial) {
    if (!nparts2_dev) return set_error(-1, "p2t_bias_grads: nparts2_dev required");
    j1.partial = db2_partial; j1.D = d_out; j1.nparts_dev = nparts2_dev; j1.nparts_max = ws_rows2;
    j1.out_bf16 = db2_bf16; j1.out_f32 = db2_f32; j1.accumulate = accumulate;
  }
  return bias_grads_final(j0, j1, S(stream));
}

int p2t_loss_fused(const float* p_f32, const float* t_f32, void* const* gather_peers, int world, int rank,
                   long long gather_bytes_per_rank, const int* labels, int R, int B, int C, int E, float tau, float w_row,
                   float w_col, float loss_scale, int all_cols_labelled, int want_col_stats, int need_grad,
                   const float* dloss, const float* pnorm, const float* stats, const int* seq_off, float* S_ws,
                   float* dp_ws, void* barrier_ws, float* loss, float* row_lse, int* argmax_row, int* argmax_col,
                   float* col_max, float* col_sum, float* c1, float* c2, void* stream) {
  if (!p_f32 || !labels || !S_ws || !barrier_ws || !loss) return set_error(-1, "p2t_loss_fused: null pointer");
  if (need_grad && (!pnorm || !stats || !seq_off || !dp_ws || !c1 || !c2))
    return set_error(-1, "p2t_loss_fused: the backward half needs pnorm, stats, seq_off, dp_ws, c1 and c2");
  return loss_fused(p_f32, t_f32, gather_peers, world, rank, gather_bytes_per_rank, labels, R, B, C, E, tau, w_row, w_col,
                    loss_scale, all_cols_labelled, want_col_stats, need_grad, dloss, pnorm, stats, seq_off, S_ws, dp_ws,
                    static_cast<unsigned*>(barrier_ws), loss, row_lse, argmax_row, argmax_col, col_max, col_sum, c1, c2,
                    S(stream));
}
int p2t_loss_fused_eligible(int R, int B, int C, int E) { return loss_fused_eligible(R, B, C, E) ? 1 : 0; }

static bool small_problem(int R, int C, int E) { return (long long)R * C * E <= (1LL << 26); }

int p2t_similarity(const void* p, const void* t, const float* p_f32, const float* t_f32, int R, int C, int E, float tau,
                   float* Sm, int cta_group, void* stream) {
  if (!Sm || (!(p && t) && !(p_f32 && t_f32))) return set_error(-1, "p2t_similarity: null pointer");
  if (tau <= 0.f) return set_error(-1, "p2t_similarity: temperature must be positive");
  if (small_problem(R, C, E) || !(p && t)) {
    if (p_f32 && t_f32) return sim_small(p_f32, t_f32, true, R, C, E, 1.f / tau, Sm, S(stream));
    return sim_small(p, t, false, R, C, E, 1.f / tau, Sm, S(stream));
  }
  GemmParams q = base_params(R, C, E);
  q.d0 = Sm; q.ldd0 = C; q.alpha = 1.f / tau;
  return launch_gemm(p, E, false, t, E, false, EPI_STORE_F32, q, cta_group, S(stream));
}

int p2t_infonce_col_stats(const float* Sm, int R, int C, float* col_max, float* col_sum, int* col_argmax,
                          int row_index_base, void* stream) {
  if (!Sm || !col_max || !col_sum) return set_error(-1, "p2t_infonce_col_stats: null pointer");
  return col_stats(Sm, R, C, col_max, col_sum, col_argmax, row_index_base, S(stream));
}

int p2t_infonce_ce(float* Sm, const int* labels, int R, int C, float w_row, float w_col, float inv_rn,
                   const float* col_max, const float* col_sum, unsigned char* col_labelled_ws, int all_cols_labelled,
                   float* row_loss,
                   float* row_lse, int* argmax_row, void* dS_bf16, int write_ds, void* stream) {
  if (!Sm || !labels || !row_loss) return set_error(-1, "p2t_infonce_ce: null pointer");
  if (w_col != 0.f) {
    if (!col_max || !col_sum || !col_labelled_ws) return set_error(-1, "p2t_infonce_ce: column statistics required");
    if (all_cols_labelled) {
      cudaError_t e = cudaMemsetAsync(col_labelled_ws, 1, C, S(stream));
      if (e != cudaSuccess) return set_error((int)e, "memset: %s", cudaGetErrorString(e));
    } else if (int rc = mark_labelled(labels, R, C, col_labelled_ws, S(stream))) {
      return rc;
    }
  }
  return ce_rows(Sm, labels, R, C, w_row, w_col, inv_rn, col_max, col_sum, col_labelled_ws, row_loss, row_lse,
                 argmax_row, dS_bf16, write_ds, S(stream));
}

int p2t_infonce_stats(const void* p, const void* t, const int* labels, int R, int C, int E, float tau, void* row_part_ws,
                      void* col_part_ws, float* pos, float* col_max, float* col_sum, int* col_argmax, int cta_group,
                      void* stream) {
  if (!p || !t || !labels || !row_part_ws || !pos) return set_error(-1, "p2t_infonce_stats: null pointer");
  if (tau <= 0.f) return set_error(-1, "p2t_infonce_stats: temperature must be positive");
  if (col_part_ws && (!col_max || !col_sum)) return set_error(-1, "p2t_infonce_stats: column outputs required with the column workspace");
  GemmParams q = base_params(R, C, E);
  q.alpha = 1.f / tau;
  q.sim_labels = labels;
  q.sim_row_part = static_cast<float4*>(row_part_ws);
  q.ld_rowsq = R;
  q.sim_col_part = static_cast<float4*>(col_part_ws);
  q.sim_pos = pos;
  if (int rc = launch_gemm(p, E, false, t, E, false, EPI_SIM_STATS, q, cta_group, S(stream))) return rc;
  if (col_part_ws)
    return sim_cols_combine(static_cast<const float4*>(col_part_ws), (R + 31) / 32, C, col_max, col_sum, col_argmax, S(stream));
  return 0;
}

int p2t_infonce_finish(const void* p, const void* t, const int* labels, int R, int C, int E, float tau, float w_row,
                       float w_col, float inv_rn, const void* row_part_ws, const float* pos, const float* col_max,
                       const float* col_sum, int all_cols_labelled, unsigned char* col_labelled_ws, float* col_lse_ws,
                       float* row_loss, float* row_lse, int* argmax_row, void* dS_bf16, int cta_group, void* stream) {
  if (!p || !t || !labels || !row_part_ws || !pos || !row_loss || !row_lse) return set_error(-1, "p2t_infonce_finish: null pointer");
  if (w_col != 0.f && (!col_max || !col_sum || !col_labelled_ws || !col_lse_ws))
    return set_error(-1, "p2t_infonce_finish: column statistics and workspaces required for the column term");
  cudaStream_t st = S(stream);
  const int npart = 4 * ((C + GEMM_BLOCK_N - 1) / GEMM_BLOCK_N);
  if (int rc = sim_finish(static_cast<const float4*>(row_part_ws), npart, R, pos, labels, R, C, w_row, w_col, col_max, col_sum,
                          w_col != 0.f ? col_lse_ws : nullptr, row_loss, row_lse, argmax_row, st)) return rc;
  if (!dS_bf16) return 0;
  if (C % 8) return set_error(-1, "p2t_infonce_finish: C must be a multiple of 8 for the bf16 dLogits");
  if (w_col != 0.f) {
    if (all_cols_labelled) {
      cudaError_t e = cudaMemsetAsync(col_labelled_ws, 1, C, st);
      if (e != cudaSuccess) return set_error((int)e, "memset: %s", cudaGetErrorString(e));
    } else if (int rc = mark_labelled(labels, R, C, col_labelled_ws, st)) {
      return rc;
    }
  }
  GemmParams q = base_params(R, C, E);
  q.alpha = 1.f / tau;
  q.d0 = dS_bf16; q.ldd0 = C;
  q.sim_labels = labels;
  q.sim_row_lse = row_lse;
  q.sim_col_lse = w_col != 0.f ? col_lse_ws : nullptr;
  q.sim_col_marks = col_labelled_ws;
  q.sim_wr = w_row * inv_rn; q.sim_wc = w_col * inv_rn;
  return launch_gemm(p, E, false, t, E, false, EPI_SIM_DS, q, cta_group, st);
}

int p2t_infonce_grad(const float* dS, const void* dS_bf16, const void* p, const void* t, const float* p_f32,
                     const float* t_f32, int R, int C, int E, float tau, float* dp, float* dt, int cta_group,
                     void* stream) {
  if (!dS && !dS_bf16) return set_error(-1, "p2t_infonce_grad: null pointer");
  cudaStream_t st = S(stream);
  const bool gemm_ok = !small_problem(R, C, E) && (C % 8 == 0) && dS_bf16 && (!dp || t) && (!dt || p);
  if (!gemm_ok) {
    if (!dS) return set_error(-1, "p2t_infonce_grad: the CUDA-core path needs the fp32 dLogits");
    if (dp) {
      if (!t && !t_f32) return set_error(-1, "p2t_infonce_grad: t required for dp");
      if (int rc = contract_small(dS, C, false, t_f32 ? (const void*)t_f32 : t, t_f32 != nullptr, R, C, E, 1.f / tau, dp, st)) return rc;
    }
    if (dt) {
      if (!p && !p_f32) return set_error(-1, "p2t_infonce_grad: p required for dt");
      if (int rc = contract_small(dS, C, true, p_f32 ? (const void*)p_f32 : p, p_f32 != nullptr, C, R, E, 1.f / tau, dt, st)) return rc;
    }
    return 0;
  }
  if (dp) {  // dp[i][e] = sum_j dS[i][j] t[j][e]: A = dS (K-major), B[n=e][k=j] = t[j][e] (MN-major)
    GemmParams q = base_params(R, E, C);
    q.d0 = dp; q.ldd0 = E; q.alpha = 1.f / tau;
    if (int rc = launch_gemm(dS_bf16, C, false, t, E, true, EPI_STORE_F32, q, cta_group, st)) return rc;
  }
  if (dt) {  // dt[j][e] = sum_i dS[i][j] p[i][e]: A[m=j][k=i] (MN-major), B[n=e][k=i] (MN-major)
    GemmParams q = base_params(C, E, R);
    q.d0 = dt; q.ldd0 = E; q.alpha = 1.f / tau;
    if (int rc = launch_gemm(dS_bf16, C, true, p, E, true, EPI_STORE_F32, q, cta_group, st)) return rc;
  }
  return 0;
}

int p2t_loss_mean(const float* row_loss, int R, float scale, float* loss, int accumulate, void* stream) {
  if (!row_loss || !loss) return set_error(-1, "p2t_loss_mean: null pointer");
  return loss_mean(row_loss, R, scale, loss, accumulate, S(stream));
}

int p2t_f32_to_bf16(const float* in, long long n, void* out, void* stream) {
  if (!in || !out) return set_error(-1, "p2t_f32_to_bf16: null pointer");
  return convert_f32_to_bf16(in, n, out, S(stream));
}
int p2t_bf16_to_f32(const void* in, long long n, float* out, void* stream) {
  if (!in || !out) return set_error(-1, "p2t_bf16_to_f32: null pointer");
  return convert_bf16_to_f32(in, n, out, S(stream));
}
int p2t_colsum(const void* x, const int* n_rows_dev, int n_static, int D, float* ws, void* out_bf16, float* out_f32,
               void* stream) {
  if (!x || !ws) return set_error(-1, "p2t_colsum: null pointer");
  return colsum(x, n_rows_dev, n_static, D, ws, out_bf16, out_f32, S(stream));
}
int p2t_dropout_mask(int rows, int cols, float dropout_p, unsigned long long seed, int layer, float* out,
                     void* stream) {
  if (!out) return set_error(-1, "p2t_dropout_mask: null pointer");
  DropoutParams d = make_dropout(dropout_p, seed, layer);
  dropout_mask_kernel<<<256, 256, 0, S(stream)>>>(rows, cols, d, out);
  return check_launch("dropout_mask_kernel", S(stream));
}

/* ---- Stage-2 hand-off: adapter rows straight into the LLM's inputs_embeds slots ---- */
int p2t_adapter_scatter_rows(const void* a, const float* rowsq, int nblk, int rows_cap, int n, int d_out, void* dst,
                             long long ld_dst, const int* row_dst, const int* n_rows_dev, const int* n_dst_dev,
                             float* inv_norm, void* stream) {
  if (!a || !rowsq || !dst || !row_dst) return set_error(-1, "p2t_adapter_scatter_rows: null pointer");
  if (n > rows_cap) return set_error(-1, "p2t_adapter_scatter_rows: n > rows_cap");
  return scatter_scaled_rows(a, rowsq, nblk, rows_cap, n, d_out, dst, ld_dst, row_dst, n_rows_dev, n_dst_dev, inv_norm, S(stream));
}

/* ---- exchange steps over NVLink peer memory ---- */
unsigned long long p2t_peer_ctrl_bytes(void) { return (unsigned long long)peer_ctrl_bytes(); }
int p2t_peer_alloc(unsigned long long bytes, void** dptr, unsigned char* handle64) { return peer_alloc((size_t)bytes, dptr, handle64); }
int p2t_peer_open(const unsigned char* handle64, void** dptr) { return peer_open(handle64, dptr); }
int p2t_peer_close(void* dptr) { return dptr ? peer_close(dptr) : set_error(-1, "p2t_peer_close: null pointer"); }
int p2t_peer_free(void* dptr) { return dptr ? peer_free(dptr) : set_error(-1, "p2t_peer_free: null pointer"); }
int p2t_peer_allgather(void* const* peers, int world, int rank, const void* src, long long bytes_per_rank, void* dst,
                       int phases, void* stream) {
  return peer_allgather(peers, world, rank, src, bytes_per_rank, dst, phases, S(stream));
}
int p2t_peer_allreduce_mean(void* const* peers, int world, int rank, long long n_bytes, long long f32_from_byte, void* dst,
                            int phases, void* stream) {
  return peer_allreduce_mean(peers, world, rank, n_bytes, f32_from_byte, dst, phases, S(stream));
}
int p2t_peer_reset(void* channel_base, void* stream) { return peer_reset(channel_base, S(stream)); }

int p2t_copy_d2d(void* dst, const void* src, unsigned long long bytes, void* stream) {
  if (!dst || !src) return set_error(-1, "p2t_copy_d2d: null pointer");
  cudaError_t e = cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDeviceToDevice, S(stream));
  if (e != cudaSuccess) return set_error((int)e, "p2t_copy_d2d: %s", cudaGetErrorString(e));
  return 0;
}
int p2t_peer_status(const void* channel_base, unsigned int* status_host) {
  if (!channel_base || !status_host) return set_error(-1, "p2t_peer_status: null pointer");
  cudaError_t e = cudaMemcpy(status_host, static_cast<const unsigned*>(channel_base) + 4, sizeof(unsigned), cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) return set_error((int)e, "p2t_peer_status: %s", cudaGetErrorString(e));
  return 0;
}

/* ---- clip_grad_norm_ + AdamW ---- */
int p2t_adamw_workspace_floats(int count, const long long* numel) {
  if (!numel || count < 1) return 0;
  int blocks = 0;
  for (int i = 0; i < count; ++i) blocks += adamw_blocks(numel[i]);
  return blocks;
}
int p2t_adamw_step(int count, void* const* params, void* const* grads, const float* const* grads_f32, float* const* exp_avg,
                   float* const* exp_avg_sq, float* const* master, const long long* numel, float* partial_ws, float* scal, const float* lr_dev,
                   long long* step_dev, float beta1, float beta2, float eps, float weight_decay, float max_norm,
                   int zero_grad, void* stream) {
  if (!params || !grads || !exp_avg || !exp_avg_sq || !numel || !partial_ws || !scal || !lr_dev || !step_dev)
    return set_error(-1, "p2t_adamw_step: null pointer");
  if (count < 1 || count > kAdamMaxTensors) return set_error(-1, "p2t_adamw_step: 1..%d tensors per call", kAdamMaxTensors);
  AdamTable t{};
  t.count = count;
  for (int i = 0; i < count; ++i) {
    t.param[i] = params[i]; t.grad[i] = grads[i]; t.exp_avg[i] = exp_avg[i]; t.exp_avg_sq[i] = exp_avg_sq[i];
    t.master[i] = master ? master[i] : nullptr;
    t.grad_f32[i] = grads_f32 ? grads_f32[i] : nullptr;
    t.numel[i] = numel[i];
  }
  return adamw_step(t, partial_ws, scal, lr_dev, step_dev, beta1, beta2, eps, weight_decay, max_norm, zero_grad, S(stream));
}

}  // extern "C"
#pragma GCC visibility pop
