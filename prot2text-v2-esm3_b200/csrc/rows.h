// Internal declarations of the host launchers (rows.cu, infonce.cu, gemm_host.cu).
#pragma once
#include <cuda_runtime.h>
#include "mathfn.cuh"

namespace p2t {

struct GemmParams;

int rows_plan(const void* mask, int mask_bytes, int B, int L, int rc, int* counts, int* seq_off, int* chunk_off,
              int* n_rows, int* row_src, int* chunk_seq, cudaStream_t st);
int rows_plan_counts(const int* counts, int B, int rc, int* seq_off, int* chunk_off, int* n_rows, int* chunk_seq,
                     cudaStream_t st);
int row_inv_norm(const float* rowsq, int nblk, const int* n_rows, int cap, float* inv_norm, cudaStream_t st);
int gather_rows(const void* src, long long ld_src, const int* row_src, const int* n_rows, int cap, int D, void* out,
                cudaStream_t st);
int pool_forward(const void* src, bool src_is_f16, long long ld_src, int src_rows, const int* row_src, const float* inv_norm,
                 const int* seq_off, const int* chunk_off, const int* chunk_seq, int B, int D, int rc, int max_chunks,
                 int mode, float2* partial, float* out, long long ld_out, void* norm_p_bf16, float* norm_p_f32,
                 float* norm_out, cudaStream_t st);
int loss_bwd_coef(const float* dS, const float* t, const float* p, const float* pnorm, const float* stats,
                  const int* seq_off, const float* dloss, int R, int B, int C, int D, float tau, float* dp_ws,
                  float* c1, float* c2, cudaStream_t st);
int l2norm_forward(const float* e, int B, int E, void* p_bf16, float* p_f32, float* norm, cudaStream_t st);
int l2norm_backward(const float* dp, const float* p, const float* norm, int B, int E, float* de, cudaStream_t st);
int pool_bwd_coef(const float* de, long long ld_de, const float* stats, long long ld_stats, const int* seq_off, int B,
                  int D, int mode, float* c1, float* c2, cudaStream_t st);
// one column-sum job of bias_grads_final (rows.cu)
struct BiasJob {
  const float* partial;   // [nparts][D] partial rows (nullptr: no job)
  int D;
  const int* nparts_dev;  // device count of partial rows, or nullptr: ceil(min(*n_rows, n_static) / block_rows)
  int nparts_max;         // rows physically present in `partial`
  const int* n_rows;
  int n_static;
  int block_rows;
  void* out_bf16;         // optional
  float* out_f32;         // optional
  int accumulate;         // add to out_f32 before storing
};
int bias_grads_final(const BiasJob& j0, const BiasJob& j1, cudaStream_t st);
int adapter_tail_backward(const void* a, const void* g, const float* inv_norm, const int* seq_off, int B, const float* c1,
                          const float* c2, const int* n_rows, int cap, int D, void* dz2, float* colsum_partial, int ws_rows,
                          int* nparts_dev, void* db2, cudaStream_t st);
int adapter_tail_backward_dy(const void* a, const void* g, const float* inv_norm, const void* dy, int n,
                             const int* n_dev, int cap, int D, void* dz2, cudaStream_t st);
int scale_rows(const void* a, const float* rowsq, int nblk, int cap, int n, int D, void* y, float* inv_norm_out,
               cudaStream_t st);
int scatter_scaled_rows(const void* a, const float* rowsq, int nblk, int cap, int n, int D, void* y, long long ld_y,
                        const int* row_dst, const int* n_dev, const int* n_dst_dev, float* inv_norm_out, cudaStream_t st);
int readout_backward(const void* x, const void* mask, int mask_bytes, int B, int S, int D, const float* c1,
                     const float* c2, void* dx, cudaStream_t st);
int colsum(const void* x, const int* n_rows, int n_static, int D, float* partial, void* out_bf16, float* out_f32,
           cudaStream_t st);
int readout_last(const void* x, const int* counts, int B, int S, int D, float* out, cudaStream_t st);
int readout_last_bwd(const void* dout, const int* counts, int B, int S, int D, void* dx, cudaStream_t st);
int stage_rows_pull(const void* host_base, const long long* table, const int* piece_prefix, int n_seg, void* dst, int ctas,
                    cudaStream_t st);

int sim_small(const void* p, const void* t, bool in_f32, int R, int C, int E, float alpha, float* S, cudaStream_t st);
int col_stats(const float* S, int R, int C, float* col_max, float* col_sum, int* col_argmax, int row_index_base,
              cudaStream_t st);
int mark_labelled(const int* labels, int R, int C, unsigned char* marks, cudaStream_t st);
int ce_rows(float* S, const int* labels, int R, int C, float w_row, float w_col, float inv_rn, const float* col_max,
            const float* col_sum, const unsigned char* col_labelled, float* row_loss, float* row_lse, int* argmax_row,
            void* dS_bf16, int write_ds, cudaStream_t st);
int loss_mean(const float* row_loss, int R, float scale, float* loss, int accumulate, cudaStream_t st);
int contract_small(const float* W, int ldw, bool transpose_w, const void* X, bool x_f32, int n_out, int n_red, int E,
                   float alpha, float* out, cudaStream_t st);
int sim_cols_combine(const float4* col_part, int nrb, int C, float* col_max, float* col_sum, int* col_argmax, cudaStream_t st);
int sim_finish(const float4* row_part, int npart, int ld, const float* pos, const int* labels, int R, int C, float w_row,
               float w_col, const float* col_max, const float* col_sum, float* col_lse, float* row_loss, float* row_lse,
               int* argmax_row, cudaStream_t st);
int convert_f32_to_bf16(const float* in, long long n, void* out, cudaStream_t st);
int convert_bf16_to_f32(const void* in, long long n, float* out, cudaStream_t st);

struct GemmCommReduce;
int launch_gemm(const void* a, long long lda, bool a_mn, const void* b, long long ldb, bool b_mn, int epi,
                GemmParams p, int cta_group, cudaStream_t stream, const GemmCommReduce* comm = nullptr);

// peer.cu: exchange steps over NVLink peer memory
size_t peer_ctrl_bytes();
int peer_alloc(size_t bytes, void** dptr, unsigned char* handle64);
int peer_open(const unsigned char* handle64, void** dptr);
int peer_close(void* dptr);
int peer_free(void* dptr);
int peer_allgather(void* const* peers, int world, int rank, const void* src, long long bytes_per_rank, void* dst,
                   int phases, cudaStream_t st);
int peer_allreduce_mean(void* const* peers, int world, int rank, long long n_bytes, long long f32_from_byte, void* dst, int phases,
                        cudaStream_t st);
int peer_reset(void* base, cudaStream_t st);

// loss_fused.cu: similarity -> InfoNCE -> pooling coefficients of the backward pass in one cooperative kernel
bool loss_fused_eligible(int R, int B, int C, int E);
int loss_fused(const float* p, const float* t, void* const* peers, int world, int rank, long long bytes_per_rank,
               const int* labels, int R, int B, int C, int E, float tau, float w_row, float w_col, float scale,
               int all_cols_labelled, int want_col, int need_grad, const float* dloss, const float* pnorm,
               const float* stats, const int* seq_off, float* S_ws, float* dp_ws, unsigned* bar_ws, float* loss,
               float* row_lse, int* argmax_row, int* argmax_col, float* col_max, float* col_sum, float* c1, float* c2,
               cudaStream_t st);

// optim.cu: clip_grad_norm_ + AdamW over a table of tensors
constexpr int kAdamMaxTensors = 8;
struct AdamTable {
  int count;
  int block_start[kAdamMaxTensors];
  long long numel[kAdamMaxTensors];
  void* param[kAdamMaxTensors];       // bf16
  void* grad[kAdamMaxTensors];        // bf16
  const float* grad_f32[kAdamMaxTensors];  // optional fp32 source of the gradient (rounded into `grad` by the norm pass)
  float* exp_avg[kAdamMaxTensors];
  float* exp_avg_sq[kAdamMaxTensors];
  float* master[kAdamMaxTensors];     // fp32 master weights or nullptr
};
int adamw_blocks(long long numel);
int adamw_step(AdamTable t, float* partial_ws, float* scal, const float* lr_dev, long long* step_dev, float beta1,
               float beta2, float eps, float weight_decay, float max_norm, int zero_grad, cudaStream_t st);

size_t gemm_streamk_workspace_bytes();
const char* last_error();
unsigned long long launches();
void reset_launches();
void launch_timing_enable(bool on);
int launch_timing_collect(double* ms, int cap, int* n_out, char* names, int names_cap);
void gemm_timing_enable(bool on);
int gemm_timing_collect(double* total_ms, int* pairs, double* each_ms, int each_cap);

}  // namespace p2t
