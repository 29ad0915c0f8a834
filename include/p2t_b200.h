/* p2t_b200.h — C ABI of the B200-native Stage-1 contrastive hot path of Prot2Text-V2.
 *
 * The reference (RockingMat/Prot2Text-V2-esm3) is pure Python and has NO plugin/FFI layer for this
 * path: its boundary is a set of Python call signatures.  Each entry point below
 * names the reference interface it stands behind (paths relative to the reference root); the
 * Python host layer in `prot2text-v2-esm3_b200/` binds them with ctypes and mirrors those
 * signatures one-for-one (see INTEGRATION.md for the reference-side stub).
 *
 * Conventions
 *   - plain C: pointers, sizes, a stream handle.  No C++/torch types.
 *   - every data pointer is a DEVICE pointer owned by the caller (the host layer allocates with
 *     torch's caching allocator); nothing is allocated, freed or retained by the library.
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing synchronises.
 *   - return value: 0 = ok, < 0 = argument/setup error, > 0 = cudaError_t.  p2t_last_error()
 *     returns the message of the last failure on the calling thread.
 *   - bf16 tensors are row-major with 16-byte aligned rows (row length multiple of 8).
 *   - "packed rows": the valid residue rows of all sequences back to back, [n_rows][D];
 *     seq_off[b] (int32, B+1 entries) is the first packed row of sequence b.  n_rows lives in
 *     device memory (`n_rows_dev`) so ragged batches never force a host sync; `rows_cap` is the
 *     allocated row count (>= n_rows rounded up to 256).
 *   - there is no CPU fallback anywhere: without a CUDA device every compute entry returns > 0.
 */
#ifndef P2T_B200_H_
#define P2T_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define P2T_ABI_VERSION 2

/* readout modes — scripts/train_contrast.py:198-248 readout_fn */
#define P2T_READOUT_MEAN 1
#define P2T_READOUT_STD 2
#define P2T_READOUT_MIX 3

int p2t_abi_version(void);
const char* p2t_last_error(void);
/* kernels launched by this library since the last reset (bench.py reports it as gpu_launches) */
unsigned long long p2t_launch_count(void);
void p2t_reset_launch_count(void);
/* per-launch CUDA-event timing of the tcgen05 GEMM kernel on its launching stream (bench.py's
 * roofline): enable, run, synchronise the stream, then collect the summed durations and (optionally, up to
 * each_cap entries in launch order) every launch's own duration. HOST pointers. */
void p2t_gemm_timing_enable(int on);
int p2t_gemm_timing_collect(double* total_ms, int* launches, double* each_ms, int each_cap);

/* per-launch timing of EVERY kernel of the library (bench.py --stages): when enabled, an event is recorded on the
 * launching stream after each launch; p2t_launch_timing_mark records a "mark" that opens a step.  After synchronising,
 * collect returns, in launch order, the kernel names (newline separated, truncated to names_cap) and the time from the
 * previous stamp to each stamp in ms (0 for marks): on an in-order stream that is the kernel's duration plus its
 * launch gap.  HOST pointers.  Not for use under stream capture. */
void p2t_launch_timing_enable(int on);
int p2t_launch_timing_mark(void* stream);
int p2t_launch_timing_collect(double* ms, int cap, int* n, char* names, int names_cap);

/* ---------------------------------------------------------------------------------------------
 * tcgen05 GEMM  D[m][n] = alpha * sum_k A[m][k] * B[n][k]      (bf16 in, fp32 accumulate in TMEM)
 * Replaces: torch.mm / nn.Linear / their autograd GEMMs (cuBLAS in the reference):
 *   scripts/train_contrast.py:87,108 (similarity), models/modeling_esm2llama_instruct.py:62,65.
 * a_mn_major = 0: A stored [m][lda] (K contiguous); 1: A stored [k][lda] (M contiguous). Same for B.
 * d_is_f32 selects fp32 or bf16 output.  dyn_m / dyn_k: optional device int32 overriding m / k
 * (must be <= the static value).  cta_group: 1 = one SM per tile (128x256), 2 = CTA pair (256x256).
 * gemm_ws: optional scratch of p2t_gemm_workspace_bytes() bytes (16-byte aligned, contents
 * irrelevant).  When given and M is static, the tiles of an incomplete last wave are cut along K
 * so that all SMs stay busy; the partial tiles are summed in a fixed order through the scratch
 * (same result layout, deterministic).  NULL = whole tiles only.
 * ------------------------------------------------------------------------------------------- */
unsigned long long p2t_gemm_workspace_bytes(void);
int p2t_gemm_bf16(const void* a, long long lda, int a_mn_major, const void* b, long long ldb, int b_mn_major,
                  void* d, long long ldd, int d_is_f32, int m, int n, int k, float alpha, const int* dyn_m,
                  const int* dyn_k, void* gemm_ws, int cta_group, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Ragged row plan from a {0,1} attention mask [B][L] (mask_bytes = 1, 4 or 8 per element; the
 * reference passes int64, scripts/train_contrast.py:270-275 / dataset collaters).  Any mask
 * pattern is accepted (right padded, left padded, holes).
 *   counts[B], seq_off[B+1], chunk_off[B+1] (pooling chunks of `chunk_rows` rows), n_rows_dev[1],
 *   row_src[>= sum(mask)] = flat source row b*L + r of each packed row (may be NULL),
 *   chunk_seq[4 * (ceil(B*L/chunk_rows)+B)] = one int4 descriptor per pooling chunk
 *       {first packed row, end row, owning sequence, 0} (may be NULL).
 * ------------------------------------------------------------------------------------------- */
int p2t_rows_plan(const void* mask, int mask_bytes, int B, int L, int chunk_rows, int* counts, int* seq_off,
                  int* chunk_off, int* n_rows_dev, int* row_src, int* chunk_seq, void* stream);

/* The same plan for rows that are already packed (the ragged hand-over format of SURVEY.md §8f-3: the encoder or
 * the host stager delivers [sum L_b][D] rows plus per-sequence counts): counts[B] int32 on the device. */
int p2t_rows_plan_counts(const int* counts, int B, int chunk_rows, int* seq_off, int* chunk_off, int* n_rows_dev,
                         int* chunk_seq, void* stream);

/* Host -> device staging of a padded host batch [B][L][row_bytes] (pinned memory) as PACKED rows: sequence b
 * contributes rows [starts[b], starts[b] + counts[b]) (HOST int arrays; contiguous valid ranges, i.e. right or
 * left padding), written back to back at dev_dst.  Only valid rows cross PCIe; one cudaMemcpyAsync per sequence,
 * sequence b on streams[b % n_streams] (HOST array of cudaStream_t: several copy streams overlap one copy's DMA
 * set-up with its neighbours' transfers; the caller joins the streams).  Replaces `tensor.to(rank)` of the padded
 * batch (scripts/train_contrast.py:329-330). */
int p2t_stage_rows_h2d(const void* host_src, long long seq_stride_bytes, long long row_bytes, const int* starts,
                       const int* counts, int B, void* dev_dst, void* const* streams, int n_streams);
/* The same staging as ONE kernel that PULLS the valid rows out of pinned host memory (the pointer must be device-
 * accessible: cudaHostAlloc / torch pin_memory under unified addressing) instead of one copy-engine transfer per
 * sequence (~3.5 us of set-up each).  DEVICE arrays: table[3 * n_seg] int64 = {byte offset in the host batch, byte
 * offset in dev_dst, byte length} per segment, all multiples of 16; piece_prefix[n_seg + 1] int32 = 32 KB pieces
 * before each segment (last entry = total).  ctas: CTAs of the launch (0 = 32); they share the SMs with whatever else
 * runs. */
int p2t_stage_rows_pull(const void* host_mapped, const long long* table_dev, const int* piece_prefix_dev, int n_seg,
                        void* dev_dst, int ctas, void* stream);

/* out[i] = src[row_src[i]] (bf16 rows of D elements), zero rows from n_rows up to the next multiple
 * of 256 (<= rows_cap).  Packs the padded (B, L, D_in) residue states that
 * models/esmc_qwen_arc.py:84-86 hands to the adapter. */
int p2t_gather_rows(const void* src, long long ld_src, const int* row_src, const int* n_rows_dev, int rows_cap,
                    int D, void* out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * ModalityAdapter.forward on packed rows — models/modeling_esm2llama_instruct.py:60-68.
 *   h1 = drop(GELU(x W1^T + b1))            [rows_cap][d_mid] bf16   (fc1 epilogue)
 *   g1 = keep * GELU'(x W1^T + b1)          [rows_cap][d_mid] fp16   (NULL when no backward)
 *   a  = drop(GELU(h1 W2^T + b2))           [rows_cap][d_out] fp16   (fc2 epilogue)
 *   g2 = keep * GELU'(h1 W2^T + b2)         [rows_cap][d_out] fp16   (NULL when no backward)
 *   (a, g1, g2 are read only by this library's streaming kernels; fp16 keeps 3 more mantissa bits
 *    than bf16 at the same bytes.  h1 and every GEMM operand are bf16.)
 *   rowsq[nblk][rows_cap] = partial sums of a^2, nblk = 4*ceil(d_out/256) — the per-residue L2 norm (:67)
 * x has x_rows rows allocated (TMA zero-fills beyond); outputs have rows_cap rows.
 * Weights in nn.Linear layout (out, in), bf16.  dropout_p = 0 is eval mode; otherwise a Philox
 * mask keyed by (seed, layer, row, column) with multiplier 1/(1-p) (nn.Dropout, :63,:66).
 * gemm_ws (optional, p2t_gemm_workspace_bytes()): split-K scratch — the tiles of the incomplete last wave are cut into
 * K ranges; the row count is ragged, so the kernel derives the cut from *n_rows_dev itself.
 * seed_dev (optional, device): its value is added to `seed` on the device — a CUDA-graph replay freezes
 * kernel arguments, so the replaying host bumps this word instead.
 * ------------------------------------------------------------------------------------------- */
int p2t_adapter_fwd(const void* x, int x_rows, const void* w1, const void* b1, const void* w2, const void* b2, int d_in,
                    int d_mid, int d_out, int rows_cap, const int* n_rows_dev, void* h1, void* g1, void* a,
                    void* g2, float* rowsq, float dropout_p, unsigned long long seed, const unsigned long long* seed_dev, void* gemm_ws,
                    int cta_group, void* stream);

/* y[row] = a[row] / max(|a[row]|, 1e-12) for the first n rows: the (B, L, d_out) tensor that
 * ModalityAdapter.forward returns (:67-68).  inv_norm[row] is saved for backward (may be NULL). */
int p2t_adapter_scale_rows(const void* a, const float* rowsq, int nblk, int rows_cap, int n, int d_out, void* y,
                           float* inv_norm, void* stream);

/* inv_norm[row] = 1 / max(sqrt(sum_j rowsq[row][j]), 1e-12) for row < n_rows: the per-residue
 * F.normalize denominator of models/modeling_esm2llama_instruct.py:67 */
int p2t_row_inv_norm(const float* rowsq, int nblk, const int* n_rows_dev, int rows_cap, float* inv_norm, void* stream);

/* ---------------------------------------------------------------------------------------------
 * readout_embeddings(..., "mean"|"std"|"mix") — scripts/train_contrast.py:217-248 — over the rows
 * listed by a plan.  `src` is bf16 (or fp16 when src_is_f16: the adapter's own `a`) [src_rows][ld_src], 16-byte
 * aligned rows (read by TMA);
 * row_src == NULL means rows are already packed.  With inv_norm != NULL every row is first scaled
 * by inv_norm[row] (adapter output: normalise fused into the pooling pass).
 * partial_ws: fp32 [max_chunks * (2 D + 1)] (float2 records [max_chunks][D], then one int32 row count per chunk);
 * out: fp32 [B][ld_out] (mean | std for mix).
 * With p_bf16 and/or p_f32 given (mode MIX, ld_out == 2*D) the following F.normalize (:354/:365) is fused into
 * the final pass: p = out / max(|out|, 1e-12) [B][2*D], norm[B] = |out| (unclamped).
 * ------------------------------------------------------------------------------------------- */
int p2t_pool_fwd(const void* src, int src_is_f16, long long ld_src, int src_rows, const int* row_src, const float* inv_norm,
                 const int* seq_off, const int* chunk_off, const int* chunk_seq, int B, int D, int chunk_rows,
                 int max_chunks, int mode, void* partial_ws, float* out, long long ld_out, void* p_bf16, float* p_f32,
                 float* norm, void* stream);

/* readout_embeddings(..., "last") — :207-215 — out fp32 [B][D] from padded x [B][S][D] */
int p2t_readout_last(const void* x, const int* counts, int B, int S, int D, float* out, void* stream);
/* its backward: dx bf16 [B, S, D] = 0 except dx[b, counts[b] - 1] = dout[b] (bf16 [B, D]); writes every element of dx */
int p2t_readout_last_bwd(const void* dout, const int* counts, int B, int S, int D, void* dx, void* stream);

/* F.normalize(p=2, dim=-1) on pooled embeddings — scripts/train_contrast.py:354,365.
 * e fp32 [B][E] -> p (bf16 and/or fp32, either may be NULL), norm[B] (unclamped). */
int p2t_l2norm_fwd(const float* e, int B, int E, void* p_bf16, float* p_f32, float* norm, void* stream);
int p2t_l2norm_bwd(const float* dp, const float* p_f32, const float* norm, int B, int E, float* de, void* stream);

/* backward of the readout: coefficient vectors with dy_r = c1[b] + c2[b] * y_r (fp32 [B][D]) */
int p2t_pool_bwd_coef(const float* de, long long ld_de, const float* stats, long long ld_stats, const int* seq_off,
                      int B, int D, int mode, float* c1, float* c2, void* stream);
/* The same coefficients straight from dLogits, for small / medium similarity blocks with fp32 embeddings (autograd of
 * :108-113 -> :365 -> :277-281): dp = dloss/tau * dS t (rows >= R get 0), F.normalize backward, 'mix' coefficients.
 * dS fp32 [R][C] (as left by p2t_infonce_ce), t_f32 [C][2D], p_f32 [B][2D], pnorm[B], stats [B][2D], dloss: device
 * scalar or NULL (= 1); dp_ws: fp32 scratch [B][2D + ceil(2D/64)]. */
int p2t_loss_bwd_coef(const float* dS, const float* t_f32, const float* p_f32, const float* pnorm, const float* stats,
                      const int* seq_off, const float* dloss, int R, int B, int C, int D, float tau, float* dp_ws,
                      float* c1, float* c2, void* stream);
/* dx[b,r] = mask[b,r] * (c1[b] + c2[b] * x[b,r]) on a padded bf16 (B, S, D) tensor */
int p2t_readout_bwd(const void* x, const void* mask, int mask_bytes, int B, int S, int D, const float* c1,
                    const float* c2, void* dx, void* stream);

/* backward through normalise -> GELU(fc2) on packed rows (autograd of :65-67):
 *   dz2 = ((dy - y (y.dy)) / |a|) * g2,  dy = c1[b] + c2[b]*y   (pooled)   or given per row (_dy)
 * The pooled form is persistent (every CTA streams an equal share of the valid rows, crossing sequence boundaries) and
 * also emits the column sums of dz2 (fc2.bias gradient) as ONE partial row per CTA: colsum_ws fp32 [ws_rows][d_out]
 * (ws_rows bounds the grid; 2 * #SM is always enough), *nparts_dev = rows written.  db2 (bf16 [d_out]) != NULL
 * finishes the sum here; the fused step passes NULL and finishes both biases with p2t_bias_grads. */
int p2t_adapter_tail_bwd(const void* a, const void* g2, const float* inv_norm, const int* seq_off, int B, const float* c1,
                         const float* c2, const int* n_rows_dev, int rows_cap, int d_out, void* dz2, float* colsum_ws,
                         int ws_rows, int* nparts_dev, void* db2, void* stream);
/* (_dy form: rows < min(n, *n_rows_dev) are computed — n_rows_dev may be NULL —, rows up to the next multiple of 256
 * are zeroed) */
int p2t_adapter_tail_bwd_dy(const void* a, const void* g2, const float* inv_norm, const void* dy, int n,
                            const int* n_rows_dev, int rows_cap, int d_out, void* dz2, void* stream);

/* ---------------------------------------------------------------------------------------------
 * backward GEMMs of the adapter (autograd of :62-65; scripts/train_contrast.py:448):
 *   dz1 = (dz2 W2) * g1;  dW2 = dz2^T h1;  db2 = colsum dz2;  dW1 = dz1^T x;  db1 = colsum dz1
 *   dx  = dz1 W1 (only when dx != NULL; the encoder is frozen in Stage 1, :186)
 * Gradients are in nn.Linear layout, bf16 — or, dw_is_f32 != 0, dw1/dw2 in fp32 (the form a gradient mean over ranks
 * should carry: a rank's gradient can be several times larger than the mean, so a bf16 rounding per rank costs several
 * times 2^-9 of the mean) —; written, or — accumulate != 0 — added to what dw1/dw2 hold (the reference
 * accumulates .grad over gradient_accumulation_steps micro-batches, scripts/train_contrast.py:448-465).
 * db1/db2/dx may be NULL (db2 is normally produced by p2t_adapter_tail_bwd).
 * colsum_ws: fp32 [ceil(rows_cap/32)][d_mid] — the dgrad GEMM's epilogue leaves the column sums of dz1 over every
 * 32-row block there (db1 = their sum: db1 != NULL finishes it here, else p2t_bias_grads does) — followed, only when
 * db2 != NULL, by [ceil(rows_cap/64)][d_out].
 * gemm_ws: optional split-K scratch for the two weight-gradient GEMMs (see p2t_gemm_bf16).
 * ------------------------------------------------------------------------------------------- */
/* phases: which GEMMs this call enqueues (0 = all three).  The fused sharded training step issues DGRAD | DW2 first,
 * finishes db2, and then DW1 with `overlap`: a gradient-mean channel (layout and meaning of p2t_peer_allreduce_mean,
 * holding this rank's dW2 / db2 contribution, all fp32: f32_from_byte = 0) whose announce + reduce phases are serviced
 * INSIDE the dW1 GEMM's launch by the epilogue warps of every CTA while their first accumulator is being computed —
 * compute step and collective in one kernel over NVLink peer memory, no SM taken from the GEMM.  The round is closed
 * afterwards with p2t_peer_allreduce_mean(phases = 4).  Needs dw_is_f32 and cta_group 2.  (DDP overlaps its bucket
 * all-reduces with the backward the same way, scripts/train_contrast.py:448 + :611-614.) */
#define P2T_BWD_DGRAD 1
#define P2T_BWD_DW2 2
#define P2T_BWD_DW1 4
typedef struct {
  void* const* peers;      /* HOST array of `world` device pointers (see p2t_peer_allreduce_mean) */
  int world, rank;
  long long n_bytes, f32_from_byte;
} p2t_overlap_reduce_t;
int p2t_adapter_bwd(const void* x, int x_rows, const void* w1, const void* w2, const void* h1, const void* g1, const void* dz2,
                    int d_in, int d_mid, int d_out, int rows_cap, const int* n_rows_dev, void* dz1, void* dw1,
                    void* db1, void* dw2, void* db2, void* dx, float* colsum_ws, void* gemm_ws, int accumulate,
                    int dw_is_f32, int phases, const p2t_overlap_reduce_t* overlap, int cta_group, void* stream);
/* Both bias gradients of the fused step in one launch, summed in a fixed order from the partial rows left by
 * p2t_adapter_bwd (db1_partial = its colsum_ws) and p2t_adapter_tail_bwd (db2_partial = its colsum_ws, nparts2_dev,
 * ws_rows2).  Outputs in fp32 (what the gradient all-reduce carries: ONE rounding to bf16, after the mean over ranks)
 * and/or bf16 (the parameter's .grad).  accumulate != 0 adds to the fp32 outputs.  Either job may be NULL. */
int p2t_bias_grads(const float* db1_partial, int rows_cap, const int* n_rows_dev, int d_mid, void* db1_bf16, float* db1_f32,
                   const float* db2_partial, const int* nparts2_dev, int ws_rows2, int d_out, void* db2_bf16,
                   float* db2_f32, int accumulate, void* stream);

/* ---------------------------------------------------------------------------------------------
 * InfoNCE — scripts/train_contrast.py:86-91 (BatchInfoNCELoss), :100-114 (Segmented...).
 * p2t_similarity : S[R][C] = (p . t^T) / tau   (fp32 out; unit-norm inputs [R][E], [C][E] as bf16 and/or
 *      fp32 — small problems use the fp32 copies on CUDA cores when given, large ones the bf16 tcgen05 GEMM)
 * p2t_infonce_col_stats : per-column online (max, sum exp, argmax row) — the text->protein term
 * p2t_infonce_ce : one pass per row: loss_i, row lse, argmax, and dS written over S
 *      loss_i = w_row (lse_j S_ij - S_i,lab) + w_col (lse_col[lab] - S_i,lab);  dS scaled by inv_rn
 *      all_cols_labelled != 0: sharded global batch, every column's positive lives on some rank
 * p2t_infonce_grad : dp = dS t / tau (fp32 [R][E]) and optionally dt = dS^T p / tau (fp32 [C][E]); same
 *      bf16 / fp32 operand convention as p2t_similarity
 * p2t_loss_mean : loss[0] (+)= scale * sum_i row_loss[i], fixed-order reduction
 * ------------------------------------------------------------------------------------------- */
int p2t_similarity(const void* p, const void* t, const float* p_f32, const float* t_f32, int R, int C, int E, float tau,
                   float* S, int cta_group, void* stream);
int p2t_infonce_col_stats(const float* S, int R, int C, float* col_max, float* col_sum, int* col_argmax,
                          int row_index_base, void* stream);
int p2t_infonce_ce(float* S, const int* labels, int R, int C, float w_row, float w_col, float inv_rn,
                   const float* col_max, const float* col_sum, unsigned char* col_labelled_ws, int all_cols_labelled,
                   float* row_loss, float* row_lse, int* argmax_row, void* dS_bf16, int write_ds, void* stream);
int p2t_infonce_grad(const float* dS, const void* dS_bf16, const void* p, const void* t, const float* p_f32,
                     const float* t_f32, int R, int C, int E, float tau, float* dp, float* dt, int cta_group,
                     void* stream);
/* Large blocks (R*C*E > 2^26), logits never materialised (north_star (3)): p2t_infonce_stats runs the similarity as a
 * tcgen05 GEMM whose epilogue reduces every 32 x 64 piece of a tile to online-softmax partials — per row (max, sum exp,
 * argmax) and, when col_part_ws != NULL, per column over 32-row blocks (max, sum exp, arg-max row) — plus the label
 * logit pos[i] = S[i][lab(i)], and merges the column partials into col_max / col_sum / col_argmax (the same
 * quantities p2t_infonce_col_stats produces, so ranks can combine them).  p2t_infonce_finish merges the row partials
 * (row_lse, argmax_row), forms row_loss as p2t_infonce_ce does, and — dS_bf16 != NULL — recomputes S tile by tile in a
 * second GEMM whose epilogue writes dLogits = inv_rn (w_row (softmax_row - onehot) + w_col (softmax_col - onehot)) in
 * bf16, the A operand of p2t_infonce_grad.  p, t: bf16 [R][E], [C][E].
 * Workspaces: row_part_ws 16 B x 4 ceil(C/256) x R; col_part_ws 16 B x ceil(R/32) x C; col_labelled_ws C bytes;
 * col_lse_ws C floats.  A label outside [0, C) yields a NaN row loss. */
int p2t_infonce_stats(const void* p, const void* t, const int* labels, int R, int C, int E, float tau, void* row_part_ws,
                      void* col_part_ws, float* pos, float* col_max, float* col_sum, int* col_argmax, int cta_group,
                      void* stream);
int p2t_infonce_finish(const void* p, const void* t, const int* labels, int R, int C, int E, float tau, float w_row,
                       float w_col, float inv_rn, const void* row_part_ws, const float* pos, const float* col_max,
                       const float* col_sum, int all_cols_labelled, unsigned char* col_labelled_ws, float* col_lse_ws,
                       float* row_loss, float* row_lse, int* argmax_row, void* dS_bf16, int cta_group, void* stream);
int p2t_loss_mean(const float* row_loss, int R, float scale, float* loss, int accumulate, void* stream);

/* The loss block of the fused step for small similarity blocks in ONE cooperative kernel (north_star (3): similarity,
 * row- and column-wise online-softmax cross-entropy, loss and dLogits in one pass, logits and probabilities never in
 * HBM as an autograd graph would keep them) — scripts/train_contrast.py:100-114, :354/:365, :237-248 and their
 * autograd down to the pooling coefficients c1, c2 (dy_r = c1[b] + c2[b] y_r).
 *   p_f32 [B][E] (rows < R enter the loss), t_f32 [C][E]; or t_f32 == NULL and (gather_peers, world, rank,
 *   gather_bytes_per_rank) name a p2t_peer_allgather channel whose push half has been issued: the kernel waits for
 *   the round itself, reads the gathered rows in place and closes the round (no arrive kernel, no copy).
 *   labels int32 [R]; loss = loss_scale * sum_i [w_row (lse_j S_ij - S_i,lab) + w_col (lse_col[lab] - S_i,lab)].
 *   need_grad != 0: dp = dloss/tau dLogits t, F.normalize backward (pnorm [B]), 'mix' coefficients (stats [B][E],
 *   seq_off [B+1]) -> c1, c2 fp32 [B][E/2]; rows >= R get zero gradient.  dloss: device scalar or NULL (= 1).
 *   Workspaces: S_ws fp32 [R][C]; dp_ws fp32 [B][E + ceil(E/64)]; barrier_ws 16 bytes, zero before the FIRST use
 *   (the kernel re-arms it).  Outputs other than loss may be NULL.  A label outside [0, C), a timed-out exchange
 *   round or a stuck grid barrier yields a NaN loss.  Eligible when R*C <= 16384 (p2t_loss_fused_eligible). */
int p2t_loss_fused(const float* p_f32, const float* t_f32, void* const* gather_peers, int world, int rank,
                   long long gather_bytes_per_rank, const int* labels, int R, int B, int C, int E, float tau, float w_row,
                   float w_col, float loss_scale, int all_cols_labelled, int want_col_stats, int need_grad,
                   const float* dloss, const float* pnorm, const float* stats, const int* seq_off, float* S_ws,
                   float* dp_ws, void* barrier_ws, float* loss, float* row_lse, int* argmax_row, int* argmax_col,
                   float* col_max, float* col_sum, float* c1, float* c2, void* stream);
int p2t_loss_fused_eligible(int R, int B, int C, int E);

/* dtype helpers used by the host layer around the fp32 <-> bf16 boundaries */
int p2t_f32_to_bf16(const float* in, long long n, void* out, void* stream);
int p2t_bf16_to_f32(const void* in, long long n, float* out, void* stream);
/* deterministic column sums of a packed bf16 matrix (bias gradients); ws: fp32 [ceil(n_static/64)][D] */
int p2t_colsum(const void* x, const int* n_rows_dev, int n_static, int D, float* ws, void* out_bf16, float* out_f32,
               void* stream);
/* debug/test aid: the dropout keep-multipliers the kernels use, as fp32 [rows][cols] */
int p2t_dropout_mask(int rows, int cols, float dropout_p, unsigned long long seed, int layer, float* out,
                     void* stream);

/* ---------------------------------------------------------------------------------------------
 * Stage-2 hand-off — replaces `inputs_embeds[placeholder_mask] = encoder_hidden_states[encoder_mask]`
 * (models/esmc_qwen_arc.py:127-144, models/modeling_esm2llama_instruct.py:134-138): the adapter's k-th valid output
 * row y_k = a_k / max(|a_k|, 1e-12) is written to row row_dst[k] of dst (bf16, row stride ld_dst elements) — i.e. into
 * the k-th placeholder slot of the LLM's input embeddings — without materialising (B, L, d_out).  Rows k < min(n,
 * *n_rows_dev, *n_dst_dev) are written (either device count may be NULL).  inv_norm[k] is saved when non-NULL.
 * ------------------------------------------------------------------------------------------- */
int p2t_adapter_scatter_rows(const void* a, const float* rowsq, int nblk, int rows_cap, int n, int d_out, void* dst,
                             long long ld_dst, const int* row_dst, const int* n_rows_dev, const int* n_dst_dev,
                             float* inv_norm, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Exchange steps of the sharded step over NVLink peer memory (the reference has none: D6) and of the
 * adapter weight gradients (DistributedDataParallel in the reference, scripts/train_contrast.py:611-614).
 * Kernels on the caller's stream, epochs in device memory: capturable in a CUDA graph, no NCCL on the data path.
 *
 * Peer buffers: p2t_peer_alloc = cudaMalloc + zero fill + cudaIpcGetMemHandle (64-byte handle, HOST pointer);
 * a peer process maps it with p2t_peer_open (cudaIpcOpenMemHandle, peer access enabled lazily) and unmaps with
 * p2t_peer_close; the owner releases with p2t_peer_free.  These four are the only entries that allocate or
 * synchronise.  Every channel buffer starts with p2t_peer_ctrl_bytes() of control words (zero = fresh).
 *
 * `peers` is a HOST array of `world` device pointers: this process's mapping of every rank's channel buffer
 * (peers[rank] = own buffer).  world <= 16.
 *
 * p2t_peer_allgather: buffer size ctrl + 2 * world * bytes_per_rank.  phases bit 0 (push): store `src`
 * (bytes_per_rank, multiple of 16) into slot `rank` of every peer and raise the arrival flags; bit 1 (arrive): wait
 * for all ranks' blocks of this round and copy the world * bytes_per_rank gathered bytes, rank-major, to `dst`.
 * Kernels between the two halves overlap the transfer.  Rounds are double-buffered: a rank may run one round ahead.
 *
 * p2t_peer_allreduce_mean: buffer size ctrl + 2 * n_bytes; the caller has written its contribution (n_bytes multiple
 * of 16: bf16 values in [0, f32_from_byte), fp32 values from f32_from_byte on — the bias gradients travel in fp32 and
 * are rounded once, after the mean; f32_from_byte < 0 or == n_bytes: all bf16) at offset ctrl.  phases bit 0: announce;
 * bit 1: sum slice `rank` over all peers in rank order in fp32, store the mean (same format) to every peer at offset
 * ctrl + n_bytes; bit 2: wait for all slices and copy the result to `dst` (may be NULL: read it in place).  Every rank
 * obtains bit-identical results.
 *
 * A rank that waits more than 20 s for a peer stores 1 + that peer's rank in control word 4 (sticky) and POISONS what
 * the round produces: the gathered rows / the reduced gradients become NaN, so the step's loss (or the next
 * optimizer step's gradient norm) is NaN instead of silently stale.  p2t_peer_reset (collective: every rank, between
 * two host-side barriers, with no round in flight) re-initialises the control block.
 * ------------------------------------------------------------------------------------------- */
unsigned long long p2t_peer_ctrl_bytes(void);
int p2t_peer_alloc(unsigned long long bytes, void** dptr, unsigned char* handle64);
int p2t_peer_open(const unsigned char* handle64, void** dptr);
int p2t_peer_close(void* dptr);
int p2t_peer_free(void* dptr);
int p2t_peer_allgather(void* const* peers, int world, int rank, const void* src, long long bytes_per_rank, void* dst,
                       int phases, void* stream);
int p2t_peer_allreduce_mean(void* const* peers, int world, int rank, long long n_bytes, long long f32_from_byte, void* dst,
                            int phases, void* stream);
int p2t_peer_reset(void* channel_base, void* stream);
/* cudaMemcpyAsync device -> device on `stream` (moving contributions into / results out of a channel buffer, which
 * torch cannot address as a tensor) */
int p2t_copy_d2d(void* dst, const void* src, unsigned long long bytes, void* stream);
/* control word 4 of this rank's channel buffer -> *status_host (synchronous; 0 = no time-out so far) */
int p2t_peer_status(const void* channel_base, unsigned int* status_host);

/* ---------------------------------------------------------------------------------------------
 * torch.nn.utils.clip_grad_norm_ + torch.optim.AdamW.step + zero_grad for up to 8 bf16 tensors per call
 * (scripts/train_contrast.py:455-465, optimizer built at :621-626 with eps=1e-6, betas=(0.9, 0.999)).
 * HOST arrays of `count` DEVICE pointers: params/grads bf16 (16-byte aligned), exp_avg/exp_avg_sq fp32,
 * master fp32 or NULL entries / NULL array (then the bf16 parameter is the state, as in the reference).
 *   partial_ws: fp32 [p2t_adamw_workspace_floats(count, numel)];  scal: fp32 [4] = {|g|, clip coefficient,
 *   lr / (1 - beta1^t), 1 / sqrt(1 - beta2^t)} written each call;  lr_dev: device float;  step_dev: device int64 step
 *   counter, incremented by the call.  max_norm <= 0 or inf: no clipping (|g| is still reported).
 *   g' = g * min(1, max_norm / (|g| + 1e-6));  p *= 1 - lr*wd;  m = b1 m + (1-b1) g';  v = b2 v + (1-b2) g'^2;
 *   p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps).   zero_grad != 0 also clears the gradients.
 *   grads_f32 (NULL, or NULL entries: none): fp32 sources of the gradients — the mean over ranks left by
 *   p2t_peer_allreduce_mean — rounded to bf16 into grads[i] by the norm pass (the sharded step's ONE rounding; saves
 *   the separate p2t_f32_to_bf16 launches), 16-byte aligned.
 * ------------------------------------------------------------------------------------------- */
int p2t_adamw_workspace_floats(int count, const long long* numel);
int p2t_adamw_step(int count, void* const* params, void* const* grads, const float* const* grads_f32, float* const* exp_avg,
                   float* const* exp_avg_sq, float* const* master, const long long* numel, float* partial_ws, float* scal, const float* lr_dev,
                   long long* step_dev, float beta1, float beta2, float eps, float weight_decay, float max_norm,
                   int zero_grad, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* P2T_B200_H_ */
