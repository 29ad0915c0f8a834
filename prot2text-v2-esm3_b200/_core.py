"""Raw-buffer host functions: allocate workspaces with torch, enqueue kernels through the C ABI.

Nothing here knows about autograd; `adapter.py`, `readout.py`, `losses.py` and `step.py` wrap
these in `torch.autograd.Function`s that mirror the reference's call signatures.
All functions enqueue on the CURRENT torch CUDA stream and never synchronise.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional

import torch

from . import _lib

CHUNK_ROWS = 64          # pooling chunk (rows per partial-moment block)
ROW_ALIGN = 256          # packed buffers are allocated in multiples of this many rows
READOUT_MODES = {"mean": 1, "std": 2, "mix": 3}


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _round_up(n: int, m: int) -> int:
    return (n + m - 1) // m * m


def require_cuda_bf16(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise _lib.P2TError(f"{name} must be a CUDA tensor: this package has no CPU path")
    if t.dtype != torch.bfloat16:
        raise _lib.P2TError(f"{name} must be bfloat16 (got {t.dtype}); the sm_100a kernels compute in bf16 "
                            "with fp32 accumulation, as the reference's default --torch_dtype bfloat16")


def default_cta_group() -> int:
    import os
    return int(os.environ.get("P2T_CTA_GROUP", "2"))


# --------------------------------------------------------------------------------------------------
# ragged plan
# --------------------------------------------------------------------------------------------------
@dataclass
class RowPlan:
    B: int
    L: int
    counts: torch.Tensor      # int32 [B]
    seq_off: torch.Tensor     # int32 [B+1]
    chunk_off: torch.Tensor   # int32 [B+1]
    n_rows: torch.Tensor      # int32 [1] (device)
    row_src: Optional[torch.Tensor]  # int32 [B*L]
    chunk_seq: torch.Tensor   # int32 [max_chunks, 4] per pooling chunk: first row, end row, sequence, 0
    rows_cap: int             # allocation size for packed buffers
    max_chunks: int           # upper bound on pooling chunks


def plan_rows(mask: torch.Tensor, want_row_src: bool = True, max_valid_rows: Optional[int] = None) -> RowPlan:
    """Device-side plan of the valid rows of a {0,1} mask (B, L); no host synchronisation.
    `max_valid_rows`: a host-side upper bound on the number of valid rows (e.g. the sum of the sequence lengths the
    data loader already knows): packed buffers are then sized for it instead of for B*L — at config 5 / 2 GPUs the
    difference between fitting 180 GB and not.  Valid rows beyond the bound would be dropped: it must be a true bound."""
    if mask.dim() != 2 or not mask.is_cuda:
        raise _lib.P2TError("attention mask must be a 2-D CUDA tensor")
    mask = mask.contiguous()
    if mask.dtype == torch.bool:
        mask = mask.view(torch.uint8)
    nbytes = mask.element_size()
    if mask.dtype.is_floating_point or nbytes not in (1, 4, 8):
        mask = mask.to(torch.int32)
        nbytes = 4
    B, L = mask.shape
    dev = mask.device
    ints = torch.empty(3 * (B + 1) + 1, dtype=torch.int32, device=dev)
    counts, seq_off, chunk_off, n_rows = ints[:B], ints[B + 1:2 * B + 2], ints[2 * B + 2:3 * B + 3], ints[3 * B + 3:]
    row_src = torch.empty(B * L, dtype=torch.int32, device=dev) if want_row_src else None
    rows_bound = B * L if max_valid_rows is None else max(1, min(B * L, int(max_valid_rows)))
    max_chunks = (rows_bound + CHUNK_ROWS - 1) // CHUNK_ROWS + B
    chunk_seq = torch.empty(max_chunks, 4, dtype=torch.int32, device=dev)
    _lib.call("p2t_rows_plan", _ptr(mask), nbytes, B, L, CHUNK_ROWS, _ptr(counts), _ptr(seq_off), _ptr(chunk_off),
              _ptr(n_rows), _ptr(row_src), _ptr(chunk_seq), _stream())
    return RowPlan(B=B, L=L, counts=counts, seq_off=seq_off, chunk_off=chunk_off, n_rows=n_rows, row_src=row_src,
                   chunk_seq=chunk_seq, rows_cap=_round_up(rows_bound, ROW_ALIGN), max_chunks=max_chunks)


def plan_packed(counts: torch.Tensor, total_rows: int) -> RowPlan:
    """Plan for rows that are already packed back to back: `counts` (B,) valid rows per sequence (any integer
    dtype, CPU or CUDA), `total_rows` = the packed buffer's row count (a host integer >= sum(counts))."""
    if not counts.is_cuda:  # host-side lengths: check them here (device-side lengths are the caller's contract)
        if int(counts.min()) < 0 or int(counts.sum()) > total_rows:
            raise _lib.P2TError(f"packed rows: lengths must be >= 0 and sum to at most the buffer's {total_rows} rows")
    dev_counts = counts.to(device="cuda" if not counts.is_cuda else counts.device, dtype=torch.int32, non_blocking=True).contiguous()
    B = dev_counts.shape[0]
    dev = dev_counts.device
    ints = torch.empty(2 * (B + 1) + 1, dtype=torch.int32, device=dev)
    seq_off, chunk_off, n_rows = ints[:B + 1], ints[B + 1:2 * B + 2], ints[2 * B + 2:]
    max_chunks = (total_rows + CHUNK_ROWS - 1) // CHUNK_ROWS + B
    chunk_seq = torch.empty(max_chunks, 4, dtype=torch.int32, device=dev)
    _lib.call("p2t_rows_plan_counts", _ptr(dev_counts), B, CHUNK_ROWS, _ptr(seq_off), _ptr(chunk_off), _ptr(n_rows),
              _ptr(chunk_seq), _stream())
    return RowPlan(B=B, L=0, counts=dev_counts, seq_off=seq_off, chunk_off=chunk_off, n_rows=n_rows, row_src=None,
                   chunk_seq=chunk_seq, rows_cap=_round_up(max(total_rows, 1), ROW_ALIGN), max_chunks=max_chunks)


def dense_plan(B: int, L: int, device) -> RowPlan:
    """Plan for an all-ones mask (every row valid) without reading a mask."""
    ones = torch.ones(B, L, dtype=torch.uint8, device=device)
    return plan_rows(ones, want_row_src=False)


def gather_rows(x2d: torch.Tensor, plan: RowPlan) -> torch.Tensor:
    """Pack the valid rows of x2d (B*L, D) into [rows_cap, D]; pad rows up to the next 256 are zeroed."""
    D = x2d.shape[1]
    out = torch.empty(plan.rows_cap, D, dtype=torch.bfloat16, device=x2d.device)
    _lib.call("p2t_gather_rows", _ptr(x2d), x2d.stride(0), _ptr(plan.row_src), _ptr(plan.n_rows), plan.rows_cap, D,
              _ptr(out), _stream())
    return out


# --------------------------------------------------------------------------------------------------
# GEMM
# --------------------------------------------------------------------------------------------------
def gemm_workspace(device) -> Optional[torch.Tensor]:
    """Scratch for split-K GEMM tails (flags + raw accumulator slots); None when P2T_STREAMK=0."""
    import os
    if os.environ.get("P2T_STREAMK", "1") == "0":
        return None
    return torch.empty(int(_lib.load().p2t_gemm_workspace_bytes()), dtype=torch.uint8, device=device)


def _fwd_splitk_workspace(device) -> Optional[torch.Tensor]:
    """Split-K tail for the GEMMs whose row count is ragged (fc1, fc2, fc2-dgrad): OFF unless P2T_SPLITK_DYN=1.
    Measured at config 2 (560 / 1120 / 560 tiles on 74 CTA pairs: 5 % of the last wave idle): cutting the last
    wave's tiles costs more than it saves there — the head pieces' heavy epilogues wait for the dumped partial
    accumulators (13 % of the fc2 kernel's stall samples sat in that wait), fc1 132 -> 159 us, fc2 208 -> 252 us."""
    import os
    if os.environ.get("P2T_SPLITK_DYN", "0") != "1":
        return None
    return gemm_workspace(device)


def gemm(a: torch.Tensor, b: torch.Tensor, m: int, n: int, k: int, *, a_mn: bool = False, b_mn: bool = False,
         out_dtype=torch.bfloat16, alpha: float = 1.0, cta_group: Optional[int] = None,
         dyn_m: Optional[torch.Tensor] = None, dyn_k: Optional[torch.Tensor] = None,
         streamk: bool = False) -> torch.Tensor:
    """D[m][n] = alpha * sum_k A[m][k] B[n][k] on the tcgen05 kernel.

    a is stored [m][k] (a_mn=False) or [k][m] (a_mn=True); b likewise with n.
    """
    require_cuda_bf16(a, "a")
    require_cuda_bf16(b, "b")
    out = torch.empty(m, n, dtype=out_dtype, device=a.device)
    _lib.call("p2t_gemm_bf16", _ptr(a), a.stride(0), int(a_mn), _ptr(b), b.stride(0), int(b_mn), _ptr(out),
              out.stride(0), int(out_dtype == torch.float32), m, n, k, float(alpha), _ptr(dyn_m), _ptr(dyn_k),
              _ptr(gemm_workspace(a.device) if streamk else None), cta_group or default_cta_group(), _stream())
    return out


# --------------------------------------------------------------------------------------------------
# adapter on packed rows
# --------------------------------------------------------------------------------------------------
@dataclass
class AdapterActs:
    x: torch.Tensor           # packed input rows [x_rows, d_in]
    x_rows: int
    h1: torch.Tensor          # [rows_cap, d_mid]
    g1: Optional[torch.Tensor]
    a: torch.Tensor           # [rows_cap, d_out]
    g2: Optional[torch.Tensor]
    rowsq: torch.Tensor       # fp32 [nblk, rows_cap]
    nblk: int
    rows_cap: int
    n_rows: torch.Tensor      # int32 [1] device


def adapter_forward(x: torch.Tensor, x_rows: int, rows_cap: int, n_rows: torch.Tensor, w1, b1, w2, b2,
                    dropout_p: float, seed: int, need_grad: bool, cta_group: Optional[int] = None,
                    seed_dev: Optional[torch.Tensor] = None) -> AdapterActs:
    d_mid, d_in = w1.shape
    d_out = w2.shape[0]
    dev = x.device
    bf = torch.bfloat16
    h1 = torch.empty(rows_cap, d_mid, dtype=bf, device=dev)
    a = torch.empty(rows_cap, d_out, dtype=torch.float16, device=dev)
    g1 = torch.empty(rows_cap, d_mid, dtype=torch.float16, device=dev) if need_grad else None
    g2 = torch.empty(rows_cap, d_out, dtype=torch.float16, device=dev) if need_grad else None
    nblk = 4 * ((d_out + 255) // 256)
    rowsq = torch.empty(nblk, rows_cap, dtype=torch.float32, device=dev)
    _lib.call("p2t_adapter_fwd", _ptr(x), x_rows, _ptr(w1), _ptr(b1), _ptr(w2), _ptr(b2), d_in, d_mid, d_out,
              rows_cap, _ptr(n_rows), _ptr(h1), _ptr(g1), _ptr(a), _ptr(g2), _ptr(rowsq), float(dropout_p),
              int(seed) & 0xFFFFFFFFFFFFFFFF, _ptr(seed_dev), _ptr(_fwd_splitk_workspace(dev)),
              cta_group or default_cta_group(), _stream())
    return AdapterActs(x=x, x_rows=x_rows, h1=h1, g1=g1, a=a, g2=g2, rowsq=rowsq, nblk=nblk, rows_cap=rows_cap,
                       n_rows=n_rows)


def adapter_backward(acts: AdapterActs, dz2: torch.Tensor, w1, w2, need_dx: bool = False, need_db2: bool = True,
                     cta_group: Optional[int] = None, need_db1: bool = True, accumulate: bool = False,
                     dw1_out: Optional[torch.Tensor] = None, dw2_out: Optional[torch.Tensor] = None,
                     dw_dtype=torch.bfloat16, mid_hook=None, overlap=None):
    """Returns (dw1, db1, dw2, db2_or_None, dx_or_None) in bf16, nn.Linear layout.

    `need_db1=False`: db1 is returned as the fp32 partial column sums [ceil(rows_cap/32)][d_mid] the dgrad GEMM's
    epilogue left behind (finish them with `bias_grads`).  `dw1_out`/`dw2_out`: write (or, with `accumulate`, add)
    the weight gradients into the caller's tensors instead of fresh ones.  `dw_dtype=torch.float32`: the weight
    gradients leave the GEMMs unrounded (what a gradient mean over ranks should carry).  `overlap`: a
    peer.PeerGradAllReduce channel holding this rank's dW2 / db2 contribution — its announce + reduce phases are then
    serviced by the idle epilogue warps inside the dW1 GEMM's launch; `mid_hook()` runs between the dW2 and the
    dW1 GEMM (the caller finishes db2 there)."""
    d_mid, d_in = w1.shape
    d_out = w2.shape[0]
    dev = dz2.device
    bf = torch.bfloat16
    dz1 = torch.empty(acts.rows_cap, d_mid, dtype=bf, device=dev)
    if accumulate and (dw1_out is None or dw2_out is None):
        raise _lib.P2TError("accumulate needs dw1_out and dw2_out")
    dw1 = dw1_out if dw1_out is not None else torch.empty(d_mid, d_in, dtype=dw_dtype, device=dev)
    dw2 = dw2_out if dw2_out is not None else torch.empty(d_out, d_mid, dtype=dw_dtype, device=dev)
    if dw1.dtype != dw2.dtype or dw1.dtype not in (torch.bfloat16, torch.float32):
        raise _lib.P2TError("weight-gradient outputs must both be bfloat16 or both float32")
    db1 = torch.empty(d_mid, dtype=bf, device=dev) if need_db1 else None
    db2 = torch.empty(d_out, dtype=bf, device=dev) if need_db2 else None
    dx = torch.empty(acts.rows_cap, d_in, dtype=bf, device=dev) if need_dx else None
    n32, n64 = (acts.rows_cap + 31) // 32, (acts.rows_cap + 63) // 64
    ws = torch.empty(n32 * d_mid + (n64 * d_out if need_db2 else 0), dtype=torch.float32, device=dev)
    gws = gemm_workspace(dev)

    def call(phases: int, ov):
        _lib.call("p2t_adapter_bwd", _ptr(acts.x), acts.x_rows, _ptr(w1), _ptr(w2), _ptr(acts.h1), _ptr(acts.g1),
                  _ptr(dz2), d_in, d_mid, d_out, acts.rows_cap, _ptr(acts.n_rows), _ptr(dz1), _ptr(dw1), _ptr(db1),
                  _ptr(dw2), _ptr(db2), _ptr(dx), _ptr(ws), _ptr(gws), int(accumulate),
                  int(dw1.dtype == torch.float32), phases, ov, cta_group or default_cta_group(), _stream())

    if mid_hook is None and overlap is None:
        call(0, None)
    else:
        # sharded training step: dgrad + dW2 first, then the caller finishes what the overlapped channel carries (db2),
        # then dW1, whose idle epilogue warps service that channel inside its launch
        call(1 | 2, None)
        if mid_hook is not None:
            mid_hook()
        ov = None
        if overlap is not None:
            ov = _lib.OverlapReduce(C.cast(overlap.buffer.table, C.POINTER(C.c_void_p)), overlap.world, overlap.rank, overlap.n_bytes, overlap.f32_from)
        call(4, ov)
    return dw1, (db1 if need_db1 else ws), dw2, db2, dx


def bias_grads(db1_partial: Optional[torch.Tensor], rows_cap: int, n_rows: Optional[torch.Tensor], d_mid: int,
               db2_partial: Optional[torch.Tensor], nparts2: Optional[torch.Tensor], d_out: int, *, accumulate: bool = False,
               out_f32: Optional[tuple] = None, out_bf16: Optional[tuple] = None):
    """(db1_bf16, db2_bf16, db1_f32, db2_f32) from the partial column sums of the dgrad GEMM's epilogue and of the
    tail backward, one launch.  Either job may be absent (partial None: its outputs come back None).  `out_f32` /
    `out_bf16`: (db1, db2) tensors to write into (fp32 is added to when `accumulate`)."""
    dev = (db1_partial if db1_partial is not None else db2_partial).device
    of = out_f32 if out_f32 is not None else (None, None)
    ob = out_bf16 if out_bf16 is not None else (None, None)
    f1 = b1 = f2 = b2 = None
    if db1_partial is not None:
        f1 = of[0] if of[0] is not None else torch.empty(d_mid, dtype=torch.float32, device=dev)
        b1 = ob[0] if ob[0] is not None else torch.empty(d_mid, dtype=torch.bfloat16, device=dev)
    if db2_partial is not None:
        f2 = of[1] if of[1] is not None else torch.empty(d_out, dtype=torch.float32, device=dev)
        b2 = ob[1] if ob[1] is not None else torch.empty(d_out, dtype=torch.bfloat16, device=dev)
    _lib.call("p2t_bias_grads", _ptr(db1_partial), rows_cap, _ptr(n_rows), d_mid, _ptr(b1), _ptr(f1), _ptr(db2_partial),
              _ptr(nparts2), db2_partial.shape[0] if db2_partial is not None else 0, d_out, _ptr(b2), _ptr(f2),
              int(accumulate), _stream())
    return b1, b2, f1, f2


# --------------------------------------------------------------------------------------------------
# pooling / normalisation
# --------------------------------------------------------------------------------------------------
def row_inv_norm(acts: "AdapterActs") -> torch.Tensor:
    """1 / max(|a_row|, 1e-12) per packed row, from the fc2 epilogue's partial sums of squares."""
    inv = torch.empty(acts.rows_cap, dtype=torch.float32, device=acts.a.device)
    _lib.call("p2t_row_inv_norm", _ptr(acts.rowsq), acts.nblk, _ptr(acts.n_rows), acts.rows_cap, _ptr(inv), _stream())
    return inv


def pool_forward(src: torch.Tensor, plan: RowPlan, D: int, *, row_src: Optional[torch.Tensor],
                 inv_norm: Optional[torch.Tensor] = None, normalize: bool = False, want_f32: bool = True):
    """(mean | std) statistics fp32 [B, 2D] of the plan's rows of `src` (bf16 or fp16, row stride src.stride(0)),
    each row optionally scaled by inv_norm[row] first.  With `normalize` the following F.normalize is fused in and
    (stats, p_bf16, p_f32 or None, norm) is returned."""
    dev = src.device
    partial = torch.empty(plan.max_chunks * (2 * D + 1), dtype=torch.float32, device=dev)  # records + per-chunk row counts
    stats = torch.empty(plan.B, 2 * D, dtype=torch.float32, device=dev)
    p_bf = p_f32 = norm = None
    if normalize:
        p_bf = torch.empty(plan.B, 2 * D, dtype=torch.bfloat16, device=dev)
        p_f32 = torch.empty(plan.B, 2 * D, dtype=torch.float32, device=dev) if want_f32 else None
        norm = torch.empty(plan.B, dtype=torch.float32, device=dev)
    _lib.call("p2t_pool_fwd", _ptr(src), int(src.dtype == torch.float16), src.stride(0), src.shape[0], _ptr(row_src), _ptr(inv_norm),
              _ptr(plan.seq_off), _ptr(plan.chunk_off), _ptr(plan.chunk_seq), plan.B, D, CHUNK_ROWS, plan.max_chunks,
              READOUT_MODES["mix"], _ptr(partial), _ptr(stats), 2 * D, _ptr(p_bf), _ptr(p_f32), _ptr(norm), _stream())
    if normalize:
        return stats, p_bf, p_f32, norm
    return stats


def loss_backward_coef(res: "InfoNCEResult", t_f32: torch.Tensor, p_f32: torch.Tensor, pnorm: torch.Tensor,
                       stats: torch.Tensor, plan: RowPlan, D: int, tau: float, dloss: Optional[torch.Tensor]):
    """dLogits -> pooling coefficients (c1, c2) in one kernel (small similarity blocks, fp32 embeddings)."""
    R, C = res.dS.shape
    c1 = torch.empty(plan.B, D, dtype=torch.float32, device=p_f32.device)
    c2 = torch.empty(plan.B, D, dtype=torch.float32, device=p_f32.device)
    ws = torch.empty(plan.B, 2 * D + (2 * D + 63) // 64, dtype=torch.float32, device=p_f32.device)
    _lib.call("p2t_loss_bwd_coef", _ptr(res.dS), _ptr(t_f32), _ptr(p_f32), _ptr(pnorm), _ptr(stats), _ptr(plan.seq_off),
              _ptr(dloss), R, plan.B, C, D, float(tau), _ptr(ws), _ptr(c1), _ptr(c2), _stream())
    return c1, c2


def l2norm_forward(e: torch.Tensor, want_f32: bool = True):
    B, E = e.shape
    p_bf = torch.empty(B, E, dtype=torch.bfloat16, device=e.device)
    p_f32 = torch.empty(B, E, dtype=torch.float32, device=e.device) if want_f32 else None
    norm = torch.empty(B, dtype=torch.float32, device=e.device)
    _lib.call("p2t_l2norm_fwd", _ptr(e), B, E, _ptr(p_bf), _ptr(p_f32), _ptr(norm), _stream())
    return p_bf, p_f32, norm


def to_bf16(x_f32: torch.Tensor) -> torch.Tensor:
    out = torch.empty(x_f32.shape, dtype=torch.bfloat16, device=x_f32.device)
    _lib.call("p2t_f32_to_bf16", _ptr(x_f32), x_f32.numel(), _ptr(out), _stream())
    return out


def l2norm_backward(dp: torch.Tensor, p_f32: torch.Tensor, norm: torch.Tensor) -> torch.Tensor:
    de = torch.empty_like(dp)
    B, E = dp.shape
    _lib.call("p2t_l2norm_bwd", _ptr(dp), _ptr(p_f32), _ptr(norm), B, E, _ptr(de), _stream())
    return de


def pool_backward_coef(de: torch.Tensor, stats: torch.Tensor, plan: RowPlan, D: int, mode: str):
    c1 = torch.empty(plan.B, D, dtype=torch.float32, device=de.device)
    c2 = torch.empty(plan.B, D, dtype=torch.float32, device=de.device)
    _lib.call("p2t_pool_bwd_coef", _ptr(de), de.stride(0), _ptr(stats), stats.stride(0), _ptr(plan.seq_off), plan.B, D,
              READOUT_MODES[mode], _ptr(c1), _ptr(c2), _stream())
    return c1, c2


_SM_COUNT = {}


def sm_count(device) -> int:
    idx = torch.device(device).index
    idx = torch.cuda.current_device() if idx is None else idx
    if idx not in _SM_COUNT:
        _SM_COUNT[idx] = torch.cuda.get_device_properties(idx).multi_processor_count
    return _SM_COUNT[idx]


def adapter_tail_backward(acts: AdapterActs, inv_norm: torch.Tensor, plan: RowPlan, c1, c2, finish_db2: bool = True):
    """dz2 (bf16 [rows_cap, d_out]) and db2 = colsum(dz2): as bf16 [d_out] (`finish_db2`), or as the tuple
    (partial fp32 [ws_rows, d_out], nparts int32 [1]) for `bias_grads`."""
    d_out = acts.a.shape[1]
    dev = acts.a.device
    dz2 = torch.empty(acts.rows_cap, d_out, dtype=torch.bfloat16, device=dev)
    ws_rows = 2 * sm_count(dev)
    ws = torch.empty(ws_rows, d_out, dtype=torch.float32, device=dev)
    nparts = torch.empty(1, dtype=torch.int32, device=dev)
    db2 = torch.empty(d_out, dtype=torch.bfloat16, device=dev) if finish_db2 else None
    _lib.call("p2t_adapter_tail_bwd", _ptr(acts.a), _ptr(acts.g2), _ptr(inv_norm), _ptr(plan.seq_off), plan.B,
              _ptr(c1), _ptr(c2), _ptr(acts.n_rows), acts.rows_cap, d_out, _ptr(dz2), _ptr(ws), ws_rows, _ptr(nparts),
              _ptr(db2), _stream())
    return dz2, (db2 if finish_db2 else (ws, nparts))


# --------------------------------------------------------------------------------------------------
# InfoNCE
# --------------------------------------------------------------------------------------------------
@dataclass
class InfoNCEResult:
    loss: torch.Tensor                 # fp32 0-dim
    dS: Optional[torch.Tensor]         # fp32 [R, C] (gradient of loss w.r.t. scaled logits), None if no grad
    dS_bf16: Optional[torch.Tensor]
    row_lse: torch.Tensor
    argmax_row: torch.Tensor           # int32 [R]
    argmax_col: Optional[torch.Tensor]  # int32 [C] (only when column statistics were computed)
    col_max: Optional[torch.Tensor]
    col_sum: Optional[torch.Tensor]


def infonce_forward(p_bf: torch.Tensor, t_bf: torch.Tensor, labels: torch.Tensor, tau: float, *, w_row: float = 1.0,
                    w_col: float = 0.0, need_grad: bool = True, want_col_argmax: bool = False,
                    p_f32: Optional[torch.Tensor] = None, t_f32: Optional[torch.Tensor] = None,
                    col_stats_hook=None, loss_scale: Optional[float] = None, all_cols_labelled: bool = False,
                    cta_group: Optional[int] = None) -> InfoNCEResult:
    """loss = mean_i [ w_row (lse_j S_ij - S_i,lab) + w_col (lse_col[lab] - S_i,lab) ], S = p t^T / tau.

    `col_stats_hook(col_max, col_sum) -> (col_max, col_sum)` lets the multi-GPU layer merge column
    statistics across ranks before the gradient pass.  `loss_scale` overrides 1/R.
    `all_cols_labelled`: every column has its positive on SOME rank (sharded global batch), so the
    column term's gradient flows to all local rows, not only to columns labelled by local rows.
    `p_f32`/`t_f32`: fp32 copies of the embeddings; when given, small problems (CUDA-core path) use
    them instead of the bf16 tensors (see infonce.cu: near-parallel embeddings make bf16 rounding of
    p/t the dominant gradient error).
    """
    require_cuda_bf16(p_bf, "p")
    if t_bf.dtype == torch.float32:  # small blocks with fp32 embeddings never touch a bf16 copy of t
        if t_f32 is None or p_f32 is None or p_bf.shape[0] * t_bf.shape[0] * t_bf.shape[1] > (1 << 26):
            raise _lib.P2TError("fp32 text embeddings need the fp32 protein embeddings and a small similarity block")
        t_bf = None
    else:
        require_cuda_bf16(t_bf, "t")
    R, E = p_bf.shape
    C = (t_bf if t_bf is not None else t_f32).shape[0]
    dev = p_bf.device
    cg = cta_group or default_cta_group()
    labels32 = labels.to(device=dev, dtype=torch.int32).contiguous()
    scale = (1.0 / R) if loss_scale is None else float(loss_scale)
    big = (R * C * E > (1 << 26)) and C % 8 == 0
    import os
    if big and t_bf is not None and os.environ.get("P2T_FUSED_LARGE", "1") != "0":
        # large block: similarity on the tensor cores with the online-softmax statistics reduced in the GEMM's epilogue,
        # then (for the gradient) S recomputed tile by tile into bf16 dLogits — no R x C fp32 tensor is ever allocated
        f32 = torch.float32
        npart = 4 * ((C + 255) // 256)
        want_cols = w_col != 0.0 or want_col_argmax
        row_part = torch.empty(npart * R * 4, dtype=f32, device=dev)
        col_part = torch.empty(((R + 31) // 32) * C * 4, dtype=f32, device=dev) if want_cols else None
        pos = torch.empty(R, dtype=f32, device=dev)
        col_max = torch.empty(C, dtype=f32, device=dev) if want_cols else None
        col_sum = torch.empty(C, dtype=f32, device=dev) if want_cols else None
        argmax_col = torch.empty(C, dtype=torch.int32, device=dev) if want_cols else None
        _lib.call("p2t_infonce_stats", _ptr(p_bf), _ptr(t_bf), _ptr(labels32), R, C, E, float(tau), _ptr(row_part),
                  _ptr(col_part), _ptr(pos), _ptr(col_max), _ptr(col_sum), _ptr(argmax_col), cg, _stream())
        if col_stats_hook is not None and want_cols:
            col_max, col_sum = col_stats_hook(col_max, col_sum)
            col_max, col_sum = col_max.contiguous(), col_sum.contiguous()
        row_loss = torch.empty(R, dtype=f32, device=dev)
        row_lse = torch.empty(R, dtype=f32, device=dev)
        argmax_row = torch.empty(R, dtype=torch.int32, device=dev)
        marks = torch.empty(C, dtype=torch.uint8, device=dev) if w_col != 0.0 else None
        col_lse = torch.empty(C, dtype=f32, device=dev) if w_col != 0.0 else None
        dS_bf16 = torch.empty(R, C, dtype=torch.bfloat16, device=dev) if need_grad else None
        _lib.call("p2t_infonce_finish", _ptr(p_bf), _ptr(t_bf), _ptr(labels32), R, C, E, float(tau), float(w_row), float(w_col),
                  scale, _ptr(row_part), _ptr(pos), _ptr(col_max), _ptr(col_sum), int(all_cols_labelled), _ptr(marks),
                  _ptr(col_lse), _ptr(row_loss), _ptr(row_lse), _ptr(argmax_row), _ptr(dS_bf16), cg, _stream())
        loss = torch.empty((), dtype=f32, device=dev)
        _lib.call("p2t_loss_mean", _ptr(row_loss), R, scale, _ptr(loss), 0, _stream())
        return InfoNCEResult(loss=loss, dS=None, dS_bf16=dS_bf16, row_lse=row_lse, argmax_row=argmax_row,
                             argmax_col=argmax_col, col_max=col_max, col_sum=col_sum)
    S = torch.empty(R, C, dtype=torch.float32, device=dev)
    _lib.call("p2t_similarity", _ptr(p_bf), _ptr(t_bf), _ptr(p_f32), _ptr(t_f32), R, C, E, float(tau), _ptr(S), cg,
              _stream())
    col_max = col_sum = argmax_col = marks = None
    if w_col != 0.0 or want_col_argmax:
        col_max = torch.empty(C, dtype=torch.float32, device=dev)
        col_sum = torch.empty(C, dtype=torch.float32, device=dev)
        argmax_col = torch.empty(C, dtype=torch.int32, device=dev)
        _lib.call("p2t_infonce_col_stats", _ptr(S), R, C, _ptr(col_max), _ptr(col_sum), _ptr(argmax_col), 0, _stream())
        if col_stats_hook is not None:
            col_max, col_sum = col_stats_hook(col_max, col_sum)
        marks = torch.empty(C, dtype=torch.uint8, device=dev)
    row_loss = torch.empty(R, dtype=torch.float32, device=dev)
    row_lse = torch.empty(R, dtype=torch.float32, device=dev)
    argmax_row = torch.empty(R, dtype=torch.int32, device=dev)
    dS_bf16 = torch.empty(R, C, dtype=torch.bfloat16, device=dev) if (need_grad and big) else None
    _lib.call("p2t_infonce_ce", _ptr(S), _ptr(labels32), R, C, float(w_row), float(w_col), scale, _ptr(col_max),
              _ptr(col_sum), _ptr(marks), int(all_cols_labelled), _ptr(row_loss), _ptr(row_lse), _ptr(argmax_row), _ptr(dS_bf16),
              int(need_grad), _stream())
    loss = torch.empty((), dtype=torch.float32, device=dev)
    _lib.call("p2t_loss_mean", _ptr(row_loss), R, scale, _ptr(loss), 0, _stream())
    return InfoNCEResult(loss=loss, dS=S if need_grad else None, dS_bf16=dS_bf16, row_lse=row_lse,
                         argmax_row=argmax_row, argmax_col=argmax_col, col_max=col_max, col_sum=col_sum)


def infonce_backward(res: InfoNCEResult, p_bf, t_bf, tau: float, need_dp: bool = True, need_dt: bool = False,
                     p_f32: Optional[torch.Tensor] = None, t_f32: Optional[torch.Tensor] = None,
                     cta_group: Optional[int] = None):
    R, E = p_bf.shape
    C = (t_bf if t_bf is not None else t_f32).shape[0]
    dev = p_bf.device
    dp = torch.empty(R, E, dtype=torch.float32, device=dev) if need_dp else None
    dt = torch.empty(C, E, dtype=torch.float32, device=dev) if need_dt else None
    _lib.call("p2t_infonce_grad", _ptr(res.dS), _ptr(res.dS_bf16), _ptr(p_bf), _ptr(t_bf), _ptr(p_f32), _ptr(t_f32),
              R, C, E, float(tau), _ptr(dp), _ptr(dt), cta_group or default_cta_group(), _stream())
    return dp, dt


_BARRIER_WS = {}


def _barrier_ws(device) -> torch.Tensor:
    """16 zero bytes per device for the fused loss kernel's grid barrier (the kernel re-arms them itself)."""
    key = str(device)
    t = _BARRIER_WS.get(key)
    if t is None:
        t = _BARRIER_WS[key] = torch.zeros(4, dtype=torch.int32, device=device)
    return t


def loss_fused_eligible(R: int, B: int, C: int, E: int) -> bool:
    return bool(_lib.load().p2t_loss_fused_eligible(R, B, C, E))


@dataclass
class FusedLossResult:
    loss: torch.Tensor
    row_lse: torch.Tensor
    argmax_row: torch.Tensor
    argmax_col: torch.Tensor
    col_max: torch.Tensor
    col_sum: torch.Tensor
    c1: Optional[torch.Tensor]
    c2: Optional[torch.Tensor]


def loss_fused(p_f32: torch.Tensor, t_f32: Optional[torch.Tensor], labels: torch.Tensor, R: int, tau: float, *,
               w_row: float = 1.0, w_col: float = 0.0, loss_scale: Optional[float] = None,
               all_cols_labelled: bool = False, need_grad: bool = True, dloss: Optional[torch.Tensor] = None,
               pnorm: Optional[torch.Tensor] = None, stats: Optional[torch.Tensor] = None,
               seq_off: Optional[torch.Tensor] = None, gather=None) -> FusedLossResult:
    """Similarity -> InfoNCE (-> pooling coefficients of the backward pass) in one cooperative kernel
    (csrc/loss_fused.cu).  `gather`: a peer.PeerAllGather whose push has been issued — the kernel waits for the round
    and reads the gathered text embeddings in place (then `t_f32` is None)."""
    B, E = p_f32.shape
    dev = p_f32.device
    if gather is not None:
        C = gather.world * gather.rows
        peers, world, rank, bpr = gather.buffer.table, gather.world, gather.rank, gather.bytes_per_rank
        if gather.cols != E or gather.dtype != torch.float32:
            raise _lib.P2TError("the gather channel must carry fp32 rows of the embedding width")
    else:
        C = t_f32.shape[0]
        peers, world, rank, bpr = None, 0, 0, 0
    labels32 = labels.to(device=dev, dtype=torch.int32).contiguous()
    D = E // 2
    f32 = torch.float32
    S_ws = torch.empty(R, C, dtype=f32, device=dev)
    nslice = (E + 63) // 64
    dp_ws = torch.empty(B, E + nslice, dtype=f32, device=dev) if need_grad else None
    loss = torch.empty((), dtype=f32, device=dev)
    row_lse = torch.empty(R, dtype=f32, device=dev)
    argmax_row = torch.empty(R, dtype=torch.int32, device=dev)
    argmax_col = torch.empty(C, dtype=torch.int32, device=dev)
    col_max = torch.empty(C, dtype=f32, device=dev)
    col_sum = torch.empty(C, dtype=f32, device=dev)
    c1 = torch.empty(B, D, dtype=f32, device=dev) if need_grad else None
    c2 = torch.empty(B, D, dtype=f32, device=dev) if need_grad else None
    scale = (1.0 / R) if loss_scale is None else float(loss_scale)
    _lib.call("p2t_loss_fused", _ptr(p_f32), _ptr(t_f32), peers, world, rank, bpr, _ptr(labels32), R, B, C, E,
              float(tau), float(w_row), float(w_col), scale, int(all_cols_labelled), 1, int(need_grad), _ptr(dloss),
              _ptr(pnorm), _ptr(stats), _ptr(seq_off), _ptr(S_ws), _ptr(dp_ws), _ptr(_barrier_ws(dev)), _ptr(loss),
              _ptr(row_lse), _ptr(argmax_row), _ptr(argmax_col), _ptr(col_max), _ptr(col_sum), _ptr(c1), _ptr(c2),
              _stream())
    if gather is not None:
        gather.note_arrived()
    return FusedLossResult(loss, row_lse, argmax_row, argmax_col, col_max, col_sum, c1, c2)


def dropout_mask(rows: int, cols: int, p: float, seed: int, layer: int, device) -> torch.Tensor:
    """The keep multipliers (0 or 1/(1-p)) the kernels apply for (seed, layer) — test aid."""
    out = torch.empty(rows, cols, dtype=torch.float32, device=device)
    _lib.call("p2t_dropout_mask", rows, cols, float(p), int(seed) & 0xFFFFFFFFFFFFFFFF, layer, _ptr(out), _stream())
    return out
