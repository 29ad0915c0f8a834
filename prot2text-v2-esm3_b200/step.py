"""The fused Stage-1 contrastive step: residue states + text hidden states -> loss, adapter grads.

Mirrors `teacher_forcing_forward_pass` of the reference (scripts/train_contrast.py:313-379) from
the point where the two frozen trunks have produced their outputs:

    text branch   : hidden_states[16] (:304) -> readout 'mix' (:306-310) -> F.normalize (:354), no grad
    protein branch: residue states -> ModalityAdapter (models/esmc_qwen_arc.py:182) -> readout 'mix'
                    (:277-281) -> F.normalize (:365)
    loss          : SegmentedBatchInfoNCELoss per segment, averaged (:356-379)

Differences in HOW (not what): the padded (B, L, D_in) batch is packed to its valid rows on the
device (no host sync), both adapter GEMMs run once over all rows, and the handoffs between stages
stay in fp32.  With equal segments the average of the segment means IS the batch mean, so the
segment loop itself costs nothing to drop.  What the loop DOES change in this fork is the pooling
mask: `get_sequence_embeddings` (:251-281) encodes every segment on its own, padded
to the segment's longest sequence, and pools with an all-ones mask over that length — pad rows
included.  The mask is an input of this module; `segment_pooling_mask` builds the fork's mask from
the sequence lengths, `protein_mask` = the true attention mask pools valid residues only.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import os

import torch

from . import _core
from .adapter import ModalityAdapter, _draw_seed


@dataclass
class StepAux:
    """Side outputs of the last contrastive step (no gradient)."""
    protein_embeddings: Optional[torch.Tensor] = None   # (B, 2*D_out) bf16, unit norm
    text_embeddings: Optional[torch.Tensor] = None      # (B_global, 2*H) unit norm (fp32, or bf16 on the tensor-core path)
    argmax_row: Optional[torch.Tensor] = None           # int32 (B,)  protein -> text retrieval
    argmax_col: Optional[torch.Tensor] = None           # int32 (B_global,) text -> protein retrieval (local rows)
    n_rows: Optional[torch.Tensor] = None               # int32 (1,) valid residue rows
    # filled by backward: the bias gradients in fp32, BEFORE their rounding to the bf16 of param.grad.  A gradient mean
    # over ranks should carry these (peer.PeerGradAllReduce.for_adapter): per-rank gradients can be several times larger
    # than their mean, so rounding each rank's value to bf16 first costs several times bf16's 2^-9 in the mean.
    bias_grads_f32: Optional[tuple] = None


@torch.no_grad()
def text_embeddings(text_hidden: torch.Tensor, text_mask: Optional[torch.Tensor] = None, dtype=torch.bfloat16, *,
                    text_lengths: Optional[torch.Tensor] = None) -> torch.Tensor:
    """readout 'mix' + L2 normalise of the frozen LLM's hidden states (reference :284-310, :354).

    text_hidden is (B, T, H) with a (B, T) mask, or — ragged hand-over — packed rows (sum T_b, H) with
    `text_lengths` (B,).  Returns (B, 2H) unit-norm embeddings in `dtype`: bfloat16 like the reference, or
    float32 — the fused step keeps the fp32 copy for the loss gradient (see csrc/infonce.cu)."""
    _core.require_cuda_bf16(text_hidden, "text_hidden")
    th = text_hidden.contiguous()
    if text_lengths is not None:
        if th.dim() != 2:
            raise ValueError("with text_lengths, text_hidden must be packed rows (sum T_b, H)")
        H = th.shape[1]
        plan = _core.plan_packed(text_lengths, th.shape[0])
        src, row_src = th, None
    else:
        B, T, H = th.shape
        plan = _core.plan_rows(text_mask)
        src, row_src = th.view(B * T, H), plan.row_src
    want_f32 = dtype == torch.float32
    _, t_bf, t_f32, _ = _core.pool_forward(src, plan, H, row_src=row_src, normalize=True, want_f32=want_f32)
    return t_f32 if want_f32 else t_bf


_SIDE_STREAMS = {}


def _side_stream(device) -> torch.cuda.Stream:
    """One extra stream per device for the text branch of the step (it does not depend on the protein side until the
    similarity, so its launch-latency-bound kernels run beside the residue-row packing instead of in front of it)."""
    key = str(device)
    st = _SIDE_STREAMS.get(key)
    if st is None:
        st = _SIDE_STREAMS[key] = torch.cuda.Stream(device=device)
    return st


def _is_gather(t_in) -> bool:
    return hasattr(t_in, "note_arrived") and hasattr(t_in, "buffer")


def step_forward(x, prot_mask, w1, b1, w2, b2, t_in, labels, cfg: dict, aux: StepAux, need_grad: bool):
    """Forward of the fused step on raw tensors (no autograd): returns (loss, state for step_backward or None).

    `t_in`: the unit-norm text embeddings (fp32 or bf16 tensor), a zero-argument callable returning them, or a
    `peer.PeerAllGather` whose push half has been issued (sharded step): the loss kernel then waits for the gathered
    rows itself and reads them in place."""
    for t, name in ((x, "residue states"), (w1, "fc1.weight"), (b1, "fc1.bias"), (w2, "fc2.weight"), (b2, "fc2.bias")):
        _core.require_cuda_bf16(t, name)
    d_out = w2.shape[0]
    w1c, b1c, w2c, b2c = w1.contiguous(), b1.contiguous(), w2.contiguous(), b2.contiguous()
    if cfg["packed"]:  # residue rows already packed, prot_mask holds the per-sequence lengths
        xp = x.contiguous()
        plan = _core.plan_packed(prot_mask, xp.shape[0])
        x_rows = xp.shape[0]
    else:
        B, L, d_in = x.shape
        plan = _core.plan_rows(prot_mask, max_valid_rows=cfg.get("max_valid_rows"))
        xp = _core.gather_rows(x.contiguous().view(B * L, d_in), plan)
        x_rows = plan.rows_cap
    join = cfg.get("text_join")
    if join is not None:
        # the text branch ran on the side stream beside the plan / pack kernels above; it must be OVER before the
        # persistent GEMMs start — they assume one CTA per SM, and another kernel still holding SMs delays a whole
        # tile share (measured with an NCCL kernel: fc1 163 -> 250 us)
        join()
    if callable(t_in) and not _is_gather(t_in) and not cfg.get("late_text", False):
        t_in = t_in()  # deferred text embeddings (NCCL all-gather of the sharded step), same reason
    acts = _core.adapter_forward(xp, x_rows, plan.rows_cap, plan.n_rows, w1c, b1c, w2c, b2c,
                                 cfg["dropout_p"], cfg["seed"], need_grad, seed_dev=cfg.get("seed_dev"))
    inv_norm = _core.row_inv_norm(acts)
    stats, p_bf, p_f32, pnorm = _core.pool_forward(acts.a, plan, d_out, row_src=None, inv_norm=inv_norm, normalize=True)
    used = cfg["rows_used"]
    hook = cfg.get("col_stats_hook")
    E = 2 * d_out
    fused = None
    if hook is None and os.environ.get("P2T_FUSED_LOSS", "1") != "0":
        if _is_gather(t_in):
            if _core.loss_fused_eligible(used, plan.B, t_in.world * t_in.rows, E):
                fused = dict(t_f32=None, gather=t_in)
        elif torch.is_tensor(t_in) and t_in.dtype == torch.float32 and _core.loss_fused_eligible(used, plan.B, t_in.shape[0], E):
            fused = dict(t_f32=t_in.contiguous(), gather=None)
    if fused is not None:
        # similarity -> online-softmax CE -> dLogits -> pooling coefficients: one cooperative kernel
        res = _core.loss_fused(p_f32, fused["t_f32"], labels[:used], used, cfg["tau"], w_row=cfg["w_row"], w_col=cfg["w_col"],
                               loss_scale=cfg.get("loss_scale"), all_cols_labelled=cfg.get("all_cols_labelled", False),
                               need_grad=need_grad, dloss=cfg.get("dloss_dev"), pnorm=pnorm, stats=stats,
                               seq_off=plan.seq_off, gather=fused["gather"])
        aux.protein_embeddings, aux.text_embeddings = p_bf, fused["t_f32"]
        aux.argmax_row, aux.argmax_col, aux.n_rows = res.argmax_row, res.argmax_col, plan.n_rows
        state = (plan, acts, inv_norm, w1c, w2c, cfg, ("coef", res.c1, res.c2)) if need_grad else None
        return res.loss, state
    if _is_gather(t_in):
        t_in = t_in.arrive()
    elif callable(t_in):
        # peer-memory exchange: its arrive kernel runs IN this stream (it never overlaps the GEMMs), so it is placed
        # where the gathered rows are first needed.
        t_in = t_in()
    if t_in.dtype == torch.float32:
        t_f32 = t_in.contiguous()
        # the bf16 copy is an operand of the tensor-core loss path only (large similarity blocks)
        big = used * t_f32.shape[0] * t_f32.shape[1] > (1 << 26)
        t_bf = _core.to_bf16(t_f32) if big else None
    else:
        _core.require_cuda_bf16(t_in, "text_embeds")
        t_f32, t_bf = None, t_in.contiguous()
    res = _core.infonce_forward(p_bf[:used], t_bf if t_bf is not None else t_f32, labels[:used], cfg["tau"],
                                w_row=cfg["w_row"], w_col=cfg["w_col"],
                                need_grad=need_grad, want_col_argmax=True, col_stats_hook=hook,
                                p_f32=p_f32[:used] if t_f32 is not None else None, t_f32=t_f32,
                                loss_scale=cfg.get("loss_scale"),
                                all_cols_labelled=cfg.get("all_cols_labelled", False))
    aux.protein_embeddings, aux.text_embeddings = p_bf, (t_bf if t_bf is not None else t_f32)
    aux.argmax_row, aux.argmax_col, aux.n_rows = res.argmax_row, res.argmax_col, plan.n_rows
    state = (plan, acts, inv_norm, w1c, w2c, cfg, ("logits", stats, p_bf, p_f32, pnorm, res, t_bf, t_f32)) if need_grad else None
    return res.loss, state


def step_backward(state, dloss: Optional[torch.Tensor], *, accumulate: bool = False, dw_out: Optional[tuple] = None,
                  db_f32_out: Optional[tuple] = None, db_bf16_out: Optional[tuple] = None, dw_f32: bool = False,
                  overlap_reduce=None):
    """Backward of the fused step: (dW1, db1, dW2, db2, db1_f32, db2_f32) — bf16 gradients in nn.Linear layout plus
    the bias gradients in fp32 (what a gradient all-reduce should carry) — for the upstream gradient `dloss` (device
    scalar; None = 1).  `dw_f32`: dW1/dW2 leave the GEMMs in fp32 instead (sharded training: the gradient mean over
    ranks then rounds once, after the mean).  `dw_out` = (dW1, dW2) / `db_f32_out` / `db_bf16_out`: write into the caller's tensors;
    `accumulate`: add to them (bf16 read-modify-write for the weights, fp32 for the biases) — the micro-batch
    accumulation of scripts/train_contrast.py:448-465.  `overlap_reduce`: a peer.PeerGradAllReduce channel over
    [dW2, db2] whose contribution area `dw_out[1]` / `db_f32_out[1]` alias: db2 is finished right after the dW2 GEMM and
    the channel's announce + reduce phases run on the idle epilogue warps inside the dW1 GEMM's launch (DDP's bucket overlap,
    scripts/train_contrast.py:448 + :611-614, as one fused GEMM + collective kernel)."""
    plan, acts, inv_norm, w1c, w2c, cfg, head = state
    used = cfg["rows_used"]
    d_out = w2c.shape[0]
    d_mid = w1c.shape[0]
    dl = None if dloss is None else dloss.to(torch.float32).contiguous()
    if head[0] == "coef":
        _, c1, c2 = head
        if dl is not None:  # the coefficients were formed for an upstream gradient of 1 (or cfg['dloss_dev']): linear in it
            c1, c2 = c1 * dl, c2 * dl
    else:
        _, stats, p_bf, p_f32, pnorm, res, t_bf, t_f32 = head
        if cfg.get("dloss_dev") is not None:
            dl = cfg["dloss_dev"] if dl is None else dl * cfg["dloss_dev"]
        if t_f32 is not None and res.dS_bf16 is None:
            # small similarity block with fp32 embeddings: dLogits -> (c1, c2) in two kernels
            c1, c2 = _core.loss_backward_coef(res, t_f32, p_f32, pnorm, stats, plan, d_out, cfg["tau"], dl)
        else:
            dp_used, _ = _core.infonce_backward(res, p_bf[:used], t_bf, cfg["tau"], need_dp=True, need_dt=False,
                                                p_f32=p_f32[:used] if t_f32 is not None else None, t_f32=t_f32)
            if used == p_bf.shape[0]:
                dp = dp_used
            else:  # rows dropped by the segment split get no gradient (reference :337, :357-359)
                dp = torch.zeros_like(p_f32)
                dp[:used] = dp_used
            if dl is not None:
                dp.mul_(dl)
            de = _core.l2norm_backward(dp, p_f32, pnorm)
            c1, c2 = _core.pool_backward_coef(de, stats, plan, d_out, "mix")
    dz2, (ws2, nparts2) = _core.adapter_tail_backward(acts, inv_norm, plan, c1, c2, finish_db2=False)
    dw1_out, dw2_out = dw_out if dw_out is not None else (None, None)
    of = db_f32_out if db_f32_out is not None else (None, None)
    ob = db_bf16_out if db_bf16_out is not None else (None, None)
    if overlap_reduce is None:
        dw1, ws1, dw2, _, _ = _core.adapter_backward(acts, dz2, w1c, w2c, need_db2=False, need_db1=False,
                                                     accumulate=accumulate, dw1_out=dw1_out, dw2_out=dw2_out,
                                                     dw_dtype=torch.float32 if dw_f32 else torch.bfloat16)
        db1, db2, db1_f32, db2_f32 = _core.bias_grads(ws1, acts.rows_cap, acts.n_rows, d_mid, ws2, nparts2, d_out,
                                                      accumulate=accumulate, out_f32=db_f32_out, out_bf16=db_bf16_out)
        return dw1, db1, dw2, db2, db1_f32, db2_f32
    late = {}

    def finish_db2():  # between the dW2 and the dW1 GEMM: the overlapped channel's contribution must be complete
        _, late["b"], _, late["f"] = _core.bias_grads(None, acts.rows_cap, None, d_mid, ws2, nparts2, d_out,
                                                      accumulate=accumulate, out_f32=(None, of[1]), out_bf16=(None, ob[1]))

    dw1, ws1, dw2, _, _ = _core.adapter_backward(acts, dz2, w1c, w2c, need_db2=False, need_db1=False,
                                                 accumulate=accumulate, dw1_out=dw1_out, dw2_out=dw2_out,
                                                 dw_dtype=torch.float32, mid_hook=finish_db2, overlap=overlap_reduce)
    db1, _, db1_f32, _ = _core.bias_grads(ws1, acts.rows_cap, acts.n_rows, d_mid, None, None, d_out,
                                          accumulate=accumulate, out_f32=(of[0], None), out_bf16=(ob[0], None))
    db2, db2_f32 = late["b"], late["f"]
    return dw1, db1, dw2, db2, db1_f32, db2_f32


class _ContrastiveStepFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, prot_mask, w1, b1, w2, b2, t_in, labels, cfg: dict, aux: StepAux):
        ctx.aux = aux
        loss, ctx.state = step_forward(x, prot_mask, w1, b1, w2, b2, t_in, labels, cfg, aux,
                                       need_grad=any(ctx.needs_input_grad[2:6]))
        return loss

    @staticmethod
    def backward(ctx, dloss):
        state, ctx.state = ctx.state, None
        dw1, db1, dw2, db2, db1_f32, db2_f32 = step_backward(state, dloss)
        ctx.aux.bias_grads_f32 = (db1_f32, db2_f32)
        return None, None, dw1, db1, dw2, db2, None, None, None, None


_LABELS = {}


def _default_labels(n: int, device) -> torch.Tensor:
    """arange(n) int32 on `device`, created once per (n, device): the diagonal pairing of the reference (:88-89)."""
    key = (n, str(device))
    t = _LABELS.get(key)
    if t is None:
        t = _LABELS[key] = torch.arange(n, device=device, dtype=torch.int32)
    return t


def _rank_labels(rank: int, n: int, device) -> torch.Tensor:
    """arange(rank*n, (rank+1)*n) int32: rank `rank`'s rows of the global pairing (sharded step)."""
    key = (rank, n, str(device))
    t = _LABELS.get(key)
    if t is None:
        t = _LABELS[key] = torch.arange(rank * n, (rank + 1) * n, device=device, dtype=torch.int32)
    return t


def _text_branch(text_hidden, text_mask, text_lengths, after=None):
    """Run `text_embeddings` (fp32) — and `after(t)` if given — on the device's side stream; returns (embeddings,
    join) where join() makes the current stream wait for the branch."""
    dev = text_hidden.device
    side, cur = _side_stream(dev), torch.cuda.current_stream(dev)
    side.wait_stream(cur)
    with torch.cuda.stream(side):
        t = text_embeddings(text_hidden, text_mask, dtype=torch.float32, text_lengths=text_lengths)
        if after is not None:
            after(t)

    def join():
        now = torch.cuda.current_stream(dev)
        now.wait_stream(side)
        t.record_stream(now)  # allocated on the side stream, consumed on this one
    return t, join


def segment_pooling_mask(lengths, num_segments: int, total_len: Optional[int] = None, device=None) -> torch.Tensor:
    """The protein pooling mask the fork's step EFFECTIVELY uses (scripts/train_contrast.py:251-281 with
    :356-377): segment s = rows [s*seg, (s+1)*seg) is encoded on its own, zero-padded to ITS longest sequence, and
    pooled with an all-ones mask over that length.  Row b of the result is therefore 1 on [0, max length of b's
    segment) — pad rows of shorter sequences included.  Rows beyond the last whole segment (dropped by the loss,
    :337) keep their own length.  `lengths`: (B,) ints; returns int64 (B, total_len or max(lengths))."""
    lens = [int(v) for v in (lengths.tolist() if torch.is_tensor(lengths) else lengths)]
    B = len(lens)
    seg = B // num_segments
    L = int(total_len) if total_len is not None else max(lens)
    mask = torch.zeros(B, L, dtype=torch.long)
    for b in range(B):
        if seg > 0 and b < seg * num_segments:
            s = b // seg
            n = max(lens[s * seg:(s + 1) * seg])
        else:
            n = lens[b]
        mask[b, :min(n, L)] = 1
    return mask.to(device) if device is not None else mask


def contrastive_step(residue_states: torch.Tensor, protein_mask: Optional[torch.Tensor], adapter: ModalityAdapter,
                     text_hidden: Optional[torch.Tensor] = None, text_mask: Optional[torch.Tensor] = None, *,
                     residue_lengths: Optional[torch.Tensor] = None, text_lengths: Optional[torch.Tensor] = None,
                     text_embeds: Optional[torch.Tensor] = None, temperature: float = 0.05,
                     contrastive_num_segments: int = 1, symmetric: bool = False,
                     labels: Optional[torch.Tensor] = None, aux: Optional[StepAux] = None,
                     col_stats_hook=None, loss_scale: Optional[float] = None,
                     all_cols_labelled: bool = False, seed_dev: Optional[torch.Tensor] = None,
                     late_text: bool = False, dloss_dev: Optional[torch.Tensor] = None,
                     max_valid_rows: Optional[int] = None, _text_join=None,
                     _raw: bool = False) -> torch.Tensor:
    """One Stage-1 step from trunk outputs to the (differentiable) fp32 loss.

    residue_states (B, L, D_in) bf16 and protein_mask (B, L) come from the frozen protein encoder
    (models/esmc_qwen_arc.py:84-86); text_hidden (B_t, T, H) bf16 is hidden_states[16] of the frozen
    LLM with its attention mask (scripts/train_contrast.py:304), or pass already normalised
    `text_embeds` (B_t, 2H), float32 (preferred) or bfloat16 — e.g. the all-gathered global negatives — or a
    zero-argument callable returning them, which is invoked after the residue rows have been packed and before
    the adapter GEMMs (so a collective on another stream overlaps the plan/pack kernels but never shares SMs with
    the persistent GEMMs) or, with `late_text`, right before the similarity (in-stream peer-memory arrive).  `labels[i]` is the text row
    paired with protein i (default: i).  Ragged hand-over: with `residue_lengths` (B,),
    `residue_states` is the PACKED (sum L_b, D_in) row buffer and `protein_mask` is ignored; likewise
    `text_lengths` with packed `text_hidden` (see host_io.HostStager).  `contrastive_num_segments` reproduces the reference's
    segment averaging including its dropping of the remainder rows; `symmetric` adds the
    text->protein term.  `seed_dev` (int64 device tensor, 1 element): dropout seed read on the device instead
    of drawn from torch's CPU generator (CUDA-graph replays, see graph.GraphedContrastiveStep).  `dloss_dev` (fp32
    device scalar): constant factor on the gradient, e.g. 1 / gradient_accumulation_steps (the reference divides the
    loss by it before backward, scripts/train_contrast.py:432); the returned loss is not scaled.
    `max_valid_rows`: host-side upper bound on sum(protein_mask) (the collater knows the sequence lengths): packed
    activations are sized for it instead of for B*L.
    """
    text_join = _text_join
    if text_embeds is None:
        if text_hidden is None or (text_mask is None and text_lengths is None):
            raise ValueError("pass either text_hidden + text_mask (or text_lengths) or text_embeds")
        if os.environ.get("P2T_TEXT_STREAM", "1") != "0" and text_hidden.is_cuda:
            # the text branch (plan, pool, normalise: no dependency on the protein side before the similarity) runs on
            # a side stream beside the packing of the residue rows and joins in front of the adapter GEMMs
            text_embeds, text_join = _text_branch(text_hidden, text_mask, text_lengths)
        else:
            text_embeds = text_embeddings(text_hidden, text_mask, dtype=torch.float32, text_lengths=text_lengths)
    packed = residue_lengths is not None
    if packed and residue_states.dim() != 2:
        raise ValueError("with residue_lengths, residue_states must be packed rows (sum L_b, D_in)")
    if not packed and protein_mask is None:
        raise ValueError("pass protein_mask, or residue_lengths with packed rows")
    B = residue_lengths.shape[0] if packed else residue_states.shape[0]
    seg = B // contrastive_num_segments
    if seg * contrastive_num_segments != B:
        print("WARNING: Given batch size is not divisible by the number of segments for contrastive learning.")
    if labels is None:
        labels = _default_labels(B, residue_states.device)
    p = adapter.dropout_p()
    cfg = dict(tau=float(temperature), w_row=0.5 if symmetric else 1.0, w_col=0.5 if symmetric else 0.0,
               dropout_p=p, seed=_draw_seed() if (p > 0 and seed_dev is None) else 0, seed_dev=seed_dev,
               rows_used=seg * contrastive_num_segments, packed=packed,
               col_stats_hook=col_stats_hook, loss_scale=loss_scale, all_cols_labelled=all_cols_labelled,
               late_text=late_text, text_join=text_join, dloss_dev=dloss_dev, max_valid_rows=max_valid_rows)
    aux = aux if aux is not None else StepAux()
    if _raw:  # graph capture: no autograd, the caller runs step_backward itself
        return step_forward(residue_states, residue_lengths if packed else protein_mask, adapter.fc1.weight.detach(),
                            adapter.fc1.bias.detach(), adapter.fc2.weight.detach(), adapter.fc2.bias.detach(),
                            text_embeds, labels, cfg, aux, need_grad=True)
    return _ContrastiveStepFunction.apply(residue_states, residue_lengths if packed else protein_mask,
                                          adapter.fc1.weight, adapter.fc1.bias,
                                          adapter.fc2.weight, adapter.fc2.bias, text_embeds, labels, cfg, aux)
