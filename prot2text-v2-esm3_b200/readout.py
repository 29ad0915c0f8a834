"""Drop-in `readout_embeddings` (reference: scripts/train_contrast.py:198-248)."""
from __future__ import annotations

from typing import Literal

import torch

from . import _core, _lib


class _ReadoutFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, embeddings, attention_mask, readout_fn: str):
        _core.require_cuda_bf16(embeddings, "embeddings")
        B, S, D = embeddings.shape
        emb = embeddings.contiguous()
        plan = _core.plan_rows(attention_mask)
        stats = _core.pool_forward(emb.view(B * S, D), plan, D, row_src=plan.row_src)
        if readout_fn == "mean":
            out = stats[:, :D]
        elif readout_fn == "std":
            out = stats[:, D:]
        else:
            out = stats
        if ctx.needs_input_grad[0]:
            ctx.plan, ctx.stats, ctx.mode = plan, stats, readout_fn
            ctx.save_for_backward(emb, attention_mask)
        return out.to(embeddings.dtype)

    @staticmethod
    def backward(ctx, dout):
        emb, mask = ctx.saved_tensors
        B, S, D = emb.shape
        de = dout.to(torch.float32).contiguous()
        c1, c2 = _core.pool_backward_coef(de, ctx.stats, ctx.plan, D, ctx.mode)
        dx = torch.empty_like(emb)
        m = mask.contiguous()
        if m.dtype == torch.bool:
            m = m.view(torch.uint8)
        if m.dtype.is_floating_point or m.element_size() not in (1, 4, 8):
            m = m.to(torch.int32)
        _lib.call("p2t_readout_bwd", _core._ptr(emb), _core._ptr(m), m.element_size(), B, S, D, _core._ptr(c1),
                  _core._ptr(c2), _core._ptr(dx), _core._stream())
        return dx, None, None


class _ReadoutLastFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, embeddings, attention_mask):
        _core.require_cuda_bf16(embeddings, "embeddings")
        B, S, D = embeddings.shape
        emb = embeddings.contiguous()
        plan = _core.plan_rows(attention_mask, want_row_src=False)
        out = torch.empty(B, D, dtype=torch.float32, device=emb.device)
        _lib.call("p2t_readout_last", _core._ptr(emb), _core._ptr(plan.counts), B, S, D, _core._ptr(out), _core._stream())
        ctx.save_for_backward(plan.counts)
        ctx.shape = (B, S, D)
        return out.to(embeddings.dtype)

    @staticmethod
    def backward(ctx, dout):
        (counts,) = ctx.saved_tensors
        B, S, D = ctx.shape
        dout = dout.contiguous()
        dx = torch.empty(B, S, D, dtype=dout.dtype, device=dout.device)
        _lib.call("p2t_readout_last_bwd", _core._ptr(dout), _core._ptr(counts), B, S, D, _core._ptr(dx), _core._stream())
        return dx, None


def readout_embeddings(
        embeddings: torch.Tensor,  # (bsz, seq_len, hidden_dim)
        attention_mask: torch.Tensor,  # (bsz, seq_len) of 0/1
        readout_fn: Literal["last", "mean", "std", "mix"],
) -> torch.Tensor:
    """Masked readout of a sequence of embeddings; same contract as the reference function.

    'mean', 'std' (population, no eps) and 'mix' = cat(mean, std) run as ONE pass over the valid
    rows (the reference makes ~13); any 0/1 mask pattern is accepted.  'last' takes the row at
    index sum(mask)-1 (right padding, as documented by the reference :208-209).
    """
    if readout_fn == "last":
        return _ReadoutLastFunction.apply(embeddings, attention_mask)
    if readout_fn not in ("mean", "std", "mix"):
        raise ValueError(f"unknown readout_fn {readout_fn!r}")
    return _ReadoutFunction.apply(embeddings, attention_mask, readout_fn)
