"""Drop-in InfoNCE loss modules (reference: scripts/train_contrast.py:72-114)."""
from __future__ import annotations

import torch

from . import _core


def _as_bf16(t: torch.Tensor) -> torch.Tensor:
    if not t.is_cuda:
        raise _core._lib.P2TError("InfoNCE inputs must be CUDA tensors: this package has no CPU path")
    return t.detach().to(torch.bfloat16).contiguous()


class _InfoNCEFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, out1, out2, labels, temperature: float, w_row: float, w_col: float):
        if out1.dim() != 2 or out2.dim() != 2 or out1.shape[1] != out2.shape[1] or out1.shape[1] % 8:
            raise _core._lib.P2TError("InfoNCE inputs must be (rows, E) and (cols, E) with E a multiple of 8")
        if labels.numel() != out1.shape[0]:
            raise _core._lib.P2TError("one label per row of the first input")
        p_bf, t_bf = _as_bf16(out1), _as_bf16(out2)
        # fp32 inputs keep their precision on the CUDA-core path (small / medium blocks): near-parallel embeddings make
        # the bf16 rounding of p and t the dominant gradient error there
        p_f32 = t_f32 = None
        if out1.dtype == torch.float32 and out2.dtype == torch.float32:
            p_f32, t_f32 = out1.detach().contiguous(), out2.detach().contiguous()
        need = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        res = _core.infonce_forward(p_bf, t_bf, labels, temperature, w_row=w_row, w_col=w_col, need_grad=need,
                                    p_f32=p_f32, t_f32=t_f32)
        if need:
            ctx.res, ctx.p_bf, ctx.t_bf, ctx.tau = res, p_bf, t_bf, temperature
            ctx.f32 = (p_f32, t_f32)
            ctx.dtypes = (out1.dtype, out2.dtype)
        return res.loss

    @staticmethod
    def backward(ctx, dloss):
        dp, dt = _core.infonce_backward(ctx.res, ctx.p_bf, ctx.t_bf, ctx.tau, need_dp=ctx.needs_input_grad[0],
                                        need_dt=ctx.needs_input_grad[1], p_f32=ctx.f32[0], t_f32=ctx.f32[1])
        if dp is not None:
            dp = (dp * dloss).to(ctx.dtypes[0])
        if dt is not None:
            dt = (dt * dloss).to(ctx.dtypes[1])
        ctx.res = None
        return dp, dt, None, None, None, None


class BatchInfoNCELoss(torch.nn.Module):
    """-mean_i log( exp(S_ii) / sum_j exp(S_ij) ), S = out1 out2^T / temperature (reference :72-91)."""

    def __init__(self, temperature: float = 0.05):
        super().__init__()
        self.temperature = temperature

    def forward(self, batch_output1: torch.Tensor, batch_output2: torch.Tensor):
        labels = torch.arange(batch_output1.size(0), device=batch_output1.device, dtype=torch.int32)
        return _InfoNCEFunction.apply(batch_output1, batch_output2, labels, self.temperature, 1.0, 0.0)


class SegmentedBatchInfoNCELoss(torch.nn.Module):
    """Segmented version: rows of a segment against the whole batch (reference :94-114)."""

    def __init__(self, temperature: float = 0.05):
        super().__init__()
        self.temperature = temperature

    def forward(self, segment_output1: torch.Tensor, batch_output2: torch.Tensor, labels: torch.Tensor):
        return _InfoNCEFunction.apply(segment_output1, batch_output2, labels, self.temperature, 1.0, 0.0)


class SymmetricInfoNCELoss(torch.nn.Module):
    """north_star extension: 0.5 * (protein->text + text->protein); the column term equals the
    reference class called with its arguments swapped."""

    def __init__(self, temperature: float = 0.05):
        super().__init__()
        self.temperature = temperature

    def forward(self, batch_output1: torch.Tensor, batch_output2: torch.Tensor):
        labels = torch.arange(batch_output1.size(0), device=batch_output1.device, dtype=torch.int32)
        return _InfoNCEFunction.apply(batch_output1, batch_output2, labels, self.temperature, 0.5, 0.5)
