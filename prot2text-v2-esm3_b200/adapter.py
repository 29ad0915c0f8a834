"""Drop-in `ModalityAdapter` / `ModalityAdapterConfig`.

Mirrors models/modeling_esm2llama_instruct.py:45-68 and models/modality_config.py:2-18 of the
reference: same constructor, same sub-module names (`fc1`, `fc2`, `activation`, `dropout`, and the
never-applied `ln1`/`ln2`, so reference checkpoints load with strict=True and PEFT's
modules_to_save=["adapter.fc1","adapter.fc2"] keeps working), same `forward(hidden_states)`
contract.  The arithmetic runs in two tcgen05 GEMMs with fused epilogues plus one row-scaling pass.
"""
from __future__ import annotations

import torch
from transformers import PretrainedConfig, PreTrainedModel

from . import _core, _lib


class ModalityAdapterConfig(PretrainedConfig):
    """Configuration of the 2-layer adapter (reference: models/modality_config.py:2-18)."""
    model_type = "modality_adapter"

    def __init__(self, input_dim: int = 0, intermediate_dim: int = 0, output_dim: int = 0,
                 dropout_rate: float = 0.3, **kwargs):
        super().__init__(**kwargs)
        self.input_dim = input_dim
        self.intermediate_dim = intermediate_dim
        self.output_dim = output_dim
        self.dropout_rate = dropout_rate


def _draw_seed() -> int:
    # CPU default generator: reproducible under torch.manual_seed, no device synchronisation
    return int(torch.empty((), dtype=torch.int64).random_().item())


class _AdapterFunction(torch.autograd.Function):
    """y = normalize(drop(GELU(fc2(drop(GELU(fc1(x))))))) for every row of x (…, D_in)."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, dropout_p: float, seed: int):
        for t, name in ((x, "hidden_states"), (w1, "fc1.weight"), (b1, "fc1.bias"), (w2, "fc2.weight"), (b2, "fc2.bias")):
            _core.require_cuda_bf16(t, name)
        d_in = w1.shape[1]
        d_out = w2.shape[0]
        x2d = x.reshape(-1, d_in)
        if not x2d.is_contiguous():
            x2d = x2d.contiguous()
        n = x2d.shape[0]
        rows_cap = _core._round_up(max(n, 1), _core.ROW_ALIGN)
        n_rows = torch.full((1,), n, dtype=torch.int32, device=x.device)
        need_grad = any(ctx.needs_input_grad[:5])
        acts = _core.adapter_forward(x2d, n, rows_cap, n_rows, w1.contiguous(), b1.contiguous(), w2.contiguous(),
                                     b2.contiguous(), dropout_p, seed, need_grad)
        y = torch.empty(n, d_out, dtype=torch.bfloat16, device=x.device)
        inv_norm = torch.empty(rows_cap, dtype=torch.float32, device=x.device) if need_grad else None
        _lib.call("p2t_adapter_scale_rows", _core._ptr(acts.a), _core._ptr(acts.rowsq), acts.nblk, rows_cap, n, d_out,
                  _core._ptr(y), _core._ptr(inv_norm), _core._stream())
        if need_grad:
            ctx.acts, ctx.inv_norm, ctx.n = acts, inv_norm, n
            ctx.save_for_backward(w1, w2)
            ctx.x_shape = x.shape
        return y.view(*x.shape[:-1], d_out)

    @staticmethod
    def backward(ctx, dy):
        w1, w2 = ctx.saved_tensors
        acts, n = ctx.acts, ctx.n
        d_out = w2.shape[0]
        dy2d = dy.reshape(-1, d_out).to(torch.bfloat16).contiguous()
        dz2 = torch.empty(acts.rows_cap, d_out, dtype=torch.bfloat16, device=dy.device)
        _lib.call("p2t_adapter_tail_bwd_dy", _core._ptr(acts.a), _core._ptr(acts.g2), _core._ptr(ctx.inv_norm),
                  _core._ptr(dy2d), n, None, acts.rows_cap, d_out, _core._ptr(dz2), _core._stream())
        need_dx = ctx.needs_input_grad[0]
        dw1, db1, dw2, db2, dx = _core.adapter_backward(acts, dz2, w1.contiguous(), w2.contiguous(), need_dx=need_dx)
        if dx is not None:
            dx = dx[:n].view(ctx.x_shape)
        ctx.acts = None
        return dx, dw1, db1, dw2, db2, None, None


class ModalityAdapter(PreTrainedModel):
    """2-layer adapter to match the hidden size of different modalities (drop-in for the reference class)."""
    config_class = ModalityAdapterConfig

    def __init__(self, config: ModalityAdapterConfig):
        super().__init__(config)
        self.config = config
        self.fc1 = torch.nn.Linear(config.input_dim, config.intermediate_dim)
        self.fc2 = torch.nn.Linear(config.intermediate_dim, config.output_dim)
        self.activation = torch.nn.GELU()  # exact-erf GELU, evaluated inside the GEMM epilogues
        self.dropout = torch.nn.Dropout(p=config.dropout_rate)  # p is read from here; the mask is a fused Philox stream
        self.ln1 = torch.nn.LayerNorm(normalized_shape=config.intermediate_dim)  # never applied (reference :56 DEPRECATED)
        self.ln2 = torch.nn.LayerNorm(normalized_shape=config.output_dim)  # never applied (reference :57 DEPRECATED)
        self.post_init()

    def dropout_p(self) -> float:
        return float(self.dropout.p) if self.training else 0.0

    def forward(self, hidden_states: torch.FloatTensor) -> torch.FloatTensor:
        # input: (bsz, seq_len, input_dim) -> (bsz, seq_len, output_dim), unit L2 norm per residue row
        p = self.dropout_p()
        seed = _draw_seed() if p > 0.0 else 0
        return _AdapterFunction.apply(hidden_states, self.fc1.weight, self.fc1.bias, self.fc2.weight, self.fc2.bias, p, seed)
