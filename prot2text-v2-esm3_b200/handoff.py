"""The callers either side of the adapter.

`llm_hidden_states_at` — text-side hand-off.  The reference runs EVERY layer of the frozen LLM with
`output_hidden_states=True` and keeps one tensor, `hidden_states[16]` (scripts/train_contrast.py:292-304).  That
tensor is the input of decoder layer 16, so layers 16.. and the final norm are dead work; this helper runs the stock
HF decoder on its first `layer` blocks only and returns the identical tensor, ready for `text_embeddings`.

`adapter_into_embeds` — Stage-2 hand-off.  `prepare_decoder_inputs` of both model classes
(models/esmc_qwen_arc.py:127-144, models/modeling_esm2llama_instruct.py:120-139) materialises the adapter output
(B, L, D_out) and then scatters its valid rows into the placeholder slots of `inputs_embeds`:
    inputs_embeds[placeholder_mask] = encoder_hidden_states[encoder_mask]
Here the adapter runs on the packed valid rows only and its normalising tail writes each row straight into its slot.
"""
from __future__ import annotations

import contextlib
from typing import Optional

import torch

from . import _core, _lib
from .adapter import ModalityAdapter, _draw_seed


# --------------------------------------------------------------------------------------------------
# text side: stop the frozen LLM at the layer whose input the contrastive step pools
# --------------------------------------------------------------------------------------------------
@contextlib.contextmanager
def _truncated(decoder, n_layers: int):
    """Temporarily make a HF decoder stack (`.layers`, `.norm`) end after `n_layers` blocks, without the final norm."""
    layers, norm = decoder.layers, decoder.norm
    if not 0 <= n_layers <= len(layers):
        raise ValueError(f"layer {n_layers} outside [0, {len(layers)}]")
    try:
        decoder.layers = layers[:n_layers]
        decoder.norm = torch.nn.Identity()
        yield decoder
    finally:
        decoder.layers = layers
        decoder.norm = norm


@torch.no_grad()
def llm_hidden_states_at(decoder, input_ids: torch.Tensor, attention_mask: Optional[torch.Tensor] = None,
                         layer: int = 16) -> torch.Tensor:
    """== decoder(input_ids, attention_mask, output_hidden_states=True).hidden_states[layer] for `layer` < depth,
    computed with `layer` decoder blocks instead of all of them.  `decoder` is the HF base model the reference calls
    (`model.llm_decoder.model`, scripts/train_contrast.py:292): any stack exposing `.layers` and `.norm`
    (LlamaModel, Qwen2Model, Qwen3Model).  `layer == depth` would need the final norm and is served by the full
    model instead."""
    depth = len(decoder.layers)
    if layer == depth:
        out = decoder(input_ids=input_ids, attention_mask=attention_mask, use_cache=False, output_hidden_states=True,
                      return_dict=True)
        return out.hidden_states[layer]
    with _truncated(decoder, layer) as d:
        out = d(input_ids=input_ids, attention_mask=attention_mask, use_cache=False, output_attentions=False,
                output_hidden_states=False, return_dict=True)
    return out.last_hidden_state


# --------------------------------------------------------------------------------------------------
# Stage 2: adapter rows straight into the LLM's input-embedding slots
# --------------------------------------------------------------------------------------------------
class _AdapterIntoEmbeds(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, enc_mask, inputs_embeds, placeholder_mask, w1, b1, w2, b2, dropout_p, seed):
        for t, name in ((x, "encoder_hidden_states"), (inputs_embeds, "inputs_embeds"), (w1, "fc1.weight"),
                        (b1, "fc1.bias"), (w2, "fc2.weight"), (b2, "fc2.bias")):
            _core.require_cuda_bf16(t, name)
        B, L, d_in = x.shape
        d_out = w2.shape[0]
        if inputs_embeds.shape[-1] != d_out or not inputs_embeds.is_contiguous():
            raise _lib.P2TError("inputs_embeds must be contiguous with last dimension = adapter output_dim")
        src_plan = _core.plan_rows(enc_mask)
        dst_plan = _core.plan_rows(placeholder_mask)
        xp = _core.gather_rows(x.contiguous().view(B * L, d_in), src_plan)
        need_grad = any(ctx.needs_input_grad[4:8])
        acts = _core.adapter_forward(xp, src_plan.rows_cap, src_plan.rows_cap, src_plan.n_rows, w1.contiguous(),
                                     b1.contiguous(), w2.contiguous(), b2.contiguous(), dropout_p, seed, need_grad)
        inv_norm = torch.empty(src_plan.rows_cap, dtype=torch.float32, device=x.device) if need_grad else None
        n_static = min(B * L, placeholder_mask.numel())
        n_moved = torch.minimum(src_plan.n_rows, dst_plan.n_rows)  # rows that have both a residue and a slot
        _lib.call("p2t_adapter_scatter_rows", _core._ptr(acts.a), _core._ptr(acts.rowsq), acts.nblk, acts.rows_cap,
                  n_static, d_out, _core._ptr(inputs_embeds), d_out, _core._ptr(dst_plan.row_src),
                  _core._ptr(src_plan.n_rows), _core._ptr(dst_plan.n_rows), _core._ptr(inv_norm), _core._stream())
        ctx.mark_dirty(inputs_embeds)
        ctx.placeholder_mask = placeholder_mask
        if need_grad:
            ctx.acts, ctx.inv_norm, ctx.dst_plan, ctx.n_moved = acts, inv_norm, dst_plan, n_moved
            ctx.save_for_backward(w1, w2)
        return inputs_embeds

    @staticmethod
    def backward(ctx, d_embeds):
        w1, w2 = ctx.saved_tensors
        acts, dst_plan = ctx.acts, ctx.dst_plan
        d_out = w2.shape[0]
        d2 = d_embeds.to(torch.bfloat16).contiguous().view(-1, d_out)
        # dy_k = d inputs_embeds[slot k]: the same gather that packs residue rows, keyed by the placeholder plan
        dy = torch.empty(acts.rows_cap, d_out, dtype=torch.bfloat16, device=d2.device)
        _lib.call("p2t_gather_rows", _core._ptr(d2), d2.stride(0), _core._ptr(dst_plan.row_src), _core._ptr(ctx.n_moved),
                  acts.rows_cap, d_out, _core._ptr(dy), _core._stream())
        dz2 = torch.empty(acts.rows_cap, d_out, dtype=torch.bfloat16, device=d2.device)
        _lib.call("p2t_adapter_tail_bwd_dy", _core._ptr(acts.a), _core._ptr(acts.g2), _core._ptr(ctx.inv_norm),
                  _core._ptr(dy), acts.rows_cap, _core._ptr(ctx.n_moved), acts.rows_cap, d_out, _core._ptr(dz2),
                  _core._stream())
        dw1, db1, dw2, db2, _ = _core.adapter_backward(acts, dz2, w1.contiguous(), w2.contiguous())
        ctx.acts = None
        # the placeholder slots were overwritten: no gradient reaches the token embeddings that stood there
        d_in_embeds = d_embeds.masked_fill(ctx.placeholder_mask.bool().unsqueeze(-1), 0)
        return None, None, d_in_embeds, None, dw1, db1, dw2, db2, None, None


def adapter_into_embeds(adapter: ModalityAdapter, encoder_hidden_states: torch.Tensor,
                        encoder_attention_mask: Optional[torch.Tensor], inputs_embeds: torch.Tensor,
                        placeholder_mask: torch.Tensor, check: bool = False) -> torch.Tensor:
    """inputs_embeds[placeholder_mask] = adapter(encoder_hidden_states)[encoder_attention_mask], in place.

    encoder_hidden_states (B, L, D_in) bf16 are the protein encoder's residue states (NOT yet adapted),
    encoder_attention_mask (B, L) marks the valid residues (None = all), inputs_embeds (B, S, D_out) bf16 are the
    LLM's token embeddings and placeholder_mask (B, S) = `input_ids == config.placeholder_id`.  As in the reference
    the k-th valid residue row (batch-major) lands in the k-th placeholder slot.  `check=True` reproduces the
    ValueError of models/esmc_qwen_arc.py:134-139 when the per-sequence counts differ (one host synchronisation);
    without it min(#residues, #placeholders) rows are written.  Differentiable w.r.t. the adapter weights and
    inputs_embeds.
    """
    if encoder_attention_mask is None:
        encoder_attention_mask = torch.ones(encoder_hidden_states.shape[:2], dtype=torch.uint8,
                                            device=encoder_hidden_states.device)
    p = adapter.dropout_p()
    seed = _draw_seed() if p > 0.0 else 0
    if check:
        n_res = encoder_attention_mask.bool().sum(dim=1)
        n_ph = placeholder_mask.bool().sum(dim=1)
        if not torch.all(n_res == n_ph):
            raise ValueError(f"Number of placeholder tokens ({n_ph.tolist()}) must match number of protein tokens "
                             f"({n_res.tolist()})")
    return _AdapterIntoEmbeds.apply(encoder_hidden_states, encoder_attention_mask, inputs_embeds, placeholder_mask,
                                    adapter.fc1.weight, adapter.fc1.bias, adapter.fc2.weight, adapter.fc2.bias, p, seed)
