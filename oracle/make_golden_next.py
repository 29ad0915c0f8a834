"""TEST INFRASTRUCTURE ONLY — golden vectors for the SURVEY.md §8f rows, dumped from the REAL reference / from the
third-party code it calls (run in the authoring container, where /root/reference is mounted):

    python -m oracle.make_golden_next       # writes tests/golden/next_*.npz

  next_scatter.npz  Esm2LlamaInstructForCausalLM.prepare_decoder_inputs (models/modeling_esm2llama_instruct.py:108-139)
                    executed unmodified on a stand-in `self` (token-embedding table + placeholder id), fed with the
                    reference ModalityAdapter's output, exactly as its forward does (:186-199).  Inputs on the bf16 grid.
  next_adamw.npz    torch.nn.utils.clip_grad_norm_ + torch.optim.AdamW(eps=1e-6, betas=(0.9, 0.999)) — the calls of
                    scripts/train_contrast.py:455-465 / :621-626 — for 4 steps on adapter-shaped fp32 tensors.
"""
from __future__ import annotations

import os
import types

import numpy as np
import torch

from oracle.reference_loader import load_reference, REFERENCE_ROOT

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _np(t):
    return t.detach().cpu().clone().numpy()


def _grid(t):
    return t.to(torch.bfloat16).to(torch.float32)


def golden_scatter(ref):
    import importlib
    modeling = importlib.import_module("models.modeling_esm2llama_instruct")
    d_in, d_mid, d_out, B, L, S, vocab, ph = 16, 32, 24, 3, 7, 12, 20, 19
    torch.manual_seed(31)
    ad = ref.ModalityAdapter(ref.ModalityAdapterConfig(input_dim=d_in, intermediate_dim=d_mid, output_dim=d_out)).eval()
    with torch.no_grad():
        for prm in (ad.fc1.weight, ad.fc2.weight):
            prm.copy_(_grid(prm * 6.0))
        ad.fc1.bias.copy_(_grid(torch.randn(d_mid) * 0.05))
        ad.fc2.bias.copy_(_grid(torch.randn(d_out) * 0.05))
    g = torch.Generator().manual_seed(32)
    x = _grid(torch.randn(B, L, d_in, generator=g))
    lens = [7, 3, 5]
    enc_mask = torch.zeros(B, L, dtype=torch.long)
    input_ids = torch.randint(0, ph, (B, S), generator=g)
    for b, n in enumerate(lens):
        enc_mask[b, :n] = 1
        start = 1 + b  # placeholders sit at different offsets in every row
        input_ids[b, start:start + n] = ph
    table = torch.nn.Embedding(vocab, d_out)
    with torch.no_grad():
        table.weight.copy_(_grid(torch.randn(vocab, d_out, generator=g)))
    fake_self = types.SimpleNamespace(
        llama_decoder=types.SimpleNamespace(get_input_embeddings=lambda: table),
        config=types.SimpleNamespace(placeholder_id=ph))
    y = ad(x)
    embeds, _ = modeling.Esm2LlamaInstructForCausalLM.prepare_decoder_inputs(
        fake_self, input_ids=input_ids, encoder_hidden_states=y, attention_mask=None, encoder_attention_mask=enc_mask)
    gy = _grid(torch.randn(B, S, d_out, generator=g))
    (embeds * gy).sum().backward()
    out = {"x": _np(x), "enc_mask": _np(enc_mask), "input_ids": _np(input_ids), "placeholder_id": np.int64(ph),
           "table": _np(table.weight), "embeds": _np(embeds), "gy": _np(gy)}
    for k, v in ad.state_dict().items():
        out["sd." + k] = _np(v)
    for k, v in ad.named_parameters():
        if v.grad is not None:
            out["grad." + k] = _np(v.grad)
    np.savez(os.path.join(GOLDEN_DIR, "next_scatter.npz"), **out)


def golden_adamw():
    shapes = [(32, 16), (32,), (24, 32), (24,)]
    g = torch.Generator().manual_seed(41)
    params = [torch.nn.Parameter(_grid(torch.randn(*s, generator=g) * 0.05)) for s in shapes]
    opt = torch.optim.AdamW(params, lr=3e-3, eps=1e-6, betas=(0.9, 0.999))
    out = {"n_steps": np.int64(4), "lr": np.float64(3e-3), "max_norm": np.float64(0.5)}
    for i, p in enumerate(params):
        out[f"p0.{i}"] = _np(p)
    for step in range(4):
        scale = [3.0, 0.3, 1.0, 0.05][step]  # steps 0 and 2 clip, 1 and 3 do not
        for i, p in enumerate(params):
            p.grad = _grid(torch.randn(*p.shape, generator=g) * scale * 0.02)
            out[f"g{step}.{i}"] = _np(p.grad)
        norm = torch.nn.utils.clip_grad_norm_(params, max_norm=0.5)
        out[f"norm{step}"] = _np(norm)
        opt.step()
        for i, p in enumerate(params):
            out[f"p{step + 1}.{i}"] = _np(p)
    np.savez(os.path.join(GOLDEN_DIR, "next_adamw.npz"), **out)


def main():
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    ref = load_reference()
    golden_scatter(ref)
    golden_adamw()
    print("wrote next_scatter.npz, next_adamw.npz from", REFERENCE_ROOT, "and torch", torch.__version__)


if __name__ == "__main__":
    main()
