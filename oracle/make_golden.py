"""TEST INFRASTRUCTURE ONLY — dump golden vectors from the REAL reference (run in the authoring
container, where /root/reference is mounted):

    python -m oracle.make_golden            # writes tests/golden/*.npz

The reference has no tests/golden vectors of its own, so these files are the pin:
inputs and outputs of the reference's own ModalityAdapter, readout_embeddings, BatchInfoNCELoss,
SegmentedBatchInfoNCELoss and teacher_forcing_forward_pass (+ autograd), executed on CPU in
float32 (and float64 for the step) with fixed seeds.  They are small on purpose (a few KB each).
"""
from __future__ import annotations

import os
import types

import numpy as np
import torch

from oracle.reference_loader import load_reference

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _np(t: torch.Tensor) -> np.ndarray:
    return t.detach().cpu().numpy()


def _adapter(ref, d_in, d_mid, d_out, seed, dtype):
    torch.manual_seed(seed)
    cfg = ref.ModalityAdapterConfig(input_dim=d_in, intermediate_dim=d_mid, output_dim=d_out)
    ad = ref.ModalityAdapter(cfg)
    # HF init leaves biases at 0; perturb them so bias handling is actually exercised
    with torch.no_grad():
        ad.fc1.bias.normal_(0, 0.05)
        ad.fc2.bias.normal_(0, 0.05)
        ad.fc1.weight.mul_(6.0)  # bring pre-activations to O(1) so GELU's curvature matters
        ad.fc2.weight.mul_(6.0)
    return ad.to(dtype).eval()


def golden_adapter(ref):
    ad = _adapter(ref, 16, 32, 24, seed=11, dtype=torch.float32)
    g = torch.Generator().manual_seed(12)
    x = torch.randn(3, 7, 16, generator=g)
    gy = torch.randn(3, 7, 24, generator=g)
    y = ad(x)
    (y * gy).sum().backward()
    out = {"x": _np(x), "gy": _np(gy), "y": _np(y)}
    for k, v in ad.state_dict().items():
        out["sd." + k] = _np(v)
    for k, v in ad.named_parameters():
        out["grad." + k] = np.zeros(0, np.float32) if v.grad is None else _np(v.grad)
    np.savez(os.path.join(GOLDEN_DIR, "adapter_eval_f32.npz"), **out)


def golden_readout(ref):
    g = torch.Generator().manual_seed(21)
    emb = torch.randn(4, 9, 8, generator=g, requires_grad=True)
    lens = [9, 4, 1, 6]
    right = torch.zeros(4, 9, dtype=torch.long)
    left = torch.zeros(4, 9, dtype=torch.long)
    for b, n in enumerate(lens):
        right[b, :n] = 1
        left[b, 9 - n:] = 1
    out = {"emb": _np(emb), "mask_right": _np(right), "mask_left": _np(left), "mask_ones": _np(torch.ones_like(right))}
    for mname, mask in (("right", right), ("left", left), ("ones", torch.ones_like(right))):
        for fn in ("last", "mean", "std", "mix"):
            if fn == "last" and mname == "left":
                continue  # reference documents 'last' as right-padding only (:208-209)
            r = ref.readout_embeddings(emb, mask, fn)
            gd = torch.randn(r.shape, generator=g)
            (gemb,) = torch.autograd.grad((r * gd).sum(), emb)
            out[f"out.{mname}.{fn}"] = _np(r)
            out[f"gout.{mname}.{fn}"] = _np(gd)
            out[f"gemb.{mname}.{fn}"] = _np(gemb)
    np.savez(os.path.join(GOLDEN_DIR, "readout_f32.npz"), **out)


def golden_losses(ref):
    g = torch.Generator().manual_seed(31)
    p = torch.nn.functional.normalize(torch.randn(6, 10, generator=g), dim=-1).requires_grad_()
    t = torch.nn.functional.normalize(torch.randn(6, 10, generator=g), dim=-1).requires_grad_()
    out = {"p": _np(p), "t": _np(t)}
    full = ref.BatchInfoNCELoss()(p, t)
    gp, gt = torch.autograd.grad(full, (p, t))
    out.update({"batch.loss": _np(full), "batch.gp": _np(gp), "batch.gt": _np(gt)})
    swapped = ref.BatchInfoNCELoss()(t, p)  # text->protein direction (north_star's column term)
    gp, gt = torch.autograd.grad(swapped, (p, t))
    out.update({"swapped.loss": _np(swapped), "swapped.gp": _np(gp), "swapped.gt": _np(gt)})
    labels = torch.tensor([4, 5, 1])  # arbitrary labels, segment of 3 rows against 6 columns
    seg = ref.SegmentedBatchInfoNCELoss(temperature=0.07)(p[1:4], t, labels)
    gp, gt = torch.autograd.grad(seg, (p, t))
    out.update({"seg.labels": _np(labels), "seg.loss": _np(seg), "seg.gp": _np(gp), "seg.gt": _np(gt),
                "seg.temperature": np.float32(0.07)})
    np.savez(os.path.join(GOLDEN_DIR, "losses_f32.npz"), **out)


class _FakeTrunks(torch.nn.Module):
    """Stands in for ESMCQwen with the two frozen trunks replaced by table look-ups, so the REAL
    teacher_forcing_forward_pass (train_contrast.py:313-379) runs end to end on CPU.

    forward(protein_sequences=[ids...], return_encoder_outputs=True) pads the segment's residue
    states to the segment's own max length with zeros and applies the adapter, exactly the
    shape contract of models/esmc_qwen_arc.py:179-186.  llm_decoder.model(...) returns an object
    whose hidden_states[16] is the stored text hidden tensor (train_contrast.py:294-304).
    """

    def __init__(self, adapter, residue_states, text_hidden):
        super().__init__()
        self.adapter = adapter
        self._x = residue_states  # list of (L_b, D_in)
        self._text = text_hidden

        def _llm_model(input_ids, attention_mask, **kw):
            hs = [torch.zeros_like(self._text)] * 16 + [self._text]
            return types.SimpleNamespace(hidden_states=hs)

        self.llm_decoder = types.SimpleNamespace(model=_llm_model)

    def forward(self, protein_sequences, return_encoder_outputs=False):
        xs = [self._x[i] for i in protein_sequences]
        lmax = max(v.shape[0] for v in xs)
        pad = torch.zeros(len(xs), lmax, xs[0].shape[1], dtype=xs[0].dtype)
        for b, v in enumerate(xs):
            pad[b, : v.shape[0]] = v
        return (self.adapter(pad),)


def golden_step(ref, dtype, tag):
    ad = _adapter(ref, 12, 20, 16, seed=41, dtype=dtype)
    g = torch.Generator().manual_seed(42)
    lens = [5, 9, 3, 7, 9, 2]
    xs = [torch.randn(n, 12, generator=g).to(dtype) for n in lens]
    tlens = [4, 6, 6, 2, 5, 3]
    text = torch.randn(6, 6, 16, generator=g).to(dtype)
    tmask = torch.zeros(6, 6, dtype=torch.long)
    for b, n in enumerate(tlens):
        tmask[b, :n] = 1
    model = _FakeTrunks(ad, xs, text)
    batch = {"protein_sequences": list(range(6)), "description_input_ids": torch.zeros(6, 6, dtype=torch.long),
             "description_attention_mask": tmask}
    out = {"lens": np.array(lens), "text": _np(text), "tmask": _np(tmask)}
    for b, v in enumerate(xs):
        out[f"x{b}"] = _np(v)
    for k, v in ad.state_dict().items():
        out["sd." + k] = _np(v)
    for nseg in (1, 2, 3):
        ad.zero_grad(set_to_none=True)
        loss = ref.teacher_forcing_forward_pass("cpu", model, batch, nseg)
        loss.backward()
        out[f"seg{nseg}.loss"] = _np(loss)
        for k, v in ad.named_parameters():
            if v.grad is not None:
                out[f"seg{nseg}.grad.{k}"] = _np(v.grad)
    np.savez(os.path.join(GOLDEN_DIR, f"step_{tag}.npz"), **out)


# ----------------------------------------------------------------------------------------------
# "grid" fixtures for the CUDA parity tests (-m gpu): same reference functions, but every input and
# weight is first rounded to the bf16 grid (so the bf16 kernels see EXACTLY the reference's inputs)
# and feature dims are multiples of 8 (16-byte rows, the C ABI's alignment rule).  The reference
# still runs in float32 on those values.
# ----------------------------------------------------------------------------------------------
def _grid(t: torch.Tensor) -> torch.Tensor:
    return t.to(torch.bfloat16).to(torch.float32)


def _grid_adapter(ref, d_in, d_mid, d_out, seed):
    ad = _adapter(ref, d_in, d_mid, d_out, seed=seed, dtype=torch.float32)
    with torch.no_grad():
        for prm in ad.parameters():
            prm.copy_(_grid(prm))
    return ad


def golden_grid_adapter(ref):
    ad = _grid_adapter(ref, 24, 40, 32, seed=51)
    g = torch.Generator().manual_seed(52)
    x = _grid(torch.randn(3, 11, 24, generator=g))
    gy = _grid(torch.randn(3, 11, 32, generator=g))
    xg = x.clone().requires_grad_()
    y = ad(xg)
    (y * gy).sum().backward()
    out = {"x": _np(x), "gy": _np(gy), "y": _np(y), "grad.x": _np(xg.grad)}
    for k, v in ad.state_dict().items():
        out["sd." + k] = _np(v)
    for k, v in ad.named_parameters():
        if v.grad is not None:
            out["grad." + k] = _np(v.grad)
    np.savez(os.path.join(GOLDEN_DIR, "grid_adapter.npz"), **out)


def golden_grid_readout(ref):
    g = torch.Generator().manual_seed(61)
    emb = _grid(torch.randn(5, 70, 16, generator=g)).requires_grad_()  # 70 rows: more than one 64-row pooling chunk
    lens = [70, 33, 2, 64, 65]
    right = torch.zeros(5, 70, dtype=torch.long)
    left = torch.zeros(5, 70, dtype=torch.long)
    for b, n in enumerate(lens):
        right[b, :n] = 1
        left[b, 70 - n:] = 1
    holes = (torch.rand(5, 70, generator=g) > 0.4).long()
    holes[:, 0] = 1
    holes[:, 5] = 1
    out = {"emb": _np(emb), "mask_right": _np(right), "mask_left": _np(left), "mask_holes": _np(holes)}
    for mname, mask in (("right", right), ("left", left), ("holes", holes)):
        for fn in ("last", "mean", "std", "mix"):
            if fn == "last" and mname != "right":
                continue
            r = ref.readout_embeddings(emb, mask, fn)
            gd = _grid(torch.randn(r.shape, generator=g))
            (gemb,) = torch.autograd.grad((r * gd).sum(), emb)
            out[f"out.{mname}.{fn}"] = _np(r)
            out[f"gout.{mname}.{fn}"] = _np(gd)
            out[f"gemb.{mname}.{fn}"] = _np(gemb)
    np.savez(os.path.join(GOLDEN_DIR, "grid_readout.npz"), **out)


def golden_grid_losses(ref):
    g = torch.Generator().manual_seed(71)
    n, e = 12, 64
    t = torch.nn.functional.normalize(torch.randn(n, e, generator=g), dim=-1)
    perm = torch.randperm(n, generator=g)  # positives planted off the diagonal for the segmented call
    p_diag = torch.nn.functional.normalize(t + 2.0 * torch.randn(n, e, generator=g) / e ** 0.5, dim=-1)
    p = _grid(p_diag).requires_grad_()
    t = _grid(t).requires_grad_()
    out = {"p": _np(p), "t": _np(t)}
    full = ref.BatchInfoNCELoss()(p, t)
    gp, gt = torch.autograd.grad(full, (p, t))
    out.update({"batch.loss": _np(full), "batch.gp": _np(gp), "batch.gt": _np(gt)})
    swapped = ref.BatchInfoNCELoss()(t, p)
    gp, gt = torch.autograd.grad(swapped, (p, t))
    out.update({"swapped.loss": _np(swapped), "swapped.gp": _np(gp), "swapped.gt": _np(gt)})
    labels = perm[:5]
    seg = ref.SegmentedBatchInfoNCELoss(temperature=0.07)(p[3:8], t, labels)
    gp, gt = torch.autograd.grad(seg, (p, t))
    out.update({"seg.labels": _np(labels), "seg.loss": _np(seg), "seg.gp": _np(gp), "seg.gt": _np(gt),
                "seg.temperature": np.float32(0.07)})
    # retrieval indices from the reference's own logits (train_contrast.py:87: torch.mm(p, t^T) / tau)
    logits = torch.mm(p.detach(), t.detach().t()) / 0.05
    top2 = logits.topk(2, dim=1).values
    top2c = logits.topk(2, dim=0).values
    assert (top2[:, 0] - top2[:, 1]).min() > 0.25 and (top2c[0] - top2c[1]).min() > 0.25, "retrieval margin too small"
    out.update({"argmax_row": _np(logits.argmax(dim=1)), "argmax_col": _np(logits.argmax(dim=0)),
                "margin_row": _np((top2[:, 0] - top2[:, 1]).min()), "margin_col": _np((top2c[0] - top2c[1]).min())})
    np.savez(os.path.join(GOLDEN_DIR, "grid_losses.npz"), **out)


def golden_grid_step(ref):
    ad = _grid_adapter(ref, 24, 40, 32, seed=81)
    g = torch.Generator().manual_seed(82)
    lens = [5, 19, 3, 70, 9, 12, 66, 8]
    xs = [_grid(torch.randn(n, 24, generator=g)) for n in lens]
    tlens = [4, 6, 6, 2, 5, 3, 6, 1]
    text = _grid(torch.randn(8, 6, 32, generator=g))
    tmask = torch.zeros(8, 6, dtype=torch.long)
    for b, n in enumerate(tlens):
        tmask[b, :n] = 1
    model = _FakeTrunks(ad, xs, text)
    batch = {"protein_sequences": list(range(8)), "description_input_ids": torch.zeros(8, 6, dtype=torch.long),
             "description_attention_mask": tmask}
    out = {"lens": np.array(lens), "text": _np(text), "tmask": _np(tmask)}
    for b, v in enumerate(xs):
        out[f"x{b}"] = _np(v)
    for k, v in ad.state_dict().items():
        out["sd." + k] = _np(v)
    for nseg in (1, 2, 3, 4):  # 3 does not divide 8: the reference drops the remainder rows (:337,:357-359)
        ad.zero_grad(set_to_none=True)
        loss = ref.teacher_forcing_forward_pass("cpu", model, batch, nseg)
        loss.backward()
        out[f"seg{nseg}.loss"] = _np(loss)
        for k, v in ad.named_parameters():
            if v.grad is not None:
                out[f"seg{nseg}.grad.{k}"] = _np(v.grad)
    # the embeddings the loss saw at nseg=1 (whole batch padded to its max length, all-ones mask — D4)
    with torch.no_grad():
        e_p = ref.get_sequence_embeddings(model, list(range(8)))
        out["p"] = _np(torch.nn.functional.normalize(e_p, p=2, dim=-1))
    np.savez(os.path.join(GOLDEN_DIR, "grid_step.npz"), **out)


def main():
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    ref = load_reference()
    golden_adapter(ref)
    golden_readout(ref)
    golden_losses(ref)
    golden_step(ref, torch.float32, "f32")
    golden_step(ref, torch.float64, "f64")
    golden_grid_adapter(ref)
    golden_grid_readout(ref)
    golden_grid_losses(ref)
    golden_grid_step(ref)
    for f in sorted(os.listdir(GOLDEN_DIR)):
        print(f, os.path.getsize(os.path.join(GOLDEN_DIR, f)), "bytes")


if __name__ == "__main__":
    main()
