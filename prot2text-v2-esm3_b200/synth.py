"""Seeded synthetic inputs for the BASELINE.json configurations.

Host-side only (CPU tensors); callers move them to the device.  There is no network for datasets
or checkpoints, so residue states / text hidden states are N(0,1) rounded to bf16 and the adapter
weights are random-init (HF default N(0, 0.02^2); `weight_gain` rescales them so pre-activations
are O(1) and GELU's curvature is exercised).
"""
from __future__ import annotations

from dataclasses import dataclass

import torch

CONFIGS = {
    # name: d_in, d_mid, d_out(=H), batch per rank, protein len range, text len range
    "cfg1_esm2_t6_llama1b": dict(d_in=320, d_mid=2048, d_out=2048, batch=16, lmin=32, lmax=256, tmin=16, tmax=128),
    "cfg2_esm2_3b_llama8b": dict(d_in=2560, d_mid=2048, d_out=4096, batch=32, lmin=50, lmax=1024, tmin=16, tmax=256),
    "cfg4_esmc600m_qwen7b": dict(d_in=1152, d_mid=2048, d_out=3584, batch=64, lmin=50, lmax=1024, tmin=16, tmax=256),
    "cfg5_sweep": dict(d_in=2560, d_mid=2048, d_out=4096, batch=512, lmin=50, lmax=2048, tmin=16, tmax=256),
    "tiny": dict(d_in=64, d_mid=128, d_out=96, batch=6, lmin=3, lmax=40, tmin=2, tmax=12),
}


@dataclass
class SynthBatch:
    x: torch.Tensor          # (B, Lmax, d_in) bf16, zero in padded rows
    prot_mask: torch.Tensor  # (B, Lmax) int64, right padded
    prot_lens: torch.Tensor
    text: torch.Tensor       # (B, Tmax, H) bf16
    text_mask: torch.Tensor  # (B, Tmax) int64
    w1: torch.Tensor
    b1: torch.Tensor
    w2: torch.Tensor
    b2: torch.Tensor


def make_batch(d_in, d_mid, d_out, batch, lmin, lmax, tmin, tmax, seed: int = 1234, rank: int = 0,
               weight_gain: float = 1.0, left_pad: bool = False, plant: float = 0.0,
               same_lengths_as_rank0: bool = False, device=None, lens=None, tlens=None) -> SynthBatch:
    """`same_lengths_as_rank0`: every rank gets rank 0's multiset of sequence lengths (in a rank-specific order) with
    its own random data — per-GPU work is then exactly fixed as ranks are added (weak-scaling benchmark).
    `device`: draw the residue / text states there with a device generator (same lengths and masks as the host
    path, different values) — the big sweep configurations are 5 GB of residue states per rank, which is minutes of
    single-threaded host randn under torchrun; everything else stays on the host.
    `lens` / `tlens`: explicit protein / text lengths instead of drawing them."""
    lens_override, tlens_override = lens, tlens
    g = torch.Generator().manual_seed(seed + rank)
    lens = torch.randint(lmin, lmax + 1, (batch,), generator=g)
    tlens = torch.randint(tmin, tmax + 1, (batch,), generator=g)
    if same_lengths_as_rank0 and rank != 0:
        g0 = torch.Generator().manual_seed(seed)
        lens0 = torch.randint(lmin, lmax + 1, (batch,), generator=g0)
        tlens0 = torch.randint(tmin, tmax + 1, (batch,), generator=g0)
        perm = torch.randperm(batch, generator=torch.Generator().manual_seed(seed + 7919 * rank))
        lens, tlens = lens0[perm], tlens0[perm]
    if lens_override is not None:  # explicit sequence lengths (e.g. a rank's share of a length-balanced global batch)
        lens = torch.as_tensor(lens_override, dtype=torch.long)
        tlens = torch.as_tensor(tlens_override, dtype=torch.long) if tlens_override is not None else tlens[:len(lens)]
        batch = len(lens)
    L, T = int(lens.max()), int(tlens.max())
    ar_l, ar_t = torch.arange(L)[None, :], torch.arange(T)[None, :]
    pm = (ar_l >= (L - lens)[:, None]) if left_pad else (ar_l < lens[:, None])
    tm = ar_t < tlens[:, None]
    if device is None:
        x = torch.randn(batch, L, d_in, generator=g).to(torch.bfloat16)
        text = torch.randn(batch, T, d_out, generator=g).to(torch.bfloat16)
        x = x * pm[..., None].to(torch.bfloat16)
        if plant > 0.0:
            # plant a per-pair signature so that retrieval has a margin (bit-exact argmax tests)
            sig = torch.randn(batch, 1, d_out, generator=g).to(torch.bfloat16)
            text = (text + plant * sig).to(torch.bfloat16)
    else:
        dg = torch.Generator(device=device).manual_seed(seed + rank)
        x = torch.randn(batch, L, d_in, generator=dg, device=device, dtype=torch.bfloat16)
        text = torch.randn(batch, T, d_out, generator=dg, device=device, dtype=torch.bfloat16)
        x.mul_(pm.to(device=device, dtype=torch.bfloat16)[..., None])
        if plant > 0.0:
            sig = torch.randn(batch, 1, d_out, generator=dg, device=device, dtype=torch.bfloat16)
            text = (text + plant * sig).to(torch.bfloat16)
    wg = torch.Generator().manual_seed(0)  # weights are the same on every rank
    w1 = (torch.randn(d_mid, d_in, generator=wg) * 0.02 * weight_gain).to(torch.bfloat16)
    w2 = (torch.randn(d_out, d_mid, generator=wg) * 0.02 * weight_gain).to(torch.bfloat16)
    b1 = (torch.randn(d_mid, generator=wg) * 0.02).to(torch.bfloat16)
    b2 = (torch.randn(d_out, generator=wg) * 0.02).to(torch.bfloat16)
    return SynthBatch(x=x, prot_mask=pm.long(), prot_lens=lens, text=text, text_mask=tm.long(), w1=w1, b1=b1, w2=w2, b2=b2)


def draw_lengths(name: str, seed: int = 1234, rank: int = 0):
    """(protein lengths, text lengths) exactly as make_config_batch(name, seed=seed, rank=rank) draws them."""
    cfg = CONFIGS[name]
    g = torch.Generator().manual_seed(seed + rank)
    lens = torch.randint(cfg["lmin"], cfg["lmax"] + 1, (cfg["batch"],), generator=g)
    tlens = torch.randint(cfg["tmin"], cfg["tmax"] + 1, (cfg["batch"],), generator=g)
    return lens, tlens


def make_config_batch(name: str, **kw) -> SynthBatch:
    cfg = dict(CONFIGS[name])
    cfg.update({k: v for k, v in kw.items() if k in cfg})
    extra = {k: v for k, v in kw.items() if k not in cfg}
    return make_batch(**cfg, **extra)
