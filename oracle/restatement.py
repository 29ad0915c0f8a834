"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's Stage-1 contrastive hot path.

This is the parity oracle for the CUDA kernels.  It is NOT part of the product: only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may import it.
The product package (`prot2text-v2-esm3_b200/`) never does, and fails loudly without its CUDA
library.

Parity status: PINNED.  The reference ships no tests or golden vectors of its own,
so the pin is the reference code itself, executed in the authoring container through
`oracle/reference_loader.py`; `oracle/make_golden.py` dumps its inputs/outputs to
`tests/golden/*.npz` and `tests/test_oracle.py` checks every function below against them (and
against the live reference when `/root/reference` is mounted).

The arithmetic of the reference lives in PyTorch ATen (torch==2.3.0 pinned in the reference
README.md:42; torch 2.11 here, same op semantics): nn.Linear, exact-erf nn.GELU, nn.Dropout,
F.normalize (eps 1e-12), exp/sum/log/sqrt/pow.  The restatement below re-derives each stage in
plain tensor algebra — forward AND hand-derived backward, so autograd on the reference and these
closed forms are two independent witnesses for every gradient.  Works in float32 or float64.

Reference sites restated (paths relative to /root/reference):
  models/modeling_esm2llama_instruct.py:60-68   ModalityAdapter.forward       -> adapter_rows
  scripts/train_contrast.py:198-248             readout_embeddings            -> readout
  scripts/train_contrast.py:354,365             F.normalize(p=2, dim=-1)      -> l2_normalize
  scripts/train_contrast.py:86-91               BatchInfoNCELoss.forward      -> infonce_rows (labels=diag)
  scripts/train_contrast.py:100-114             SegmentedBatchInfoNCELoss     -> infonce_rows
  scripts/train_contrast.py:345-379             segment loop / averaging      -> step_forward
  scripts/train_contrast.py:448                 loss.backward()               -> step_backward
SURVEY.md §8f rows (callers either side of the path):
  scripts/train_contrast.py:455-465,621-626     clip_grad_norm_ + AdamW.step  -> clip_grad_norm, adamw_step
                                                (third-party: torch.nn.utils.clip_grad_norm_, torch.optim.AdamW of
                                                 torch==2.3.0 per the reference README; pinned against the installed
                                                 torch 2.11 in tests/test_oracle_next.py and tests/golden/next_adamw.npz)
  scripts/train_contrast.py:611-614             DDP gradient averaging        -> mean_allreduce_bf16
  models/modeling_esm2llama_instruct.py:134-138 placeholder replacement       -> placeholder_scatter
  models/esmc_qwen_arc.py:127-144               (same, with the count check)
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, Optional

import torch

EPS_NORM = 1e-12  # F.normalize default eps (clamp_min on the norm)
SQRT1_2 = 1.0 / math.sqrt(2.0)
INV_SQRT_2PI = 1.0 / math.sqrt(2.0 * math.pi)


# ----------------------------------------------------------------------------------------------
# element functions
# ----------------------------------------------------------------------------------------------
def gelu_erf(z: torch.Tensor) -> torch.Tensor:
    """Exact GELU z*Phi(z) — torch.nn.GELU() default (modeling_esm2llama_instruct.py:54)."""
    return z * 0.5 * (1.0 + torch.erf(z * SQRT1_2))


def gelu_erf_grad(z: torch.Tensor) -> torch.Tensor:
    """d/dz [z Phi(z)] = Phi(z) + z phi(z)."""
    cdf = 0.5 * (1.0 + torch.erf(z * SQRT1_2))
    pdf = torch.exp(-0.5 * z * z) * INV_SQRT_2PI
    return cdf + z * pdf


# ----------------------------------------------------------------------------------------------
# adapter (per residue row)
# ----------------------------------------------------------------------------------------------
@dataclass
class AdapterTrace:
    x: torch.Tensor
    z1: torch.Tensor
    h1: torch.Tensor
    z2: torch.Tensor
    a: torch.Tensor
    norm: torch.Tensor  # (rows, 1) clamped
    y: torch.Tensor
    keep1: Optional[torch.Tensor] = None  # dropout scale masks (0 or 1/(1-p)); None = eval
    keep2: Optional[torch.Tensor] = None


def adapter_rows(x, w1, b1, w2, b2, keep1=None, keep2=None) -> AdapterTrace:
    """fc1 -> GELU -> dropout -> fc2 -> GELU -> dropout -> per-row L2 normalise.

    Follows ModalityAdapter.forward (modeling_esm2llama_instruct.py:60-68).  ln1/ln2 exist in the
    module but are never applied (:56-57 'DEPRECATED').  `keep*` are the dropout multipliers
    (mask/(1-p)); None means eval mode.  x: (..., D_in); weights in nn.Linear layout (out, in).
    """
    z1 = x @ w1.t() + b1
    h1 = gelu_erf(z1)
    if keep1 is not None:
        h1 = h1 * keep1
    z2 = h1 @ w2.t() + b2
    a = gelu_erf(z2)
    if keep2 is not None:
        a = a * keep2
    norm = a.pow(2).sum(dim=-1, keepdim=True).sqrt().clamp_min(EPS_NORM)
    y = a / norm
    return AdapterTrace(x=x, z1=z1, h1=h1, z2=z2, a=a, norm=norm, y=y, keep1=keep1, keep2=keep2)


def adapter_rows_backward(tr: AdapterTrace, dy: torch.Tensor, w1, w2, need_dx: bool = False):
    """Closed-form backward of adapter_rows for flattened rows (R, D)."""
    y, norm = tr.y, tr.norm
    # y = a / max(|a|, eps): where the clamp is active the Jacobian is I/eps (ATen's clamp_min has
    # zero slope below the threshold), otherwise the usual projection.
    clamp_active = (tr.a.pow(2).sum(dim=-1, keepdim=True).sqrt() < EPS_NORM)
    proj = (y * dy).sum(dim=-1, keepdim=True)
    da = torch.where(clamp_active, dy / norm, (dy - y * proj) / norm)
    if tr.keep2 is not None:
        da = da * tr.keep2
    dz2 = da * gelu_erf_grad(tr.z2)
    dw2 = dz2.t() @ tr.h1
    db2 = dz2.sum(dim=0)
    dh1 = dz2 @ w2
    if tr.keep1 is not None:
        dh1 = dh1 * tr.keep1
    dz1 = dh1 * gelu_erf_grad(tr.z1)
    dw1 = dz1.t() @ tr.x
    db1 = dz1.sum(dim=0)
    out = {"fc1.weight": dw1, "fc1.bias": db1, "fc2.weight": dw2, "fc2.bias": db2,
           "dz2": dz2, "dz1": dz1}
    if need_dx:
        out["dx"] = dz1 @ w1
    return out


# ----------------------------------------------------------------------------------------------
# readout
# ----------------------------------------------------------------------------------------------
def readout(emb: torch.Tensor, mask: torch.Tensor, fn: str) -> torch.Tensor:
    """readout_embeddings (train_contrast.py:198-248).

    emb (B, S, D), mask (B, S) of 0/1.  'last' picks index sum(mask)-1 (right padding assumed,
    :207-215); 'mean' = sum(x*m)/sum(m) (:217-221); 'std' = sqrt(sum((x-mean)^2*m)/sum(m)),
    population, no eps (:223-235); 'mix' = cat(mean, std) (:237-248).
    """
    m = mask.to(emb.dtype).unsqueeze(-1)  # (B, S, 1)
    n = m.sum(dim=1)  # (B, 1)
    if fn == "last":
        idx = mask.sum(dim=1).long() - 1
        return emb[torch.arange(emb.shape[0]), idx, :]
    mean = (emb * m).sum(dim=1) / n
    if fn == "mean":
        return mean
    var = (((emb - mean.unsqueeze(1)) ** 2) * m).sum(dim=1) / n
    std = var.sqrt()
    if fn == "std":
        return std
    if fn == "mix":
        return torch.cat([mean, std], dim=1)
    raise ValueError(f"unknown readout_fn {fn!r}")


def readout_backward(emb, mask, fn: str, dout: torch.Tensor) -> torch.Tensor:
    """Closed-form d(readout)/d(emb) contracted with dout.  NaN where std == 0, as in autograd."""
    m = mask.to(emb.dtype).unsqueeze(-1)
    n = m.sum(dim=1, keepdim=True)  # (B,1,1)
    d = emb.shape[-1]
    if fn == "last":
        g = torch.zeros_like(emb)
        idx = mask.sum(dim=1).long() - 1
        g[torch.arange(emb.shape[0]), idx, :] = dout
        return g
    mean = ((emb * m).sum(dim=1, keepdim=True)) / n
    if fn == "mean":
        return m * dout.unsqueeze(1) / n
    var = (((emb - mean) ** 2) * m).sum(dim=1, keepdim=True) / n
    std = var.sqrt()
    if fn == "std":
        dmean, dstd = torch.zeros_like(dout), dout
    elif fn == "mix":
        dmean, dstd = dout[:, :d], dout[:, d:]
    else:
        raise ValueError(fn)
    # d std / d x_r = m_r (x_r - mean) / (n std)  [the mean's own dependence cancels: sum m (x-mean) = 0]
    # but the reference's `emb - mean` term is NOT masked before pow, and the masked sum of
    # (x - mean) is exactly zero only over valid rows, so the extra path through mean vanishes.
    g = m * (dmean.unsqueeze(1) / n + dstd.unsqueeze(1) * (emb - mean) / (n * std))
    return g


def l2_normalize(e: torch.Tensor):
    nrm = e.pow(2).sum(dim=-1, keepdim=True).sqrt().clamp_min(EPS_NORM)
    return e / nrm, nrm


def l2_normalize_backward(p, nrm, dp):
    return (dp - p * (p * dp).sum(dim=-1, keepdim=True)) / nrm


# ----------------------------------------------------------------------------------------------
# InfoNCE
# ----------------------------------------------------------------------------------------------
def similarity(p: torch.Tensor, t: torch.Tensor, tau: float) -> torch.Tensor:
    return (p @ t.t()) / tau


def infonce_rows(p, t, labels, tau: float = 0.05) -> torch.Tensor:
    """SegmentedBatchInfoNCELoss.forward (train_contrast.py:100-114); with labels = arange(B) it
    is BatchInfoNCELoss.forward (:86-91).  No max-subtraction in the reference; |S| <= 1/tau = 20
    for unit-norm inputs so exp is safe — the LSE form used here is algebraically identical."""
    s = similarity(p, t, tau)
    pos = s[torch.arange(s.shape[0]), labels]
    return (torch.logsumexp(s, dim=1) - pos).mean()


def infonce_cols(p, t, labels, tau: float = 0.05) -> torch.Tensor:
    """Text->protein direction over the columns that have a positive among these rows:
    the reference class called with arguments swapped.  Requires the row set to
    cover the labelled columns (labels index columns of S)."""
    s = similarity(p, t, tau)  # (R, C)
    pos = s[torch.arange(s.shape[0]), labels]
    lse_c = torch.logsumexp(s, dim=0)  # (C,)
    return (lse_c[labels] - pos).mean()


def infonce_backward(p, t, labels, tau: float, w_row: float = 1.0, w_col: float = 0.0):
    """dLoss/dS, dLoss/dp, dLoss/dt for loss = w_row*rows + w_col*cols."""
    s = similarity(p, t, tau)
    r = s.shape[0]
    onehot = torch.zeros_like(s)
    onehot[torch.arange(r), labels] = 1.0
    ds = torch.zeros_like(s)
    if w_row:
        ds = ds + w_row * (torch.softmax(s, dim=1) - onehot) / r
    if w_col:
        sm_c = torch.softmax(s, dim=0)
        colmask = torch.zeros(s.shape[1], dtype=s.dtype)
        colmask[labels] = 1.0
        ds = ds + w_col * (sm_c * colmask.unsqueeze(0) - onehot) / r
    dp = ds @ t / tau
    dt = ds.t() @ p / tau
    return ds, dp, dt


def retrieval_argmax(p, t):
    """Row-wise (protein->text) and column-wise (text->protein) argmax of the similarity;
    ties -> lowest index (torch.argmax returns the first maximal index on CPU)."""
    s = p @ t.t()
    return torch.argmax(s, dim=1), torch.argmax(s, dim=0)


# ----------------------------------------------------------------------------------------------
# whole step
# ----------------------------------------------------------------------------------------------
@dataclass
class StepTrace:
    loss: torch.Tensor
    p: torch.Tensor
    t: torch.Tensor
    e_prot: torch.Tensor
    e_text: torch.Tensor
    logits: torch.Tensor
    extras: Dict[str, torch.Tensor] = field(default_factory=dict)


def step_forward(x, prot_mask, w1, b1, w2, b2, text_hidden, text_mask, tau: float = 0.05,
                 num_segments: int = 1, symmetric: bool = False, keep1=None, keep2=None,
                 readout_fn: str = "mix") -> StepTrace:
    """teacher_forcing_forward_pass (train_contrast.py:313-379) from residue states to loss.

    x (B, L, D_in) residue states, prot_mask (B, L); text_hidden (B, T, H) = hidden_states[16] of
    the frozen LLM (:304), text_mask (B, T).  The segment loop (:356-377) averages per-segment
    means; rows beyond num_segments*(B//num_segments) are dropped exactly as the reference does
    (:337, :357-359).  `symmetric` adds the column-direction term (north_star extension, D5).
    """
    bsz = x.shape[0]
    tr = adapter_rows(x, w1, b1, w2, b2, keep1, keep2)
    e_p = readout(tr.y, prot_mask, readout_fn)
    p, pn = l2_normalize(e_p)
    e_t = readout(text_hidden, text_mask, readout_fn)
    t, tn = l2_normalize(e_t)
    seg = bsz // num_segments
    loss = x.new_zeros(())
    for s_id in range(num_segments):
        rows = slice(s_id * seg, (s_id + 1) * seg)
        labels = torch.arange(s_id * seg, (s_id + 1) * seg)
        part = infonce_rows(p[rows], t, labels, tau)
        if symmetric:
            part = 0.5 * (part + infonce_cols(p[rows], t, labels, tau))
        loss = loss + part
    loss = loss / num_segments
    return StepTrace(loss=loss, p=p, t=t, e_prot=e_p, e_text=e_t, logits=similarity(p, t, tau),
                     extras={"adapter": tr, "pn": pn, "tn": tn})


def step_backward(st: StepTrace, x, prot_mask, w1, w2, tau: float = 0.05, num_segments: int = 1,
                  symmetric: bool = False, readout_fn: str = "mix"):
    """Closed-form gradients of step_forward's loss w.r.t. the adapter parameters."""
    tr: AdapterTrace = st.extras["adapter"]
    bsz = x.shape[0]
    seg = bsz // num_segments
    dp = torch.zeros_like(st.p)
    for s_id in range(num_segments):
        rows = slice(s_id * seg, (s_id + 1) * seg)
        labels = torch.arange(s_id * seg, (s_id + 1) * seg)
        w_r, w_c = (0.5, 0.5) if symmetric else (1.0, 0.0)
        _, dps, _ = infonce_backward(st.p[rows], st.t, labels, tau, w_r, w_c)
        dp[rows] = dps / num_segments
    de = l2_normalize_backward(st.p, st.extras["pn"], dp)
    dy = readout_backward(tr.y, prot_mask, readout_fn, de)
    d_in, d_mid, d_out = x.shape[-1], w1.shape[0], w2.shape[0]
    flat = AdapterTrace(
        x=tr.x.reshape(-1, d_in), z1=tr.z1.reshape(-1, d_mid), h1=tr.h1.reshape(-1, d_mid),
        z2=tr.z2.reshape(-1, d_out), a=tr.a.reshape(-1, d_out), norm=tr.norm.reshape(-1, 1),
        y=tr.y.reshape(-1, d_out),
        keep1=None if tr.keep1 is None else tr.keep1.reshape(-1, d_mid),
        keep2=None if tr.keep2 is None else tr.keep2.reshape(-1, d_out))
    grads = adapter_rows_backward(flat, dy.reshape(-1, d_out), w1, w2)
    grads["dp"] = dp
    grads["de"] = de
    return grads


# ----------------------------------------------------------------------------------------------
# SURVEY.md §8f rows
# ----------------------------------------------------------------------------------------------
def clip_grad_norm(grads, max_norm: float):
    """torch.nn.utils.clip_grad_norm_ (L2): returns (total_norm, clipped grads).  coefficient =
    clamp(max_norm / (total_norm + 1e-6), max=1); max_norm = inf leaves the gradients untouched
    (scripts/train_contrast.py:456-463)."""
    total = torch.sqrt(sum((g.double() ** 2).sum() for g in grads))
    if max_norm is None or math.isinf(max_norm):
        return total, [g.clone() for g in grads]
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    return total, [g * coef.to(g.dtype) for g in grads]


def adamw_step(params, grads, exp_avg, exp_avg_sq, step: int, lr: float, betas=(0.9, 0.999), eps: float = 1e-6,
               weight_decay: float = 1e-2):
    """One torch.optim.AdamW step (decoupled weight decay, `step` counts from 1), out of place:
    returns (params, exp_avg, exp_avg_sq).  Reference optimizer: AdamW(lr, eps=1e-6, betas=(0.9, 0.999)),
    scripts/train_contrast.py:621-626 (weight_decay at torch's default 1e-2)."""
    b1, b2 = betas
    bc1, bc2 = 1.0 - b1 ** step, 1.0 - b2 ** step
    out_p, out_m, out_v = [], [], []
    for p, g, m, v in zip(params, grads, exp_avg, exp_avg_sq):
        p = p * (1.0 - lr * weight_decay)
        m = b1 * m + (1.0 - b1) * g
        v = b2 * v + (1.0 - b2) * g * g
        denom = v.sqrt() / math.sqrt(bc2) + eps
        p = p - (lr / bc1) * m / denom
        out_p.append(p); out_m.append(m); out_v.append(v)
    return out_p, out_m, out_v


def mean_allreduce_bf16(per_rank):
    """What DDP's gradient averaging yields for bf16 gradients, evaluated the way csrc/peer.cu does: fp32 sum in rank
    order, times 1/W, rounded to bf16 once."""
    acc = torch.zeros_like(per_rank[0], dtype=torch.float32)
    for g in per_rank:
        acc = acc + g.to(torch.float32)
    return (acc * (1.0 / len(per_rank))).to(torch.bfloat16)


def placeholder_scatter(inputs_embeds, placeholder_mask, encoder_hidden_states, encoder_mask):
    """inputs_embeds[placeholder_mask] = encoder_hidden_states[encoder_mask]
    (models/modeling_esm2llama_instruct.py:134-138): the k-th valid encoder row, batch-major, goes to the k-th
    placeholder slot, batch-major."""
    out = inputs_embeds.clone()
    out[placeholder_mask.bool()] = encoder_hidden_states[encoder_mask.bool()].to(out.dtype)
    return out
