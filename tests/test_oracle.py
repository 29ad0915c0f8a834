"""CPU: pin the oracle (oracle/restatement.py) against golden vectors dumped from the REAL
reference by oracle/make_golden.py, and against the live reference when it is mounted."""
import os

import numpy as np
import pytest
import torch

from oracle import restatement as R
from oracle.reference_loader import load_reference, reference_available


def _load(golden_dir, name):
    z = np.load(os.path.join(golden_dir, name))
    return {k: torch.from_numpy(np.asarray(z[k])) for k in z.files}


def _close(a, b, rtol=2e-5, atol=2e-6):
    torch.testing.assert_close(a, b, rtol=rtol, atol=atol, equal_nan=True)


def test_adapter_forward_backward_matches_reference_golden(golden_dir):
    g = _load(golden_dir, "adapter_eval_f32.npz")
    w1, b1, w2, b2 = g["sd.fc1.weight"], g["sd.fc1.bias"], g["sd.fc2.weight"], g["sd.fc2.bias"]
    x = g["x"]
    tr = R.adapter_rows(x.reshape(-1, x.shape[-1]), w1, b1, w2, b2)
    _close(tr.y.reshape(g["y"].shape), g["y"])
    # rows are unit-norm (known-answer fact)
    _close(tr.y.norm(dim=-1), torch.ones(tr.y.shape[0]))
    grads = R.adapter_rows_backward(tr, g["gy"].reshape(-1, g["gy"].shape[-1]), w1, w2)
    for k in ("fc1.weight", "fc1.bias", "fc2.weight", "fc2.bias"):
        _close(grads[k], g["grad." + k], rtol=1e-4, atol=1e-5)
    # ln1/ln2 exist in the state dict but get no gradient (D1)
    assert {"sd.ln1.weight", "sd.ln1.bias", "sd.ln2.weight", "sd.ln2.bias"} <= set(g)
    for k in ("ln1.weight", "ln1.bias", "ln2.weight", "ln2.bias"):
        assert g["grad." + k].numel() == 0


@pytest.mark.parametrize("mname", ["right", "left", "ones"])
@pytest.mark.parametrize("fn", ["last", "mean", "std", "mix"])
def test_readout_matches_reference_golden(golden_dir, mname, fn):
    if fn == "last" and mname == "left":
        pytest.skip("reference defines 'last' for right padding only")
    g = _load(golden_dir, "readout_f32.npz")
    emb, mask = g["emb"], g["mask_" + mname]
    _close(R.readout(emb, mask, fn), g[f"out.{mname}.{fn}"])
    _close(R.readout_backward(emb, mask, fn, g[f"gout.{mname}.{fn}"]), g[f"gemb.{mname}.{fn}"],
           rtol=1e-4, atol=1e-5)


def test_readout_std_of_single_token_is_zero_with_nan_grad(golden_dir):
    # row 2 of the right-padded mask has exactly one valid token -> std 0, grad NaN
    g = _load(golden_dir, "readout_f32.npz")
    assert torch.all(g["out.right.std"][2] == 0)
    assert torch.isnan(g["gemb.right.std"][2, 0]).all()
    mine = R.readout_backward(g["emb"], g["mask_right"], "std", g["gout.right.std"])
    assert torch.isnan(mine[2, 0]).all()


def test_losses_match_reference_golden(golden_dir):
    g = _load(golden_dir, "losses_f32.npz")
    p, t = g["p"], g["t"]
    diag = torch.arange(p.shape[0])
    _close(R.infonce_rows(p, t, diag, 0.05), g["batch.loss"])
    _, dp, dt = R.infonce_backward(p, t, diag, 0.05)
    _close(dp, g["batch.gp"], rtol=1e-4, atol=1e-6)
    _close(dt, g["batch.gt"], rtol=1e-4, atol=1e-6)
    # column direction == reference class with swapped arguments
    _close(R.infonce_cols(p, t, diag, 0.05), g["swapped.loss"])
    _, dp, dt = R.infonce_backward(p, t, diag, 0.05, w_row=0.0, w_col=1.0)
    _close(dp, g["swapped.gp"], rtol=1e-4, atol=1e-6)
    _close(dt, g["swapped.gt"], rtol=1e-4, atol=1e-6)
    # segmented, arbitrary labels, non-default temperature
    tau = float(g["seg.temperature"])
    labels = g["seg.labels"].long()
    _close(R.infonce_rows(p[1:4], t, labels, tau), g["seg.loss"])
    _, dps, dts = R.infonce_backward(p[1:4], t, labels, tau)
    _close(dps, g["seg.gp"][1:4], rtol=1e-4, atol=1e-6)
    assert torch.all(g["seg.gp"][0] == 0) and torch.all(g["seg.gp"][4:] == 0)
    _close(dts, g["seg.gt"], rtol=1e-4, atol=1e-6)


def _fork_step(g, dtype, nseg):
    """The reference step as this fork runs it: each segment zero-padded to its own max
    length, all-ones mask, 'mix' readout, loss averaged over segments — built from oracle pieces."""
    lens = [int(v) for v in g["lens"]]
    xs = [g[f"x{b}"].to(dtype) for b in range(len(lens))]
    w1, b1, w2, b2 = (g["sd." + k].to(dtype).requires_grad_() for k in
                      ("fc1.weight", "fc1.bias", "fc2.weight", "fc2.bias"))
    e_t = R.readout(g["text"].to(dtype), g["tmask"], "mix")
    t, _ = R.l2_normalize(e_t)
    bsz = len(lens)
    seg = bsz // nseg
    loss = torch.zeros((), dtype=dtype)
    for s in range(nseg):
        ids = list(range(s * seg, (s + 1) * seg))
        lmax = max(lens[i] for i in ids)
        x = torch.zeros(len(ids), lmax, xs[0].shape[1], dtype=dtype)
        for r, i in enumerate(ids):
            x[r, : lens[i]] = xs[i]
        tr = R.adapter_rows(x, w1, b1, w2, b2)
        e = R.readout(tr.y, torch.ones(len(ids), lmax, dtype=torch.long), "mix")
        p, _ = R.l2_normalize(e)
        loss = loss + R.infonce_rows(p, t, torch.tensor(ids), 0.05)
    loss = loss / nseg
    grads = torch.autograd.grad(loss, (w1, b1, w2, b2))
    return loss.detach(), dict(zip(("fc1.weight", "fc1.bias", "fc2.weight", "fc2.bias"), grads))


@pytest.mark.parametrize("tag,dtype,rtol", [("f32", torch.float32, 2e-4), ("f64", torch.float64, 1e-8)])
@pytest.mark.parametrize("nseg", [1, 2, 3])
def test_step_matches_reference_teacher_forcing_forward_pass(golden_dir, tag, dtype, rtol, nseg):
    g = _load(golden_dir, f"step_{tag}.npz")
    loss, grads = _fork_step(g, dtype, nseg)
    # the reference accumulates the loss into an fp32 0-dim tensor (train_contrast.py:345), so even
    # the float64 run carries fp32 rounding on the scalar and on 1/num_segments in its backward
    _close(loss, g[f"seg{nseg}.loss"].to(dtype), rtol=max(rtol, 3e-7), atol=0)
    for k, v in grads.items():
        ref = g[f"seg{nseg}.grad.{k}"].to(dtype)
        _close(v, ref, rtol=rtol * 10, atol=ref.abs().max().item() * rtol)


def _fork_mask_step(g, nseg):
    """Same step through R.step_forward with ONE padded batch and the mask the fork effectively applies
    (every segment pooled over its own padded length) — the form the CUDA path is driven with."""
    lens = [int(v) for v in g["lens"]]
    bsz, d_in = len(lens), g["x0"].shape[1]
    seg = bsz // nseg
    lmax = max(lens)
    x = torch.zeros(bsz, lmax, d_in)
    mask = torch.zeros(bsz, lmax, dtype=torch.long)
    for b, n in enumerate(lens):
        x[b, :n] = g[f"x{b}"]
        ids = range((b // seg) * seg, (b // seg + 1) * seg) if b < seg * nseg else [b]
        mask[b, : max(lens[i] for i in ids)] = 1
    w1, b1, w2, b2 = (g["sd." + k] for k in ("fc1.weight", "fc1.bias", "fc2.weight", "fc2.bias"))
    st = R.step_forward(x, mask, w1, b1, w2, b2, g["text"], g["tmask"], 0.05, nseg)
    return st, R.step_backward(st, x, mask, w1, w2, 0.05, nseg)


@pytest.mark.parametrize("nseg", [1, 2, 3, 4])
def test_grid_step_fixture_matches_oracle(golden_dir, nseg):
    """bf16-grid fixture of the real teacher_forcing_forward_pass (incl. nseg=3, which drops 2 of 8 rows)."""
    g = _load(golden_dir, "grid_step.npz")
    st, grads = _fork_mask_step(g, nseg)
    _close(st.loss, g[f"seg{nseg}.loss"], rtol=2e-5, atol=0)
    for k in ("fc1.weight", "fc1.bias", "fc2.weight", "fc2.bias"):
        ref = g[f"seg{nseg}.grad.{k}"]
        _close(grads[k], ref, rtol=2e-3, atol=ref.abs().max().item() * 2e-4)
    if nseg == 1:
        _close(st.p, g["p"], rtol=1e-4, atol=1e-6)


def test_grid_fixtures_match_oracle(golden_dir):
    g = _load(golden_dir, "grid_adapter.npz")
    w1, b1, w2, b2 = (g["sd." + k] for k in ("fc1.weight", "fc1.bias", "fc2.weight", "fc2.bias"))
    tr = R.adapter_rows(g["x"].reshape(-1, 24), w1, b1, w2, b2)
    _close(tr.y.reshape(g["y"].shape), g["y"])
    grads = R.adapter_rows_backward(tr, g["gy"].reshape(-1, 32), w1, w2, need_dx=True)
    for k in ("fc1.weight", "fc1.bias", "fc2.weight", "fc2.bias"):
        _close(grads[k], g["grad." + k], rtol=1e-4, atol=1e-5)
    _close(grads["dx"].reshape(g["grad.x"].shape), g["grad.x"], rtol=1e-4, atol=1e-5)
    g = _load(golden_dir, "grid_readout.npz")
    for mname in ("right", "left", "holes"):
        for fn in ("last", "mean", "std", "mix"):
            if f"out.{mname}.{fn}" not in g:
                continue
            _close(R.readout(g["emb"], g["mask_" + mname], fn), g[f"out.{mname}.{fn}"])
            _close(R.readout_backward(g["emb"], g["mask_" + mname], fn, g[f"gout.{mname}.{fn}"]), g[f"gemb.{mname}.{fn}"],
                   rtol=1e-4, atol=1e-5)
    g = _load(golden_dir, "grid_losses.npz")
    p, t = g["p"], g["t"]
    diag = torch.arange(p.shape[0])
    _close(R.infonce_rows(p, t, diag, 0.05), g["batch.loss"], rtol=1e-4)
    _close(R.infonce_cols(p, t, diag, 0.05), g["swapped.loss"], rtol=1e-4)
    _close(R.infonce_rows(p[3:8], t, g["seg.labels"].long(), float(g["seg.temperature"])), g["seg.loss"], rtol=1e-4)
    am_r, am_c = R.retrieval_argmax(p, t)
    assert torch.equal(am_r, g["argmax_row"]) and torch.equal(am_c, g["argmax_col"])  # index work: bit-exact


def test_step_closed_form_backward_agrees_with_autograd():
    """Second witness: the hand-derived backward equals autograd on the restatement (fp64)."""
    torch.manual_seed(5)
    dt = torch.float64
    b, l, din, dmid, dout, tlen = 6, 16, 10, 14, 12, 7
    x = torch.randn(b, l, din, dtype=dt)
    lens = torch.tensor([16, 9, 12, 8, 14, 10])  # long enough that dropout never zeroes a whole column (std=0 -> NaN grad)
    mask = (torch.arange(l)[None, :] < lens[:, None]).long()
    w1 = (torch.randn(dmid, din, dtype=dt) * 0.3).requires_grad_()
    b1 = (torch.randn(dmid, dtype=dt) * 0.1).requires_grad_()
    w2 = (torch.randn(dout, dmid, dtype=dt) * 0.3).requires_grad_()
    b2 = (torch.randn(dout, dtype=dt) * 0.1).requires_grad_()
    text = torch.randn(b, tlen, dout, dtype=dt)
    tmask = (torch.arange(tlen)[None, :] < torch.tensor([7, 2, 5, 4, 7, 3])[:, None]).long()
    keep1 = (torch.rand(b, l, dmid) > 0.2).to(dt) / 0.8
    keep2 = (torch.rand(b, l, dout) > 0.2).to(dt) / 0.8
    for sym in (False, True):
        for nseg in ((1,) if sym else (1, 2, 3)):
            st = R.step_forward(x, mask, w1, b1, w2, b2, text, tmask, 0.05, nseg, sym, keep1, keep2)
            auto = torch.autograd.grad(st.loss, (w1, b1, w2, b2))
            mine = R.step_backward(st, x, mask, w1.detach(), w2.detach(), 0.05, nseg, sym)
            for k, a in zip(("fc1.weight", "fc1.bias", "fc2.weight", "fc2.bias"), auto):
                torch.testing.assert_close(mine[k], a, rtol=1e-8, atol=1e-12)


def test_known_answer_facts():
    """Facts derivable from the reference code: segment-average == full batch ==
    cross_entropy(p t^T / 0.05, arange); random unit vectors give loss ~ ln B."""
    torch.manual_seed(0)
    p = torch.nn.functional.normalize(torch.randn(16, 4096), dim=-1)
    t = torch.nn.functional.normalize(torch.randn(16, 4096), dim=-1)
    full = R.infonce_rows(p, t, torch.arange(16), 0.05)
    halves = 0.5 * (R.infonce_rows(p[:8], t, torch.arange(8), 0.05) + R.infonce_rows(p[8:], t, torch.arange(8, 16), 0.05))
    ce = torch.nn.functional.cross_entropy(p @ t.t() / 0.05, torch.arange(16))
    torch.testing.assert_close(full, halves)
    torch.testing.assert_close(full, ce)
    assert abs(full.item() - np.log(16)) < 0.15


@pytest.mark.skipif(not reference_available(), reason="reference tree not mounted (GPU box)")
def test_live_reference_agrees_on_fresh_random_inputs():
    ref = load_reference()
    torch.manual_seed(77)
    cfg = ref.ModalityAdapterConfig(input_dim=20, intermediate_dim=24, output_dim=28)
    ad = ref.ModalityAdapter(cfg).eval()
    with torch.no_grad():
        ad.fc1.weight.mul_(8); ad.fc2.weight.mul_(8); ad.fc1.bias.normal_(0, .1); ad.fc2.bias.normal_(0, .1)
    x = torch.randn(5, 13, 20)
    lens = torch.tensor([13, 1, 6, 9, 4])
    mask = (torch.arange(13)[None, :] < lens[:, None]).long()
    y_ref = ad(x)
    tr = R.adapter_rows(x, ad.fc1.weight, ad.fc1.bias, ad.fc2.weight, ad.fc2.bias)
    _close(tr.y, y_ref)
    for fn in ("last", "mean", "std", "mix"):
        _close(R.readout(y_ref, mask, fn), ref.readout_embeddings(y_ref, mask, fn))
    p = torch.nn.functional.normalize(ref.readout_embeddings(y_ref, mask, "mix"), dim=-1)
    t = torch.nn.functional.normalize(torch.randn(5, 56), dim=-1)
    _close(R.infonce_rows(p, t, torch.arange(5), 0.05), ref.BatchInfoNCELoss()(p, t))
    mine, _ = R.l2_normalize(R.readout(tr.y, mask, "mix"))
    _close(mine, p)
