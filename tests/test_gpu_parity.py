"""GPU parity tests (run on the B200 box: `pytest -m gpu`).  Every computation under test goes
through the C-ABI CUDA library (include/p2t_b200.h), either directly via the ctypes layer
(`p2t_b200._core`) or through the reference-shaped public API; the checker is

  * tests/golden/grid_*.npz  — outputs of the REAL reference (float32, CPU) on inputs that lie on
    the bf16 grid, dumped by oracle/make_golden.py, and
  * oracle/restatement.py    — the CPU restatement, on seeded synthetic inputs.

Tolerances are BASELINE.json's north_star: loss within 1e-3 relative, gradients cosine >= 0.999 and
max relative error <= 1e-2 (max |g - g_ref| / max |g_ref|), integer/index results (row plan,
gather, argmax retrieval) bit-exact.
"""
import importlib
import math
import os
import sys

import numpy as np
import pytest
import torch

from oracle import restatement as R

pytestmark = pytest.mark.gpu

LOSS_RTOL = 1e-3
GRAD_COS = 0.999
GRAD_MAXREL = 1e-2
PARAMS = ("fc1.weight", "fc1.bias", "fc2.weight", "fc2.bias")


# --------------------------------------------------------------------------------------------------
# helpers
# --------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.fail("-m gpu tests need a CUDA device: the product path has no CPU fallback")
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def core(p2t):
    p2t._lib.load()
    return sys.modules["p2t_b200._core"]


@pytest.fixture(scope="module")
def synth(p2t):
    return importlib.import_module("p2t_b200.synth")


def _load(golden_dir, name):
    z = np.load(os.path.join(golden_dir, name))
    return {k: torch.from_numpy(np.asarray(z[k])) for k in z.files}


def maxrel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-300)).item()


def cosine(a, b):
    a, b = a.detach().double().cpu().flatten(), b.detach().double().cpu().flatten()
    return torch.nn.functional.cosine_similarity(a, b, dim=0).item()


def assert_grads(mine: dict, ref: dict, what=""):
    for k in PARAMS:
        c, m = cosine(mine[k], ref[k]), maxrel(mine[k], ref[k])
        assert c >= GRAD_COS and m <= GRAD_MAXREL, f"{what} grad {k}: cosine {c:.6f} maxrel {m:.3e}"


def make_adapter(p2t, dev, w1, b1, w2, b2, train=False, p=0.3):
    cfg = p2t.ModalityAdapterConfig(input_dim=w1.shape[1], intermediate_dim=w1.shape[0], output_dim=w2.shape[0],
                                    dropout_rate=p)
    ad = p2t.ModalityAdapter(cfg).to(dev).to(torch.bfloat16)
    with torch.no_grad():
        ad.fc1.weight.copy_(w1); ad.fc1.bias.copy_(b1); ad.fc2.weight.copy_(w2); ad.fc2.bias.copy_(b2)
    return ad.train() if train else ad.eval()


def adapter_grads(ad):
    return {"fc1.weight": ad.fc1.weight.grad, "fc1.bias": ad.fc1.bias.grad, "fc2.weight": ad.fc2.weight.grad,
            "fc2.bias": ad.fc2.bias.grad}


def bf(t):
    return t.to(torch.bfloat16)


# --------------------------------------------------------------------------------------------------
# the extension really is what runs
# --------------------------------------------------------------------------------------------------
def test_cuda_library_is_loaded_and_launches_kernels(p2t, core, dev):
    lib = p2t._lib
    assert os.path.basename(lib.LIB_PATH) == "libp2t_b200.so" and lib.load().p2t_abi_version() == 2
    lib.reset_launch_count()
    out = core.gemm(bf(torch.randn(128, 64, device=dev)), bf(torch.randn(256, 64, device=dev)), 128, 256, 64)
    torch.cuda.synchronize()
    assert lib.launch_count() == 1 and out.shape == (128, 256)
    assert torch.cuda.get_device_capability(dev)[0] == 10, "built for sm_100a only"


def test_abi_rejects_bad_arguments_on_the_device(p2t, core, dev):
    a = bf(torch.randn(64, 72, device=dev))
    with pytest.raises(p2t.P2TError, match="16-byte aligned"):  # row stride 2*(72+1) bytes is not a multiple of 16
        core.gemm(torch.empty(64, 73, dtype=torch.bfloat16, device=dev)[:, :72], a, 64, 64, 72)
    with pytest.raises(p2t.P2TError, match="multiples of 8"):
        x = bf(torch.randn(256, 12, device=dev))
        n = torch.tensor([256], dtype=torch.int32, device=dev)
        core.adapter_forward(x, 256, 256, n, bf(torch.randn(16, 12, device=dev)), bf(torch.zeros(16, device=dev)),
                             bf(torch.randn(8, 16, device=dev)), bf(torch.zeros(8, device=dev)), 0.0, 0, False)
    with pytest.raises(p2t.P2TError, match="bfloat16"):
        p2t.readout_embeddings(torch.randn(2, 4, 8, device=dev), torch.ones(2, 4, dtype=torch.long, device=dev), "mix")
    with pytest.raises(ValueError):
        p2t.readout_embeddings(bf(torch.randn(2, 4, 8, device=dev)), torch.ones(2, 4, dtype=torch.long, device=dev), "max")


# --------------------------------------------------------------------------------------------------
# tcgen05 GEMM
# --------------------------------------------------------------------------------------------------
GEMM_CASES = [
    # m, n, k, a_mn, b_mn, f32
    (256, 256, 128, 0, 0, 0), (128, 256, 64, 0, 0, 1), (1000, 520, 328, 0, 0, 0), (8, 8, 8, 0, 0, 1),
    (512, 512, 256, 0, 1, 0), (512, 512, 256, 1, 0, 0), (512, 512, 256, 1, 1, 0), (304, 264, 1000, 1, 1, 1),
    (2195, 2048, 320, 0, 0, 0), (2048, 320, 2195 // 8 * 8, 1, 1, 0), (4096, 4096, 4096, 0, 0, 0),
]


@pytest.mark.parametrize("cta_group", [1, 2])
@pytest.mark.parametrize("m,n,k,a_mn,b_mn,f32", GEMM_CASES)
def test_gemm_matches_fp64_matmul(core, dev, cta_group, m, n, k, a_mn, b_mn, f32):
    g = torch.Generator(device="cpu").manual_seed(m * 7 + n * 3 + k)
    A = bf(torch.randn(m, k, generator=g)).to(dev)
    B = bf(torch.randn(n, k, generator=g)).to(dev)
    a_store = A.t().contiguous() if a_mn else A
    b_store = B.t().contiguous() if b_mn else B
    out = core.gemm(a_store, b_store, m, n, k, a_mn=bool(a_mn), b_mn=bool(b_mn),
                    out_dtype=torch.float32 if f32 else torch.bfloat16, alpha=0.5, cta_group=cta_group)
    ref = 0.5 * (A.double() @ B.double().t())
    scale = ref.abs().max().item()
    err = (out.double() - ref).abs().max().item()
    # fp32 accumulation of exact bf16 products; a bf16 result adds one rounding (2^-9 relative)
    assert err <= (2e-5 if f32 else 4.5e-3) * scale, f"max err {err:.3e} vs scale {scale:.3e}"


@pytest.mark.parametrize("cta_group", [1, 2])
@pytest.mark.parametrize("m,n,k,a_mn,b_mn,f32", [
    (2048, 2560, 18143 // 8 * 8, 1, 1, 0),   # dW1 at config 2: 80 tiles of 256x256 for 74 CTA pairs
    (4096, 2048, 4000, 1, 1, 0),             # dW2 shape, shorter K
    (512, 512, 8192, 0, 0, 1),               # 4 tiles: every tile is cut 16 ways
    (300, 704, 2048, 0, 1, 1),               # ragged M and N
])
def test_gemm_split_k_tail_matches_fp64_matmul_and_is_deterministic(core, dev, cta_group, m, n, k, a_mn, b_mn, f32):
    g = torch.Generator(device="cpu").manual_seed(m + n + k)
    A = bf(torch.randn(m, k, generator=g)).to(dev)
    B = bf(torch.randn(n, k, generator=g)).to(dev)
    a_store = A.t().contiguous() if a_mn else A
    b_store = B.t().contiguous() if b_mn else B
    kw = dict(a_mn=bool(a_mn), b_mn=bool(b_mn), out_dtype=torch.float32 if f32 else torch.bfloat16, cta_group=cta_group)
    out = core.gemm(a_store, b_store, m, n, k, streamk=True, **kw)
    again = core.gemm(a_store, b_store, m, n, k, streamk=True, **kw)
    ref = A.double() @ B.double().t()
    scale = ref.abs().max().item()
    err = (out.double() - ref).abs().max().item()
    assert err <= (2e-5 if f32 else 4.5e-3) * scale, f"max err {err:.3e} vs scale {scale:.3e}"
    assert torch.equal(out, again)  # fixed summation order of the partial tiles


@pytest.mark.parametrize("m_actual", [700, 2195, 4000])
def test_gemm_split_k_tail_chosen_on_the_device_for_ragged_m(p2t, core, dev, m_actual):
    """M read from device memory + split-K scratch: the kernel derives the tile count and the cut itself (same cost
    model as the launcher) — opt-in for the ragged adapter GEMMs (P2T_SPLITK_DYN=1), exact and deterministic."""
    m_cap, n, k = 4096, 512, 1024
    g = torch.Generator().manual_seed(m_actual)
    A = bf(torch.randn(m_cap, k, generator=g)).to(dev)
    Bm = bf(torch.randn(n, k, generator=g)).to(dev)
    dyn = torch.tensor([m_actual], dtype=torch.int32, device=dev)
    outs = []
    for _ in range(2):
        out = torch.zeros(m_cap, n, dtype=torch.float32, device=dev)
        ws = core.gemm_workspace(dev)
        p2t._lib.call("p2t_gemm_bf16", A.data_ptr(), k, 0, Bm.data_ptr(), k, 0, out.data_ptr(), n, 1, m_cap, n, k, 1.0,
                      dyn.data_ptr(), None, ws.data_ptr(), 2, None)
        outs.append(out)
    ref = A[:m_actual].double() @ Bm.double().t()
    assert (outs[0][:m_actual].double() - ref).abs().max().item() <= 2e-5 * ref.abs().max().item()
    assert torch.equal(outs[0], outs[1])


def test_gemm_device_side_extents(core, dev):
    """dyn_m / dyn_k are read from device memory: rows >= dyn_m are not written, the K loop stops at dyn_k."""
    m, n, k = 700, 256, 512
    A = bf(torch.randn(m, k, device=dev))
    B = bf(torch.randn(n, k, device=dev))
    dyn_m = torch.tensor([333], dtype=torch.int32, device=dev)
    out = torch.full((m, n), 7.0, dtype=torch.bfloat16, device=dev)
    core._lib.call("p2t_gemm_bf16", A.data_ptr(), k, 0, B.data_ptr(), k, 0, out.data_ptr(), n, 0, m, n, k, 1.0,
                   dyn_m.data_ptr(), None, None, 2, torch.cuda.current_stream().cuda_stream)
    ref = A[:333].double() @ B.double().t()
    assert maxrel(out[:333], ref) < 4.5e-3 and bool((out[333:] == 7.0).all())
    # K from the device: operands are [K][rows] (the weight-gradient layout), rows >= dyn_k zero up to the 64-block
    kk, m2, n2 = 1024, 256, 264
    At = bf(torch.randn(kk, m2, device=dev))
    Bt = bf(torch.randn(kk, n2, device=dev))
    real_k = 517
    At[real_k:] = 0
    Bt[real_k:] = 0
    dyn_k = torch.tensor([real_k], dtype=torch.int32, device=dev)
    out2 = core.gemm(At, Bt, m2, n2, kk, a_mn=True, b_mn=True, out_dtype=torch.float32, dyn_k=dyn_k)
    ref2 = At[:real_k].double().t() @ Bt[:real_k].double()
    assert maxrel(out2, ref2) < 2e-5


# --------------------------------------------------------------------------------------------------
# ragged row plan + gather: integer / byte work, bit-exact
# --------------------------------------------------------------------------------------------------
def _masks(B, L, seed):
    g = torch.Generator().manual_seed(seed)
    lens = torch.randint(0, L + 1, (B,), generator=g)
    lens[0], lens[-1] = L, 0  # a full row and an EMPTY sequence
    ar = torch.arange(L)[None, :]
    return {"right": ar < lens[:, None], "left": ar >= (L - lens)[:, None],
            "holes": torch.rand(B, L, generator=g) > 0.5, "ones": torch.ones(B, L, dtype=torch.bool),
            "none": torch.zeros(B, L, dtype=torch.bool)}


@pytest.mark.parametrize("kind", ["right", "left", "holes", "ones", "none"])
@pytest.mark.parametrize("dtype", [torch.int64, torch.int32, torch.uint8, torch.bool, torch.float32])
def test_row_plan_and_gather_are_bit_exact(core, dev, kind, dtype):
    B, L, D = 7, 333, 40
    mask = _masks(B, L, 5)[kind]
    plan = core.plan_rows(mask.to(dtype).to(dev))
    counts = mask.sum(1)
    assert torch.equal(plan.counts.cpu().long(), counts)
    assert torch.equal(plan.seq_off.cpu().long(), torch.cat([torch.zeros(1, dtype=torch.long), counts.cumsum(0)]))
    chunks = (counts + core.CHUNK_ROWS - 1) // core.CHUNK_ROWS
    assert torch.equal(plan.chunk_off.cpu().long(), torch.cat([torch.zeros(1, dtype=torch.long), chunks.cumsum(0)]))
    n = int(plan.n_rows.item())
    assert n == int(mask.sum())
    flat = mask.flatten().nonzero().flatten()
    assert torch.equal(plan.row_src[:n].cpu().long(), flat)
    nch = int(chunks.sum())
    desc = plan.chunk_seq[:nch].cpu().long()  # {first packed row, end row, sequence, 0} per 64-row pooling chunk
    seq = torch.repeat_interleave(torch.arange(B), chunks)
    first = torch.cat([torch.zeros(1, dtype=torch.long), counts.cumsum(0)])[seq] + 64 * (torch.arange(nch) - torch.cat([torch.zeros(1, dtype=torch.long), chunks.cumsum(0)])[seq])
    end = torch.minimum(first + 64, counts.cumsum(0)[seq])
    assert torch.equal(desc[:, 2], seq) and torch.equal(desc[:, 0], first) and torch.equal(desc[:, 1], end)
    x = bf(torch.randn(B * L, D)).to(dev)
    xp = core.gather_rows(x, plan)
    assert torch.equal(xp[:n].cpu(), x.cpu()[flat])
    n_pad = min(plan.rows_cap, (n + 255) // 256 * 256)
    assert bool((xp[n:n_pad] == 0).all())


# --------------------------------------------------------------------------------------------------
# ModalityAdapter (module API) — models/modeling_esm2llama_instruct.py:45-68
# --------------------------------------------------------------------------------------------------
def test_adapter_matches_reference_golden(p2t, dev, golden_dir):
    g = _load(golden_dir, "grid_adapter.npz")
    ad = make_adapter(p2t, dev, *(g["sd." + k] for k in PARAMS))
    x = bf(g["x"]).to(dev).requires_grad_()
    y = ad(x)
    assert y.dtype == torch.bfloat16 and y.shape == g["y"].shape
    assert maxrel(y, g["y"]) <= 6e-3  # bf16 output: 2^-9 relative rounding on top of the bf16 h1 operand
    (y.float() * g["gy"].to(dev)).sum().backward()
    assert_grads(adapter_grads(ad), {k: g["grad." + k] for k in PARAMS}, "golden adapter")
    assert cosine(x.grad, g["grad.x"]) >= GRAD_COS and maxrel(x.grad, g["grad.x"]) <= GRAD_MAXREL
    assert ad.ln1.weight.grad is None and ad.ln2.weight.grad is None  # never applied (reference :56-57)


@pytest.mark.parametrize("shape", [(1, 1, 64, 128, 96), (3, 50, 320, 512, 264), (2, 700, 1152, 2048, 3584)])
def test_adapter_matches_oracle_random(p2t, dev, shape):
    B, L, d_in, d_mid, d_out = shape
    g = torch.Generator().manual_seed(sum(shape))
    gain = 1.0 / (0.02 * math.sqrt(d_in))
    w1, w2 = bf(torch.randn(d_mid, d_in, generator=g) * 0.02 * gain), bf(torch.randn(d_out, d_mid, generator=g) * 0.04)
    b1, b2 = bf(torch.randn(d_mid, generator=g) * 0.1), bf(torch.randn(d_out, generator=g) * 0.1)
    x, gy = bf(torch.randn(B, L, d_in, generator=g)), bf(torch.randn(B, L, d_out, generator=g))
    ad = make_adapter(p2t, dev, w1, b1, w2, b2)
    xg = x.to(dev).requires_grad_()
    y = ad(xg)
    (y.float() * gy.to(dev).float()).sum().backward()
    f = torch.float32
    tr = R.adapter_rows(x.to(f).reshape(-1, d_in), w1.to(f), b1.to(f), w2.to(f), b2.to(f))
    ref = R.adapter_rows_backward(tr, gy.to(f).reshape(-1, d_out), w1.to(f), w2.to(f), need_dx=True)
    assert maxrel(y, tr.y.view(B, L, d_out)) <= 6e-3
    torch.testing.assert_close(y.float().norm(dim=-1).cpu(), torch.ones(B, L), rtol=0, atol=4e-3)  # unit rows (:67)
    assert_grads(adapter_grads(ad), ref, f"adapter {shape}")
    assert cosine(xg.grad, ref["dx"]) >= GRAD_COS and maxrel(xg.grad, ref["dx"].view(B, L, d_in)) <= GRAD_MAXREL


def test_adapter_no_grad_and_eval_train_switch(p2t, dev):
    g = torch.Generator().manual_seed(3)
    ad = make_adapter(p2t, dev, bf(torch.randn(64, 32, generator=g) * 0.2), bf(torch.zeros(64)),
                      bf(torch.randn(48, 64, generator=g) * 0.2), bf(torch.zeros(48)), p=0.5)
    x = bf(torch.randn(2, 9, 32, generator=g)).to(dev)
    with torch.no_grad():
        y0, y1 = ad(x), ad(x)
    assert torch.equal(y0, y1) and not y0.requires_grad
    ad.train()
    torch.manual_seed(11); ya = ad(x)
    torch.manual_seed(11); yb = ad(x)
    yc = ad(x)
    assert torch.equal(ya, yb) and not torch.equal(ya, yc) and not torch.equal(ya, y0)  # Philox mask follows torch's seed
    torch.testing.assert_close(ya.float().norm(dim=-1).cpu(), torch.ones(2, 9), rtol=0, atol=4e-3)


# --------------------------------------------------------------------------------------------------
# readout_embeddings — scripts/train_contrast.py:198-248
# --------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("mname", ["right", "left", "holes"])
@pytest.mark.parametrize("fn", ["last", "mean", "std", "mix"])
def test_readout_matches_reference_golden(p2t, dev, golden_dir, mname, fn):
    if fn == "last" and mname != "right":
        pytest.skip("reference defines 'last' for right padding only (:208-209)")
    g = _load(golden_dir, "grid_readout.npz")
    emb = bf(g["emb"]).to(dev).requires_grad_()
    mask = g["mask_" + mname].to(dev)
    out = p2t.readout_embeddings(emb, mask, fn)
    ref = g[f"out.{mname}.{fn}"]
    assert out.dtype == torch.bfloat16 and out.shape == ref.shape
    assert maxrel(out, ref) <= 4e-3  # one bf16 rounding of the result
    (gemb,) = torch.autograd.grad((out.float() * g[f"gout.{mname}.{fn}"].to(dev)).sum(), emb)
    gref = g[f"gemb.{mname}.{fn}"]
    assert cosine(gemb, gref) >= GRAD_COS and maxrel(gemb, gref) <= GRAD_MAXREL
    assert bool((gemb.cpu()[g["mask_" + mname] == 0] == 0).all())  # padded positions get exactly zero


def test_readout_last_backward_is_an_exact_scatter(p2t, dev):
    """'last' backward (csrc/rows.cu readout_last_bwd_kernel): dx[b, len_b - 1] = dout[b], zero elsewhere, bit-exact;
    an all-masked row indexes -1 like the reference's advanced indexing does (scripts/train_contrast.py:207-215)."""
    gen = torch.Generator().manual_seed(3)
    B, S, D = 5, 37, 72
    emb = bf(torch.randn(B, S, D, generator=gen)).to(dev).requires_grad_()
    lens = torch.tensor([37, 1, 20, 0, 9])
    mask = (torch.arange(S)[None, :] < lens[:, None]).long().to(dev)
    out = p2t.readout_embeddings(emb, mask, "last")
    dout = bf(torch.randn(B, D, generator=gen)).to(dev)
    (gx,) = torch.autograd.grad(out, emb, dout)
    want = torch.zeros(B, S, D, dtype=torch.bfloat16, device=dev)
    want[torch.arange(B), (lens - 1) % S] = dout
    assert torch.equal(gx, want)
    assert torch.equal(out, emb.detach()[torch.arange(B), (lens - 1) % S])


def test_readout_degenerate_sequences_follow_the_reference(p2t, dev):
    """std of a constant / single-token sequence is 0 with a NaN gradient; an all-masked row is 0/0 = NaN."""
    emb = bf(torch.randn(3, 6, 16, generator=torch.Generator().manual_seed(1)))
    emb[1] = emb[1, :1]  # constant sequence
    mask = torch.tensor([[1, 0, 0, 0, 0, 0], [1, 1, 1, 1, 0, 0], [0, 0, 0, 0, 0, 0]])
    e = emb.to(dev).requires_grad_()
    out = p2t.readout_embeddings(e, mask.to(dev), "mix")
    ref = R.readout(emb.float(), mask, "mix")
    torch.testing.assert_close(out.float().cpu(), ref.to(torch.bfloat16).float(), rtol=0, atol=0, equal_nan=True)
    assert bool((out[:2, 16:] == 0).all()) and bool(out[2].isnan().all())
    (gr,) = torch.autograd.grad(out[:2].float().sum(), e)
    assert bool(gr[0, 0].isnan().all()) and bool(gr[1, :4].isnan().all())


def test_readout_full_size_mean_is_linear_and_std_is_shift_invariant(p2t, dev):
    """Size-independent properties at config-2 text size (32 x 256 x 4096)."""
    g = torch.Generator().manual_seed(9)
    x1, x2 = bf(torch.randn(32, 256, 4096, generator=g)).to(dev), bf(torch.randn(32, 256, 4096, generator=g)).to(dev)
    lens = torch.randint(16, 257, (32,), generator=g)
    mask = (torch.arange(256)[None, :] < lens[:, None]).long().to(dev)
    m1 = p2t.readout_embeddings(x1, mask, "mean").float()
    m2 = p2t.readout_embeddings(x2, mask, "mean").float()
    m12 = p2t.readout_embeddings(x1 + x2, mask, "mean").float()  # x1 + x2 rounds to bf16: 2^-9 relative per element
    assert maxrel(m12, m1 + m2) <= 1e-2
    shifted = (x1.float() + 3.0).to(torch.bfloat16)  # shift by a bf16-exact constant (values re-round, std must agree)
    s1 = p2t.readout_embeddings(x1, mask, "std").float()
    s2 = p2t.readout_embeddings(shifted, mask, "std").float()
    assert maxrel(s2, s1) <= 1e-2
    mix = p2t.readout_embeddings(x1, mask, "mix")
    assert torch.equal(mix[:, :4096], p2t.readout_embeddings(x1, mask, "mean"))
    assert torch.equal(mix[:, 4096:], p2t.readout_embeddings(x1, mask, "std"))


# --------------------------------------------------------------------------------------------------
# InfoNCE — scripts/train_contrast.py:72-114
# --------------------------------------------------------------------------------------------------
def test_losses_match_reference_golden(p2t, dev, golden_dir):
    g = _load(golden_dir, "grid_losses.npz")
    p = bf(g["p"]).to(dev).requires_grad_()
    t = bf(g["t"]).to(dev).requires_grad_()
    for name, a, b in (("batch", p, t), ("swapped", t, p)):
        loss = p2t.BatchInfoNCELoss()(a, b)
        assert loss.dtype == torch.float32 and loss.dim() == 0
        assert abs(loss.item() - g[name + ".loss"].item()) <= 1e-5 * abs(g[name + ".loss"].item()) + 1e-7
        gp, gt = torch.autograd.grad(loss, (p, t))
        assert gp.dtype == torch.bfloat16
        assert maxrel(gp, g[name + ".gp"]) <= 4e-3 and maxrel(gt, g[name + ".gt"]) <= 4e-3  # one bf16 rounding
    labels = g["seg.labels"].to(dev)
    seg = p2t.SegmentedBatchInfoNCELoss(temperature=float(g["seg.temperature"]))(p[3:8], t, labels)
    assert abs(seg.item() - g["seg.loss"].item()) <= 1e-5 * g["seg.loss"].item()
    gp, gt = torch.autograd.grad(seg, (p, t))
    assert maxrel(gp, g["seg.gp"]) <= 4e-3 and maxrel(gt, g["seg.gt"]) <= 4e-3
    assert bool((gp[:3] == 0).all()) and bool((gp[8:] == 0).all())
    sym = p2t.SymmetricInfoNCELoss()(p, t)
    assert abs(sym.item() - 0.5 * (g["batch.loss"].item() + g["swapped.loss"].item())) <= 1e-5


def test_retrieval_argmax_is_bit_exact(core, dev, golden_dir):
    g = _load(golden_dir, "grid_losses.npz")
    p, t = bf(g["p"]).to(dev), bf(g["t"]).to(dev)
    labels = torch.arange(p.shape[0], dtype=torch.int32, device=dev)
    res = core.infonce_forward(p, t, labels, 0.05, need_grad=False, want_col_argmax=True)
    assert torch.equal(res.argmax_row.cpu().long(), g["argmax_row"])
    assert torch.equal(res.argmax_col.cpu().long(), g["argmax_col"])


@pytest.mark.parametrize("R_,C_,E_", [(5, 7, 64), (33, 130, 256), (512, 512, 1024), (256, 1024, 2048), (300, 1000, 512),
                                      (520, 1000, 264)])
@pytest.mark.parametrize("sym", [False, True])
def test_infonce_matches_oracle(core, dev, R_, C_, E_, sym):
    """Small problems run on CUDA cores in fp32; large ones (R*C*E > 2^26) on the tcgen05 GEMM with the online-softmax
    statistics reduced in its epilogue (no fp32 logits in memory) and bf16 dLogits from a recomputing second GEMM —
    including blocks whose extents are not multiples of the 256 x 256 tile."""
    g = torch.Generator().manual_seed(R_ + C_)
    t = torch.nn.functional.normalize(torch.randn(C_, E_, generator=g), dim=-1)
    labels = torch.randperm(C_, generator=g)[:R_]
    p = torch.nn.functional.normalize(t[labels] + 2.5 * torch.randn(R_, E_, generator=g) / math.sqrt(E_), dim=-1)
    p, t = bf(p), bf(t)
    wr, wc = (0.5, 0.5) if sym else (1.0, 0.0)
    res = core.infonce_forward(p.to(dev), t.to(dev), labels.to(dev), 0.05, w_row=wr, w_col=wc, want_col_argmax=True)
    dp, dt = core.infonce_backward(res, p.to(dev), t.to(dev), 0.05, need_dt=True)
    pf, tf = p.double(), t.double()
    ref = wr * R.infonce_rows(pf, tf, labels, 0.05) + (wc * R.infonce_cols(pf, tf, labels, 0.05) if wc else 0.0)
    _, dpo, dto = R.infonce_backward(pf, tf, labels, 0.05, wr, wc)
    assert abs(res.loss.item() - float(ref)) <= 1e-4 * abs(float(ref)) + 1e-6
    big = R_ * C_ * E_ > (1 << 26)
    if big:
        assert res.dS is None and res.dS_bf16 is not None and res.dS_bf16.shape == (R_, C_)  # logits never materialised
    tol = 8e-3 if big else 5e-4  # bf16 dS operand on the tensor-core path; fp32 cancellation in (softmax - 1) otherwise
    assert cosine(dp, dpo) >= 0.9999 and maxrel(dp, dpo) <= tol
    assert cosine(dt, dto) >= 0.9999 and maxrel(dt, dto) <= tol
    am_r, am_c = R.retrieval_argmax(pf, tf)
    assert torch.equal(res.argmax_row.cpu().long(), am_r)  # planted positives: margins >> fp32 noise
    # unlabelled columns hold only random negatives (no margin): compare the columns that have a partner row
    assert torch.equal(res.argmax_col.cpu().long()[labels], am_c[labels])


# --------------------------------------------------------------------------------------------------
# the fused step — scripts/train_contrast.py:313-379 + :448
# --------------------------------------------------------------------------------------------------
def _golden_step_inputs(g, nseg):
    """The padded batch and the mask the fork's step effectively uses: every segment is
    zero-padded to ITS OWN max length and pooled with an all-ones mask."""
    lens = [int(v) for v in g["lens"]]
    B, d_in = len(lens), g["x0"].shape[1]
    seg = B // nseg
    L = max(lens)
    x = torch.zeros(B, L, d_in)
    mask = torch.zeros(B, L, dtype=torch.long)
    for b, n in enumerate(lens):
        x[b, :n] = g[f"x{b}"]
        s = min(b // seg, nseg - 1) if b < seg * nseg else None
        ids = range(s * seg, (s + 1) * seg) if s is not None else [b]
        mask[b, : max(lens[i] for i in ids)] = 1
    return x, mask


@pytest.mark.parametrize("nseg", [1, 2, 3, 4])
def test_step_matches_reference_teacher_forcing_forward_pass(p2t, dev, golden_dir, nseg):
    g = _load(golden_dir, "grid_step.npz")
    ad = make_adapter(p2t, dev, *(g["sd." + k] for k in PARAMS))
    x, mask = _golden_step_inputs(g, nseg)
    aux = p2t.StepAux()
    loss = p2t.contrastive_step(bf(x).to(dev), mask.to(dev), ad, bf(g["text"]).to(dev), g["tmask"].to(dev),
                                contrastive_num_segments=nseg, aux=aux)
    loss.backward()
    ref = g[f"seg{nseg}.loss"].item()
    assert loss.dtype == torch.float32 and abs(loss.item() - ref) <= LOSS_RTOL * abs(ref)
    assert_grads(adapter_grads(ad), {k: g[f"seg{nseg}.grad.{k}"] for k in PARAMS}, f"golden step nseg={nseg}")
    if nseg == 1:
        assert maxrel(aux.protein_embeddings, g["p"]) <= 6e-3


@pytest.mark.parametrize("nseg", [1, 2, 3, 4])
def test_integration_recipe_reproduces_the_reference_step(p2t, dev, golden_dir, nseg):
    """INTEGRATION.md §2 verbatim: the whole batch padded once to the batch's longest sequence, the pooling mask built
    by `segment_pooling_mask(lengths, num_segments)` (NOT torch.ones_like: the fork pools each segment over the
    segment's own padded length), against the outputs of the REAL teacher_forcing_forward_pass."""
    g = _load(golden_dir, "grid_step.npz")
    ad = make_adapter(p2t, dev, *(g["sd." + k] for k in PARAMS))
    lens = [int(v) for v in g["lens"]]
    B, L, d_in = len(lens), max(lens), g["x0"].shape[1]
    residues = torch.zeros(B, L, d_in)
    rmask = torch.zeros(B, L, dtype=torch.long)
    for b, n in enumerate(lens):
        residues[b, :n] = g[f"x{b}"]
        rmask[b, :n] = 1
    residues, rmask = bf(residues).to(dev), rmask.to(dev)
    mask = p2t.segment_pooling_mask(rmask.sum(dim=1), nseg, rmask.shape[1], device=rmask.device)
    assert torch.equal(mask.cpu(), _golden_step_inputs(g, nseg)[1])
    loss = p2t.contrastive_step(residues, mask, ad, bf(g["text"]).to(dev), g["tmask"].to(dev),
                                temperature=0.05, contrastive_num_segments=nseg)
    loss.backward()
    ref = g[f"seg{nseg}.loss"].item()
    assert abs(loss.item() - ref) <= LOSS_RTOL * abs(ref)
    assert_grads(adapter_grads(ad), {k: g[f"seg{nseg}.grad.{k}"] for k in PARAMS}, f"INTEGRATION recipe nseg={nseg}")
    if nseg == 4:  # ... and the all-ones mask over the BATCH length is a different (wrong) computation
        ad.zero_grad(set_to_none=True)
        wrong = p2t.contrastive_step(residues, torch.ones_like(rmask), ad, bf(g["text"]).to(dev), g["tmask"].to(dev),
                                     contrastive_num_segments=nseg)
        assert abs(wrong.item() - ref) > LOSS_RTOL * abs(ref)


STEP_CASES = [
    # workload, weight_gain, kwargs
    ("tiny", 6.0, {}), ("tiny", 6.0, {"left_pad": True}), ("tiny", 6.0, {"symmetric": True}),
    ("tiny", 6.0, {"nseg": 3}), ("cfg1_esm2_t6_llama1b", 2.5, {}), ("cfg1_esm2_t6_llama1b", 2.5, {"symmetric": True}),
    ("cfg2_esm2_3b_llama8b", 1.0, {}), ("cfg4_esmc600m_qwen7b", 1.5, {"batch": 16}),
    ("cfg4_esmc600m_qwen7b", 1.5, {}),  # BASELINE config 4 at its own per-rank size: 64 pairs
]


@pytest.mark.parametrize("workload,gain,kw", STEP_CASES)
def test_step_matches_oracle(p2t, synth, dev, workload, gain, kw):
    kw = dict(kw)
    sym, nseg = kw.pop("symmetric", False), kw.pop("nseg", 1)
    sb = synth.make_config_batch(workload, weight_gain=gain, **kw)
    ad = make_adapter(p2t, dev, sb.w1, sb.b1, sb.w2, sb.b2)
    aux = p2t.StepAux()
    loss = p2t.contrastive_step(sb.x.to(dev), sb.prot_mask.to(dev), ad, sb.text.to(dev), sb.text_mask.to(dev),
                                symmetric=sym, contrastive_num_segments=nseg, aux=aux)
    loss.backward()
    f = torch.float32
    st = R.step_forward(sb.x.to(f), sb.prot_mask, sb.w1.to(f), sb.b1.to(f), sb.w2.to(f), sb.b2.to(f), sb.text.to(f),
                        sb.text_mask, 0.05, nseg, sym)
    ref = R.step_backward(st, sb.x.to(f), sb.prot_mask, sb.w1.to(f), sb.w2.to(f), 0.05, nseg, sym)
    assert not any(v.isnan().any() for v in ref.values()), "test inputs must keep the oracle finite"
    assert abs(loss.item() - st.loss.item()) <= LOSS_RTOL * abs(st.loss.item())
    assert_grads(adapter_grads(ad), ref, f"{workload} {kw}")
    assert int(aux.n_rows.item()) == int(sb.prot_mask.sum())
    assert maxrel(aux.protein_embeddings, st.p) <= 6e-3


def test_step_retrieval_indices_are_bit_exact_with_planted_pairs(p2t, synth, dev):
    """argmax retrieval (protein->text rows, text->protein columns) against torch.argmax on the oracle
    logits; the text side carries a planted copy of the protein embedding so margins are real."""
    sb = synth.make_config_batch("cfg1_esm2_t6_llama1b", weight_gain=2.5)
    f = torch.float32
    tr = R.adapter_rows(sb.x.to(f), sb.w1.to(f), sb.b1.to(f), sb.w2.to(f), sb.b2.to(f))
    p, _ = R.l2_normalize(R.readout(tr.y, sb.prot_mask, "mix"))
    g = torch.Generator().manual_seed(4)
    perm = torch.randperm(p.shape[0], generator=g)  # text j is the partner of protein perm^-1(j)
    t = torch.nn.functional.normalize(p[perm] + 0.02 * torch.randn(p.shape, generator=g) / math.sqrt(p.shape[1]), dim=-1)
    ad = make_adapter(p2t, dev, sb.w1, sb.b1, sb.w2, sb.b2)
    aux = p2t.StepAux()
    labels = torch.argsort(perm).to(torch.int32).to(dev)
    loss = p2t.contrastive_step(sb.x.to(dev), sb.prot_mask.to(dev), ad, text_embeds=t.to(dev), labels=labels, aux=aux)
    am_r, am_c = R.retrieval_argmax(p, t)
    assert torch.equal(am_r, torch.argsort(perm)) and torch.equal(am_c, perm)
    top2 = (p @ t.t() / 0.05).topk(2, dim=1).values
    assert (top2[:, 0] - top2[:, 1]).min() > 0.02, "planted margin too small for a meaningful bit-exact check"
    assert torch.equal(aux.argmax_row.cpu().long(), am_r) and torch.equal(aux.argmax_col.cpu().long(), am_c)
    ref = R.infonce_rows(p, t, torch.argsort(perm), 0.05)
    assert abs(loss.item() - ref.item()) <= LOSS_RTOL * abs(ref.item()) + 1e-5


def test_step_with_dropout_matches_oracle_given_the_kernel_masks(p2t, core, synth, dev):
    """Training mode: the fused Philox masks are exported (p2t_dropout_mask) and fed to the oracle."""
    sb = synth.make_config_batch("cfg1_esm2_t6_llama1b", weight_gain=2.5)
    ad = make_adapter(p2t, dev, sb.w1, sb.b1, sb.w2, sb.b2, train=True, p=0.3)
    adapter_mod = importlib.import_module("p2t_b200.adapter")
    torch.manual_seed(2024)
    seed = adapter_mod._draw_seed()
    torch.manual_seed(2024)
    loss = p2t.contrastive_step(sb.x.to(dev), sb.prot_mask.to(dev), ad, sb.text.to(dev), sb.text_mask.to(dev))
    loss.backward()
    n = int(sb.prot_mask.sum())
    d_mid, d_out = sb.w1.shape[0], sb.w2.shape[0]
    k1 = core.dropout_mask(n, d_mid, 0.3, seed, 1, dev).cpu()
    k2 = core.dropout_mask(n, d_out, 0.3, seed, 2, dev).cpu()
    assert abs((k1 > 0).float().mean().item() - 0.7) < 5e-3 and abs((k2 > 0).float().mean().item() - 0.7) < 5e-3
    assert k1.unique().tolist() == pytest.approx([0.0, 1 / 0.7])
    B, L = sb.prot_mask.shape
    valid = sb.prot_mask.bool()
    keep1, keep2 = torch.ones(B, L, d_mid), torch.ones(B, L, d_out)
    keep1[valid], keep2[valid] = k1, k2  # packed-row order == row-major order of the valid positions
    f = torch.float32
    st = R.step_forward(sb.x.to(f), sb.prot_mask, sb.w1.to(f), sb.b1.to(f), sb.w2.to(f), sb.b2.to(f), sb.text.to(f),
                        sb.text_mask, keep1=keep1, keep2=keep2)
    ref = R.step_backward(st, sb.x.to(f), sb.prot_mask, sb.w1.to(f), sb.w2.to(f))
    assert abs(loss.item() - st.loss.item()) <= LOSS_RTOL * abs(st.loss.item())
    assert_grads(adapter_grads(ad), ref, "dropout step")


def test_step_full_size_properties(p2t, synth, dev):
    """Config 2 at full size (32 pairs, ~18 k residue rows): size-independent properties."""
    sb = synth.make_config_batch("cfg2_esm2_3b_llama8b")
    ad = make_adapter(p2t, dev, sb.w1, sb.b1, sb.w2, sb.b2)
    x, pm, text, tm = sb.x.to(dev), sb.prot_mask.to(dev), sb.text.to(dev), sb.text_mask.to(dev)

    def run(x_, pm_, text_, tm_, **kw):
        ad.zero_grad(set_to_none=True)
        aux = p2t.StepAux()
        loss = p2t.contrastive_step(x_, pm_, ad, text_, tm_, aux=aux, **kw)
        loss.backward()
        return loss.detach().clone(), {k: v.clone() for k, v in adapter_grads(ad).items()}, aux

    l0, g0, a0 = run(x, pm, text, tm)
    l1, g1, _ = run(x, pm, text, tm)
    assert torch.equal(l0, l1) and all(torch.equal(g0[k], g1[k]) for k in PARAMS)  # run-to-run bit-identical
    assert abs(l0.item() - math.log(32)) < 0.2  # random pairs: loss ~ ln B
    # extra padding columns change nothing: the packed rows are the same
    xpad = torch.cat([x, torch.zeros(32, 77, x.shape[2], dtype=x.dtype, device=dev)], dim=1)
    mpad = torch.cat([pm, torch.zeros(32, 77, dtype=pm.dtype, device=dev)], dim=1)
    l2, g2, _ = run(xpad, mpad, text, tm)
    assert torch.equal(l0, l2) and all(torch.equal(g0[k], g2[k]) for k in PARAMS)
    # equal segments average to the full-batch loss (reference :356-379)
    l3, g3, _ = run(x, pm, text, tm, contrastive_num_segments=2)
    assert torch.equal(l0, l3)
    # permuting the pairs permutes the embeddings and leaves loss and weight gradients unchanged
    perm = torch.randperm(32, generator=torch.Generator().manual_seed(0)).to(dev)
    l4, g4, a4 = run(x[perm], pm[perm], text[perm], tm[perm])
    assert abs(l4.item() - l0.item()) <= 1e-5 * abs(l0.item())
    assert maxrel(a4.protein_embeddings, a0.protein_embeddings[perm]) <= 4e-3
    for k in PARAMS:
        assert cosine(g4[k], g0[k]) >= 0.9999 and maxrel(g4[k], g0[k]) <= GRAD_MAXREL
    # unit-norm embeddings on both sides
    torch.testing.assert_close(a0.protein_embeddings.float().norm(dim=-1).cpu(), torch.ones(32), rtol=0, atol=4e-3)
    torch.testing.assert_close(a0.text_embeddings.float().norm(dim=-1).cpu(), torch.ones(32), rtol=0, atol=4e-3)
    # symmetric loss of swapped roles is bounded the same way and finite
    l5, g5, _ = run(x, pm, text, tm, symmetric=True)
    assert math.isfinite(l5.item()) and all(torch.isfinite(v.float()).all() for v in g5.values())


def test_step_scales_with_upstream_gradient_and_accumulates_like_the_reference(p2t, synth, dev):
    """The caller adds losses into an fp32 accumulator, divides by accumulation steps and calls backward once
    (train_contrast.py:345,432,448): gradients must scale linearly and accumulate into .grad."""
    sb = synth.make_config_batch("tiny", weight_gain=6.0)
    ad = make_adapter(p2t, dev, sb.w1, sb.b1, sb.w2, sb.b2)
    args = (sb.x.to(dev), sb.prot_mask.to(dev), ad, sb.text.to(dev), sb.text_mask.to(dev))
    p2t.contrastive_step(*args).backward()
    g1 = {k: v.float().clone() for k, v in adapter_grads(ad).items()}
    ad.zero_grad(set_to_none=True)
    acc = torch.zeros((), dtype=torch.float32, device=dev)
    acc = acc + p2t.contrastive_step(*args) + p2t.contrastive_step(*args)
    (acc / 4).backward()
    for k in PARAMS:
        assert maxrel(adapter_grads(ad)[k], 0.5 * g1[k]) <= 1e-2


# --------------------------------------------------------------------------------------------------
# ragged hand-over format: packed rows + lengths, and the host stager
# --------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("left_pad", [False, True])
def test_packed_input_step_is_bit_identical_to_the_padded_step(p2t, synth, dev, left_pad):
    sb = synth.make_config_batch("cfg1_esm2_t6_llama1b", weight_gain=2.5, left_pad=left_pad)
    ad = make_adapter(p2t, dev, sb.w1, sb.b1, sb.w2, sb.b2)
    loss = p2t.contrastive_step(sb.x.to(dev), sb.prot_mask.to(dev), ad, sb.text.to(dev), sb.text_mask.to(dev))
    loss.backward()
    g_ref = {k: v.clone() for k, v in adapter_grads(ad).items()}
    ad.zero_grad(set_to_none=True)
    # (a) packed by hand on the host
    xr = sb.x[sb.prot_mask.bool()].to(dev)
    tr = sb.text[sb.text_mask.bool()].to(dev)
    l2 = p2t.contrastive_step(xr, None, ad, tr, None, residue_lengths=sb.prot_mask.sum(1), text_lengths=sb.text_mask.sum(1).to(dev))
    l2.backward()
    assert torch.equal(loss, l2) and all(torch.equal(g_ref[k], adapter_grads(ad)[k]) for k in PARAMS)
    ad.zero_grad(set_to_none=True)
    # (b) through the host stager: pinned padded host batch -> packed device rows on a copy stream
    stager = p2t.HostStager(dev)
    stager.submit(sb.x.pin_memory(), sb.prot_mask.pin_memory(), sb.text.pin_memory(), sb.text_mask.pin_memory())
    batch = stager.take()
    assert torch.equal(batch.residue_rows, xr) and torch.equal(batch.text_rows, tr)  # byte moves: exact
    assert torch.equal(batch.residue_lengths.cpu().long(), sb.prot_mask.sum(1))
    assert batch.h2d_bytes == xr.numel() * 2 + tr.numel() * 2 + 2 * 4 * sb.x.shape[0]
    l3 = p2t.contrastive_step(batch.residue_rows, None, ad, batch.text_rows, None,
                              residue_lengths=batch.residue_lengths, text_lengths=batch.text_lengths)
    l3.backward()
    assert torch.equal(loss, l3) and all(torch.equal(g_ref[k], adapter_grads(ad)[k]) for k in PARAMS)


def test_host_stager_rejects_masks_with_holes_and_device_tensors(p2t, dev):
    stager = p2t.HostStager(dev)
    x = bf(torch.randn(2, 6, 16))
    holes = torch.tensor([[1, 0, 1, 1, 0, 0], [1, 1, 1, 0, 0, 0]])
    ok = torch.tensor([[1, 1, 1, 1, 0, 0], [0, 0, 0, 1, 1, 1]])
    with pytest.raises(p2t.P2TError, match="contiguous"):
        stager.submit(x, holes, x, ok)
    with pytest.raises(p2t.P2TError, match="host tensors"):
        stager.submit(x.to(dev), ok.to(dev), x, ok)
    with pytest.raises(p2t.P2TError, match="without a submitted"):
        stager.take()


# --------------------------------------------------------------------------------------------------
# CUDA-graph replay of the step
# --------------------------------------------------------------------------------------------------
def test_graphed_step_replays_bit_identically_to_the_eager_step(p2t, synth, dev):
    sb = synth.make_config_batch("cfg1_esm2_t6_llama1b", weight_gain=2.5)
    ad = make_adapter(p2t, dev, sb.w1, sb.b1, sb.w2, sb.b2)  # eval mode: no dropout, results comparable bit for bit
    x, pm, text, tm = sb.x.to(dev), sb.prot_mask.to(dev), sb.text.to(dev), sb.text_mask.to(dev)
    loss = p2t.contrastive_step(x, pm, ad, text, tm)
    loss.backward()
    g_ref = {k: v.clone() for k, v in adapter_grads(ad).items()}
    l_ref = loss.detach().clone()
    ad.zero_grad(set_to_none=True)
    step = p2t.GraphedContrastiveStep(ad, x, pm, text, tm)
    assert 8 <= step.launches_per_replay <= 20  # round 1: 25 launches; the loss block is one cooperative kernel now
    for _ in range(3):
        out = step.replay()
    assert torch.equal(out, l_ref) and all(torch.equal(g_ref[k], adapter_grads(ad)[k]) for k in PARAMS)
    # new data in the SAME input buffers (shorter sequences: the ragged extents are read on the device)
    sb2 = synth.make_config_batch("cfg1_esm2_t6_llama1b", weight_gain=2.5, seed=77)
    L2, T2 = min(sb2.x.shape[1], x.shape[1]), min(sb2.text.shape[1], text.shape[1])
    x.zero_(); pm.zero_(); text.zero_(); tm.zero_()
    x[:, :L2] = sb2.x[:, :L2].to(dev); pm[:, :L2] = sb2.prot_mask[:, :L2].to(dev)
    text[:, :T2] = sb2.text[:, :T2].to(dev); tm[:, :T2] = sb2.text_mask[:, :T2].to(dev)
    x.mul_(pm[..., None].to(x.dtype))
    out2 = step.replay().clone()
    g2 = {k: v.clone() for k, v in adapter_grads(ad).items()}
    ad.zero_grad(set_to_none=True)
    l_eager = p2t.contrastive_step(x, pm, ad, text, tm)
    l_eager.backward()
    assert torch.equal(out2, l_eager) and all(torch.equal(g2[k], adapter_grads(ad)[k]) for k in PARAMS)


def test_graphed_step_draws_a_new_dropout_mask_every_replay(p2t, synth, dev):
    sb = synth.make_config_batch("tiny", weight_gain=6.0)
    ad = make_adapter(p2t, dev, sb.w1, sb.b1, sb.w2, sb.b2, train=True, p=0.3)
    step = p2t.GraphedContrastiveStep(ad, sb.x.to(dev), sb.prot_mask.to(dev), sb.text.to(dev), sb.text_mask.to(dev), seed=5)
    a = step.replay().item()
    b = step.replay().item()
    assert a != b and math.isfinite(a) and math.isfinite(b)
    again = p2t.GraphedContrastiveStep(ad, sb.x.to(dev), sb.prot_mask.to(dev), sb.text.to(dev), sb.text_mask.to(dev), seed=5)
    assert again.replay().item() == a  # same seed, same replay index: same mask


def test_step_with_many_negatives_uses_the_tensor_core_loss_path(p2t, synth, dev):
    """B = 160 pairs against 2048 text embeddings of width 2*256: R*C*E > 2^26, so the similarity, dLogits (bf16) and
    dp = dS t run on the tcgen05 GEMM — the first CUDA call of autograd's backward thread is then a TMA descriptor."""
    sb = synth.make_batch(d_in=64, d_mid=128, d_out=256, batch=160, lmin=3, lmax=24, tmin=2, tmax=8, weight_gain=8.0)
    f = torch.float32
    w1, b1, w2, b2 = (t.to(f).requires_grad_() for t in (sb.w1, sb.b1, sb.w2, sb.b2))
    tr = R.adapter_rows(sb.x.to(f), w1, b1, w2, b2)
    p, _ = R.l2_normalize(R.readout(tr.y, sb.prot_mask, "mix"))
    g = torch.Generator().manual_seed(11)
    t_all = torch.nn.functional.normalize(torch.randn(2048, 512, generator=g), dim=-1)
    labels = torch.randperm(2048, generator=g)[:160]
    # plant the partners (retrieval margins far above the bf16 noise of the tensor-core similarity)
    t_all[labels] = torch.nn.functional.normalize(p.detach() + 0.5 * torch.randn(160, 512, generator=g) / math.sqrt(512), dim=-1)
    ref = R.infonce_rows(p, t_all, labels, 0.05)
    gs = torch.autograd.grad(ref, (w1, b1, w2, b2))
    ad = make_adapter(p2t, dev, sb.w1, sb.b1, sb.w2, sb.b2)
    aux = p2t.StepAux()
    loss = p2t.contrastive_step(sb.x.to(dev), sb.prot_mask.to(dev), ad, text_embeds=t_all.to(dev),
                                labels=labels.to(torch.int32).to(dev), aux=aux)
    loss.backward()
    assert abs(loss.item() - ref.item()) <= 5 * LOSS_RTOL * abs(ref.item()) + 1e-4  # bf16 p and t in the similarity
    for k, gref in zip(PARAMS, gs):
        c, m = cosine(adapter_grads(ad)[k], gref), maxrel(adapter_grads(ad)[k], gref)
        assert c >= GRAD_COS and m <= 2e-2, f"{k}: cosine {c:.6f} maxrel {m:.3e}"  # bf16 dS and bf16 p/t operands
    am = torch.argmax(p.detach() @ t_all.t(), dim=1)
    assert torch.equal(am, labels) and torch.equal(aux.argmax_row.cpu().long(), am)
