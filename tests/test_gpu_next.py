"""GPU parity tests of the SURVEY.md §8 "next" rows and of the peer-memory exchange (run on the B200 box:
`pytest -m gpu`).  Everything under test goes through the C-ABI CUDA library; the checkers are

  * tests/golden/next_scatter.npz — the reference's own prepare_decoder_inputs fed by its ModalityAdapter,
  * tests/golden/next_adamw.npz   — torch.nn.utils.clip_grad_norm_ + torch.optim.AdamW, the calls of
                                    scripts/train_contrast.py:455-465,
  * oracle/restatement.py         — CPU restatements (mean all-reduce, AdamW at adapter size, scatter).

The exchange kernels are exercised on ONE device with several simulated ranks (`PeerBuffer.virtual`): every rank's
launch of a phase is issued before any rank's launch of the next phase, which is the order the flag protocol needs on a
single stream; tests/dist_check.py runs the same kernels across real processes over NVLink (tests/test_gpu_dist.py).
Bar: bit-exact for the exchange (byte moves, fixed-order fp32 sums), reference tolerances for floating point."""
import importlib
import math
import os
import sys

import numpy as np
import pytest
import torch

from oracle import restatement as R

pytestmark = pytest.mark.gpu
PARAMS = ("fc1.weight", "fc1.bias", "fc2.weight", "fc2.bias")


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.fail("-m gpu tests need a CUDA device: the product path has no CPU fallback")
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def mods(p2t):
    p2t._lib.load()
    return {n: importlib.import_module("p2t_b200." + n) for n in ("peer", "optim", "handoff", "dist", "synth", "graph")}


def _load(golden_dir, name):
    z = np.load(os.path.join(golden_dir, name))
    return {k: torch.from_numpy(np.asarray(z[k])) for k in z.files}


def bf(t):
    return t.to(torch.bfloat16)


def maxrel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-300)).item()


def cosine(a, b):
    a, b = a.detach().double().cpu().flatten(), b.detach().double().cpu().flatten()
    return torch.nn.functional.cosine_similarity(a, b, dim=0).item()


def make_adapter(p2t, dev, w1, b1, w2, b2, train=False, p=0.3):
    cfg = p2t.ModalityAdapterConfig(input_dim=w1.shape[1], intermediate_dim=w1.shape[0], output_dim=w2.shape[0], dropout_rate=p)
    ad = p2t.ModalityAdapter(cfg).to(dev).to(torch.bfloat16)
    with torch.no_grad():
        ad.fc1.weight.copy_(w1); ad.fc1.bias.copy_(b1); ad.fc2.weight.copy_(w2); ad.fc2.bias.copy_(b2)
    return ad.train() if train else ad.eval()


# --------------------------------------------------------------------------------------------------
# exchange over peer memory
# --------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("world,rows,cols,dtype", [(1, 32, 8192, torch.float32), (2, 32, 8192, torch.float32),
                                                   (3, 5, 24, torch.float32), (8, 16, 4096, torch.bfloat16),
                                                   (16, 2, 64, torch.float32)])
def test_peer_allgather_moves_every_block_bit_exactly(mods, dev, world, rows, cols, dtype):
    peer = mods["peer"]
    bufs = peer.PeerBuffer.virtual(peer.PeerAllGather.buffer_bytes(rows, cols, dtype, world), world)
    chans = [peer.PeerAllGather(rows, cols, dtype, _buffer=b) for b in bufs]
    g = torch.Generator().manual_seed(world * 1000 + rows)
    for rnd in range(5):  # odd and even epochs: both halves of the double buffer, and the flag comparison across rounds
        blocks = [torch.randn(rows, cols, generator=g).to(dtype).to(dev) for _ in range(world)]
        for c, blk in zip(chans, blocks):
            c.push(blk)
        outs = [c.arrive() for c in chans]
        want = torch.cat(blocks)
        for r, o in enumerate(outs):
            assert torch.equal(o, want), (rnd, r)
    for b in bufs:
        b.check()


def test_peer_allgather_rejects_wrong_blocks(p2t, mods, dev):
    peer = mods["peer"]
    ch = peer.PeerAllGather(4, 8, torch.float32)
    with pytest.raises(ValueError):
        ch.push(torch.zeros(4, 9, device=dev))
    with pytest.raises(ValueError):
        ch.push(torch.zeros(4, 8, device=dev, dtype=torch.bfloat16))
    with pytest.raises(ValueError):
        peer.PeerAllGather(1, 3, torch.float32)  # 12 bytes: not a multiple of 16
    with pytest.raises(p2t.P2TError, match="world"):
        p2t._lib.call("p2t_peer_allgather", ch.buffer.table, 17, 0, None, 16, None, 1, None)
    ch.close()


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_peer_allreduce_is_the_fixed_order_fp32_mean(mods, dev, world):
    peer = mods["peer"]
    shapes = [(64, 40), (64,), (24, 64), (20,)]  # (20,) = 40 bytes: padded to 48 inside the channel
    like = [torch.empty(s, dtype=torch.bfloat16) for s in shapes]
    bufs = peer.PeerBuffer.virtual(peer.PeerGradAllReduce.buffer_bytes(like), world)
    reds = [peer.PeerGradAllReduce(like, _buffer=b) for b in bufs]
    g = torch.Generator().manual_seed(7 + world)
    for rnd in range(3):
        per_rank = [[(torch.randn(s, generator=g) * (1 + r)).to(torch.bfloat16) for s in shapes] for r in range(world)]
        on_dev = [[t.to(dev) for t in grads] for grads in per_rank]
        for red, grads in zip(reds, on_dev):
            red.stage(grads)
        for red in reds:
            red.reduce()
        for red, grads in zip(reds, on_dev):
            red.finish(grads)
        for i in range(len(shapes)):
            want = R.mean_allreduce_bf16([per_rank[r][i] for r in range(world)])
            for r in range(world):
                assert torch.equal(on_dev[r][i].cpu(), want), (rnd, r, i)
    for b in bufs:
        b.check()


def test_peer_allreduce_adapter_sized(mods, dev):
    """config 2's adapter: 13.6 M gradient elements, 4 simulated ranks — bit-exact against the CPU restatement."""
    peer = mods["peer"]
    shapes = [(2048, 2560), (2048,), (4096, 2048), (4096,)]
    like = [torch.empty(s, dtype=torch.bfloat16) for s in shapes]
    world = 4
    bufs = peer.PeerBuffer.virtual(peer.PeerGradAllReduce.buffer_bytes(like), world)
    reds = [peer.PeerGradAllReduce(like, _buffer=b) for b in bufs]
    g = torch.Generator(device=dev).manual_seed(3)
    on_dev = [[torch.randn(s, generator=g, device=dev).to(torch.bfloat16) for s in shapes] for _ in range(world)]
    per_rank = [[t.cpu() for t in grads] for grads in on_dev]
    for red, grads in zip(reds, on_dev):
        red.stage(grads)
    for red in reds:
        red.reduce()
    for red, grads in zip(reds, on_dev):
        red.finish(grads)
    for i in range(4):
        want = R.mean_allreduce_bf16([per_rank[r][i] for r in range(world)])
        for r in range(world):
            assert torch.equal(on_dev[r][i].cpu(), want)


def test_graphed_step_through_the_exchange_equals_the_plain_step(p2t, mods, dev):
    """World of one: the sharded step's exchange (push, arrive, gradient mean) inside the captured graph must leave the
    step's numbers untouched — the same kernels then run across ranks in tests/dist_check.py."""
    synth, pdist, peer = mods["synth"], mods["dist"], mods["peer"]
    sb = synth.make_config_batch("tiny", weight_gain=8.0)
    ad = make_adapter(p2t, dev, sb.w1, sb.b1, sb.w2, sb.b2)
    x, pm, th, tm = (t.to(dev) for t in (sb.x, sb.prot_mask, sb.text, sb.text_mask))
    plain = p2t.GraphedContrastiveStep(ad, x, pm, th, tm)
    l0 = plain.replay().clone()
    g0 = [g.clone() for g in plain.grads]
    ex = pdist.ShardedExchange(x.shape[0], 2 * sb.w2.shape[0], symmetric=True)
    red = peer.PeerGradAllReduce.for_adapter(ad)
    for sym in (False, True):
        ref = p2t.GraphedContrastiveStep(ad, x, pm, th, tm, symmetric=sym)
        sharded = p2t.GraphedContrastiveStep(ad, x, pm, th, tm, symmetric=sym, exchange=ex, grad_reducer=red)
        for _ in range(3):
            la, lb = ref.replay(), sharded.replay()
            assert torch.equal(la, lb)
            for a, b in zip(ref.grads, sharded.grads):
                assert torch.equal(a, b)
        if not sym:
            assert torch.equal(l0, la) and all(torch.equal(a, b) for a, b in zip(g0, ref.grads))
    ex.check()
    assert sharded.launches_per_replay > ref.launches_per_replay
    ex.close(); red.close()


def test_gradient_mean_inside_the_dw1_gemm_launch_equals_the_plain_step(p2t, mods, dev):
    """OverlappedGradReduce: the mean of dW2 / db2 is formed by the idle epilogue warps inside the dW1 GEMM's launch (fused
    tcgen05 GEMM + peer-memory reduce), dW1 / db1 after it.  World of one: the numbers must equal the plain captured
    step's bit for bit (the GEMM keeps every SM and the same split-K cut); across real ranks tests/dist_check.py runs
    the same form against the global-batch oracle."""
    synth, peer = mods["synth"], mods["peer"]
    for workload, gain in (("tiny", 8.0), ("cfg1_esm2_t6_llama1b", 2.5)):
        sb = synth.make_config_batch(workload, weight_gain=gain)
        ad = make_adapter(p2t, dev, sb.w1, sb.b1, sb.w2, sb.b2)
        x, pm, th, tm = (t.to(dev) for t in (sb.x, sb.prot_mask, sb.text, sb.text_mask))
        plain = p2t.GraphedContrastiveStep(ad, x, pm, th, tm)
        red = peer.OverlappedGradReduce(ad)
        fused = p2t.GraphedContrastiveStep(ad, x, pm, th, tm, grad_reducer=red)
        for _ in range(3):
            la, lb = plain.replay(), fused.replay()
            assert torch.equal(la, lb)
            for i, (a, b) in enumerate(zip(plain.grads, fused.grads)):
                assert torch.equal(a, b), i
        red.buffer.check()
        # micro-batch accumulation: only the boundary replay runs the fused GEMM + reduce
        acc = p2t.GraphedContrastiveStep(ad, x, pm, th, tm, grad_reducer=red, accumulation_steps=2)
        ref = p2t.GraphedContrastiveStep(ad, x, pm, th, tm, accumulation_steps=2)
        for _ in range(4):
            acc.replay(); ref.replay()
        for a, b in zip(ref.grads, acc.grads):
            assert cosine(a, b) >= 0.99999 and maxrel(a, b) <= 1e-2  # bf16 vs fp32 running sums
        red.buffer.check()
        red.close()


def test_host_stager_pull_mode_equals_copy_mode(p2t, mods, dev):
    """HostStager(mode="pull"): one kernel per side reads the valid rows out of the pinned host batch; same packed rows,
    bit for bit, as the per-sequence copies — right and left padding, an empty sequence, a piece boundary inside a row."""
    gen = torch.Generator().manual_seed(21)
    B, L, D, T, H = 7, 300, 200, 40, 136
    x = bf(torch.randn(B, L, D, generator=gen)).pin_memory()
    text = bf(torch.randn(B, T, H, generator=gen)).pin_memory()
    lens = torch.tensor([300, 1, 0, 157, 82, 299, 41])
    tl = torch.tensor([40, 3, 17, 1, 40, 22, 9])
    for left in (False, True):
        ar, at = torch.arange(L)[None, :], torch.arange(T)[None, :]
        pm = ((ar >= L - lens[:, None]) if left else (ar < lens[:, None])).long()
        tm = ((at >= T - tl[:, None]) if left else (at < tl[:, None])).long()
        got = {}
        for mode in ("copy", "pull"):
            st = p2t.HostStager(dev, mode=mode, pull_ctas=5)
            st.submit(x, pm, text, tm)
            b = st.take()
            torch.cuda.synchronize()
            got[mode] = (b.residue_rows.clone(), b.residue_lengths.clone(), b.text_rows.clone(), b.text_lengths.clone(), b.h2d_bytes)
        for a, c in zip(got["copy"][:4], got["pull"][:4]):
            assert torch.equal(a, c)
        assert got["copy"][4] == got["pull"][4]
        want = torch.cat([x[j][pm[j] != 0] for j in range(B)])
        assert torch.equal(got["pull"][0].cpu(), want)
    with pytest.raises(p2t.P2TError, match="pinned"):
        p2t.HostStager(dev, mode="pull").submit(x.clone(), pm, text, tm)


# --------------------------------------------------------------------------------------------------
# clip_grad_norm_ + AdamW
# --------------------------------------------------------------------------------------------------
def test_fused_adamw_matches_torch_golden(p2t, mods, dev, golden_dir):
    g = _load(golden_dir, "next_adamw.npz")
    n, lr, max_norm = int(g["n_steps"]), float(g["lr"]), float(g["max_norm"])
    params = [torch.nn.Parameter(bf(g[f"p0.{i}"]).to(dev)) for i in range(4)]  # the fixture's p0 lies on the bf16 grid
    opt = mods["optim"].FusedAdamW(params, lr=lr, eps=1e-6, betas=(0.9, 0.999), max_grad_norm=max_norm)
    for step in range(n):
        for i, p in enumerate(params):
            p.grad = bf(g[f"g{step}.{i}"]).to(dev)
            assert torch.equal(p.grad.float().cpu(), g[f"g{step}.{i}"])  # gradients on the bf16 grid too
        opt.step()
        torch.testing.assert_close(opt.grad_norm.cpu(), g[f"norm{step}"], rtol=1e-5, atol=0)
        for i, p in enumerate(params):
            master = opt.state[p]["master"]
            torch.testing.assert_close(master.cpu(), g[f"p{step + 1}.{i}"], rtol=1e-5, atol=1e-7)
            assert torch.equal(p.detach(), master.to(torch.bfloat16))  # the parameter is the rounding of its master
            assert torch.equal(p.grad.float().cpu(), g[f"g{step}.{i}"])  # gradients are left as they were


def test_fused_adamw_adapter_sized_against_the_oracle(p2t, mods, dev):
    shapes = [(2048, 2560), (2048,), (4096, 2048), (4096,)]
    gen = torch.Generator().manual_seed(5)
    p0 = [bf(torch.randn(s, generator=gen) * 0.02) for s in shapes]
    for master in (True, False):
        params = [torch.nn.Parameter(p.clone().to(dev)) for p in p0]
        opt = mods["optim"].FusedAdamW(params, lr=1e-3, eps=1e-6, weight_decay=0.01, max_grad_norm=1.0, master_weights=master,
                                       zero_grad_in_step=True)
        ref_p = [p.float() for p in p0]
        m = [torch.zeros_like(p) for p in ref_p]
        v = [torch.zeros_like(p) for p in ref_p]
        for step in range(1, 4):
            grads = [bf(torch.randn(s, generator=gen) * (0.01 if step != 2 else 1e-5)) for s in shapes]
            for p, gr in zip(params, grads):
                p.grad = gr.clone().to(dev)
            opt.step()
            norm, cg = R.clip_grad_norm([gr.float() for gr in grads], 1.0)
            assert abs(opt.grad_norm.item() - norm.item()) <= 1e-5 * norm.item()
            ref_p, m, v = R.adamw_step(ref_p, cg, m, v, step, 1e-3, weight_decay=0.01)
            for p, rp in zip(params, ref_p):
                if master:
                    torch.testing.assert_close(opt.state[p]["master"].cpu(), rp, rtol=2e-5, atol=1e-8)
                    assert torch.equal(p.detach().cpu(), rp.to(torch.bfloat16)) or maxrel(p, rp) <= 2 ** -8
                else:  # bf16 parameter is the state: one rounding per step on top of the fp32 trajectory
                    assert maxrel(p, rp) <= step * 2 ** -8
                assert not p.grad.any()  # zero_grad_in_step
            if not master:
                ref_p = [p.detach().float().cpu() for p in params]  # follow the bf16 trajectory like the kernel does


def test_fused_adamw_rounds_fp32_gradient_sources_itself(p2t, mods, dev):
    """With fp32 sources (the mean left by the gradient all-reduce) the norm pass does the rounding to bf16: same
    .grad, norm, moments and weights, bit for bit, as converting first and stepping on the bf16 gradients."""
    shapes = [(96, 160), (96,), (130, 96), (133,)]  # the last one has a ragged tail (133 % 8 != 0)
    gen = torch.Generator().manual_seed(11)
    p0 = [bf(torch.randn(s, generator=gen) * 0.02) for s in shapes]
    pa = [torch.nn.Parameter(p.clone().to(dev)) for p in p0]
    pb = [torch.nn.Parameter(p.clone().to(dev)) for p in p0]
    oa = mods["optim"].FusedAdamW(pa, lr=1e-3, eps=1e-6, weight_decay=0.01, max_grad_norm=0.05)
    ob = mods["optim"].FusedAdamW(pb, lr=1e-3, eps=1e-6, weight_decay=0.01, max_grad_norm=0.05)
    for step in range(3):
        f32 = [(torch.randn(s, generator=gen) * 0.01).to(dev) for s in shapes]
        for p, g in zip(pa, f32):
            p.grad = g.to(torch.bfloat16)
        oa.step()
        for p in pb:
            p.grad = torch.full_like(p, float("nan"))  # must be overwritten by the norm pass
        ob.fp32_grad_sources = dict(zip(pb, f32))
        ob.step()
        ob.fp32_grad_sources = {}
        assert torch.equal(oa.grad_norm, ob.grad_norm)
        for a, b in zip(pa, pb):
            assert torch.equal(a.grad, b.grad) and torch.equal(a, b)
            assert torch.equal(oa.state[a]["exp_avg_sq"], ob.state[b]["exp_avg_sq"])
            assert torch.equal(oa.state[a]["master"], ob.state[b]["master"])
    ob.fp32_grad_sources = {pb[0]: torch.zeros(3, device=dev)}
    with pytest.raises(p2t.P2TError, match="fp32 gradient source"):
        ob.step()


def test_fused_adamw_is_capturable_and_follows_the_scheduler(p2t, mods, dev):
    w = torch.nn.Parameter(bf(torch.ones(64, 64)).to(dev))
    w.grad = bf(torch.full((64, 64), 0.5)).to(dev)
    opt = mods["optim"].FusedAdamW([w], lr=1e-2, weight_decay=0.0)
    opt.step()  # eager: state allocated, lr written
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        opt.step()
    before = opt.state[w]["master"].clone()
    opt.set_lr(0.0)
    graph.replay()
    assert torch.equal(opt.state[w]["master"], before)  # lr is read from the device at replay time
    opt.set_lr(1e-2)
    graph.replay()
    torch.cuda.synchronize()
    assert opt.step_count() == 3 and (opt.state[w]["master"] < before).all()  # 1 eager + 2 replays
    # checkpoint round trip: fp32 moments / master weights and the step counter survive (torch would cast them to bf16)
    sd = opt.state_dict()
    w2 = torch.nn.Parameter(w.detach().clone())
    w2.grad = w.grad.clone()
    opt2 = mods["optim"].FusedAdamW([w2], lr=1e-2, weight_decay=0.0)
    opt2.load_state_dict(sd)
    assert opt2.state[w2]["master"].dtype == torch.float32 and torch.equal(opt2.state[w2]["master"], opt.state[w]["master"])
    opt.step(); opt2.step()
    assert opt2.step_count() == 4 and torch.equal(w2, w) and torch.equal(opt2.state[w2]["exp_avg_sq"], opt.state[w]["exp_avg_sq"])
    with pytest.raises(p2t.P2TError, match="bfloat16"):
        f = torch.nn.Parameter(torch.ones(8, 8, device=dev))
        f.grad = torch.ones_like(f)
        mods["optim"].FusedAdamW([f]).step()


# --------------------------------------------------------------------------------------------------
# Stage-2 hand-off: adapter rows straight into the placeholder slots
# --------------------------------------------------------------------------------------------------
def test_adapter_into_embeds_matches_reference_golden(p2t, mods, dev, golden_dir):
    g = _load(golden_dir, "next_scatter.npz")
    ad = make_adapter(p2t, dev, *(g["sd." + k] for k in PARAMS))
    ids = g["input_ids"]
    ph_mask = (ids == int(g["placeholder_id"])).to(dev)
    base = bf(g["table"])[ids].to(dev)  # what llm.get_input_embeddings()(input_ids) returns (table on the bf16 grid)
    embeds = base.clone().requires_grad_()
    out = mods["handoff"].adapter_into_embeds(ad, bf(g["x"]).to(dev), g["enc_mask"].to(dev), embeds * 1, ph_mask, check=True)
    assert out.dtype == torch.bfloat16 and out.shape == g["embeds"].shape
    assert torch.equal(out[~ph_mask].cpu(), bf(g["embeds"])[~ph_mask.cpu()])  # token rows: byte moves
    assert maxrel(out[ph_mask], g["embeds"][ph_mask.cpu()]) <= 6e-3  # adapter rows: bf16 output of the adapter
    (out.float() * g["gy"].to(dev)).sum().backward()
    grads = {"fc1.weight": ad.fc1.weight.grad, "fc1.bias": ad.fc1.bias.grad, "fc2.weight": ad.fc2.weight.grad,
             "fc2.bias": ad.fc2.bias.grad}
    for k in PARAMS:
        c, m = cosine(grads[k], g["grad." + k]), maxrel(grads[k], g["grad." + k])
        assert c >= 0.999 and m <= 1e-2, (k, c, m)
    # gradient w.r.t. the token embeddings: upstream gradient outside the placeholders, zero inside
    want = g["gy"].clone()
    want[ph_mask.cpu()] = 0
    assert torch.equal(embeds.grad.float().cpu(), bf(want).float())
    with pytest.raises(ValueError, match="must match"):
        bad = ph_mask.clone()
        bad[0, 0] = ~bad[0, 0]
        mods["handoff"].adapter_into_embeds(ad, bf(g["x"]).to(dev), g["enc_mask"].to(dev), base.clone(), bad, check=True)


@pytest.mark.parametrize("left_pad", [False, True])
def test_adapter_into_embeds_equals_adapter_then_scatter(p2t, mods, dev, left_pad):
    """Against the composition the reference performs — adapter on the padded batch, then the masked assignment
    (oracle: placeholder_scatter) — at a realistic width; the rows that move are bit-identical to the module's."""
    synth = mods["synth"]
    sb = synth.make_batch(320, 512, 264, 5, 3, 90, 2, 8, seed=9, weight_gain=6.0, left_pad=left_pad)
    ad = make_adapter(p2t, dev, sb.w1, sb.b1, sb.w2, sb.b2)
    B, L, _ = sb.x.shape
    S = L + 7
    gen = torch.Generator().manual_seed(2)
    base = bf(torch.randn(B, S, 264, generator=gen))
    ph = torch.zeros(B, S, dtype=torch.bool)
    for b, n in enumerate(sb.prot_lens.tolist()):
        ph[b, 2 + b:2 + b + n] = True
    with torch.no_grad():
        y = ad(sb.x.to(dev))
        out = mods["handoff"].adapter_into_embeds(ad, sb.x.to(dev), sb.prot_mask.to(dev), base.clone().to(dev), ph.to(dev))
    want = R.placeholder_scatter(base, ph, y.cpu(), sb.prot_mask)
    assert torch.equal(out.cpu(), want)


# --------------------------------------------------------------------------------------------------
# medium similarity blocks (the sharded step's B x B_global block with fp32 embeddings)
# --------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("R_,B_,C_,D_", [(32, 32, 256, 4096), (30, 32, 200, 96), (70, 70, 130, 64), (16, 16, 32, 2048)])
def test_fp32_similarity_and_loss_backward_head_at_sharded_sizes(p2t, dev, R_, B_, C_, D_):
    """p2t_similarity (fp32 operands: register-tiled kernel from 512 logits up) and p2t_loss_bwd_coef (t slab staged in
    shared memory) against fp64: S = p t^T / tau; dp = dS t / tau; F.normalize backward; 'mix' coefficients."""
    core = sys.modules["p2t_b200._core"]
    E, tau = 2 * D_, 0.05
    g = torch.Generator().manual_seed(R_ * 7 + C_)
    t = torch.nn.functional.normalize(torch.randn(C_, E, generator=g), dim=-1)
    e = torch.randn(B_, E, generator=g) + 3.0  # un-normalised pooled embeddings (mean | std), std part positive
    e[:, D_:] = e[:, D_:].abs() + 0.5
    pnorm = e.norm(dim=-1)
    p = e / pnorm[:, None]
    S = torch.empty(R_, C_, dtype=torch.float32, device=dev)
    p_dev, t_dev = p[:R_].contiguous().to(dev), t.to(dev)  # keep both alive: the call only sees raw pointers
    p2t._lib.call("p2t_similarity", None, None, p_dev.data_ptr(), t_dev.data_ptr(), R_, C_, E, tau, S.data_ptr(), 2, None)
    want = p[:R_].double() @ t.double().T / tau
    assert maxrel(S, want) <= 2e-6
    # backward head
    dS = (torch.randn(R_, C_, generator=g) / R_).float()
    lens = torch.randint(3, 40, (B_,), generator=g)
    seq_off = torch.cat([torch.zeros(1, dtype=torch.long), lens.cumsum(0)]).to(torch.int32)
    ws = torch.empty(B_, E + (E + 63) // 64, dtype=torch.float32, device=dev)
    c1 = torch.empty(B_, D_, dtype=torch.float32, device=dev)
    c2 = torch.empty(B_, D_, dtype=torch.float32, device=dev)
    dl = torch.tensor([0.7], dtype=torch.float32, device=dev)
    td, pd, nd, ed, so, dSd = (x.to(dev).contiguous() for x in (t, p, pnorm, e, seq_off, dS))
    p2t._lib.call("p2t_loss_bwd_coef", dSd.data_ptr(), td.data_ptr(), pd.data_ptr(), nd.data_ptr(), ed.data_ptr(),
                  so.data_ptr(), dl.data_ptr(), R_, B_, C_, D_, tau, ws.data_ptr(), c1.data_ptr(), c2.data_ptr(), None)
    dp = torch.zeros(B_, E, dtype=torch.float64)
    dp[:R_] = 0.7 * (dS.double() @ t.double()) / tau
    got_dp = ws.flatten()[:B_ * E].view(B_, E)
    assert maxrel(got_dp, dp) <= 2e-6
    pdd = p.double()
    de = (dp - pdd * (pdd * dp).sum(-1, keepdim=True)) / pnorm.double()[:, None]
    n = lens.double()[:, None]
    k2 = de[:, D_:] / (n * e[:, D_:].double())
    k1 = de[:, :D_] / n - k2 * e[:, :D_].double()
    assert maxrel(c2, k2) <= 1e-4 and maxrel(c1, k1) <= 1e-4  # (dp - p (p.dp)) cancels ~2 digits in fp32


def test_graphed_training_step_with_optimizer_equals_eager_step_then_optimizer(p2t, mods, dev):
    """(r2-prep, not yet run on a GPU) one replay = forward + backward + clip + AdamW."""
    synth = mods["synth"]
    sb = synth.make_config_batch("tiny", weight_gain=8.0)
    x, pm, th, tm = (t.to(dev) for t in (sb.x, sb.prot_mask, sb.text, sb.text_mask))
    a1 = make_adapter(p2t, dev, sb.w1, sb.b1, sb.w2, sb.b2)
    a2 = make_adapter(p2t, dev, sb.w1, sb.b1, sb.w2, sb.b2)
    prm = lambda a: [a.fc1.weight, a.fc1.bias, a.fc2.weight, a.fc2.bias]
    o1 = mods["optim"].FusedAdamW(prm(a1), lr=1e-2, eps=1e-6, max_grad_norm=1.0)
    o2 = mods["optim"].FusedAdamW(prm(a2), lr=1e-2, eps=1e-6, max_grad_norm=1.0)
    fused = p2t.GraphedContrastiveStep(a1, x, pm, th, tm, optimizer=o1)
    for w_a, w_b in zip(prm(a1), (sb.w1, sb.b1, sb.w2, sb.b2)):
        assert torch.equal(w_a.detach().cpu(), w_b)  # building the graph moved nothing
    plain = p2t.GraphedContrastiveStep(a2, x, pm, th, tm)
    for _ in range(3):
        l1 = fused.replay()
        l2 = plain.replay()
        o2.step()
        assert torch.equal(l1, l2)
        for w_a, w_b in zip(prm(a1), prm(a2)):
            assert torch.equal(w_a, w_b)
    assert o1.step_count() == 3 == o2.step_count()


# --------------------------------------------------------------------------------------------------
# round 2: the loss block as one cooperative kernel (csrc/loss_fused.cu)
# --------------------------------------------------------------------------------------------------
FUSED_CASES = [
    # R, B, C, D, symmetric
    (32, 32, 32, 4096, False), (32, 32, 256, 4096, False), (32, 32, 256, 4096, True), (30, 32, 200, 96, True),
    (70, 70, 130, 64, True), (16, 16, 32, 2048, False), (6, 6, 6, 96, True), (3, 5, 7, 32, False), (64, 64, 256, 1792, False),
]


def _fused_inputs(R_, B_, C_, D_, seed):
    """Un-normalised pooled embeddings e = (mean | std) (std half positive), p = e / |e|, unit-norm t, labels.  The
    positives are planted with the strength (searched in fp64) that leaves the row loss closest to 1: real retrieval
    margins, but a softmax far from saturation — a saturated one makes dLogits = softmax - onehot a difference of
    nearly equal numbers, which tests the test's conditioning, not the kernel."""
    E = 2 * D_
    g = torch.Generator().manual_seed(seed)
    t = torch.nn.functional.normalize(torch.randn(C_, E, generator=g), dim=-1)
    labels = torch.randperm(C_, generator=g)[:R_]
    base = torch.randn(B_, E, generator=g) + 3.0
    base[:, D_:] = base[:, D_:].abs() + 0.5

    def build(alpha):
        e = base.clone()
        e[:R_] = e[:R_] + alpha * t[labels] * e[:R_].norm(dim=-1, keepdim=True)
        e[:, D_:] = e[:, D_:].abs() + 0.1
        return e

    best = None
    for alpha in (0.05, 0.1, 0.15, 0.2, 0.3, 0.4, 0.6, 0.8, 1.0, 1.5):
        e = build(alpha)
        p = (e / e.norm(dim=-1, keepdim=True)).float().double()
        l = float(R.infonce_rows(p[:R_], t.float().double(), labels, 0.05))
        if best is None or abs(l - 1.0) < abs(best[0] - 1.0):
            best = (l, e)
    e = best[1]
    pnorm = e.norm(dim=-1)
    p = e / pnorm[:, None]
    lens = torch.randint(3, 40, (B_,), generator=g)
    seq_off = torch.cat([torch.zeros(1, dtype=torch.long), lens.cumsum(0)]).to(torch.int32)
    return t, labels, e, pnorm, p, lens, seq_off


@pytest.mark.parametrize("R_,B_,C_,D_,sym", FUSED_CASES)
def test_fused_loss_kernel_matches_fp64(p2t, dev, R_, B_, C_, D_, sym):
    """similarity -> online-softmax CE (rows, and columns for the symmetric term) -> dLogits -> dp -> F.normalize
    backward -> 'mix' coefficients in ONE cooperative kernel, against the fp64 restatement of
    scripts/train_contrast.py:100-114 and its autograd.  Retrieval argmax bit-exact (planted positives)."""
    core = sys.modules["p2t_b200._core"]
    E, tau = 2 * D_, 0.05
    t, labels, e, pnorm, p, lens, seq_off = _fused_inputs(R_, B_, C_, D_, R_ * 7 + C_ + D_)
    wr, wc = (0.5, 0.5) if sym else (1.0, 0.0)
    assert core.loss_fused_eligible(R_, B_, C_, E)
    pd, td, nd, ed, so = (x.to(dev).contiguous() for x in (p.float(), t.float(), pnorm.float(), e.float(), seq_off))
    dl = torch.tensor(0.7, dtype=torch.float32, device=dev)
    res = core.loss_fused(pd, td, labels.to(dev), R_, tau, w_row=wr, w_col=wc, need_grad=True, dloss=dl, pnorm=nd,
                          stats=ed, seq_off=so)
    again = core.loss_fused(pd, td, labels.to(dev), R_, tau, w_row=wr, w_col=wc, need_grad=True, dloss=dl, pnorm=nd,
                            stats=ed, seq_off=so)
    pf, tf = p.float().double(), t.float().double()
    ref = wr * R.infonce_rows(pf[:R_], tf, labels, tau) + (wc * R.infonce_cols(pf[:R_], tf, labels, tau) if wc else 0.0)
    assert abs(res.loss.item() - float(ref)) <= 2e-5 * abs(float(ref)) + 1e-6
    _, dpo, _ = R.infonce_backward(pf[:R_], tf, labels, tau, wr, wc)
    dp = torch.zeros(B_, E, dtype=torch.float64)
    dp[:R_] = 0.7 * dpo
    de = (dp - pf * (pf * dp).sum(-1, keepdim=True)) / pnorm.double()[:, None]
    n = lens.double()[:, None]
    k2 = de[:, D_:] / (n * e[:, D_:].float().double())
    k1 = de[:, :D_] / n - k2 * e[:, :D_].float().double()
    assert maxrel(res.c2, k2) <= 5e-4 and maxrel(res.c1, k1) <= 5e-4  # (softmax - 1) and (dp - p (p.dp)) cancel digits in fp32
    assert cosine(res.c1, k1) >= 0.99999 and cosine(res.c2, k2) >= 0.99999
    am_r, am_c = R.retrieval_argmax(pf[:R_], tf)
    assert torch.equal(res.argmax_row.cpu().long(), am_r)
    assert torch.equal(res.argmax_col.cpu().long()[labels], am_c[labels])
    lse = torch.logsumexp(pf[:R_] @ tf.T / tau, dim=1)
    assert maxrel(res.row_lse, lse) <= 1e-5
    # deterministic: fixed reduction orders everywhere
    assert torch.equal(res.loss, again.loss) and torch.equal(res.c1, again.c1) and torch.equal(res.c2, again.c2)


def test_fused_loss_kernel_forward_only_and_bad_labels(p2t, dev):
    core = sys.modules["p2t_b200._core"]
    t, labels, e, pnorm, p, lens, seq_off = _fused_inputs(8, 8, 24, 64, 5)
    pd, td = p.float().to(dev), t.float().to(dev)
    res = core.loss_fused(pd, td, labels.to(dev), 8, 0.05, need_grad=False)
    ref = R.infonce_rows(p.float().double(), t.float().double(), labels, 0.05)
    assert res.c1 is None and abs(res.loss.item() - float(ref)) <= 2e-5 * abs(float(ref))
    bad = labels.clone()
    bad[3] = 24  # outside [0, C): the loss is poisoned instead of reading out of bounds
    assert math.isnan(core.loss_fused(pd, td, bad.to(dev), 8, 0.05, need_grad=False).loss.item())
    assert not core.loss_fused_eligible(512, 512, 4096, 8192)  # large blocks run on the tensor cores
    with pytest.raises(p2t.P2TError, match="not eligible"):
        big_p = torch.zeros(512, 64, device=dev)
        core.loss_fused(big_p, torch.zeros(4096, 64, device=dev), torch.zeros(512, dtype=torch.int32, device=dev), 512, 0.05,
                        need_grad=False)
    # the kernel re-armed its barrier words: another launch right after works
    assert math.isfinite(core.loss_fused(pd, td, labels.to(dev), 8, 0.05, need_grad=False).loss.item())


def test_step_bias_gradients_in_fp32_round_to_the_bf16_gradients(p2t, mods, dev):
    """step_backward hands the bias gradients out twice: bf16 (param.grad) and fp32 (what the gradient mean over ranks
    carries, so that the only rounding happens after the mean)."""
    synth = mods["synth"]
    step_mod = sys.modules["p2t_b200.step"]
    sb = synth.make_config_batch("tiny", weight_gain=6.0)
    ad = make_adapter(p2t, dev, sb.w1, sb.b1, sb.w2, sb.b2)
    with torch.no_grad():
        _, state = p2t.contrastive_step(sb.x.to(dev), sb.prot_mask.to(dev), ad, sb.text.to(dev), sb.text_mask.to(dev), _raw=True)
        dw1, db1, dw2, db2, db1_f32, db2_f32 = step_mod.step_backward(state, None)
    assert db1_f32.dtype == torch.float32 and torch.equal(db1, db1_f32.to(torch.bfloat16))
    assert torch.equal(db2, db2_f32.to(torch.bfloat16))
    f = torch.float32
    st = R.step_forward(sb.x.to(f), sb.prot_mask, sb.w1.to(f), sb.b1.to(f), sb.w2.to(f), sb.b2.to(f), sb.text.to(f), sb.text_mask)
    ref = R.step_backward(st, sb.x.to(f), sb.prot_mask, sb.w1.to(f), sb.w2.to(f))
    assert maxrel(db1_f32, ref["fc1.bias"]) <= 1e-2 and maxrel(db2_f32, ref["fc2.bias"]) <= 1e-2
    assert cosine(dw1, ref["fc1.weight"]) >= 0.999 and cosine(dw2, ref["fc2.weight"]) >= 0.999


@pytest.mark.parametrize("k", [2, 3])
def test_graphed_step_accumulates_micro_batches_like_the_reference(p2t, mods, dev, k):
    """scripts/train_contrast.py:432,448-465: loss / k, backward on every micro-batch, optimizer.step() on the k-th.
    The captured step adds into its static gradient buffers on the non-first micro-steps and binds param.grad on the
    boundary step only; the result must equal eager autograd accumulation of the same micro-batches."""
    synth = mods["synth"]
    sbs = [synth.make_config_batch("tiny", weight_gain=6.0, seed=50 + i) for i in range(k)]
    L = max(s.x.shape[1] for s in sbs)
    T = max(s.text.shape[1] for s in sbs)
    B, d_in, H = sbs[0].x.shape[0], sbs[0].x.shape[2], sbs[0].text.shape[2]
    ad = make_adapter(p2t, dev, sbs[0].w1, sbs[0].b1, sbs[0].w2, sbs[0].b2)
    x = torch.zeros(B, L, d_in, dtype=torch.bfloat16, device=dev)
    pm = torch.zeros(B, L, dtype=torch.long, device=dev)
    th = torch.zeros(B, T, H, dtype=torch.bfloat16, device=dev)
    tm = torch.zeros(B, T, dtype=torch.long, device=dev)

    def fill(s):
        x.zero_(); pm.zero_(); th.zero_(); tm.zero_()
        x[:, :s.x.shape[1]] = s.x.to(dev); pm[:, :s.x.shape[1]] = s.prot_mask.to(dev)
        th[:, :s.text.shape[1]] = s.text.to(dev); tm[:, :s.text.shape[1]] = s.text_mask.to(dev)

    fill(sbs[0])
    step = p2t.GraphedContrastiveStep(ad, x, pm, th, tm, accumulation_steps=k)
    for window in range(2):
        ad.zero_grad(set_to_none=True)
        losses = []
        for i, s in enumerate(sbs):
            fill(s)
            assert step.is_boundary == (i == k - 1)
            losses.append(step.replay().item())
            if i < k - 1:
                assert ad.fc1.weight.grad is None  # non-boundary micro-steps bind nothing (and reduce nothing)
        got = {n: v.clone() for n, v in zip(("fc1.weight", "fc1.bias", "fc2.weight", "fc2.bias"),
                                            (ad.fc1.weight.grad, ad.fc1.bias.grad, ad.fc2.weight.grad, ad.fc2.bias.grad))}
        ad.zero_grad(set_to_none=True)
        want_losses = []
        for s in sbs:
            fill(s)
            l = p2t.contrastive_step(x, pm, ad, th, tm)
            want_losses.append(l.item())
            (l / k).backward()
        assert losses == want_losses
        for n, prm in (("fc1.weight", ad.fc1.weight), ("fc1.bias", ad.fc1.bias), ("fc2.weight", ad.fc2.weight), ("fc2.bias", ad.fc2.bias)):
            # same kernels, same micro-batches; the only difference is where the bf16 roundings of the running sum fall
            assert cosine(got[n], prm.grad) >= 0.99999 and maxrel(got[n], prm.grad) <= 1.2e-2, n


# --------------------------------------------------------------------------------------------------
# round 2: the sharded step at config-3 shape on ONE device (8 simulated ranks), config-5-size loss blocks
# --------------------------------------------------------------------------------------------------
def _global_batch(shards):
    f = torch.float32
    B = shards[0].x.shape[0]
    W = len(shards)
    Lmax = max(s.x.shape[1] for s in shards)
    Tmax = max(s.text.shape[1] for s in shards)
    X = torch.zeros(W * B, Lmax, shards[0].x.shape[2])
    PM = torch.zeros(W * B, Lmax, dtype=torch.long)
    TX = torch.zeros(W * B, Tmax, shards[0].text.shape[2])
    TM = torch.zeros(W * B, Tmax, dtype=torch.long)
    for r, s in enumerate(shards):
        X[r * B:(r + 1) * B, :s.x.shape[1]] = s.x.to(f)
        PM[r * B:(r + 1) * B, :s.x.shape[1]] = s.prot_mask
        TX[r * B:(r + 1) * B, :s.text.shape[1]] = s.text.to(f)
        TM[r * B:(r + 1) * B, :s.text.shape[1]] = s.text_mask
    return X, PM, TX, TM


def _virtual_sharded_step(p2t, mods, dev, shards, ad, sym):
    """The sharded step of every simulated rank on one stream: all pushes, then every rank's step (whose loss kernel —
    or arrive kernel in the symmetric form — waits for the gathered rows).  The symmetric form needs the column
    statistics of ALL ranks before any rank's dLogits, so it runs in two passes: pass 1 publishes every rank's
    statistics, pass 2 repeats the step with the merged ones (eval mode: both passes compute the same numbers)."""
    step_mod = sys.modules["p2t_b200.step"]
    W, B = len(shards), shards[0].x.shape[0]
    E = 2 * shards[0].w2.shape[0]
    exs = mods["dist"].ShardedExchange.virtual(B, E, W, symmetric=sym)
    dev_in = [tuple(t.to(dev) for t in (s.x, s.prot_mask, s.text, s.text_mask)) for s in shards]

    def push_all():
        for ex, (x, pm, th, tm) in zip(exs, dev_in):
            ex.text.push(step_mod.text_embeddings(th, tm, dtype=torch.float32))

    def run_rank(r, hook):
        x, pm, th, tm = dev_in[r]
        aux = p2t.StepAux()
        with torch.no_grad():
            loss, state = p2t.contrastive_step(x, pm, ad, text_embeds=exs[r].text, symmetric=sym,
                                               labels=step_mod._rank_labels(r, B, dev), aux=aux, col_stats_hook=hook,
                                               all_cols_labelled=sym, late_text=True, _raw=True)
            grads = step_mod.step_backward(state, None, dw_f32=True)  # unrounded: what the gradient reducer carries
        return loss, grads, aux

    outs = []
    if sym:
        push_all()
        for r in range(W):  # pass 1: publish this rank's column statistics, results discarded
            run_rank(r, lambda m, s, ex=exs[r]: (ex.push_column_stats(m, s), (m, s))[1])
        push_all()
        for r in range(W):
            outs.append(run_rank(r, lambda m, s, ex=exs[r]: ex.arrive_column_stats()))
    else:
        push_all()
        for r in range(W):
            outs.append(run_rank(r, None))
    torch.cuda.synchronize()
    for ex in exs:
        ex.check()
    loss = sum(o[0].item() for o in outs) / W
    mean = lambda i: sum(o[1][i].float() for o in outs) / W
    # the mean over ranks of the fp32 gradients, as peer.PeerGradAllReduce.for_adapter forms it (one rounding, at the end)
    grads = {"fc1.weight": mean(0).to(torch.bfloat16), "fc1.bias": mean(4).to(torch.bfloat16),
             "fc2.weight": mean(2).to(torch.bfloat16), "fc2.bias": mean(5).to(torch.bfloat16)}
    am_row = torch.cat([o[2].argmax_row.cpu().long() for o in outs])
    return loss, grads, am_row


def _oracle_global_batch_from_shards(shards, variants, tau=0.05):
    """The oracle on the concatenated global batch, evaluated shard by shard so that the CPU never
    holds more than one shard's activations: the oracle's own stage functions (adapter_rows, readout, l2_normalize,
    infonce_*, and their closed-form backwards) composed exactly as R.step_forward / R.step_backward compose them —
    pairs only couple through the (global) similarity.  Sequences are processed in length-sorted groups of 8 to keep
    pad rows out of the CPU GEMMs.  Returns {symmetric: (loss, grads, P, T)}."""
    f = torch.float32
    sb = shards[0]
    w1, b1, w2, b2 = (t.to(f) for t in (sb.w1, sb.b1, sb.w2, sb.b2))

    def groups(s):
        order = torch.argsort(s.prot_lens)
        for i in range(0, len(order), 8):
            idx = order[i:i + 8]
            L = int(s.prot_lens[idx].max())
            yield idx, s.x[idx, :L].to(f), s.prot_mask[idx, :L]

    P, PN, T = [], [], []
    for s in shards:
        p_s = torch.zeros(s.x.shape[0], 2 * w2.shape[0])
        pn_s = torch.zeros(s.x.shape[0], 1)
        for idx, x, m in groups(s):
            tr = R.adapter_rows(x, w1, b1, w2, b2)
            p_g, pn_g = R.l2_normalize(R.readout(tr.y, m, "mix"))
            p_s[idx], pn_s[idx] = p_g, pn_g
        P.append(p_s); PN.append(pn_s)
        T.append(R.l2_normalize(R.readout(s.text.to(f), s.text_mask, "mix"))[0])
    P, PN, T = torch.cat(P), torch.cat(PN), torch.cat(T)
    labels = torch.arange(P.shape[0])
    out, dPs = {}, {}
    for sym in variants:
        wr, wc = (0.5, 0.5) if sym else (1.0, 0.0)
        loss = wr * R.infonce_rows(P, T, labels, tau) + (wc * R.infonce_cols(P, T, labels, tau) if wc else 0.0)
        _, dP, _ = R.infonce_backward(P, T, labels, tau, wr, wc)
        dPs[sym] = dP
        out[sym] = [float(loss), None, P, T]
    grads = {sym: None for sym in variants}
    B = shards[0].x.shape[0]
    for r, s in enumerate(shards):
        rows = slice(r * B, (r + 1) * B)
        for idx, x, m in groups(s):
            tr = R.adapter_rows(x, w1, b1, w2, b2)
            flat = R.AdapterTrace(x=tr.x.reshape(-1, x.shape[-1]), z1=tr.z1.reshape(-1, w1.shape[0]), h1=tr.h1.reshape(-1, w1.shape[0]),
                                  z2=tr.z2.reshape(-1, w2.shape[0]), a=tr.a.reshape(-1, w2.shape[0]), norm=tr.norm.reshape(-1, 1),
                                  y=tr.y.reshape(-1, w2.shape[0]))
            for sym in variants:
                de = R.l2_normalize_backward(P[rows][idx], PN[rows][idx], dPs[sym][rows][idx])
                dy = R.readout_backward(tr.y, m, "mix", de)
                g = R.adapter_rows_backward(flat, dy.reshape(-1, w2.shape[0]), w1, w2)
                if grads[sym] is None:
                    grads[sym] = {k: g[k].clone() for k in PARAMS}
                else:
                    for k in PARAMS:
                        grads[sym][k] += g[k]
    for sym in variants:
        out[sym][1] = grads[sym]
    return out


def test_oracle_sharded_composition_equals_the_global_step():
    """The shard-by-shard evaluation above IS R.step_forward / R.step_backward on the concatenated batch (small case)."""
    synth = importlib.import_module("p2t_b200.synth")
    shards = [synth.make_config_batch("tiny", weight_gain=6.0, rank=r) for r in range(3)]
    f = torch.float32
    sb = shards[0]
    X, PM, TX, TM = _global_batch(shards)
    for sym in (False, True):
        st = R.step_forward(X, PM, sb.w1.to(f), sb.b1.to(f), sb.w2.to(f), sb.b2.to(f), TX, TM, 0.05, 1, sym)
        ref = R.step_backward(st, X, PM, sb.w1.to(f), sb.w2.to(f), 0.05, 1, sym)
        loss, grads, _, _ = _oracle_global_batch_from_shards(shards, [sym])[sym]
        assert abs(loss - st.loss.item()) <= 1e-5 * abs(st.loss.item())
        for k in PARAMS:
            assert maxrel(grads[k], ref[k]) <= 2e-4 and cosine(grads[k], ref[k]) >= 0.999999, k


def test_sharded_step_at_config3_shape_on_one_device(p2t, mods, dev):
    """BASELINE config 3 — config-2 dimensions (2560 -> 2048 -> 4096), 8 ranks x 32 pairs, global batch 256 — through
    the sharded step's own kernels (peer-memory gather with the loss kernel's in-kernel arrival; column-statistics
    exchange for the symmetric form), all 8 ranks simulated on this device, against the oracle on the concatenated
    global batch.  Mean of the local losses == global loss; mean of the rank gradients == global
    gradients at north_star's tolerances, the symmetric form included; retrieval argmax exact wherever the oracle's
    own top-2 margin is above fp32 noise."""
    synth = mods["synth"]
    W = 8
    shards = [synth.make_config_batch("cfg2_esm2_3b_llama8b", weight_gain=1.0, rank=r) for r in range(W)]
    sb = shards[0]
    ad = make_adapter(p2t, dev, sb.w1, sb.b1, sb.w2, sb.b2)
    got = {sym: _virtual_sharded_step(p2t, mods, dev, shards, ad, sym) for sym in (False, True)}
    want = _oracle_global_batch_from_shards(shards, [False, True])
    for sym in (False, True):
        loss, grads, am_row = got[sym]
        ref_loss, ref, P, T = want[sym]
        assert abs(loss - ref_loss) <= 1e-3 * abs(ref_loss), (sym, loss, ref_loss)
        for k in PARAMS:
            c, m = cosine(grads[k], ref[k]), maxrel(grads[k], ref[k])
            assert c >= 0.999 and m <= 1e-2, f"config 3, symmetric={sym}, {k}: cosine {c:.6f} maxrel {m:.3e}"
        logits = P.double() @ T.double().T
        top2 = logits.topk(2, dim=1).values
        sure = (top2[:, 0] - top2[:, 1]) > 1e-5
        assert torch.equal(am_row[sure], logits.argmax(dim=1)[sure]) and int(sure.sum()) >= 128


BIG_LOSS_CASES = [
    # R, C, E, symmetric: the per-rank block of config 5 at W = 8, and its single-GPU 4096 x 4096 block
    (512, 4096, 8192, False), (512, 4096, 8192, True), (4096, 4096, 8192, False), (4096, 4096, 8192, True),
]


@pytest.mark.parametrize("R_,C_,E_,sym", BIG_LOSS_CASES)
def test_infonce_at_config5_size(p2t, dev, R_, C_, E_, sym):
    """scripts/train_contrast.py:100-114 at BASELINE config 5's size (global batch 4096, E = 2 * 4096): loss, dp, dt,
    row / column retrieval argmax (planted positives, bit-exact) against the fp64 restatement."""
    core = sys.modules["p2t_b200._core"]
    g = torch.Generator().manual_seed(R_ + C_ + int(sym))
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
    assert cosine(dp, dpo) >= 0.9999 and maxrel(dp, dpo) <= 8e-3   # bf16 dLogits operand on the tensor-core path
    assert cosine(dt, dto) >= 0.9999 and maxrel(dt, dto) <= 8e-3
    am_r, am_c = R.retrieval_argmax(pf, tf)
    assert torch.equal(res.argmax_row.cpu().long(), am_r)
    assert torch.equal(res.argmax_col.cpu().long()[labels], am_c[labels])


def test_adapter_under_distributed_data_parallel_with_unused_layer_norms(p2t, mods, dev):
    """scripts/train_contrast.py:611-614 wraps the model in DistributedDataParallel(find_unused_parameters=True) because
    ln1/ln2 never take part in the forward (models/modeling_esm2llama_instruct.py:56-57).  World of one on this GPU:
    the wrapped repo adapter must run forward/backward, give fc1/fc2 their gradients and leave ln1/ln2 without."""
    import torch.distributed as dist
    synth = mods["synth"]
    sb = synth.make_config_batch("tiny", weight_gain=6.0)
    ad = make_adapter(p2t, dev, sb.w1, sb.b1, sb.w2, sb.b2)
    ref_ad = make_adapter(p2t, dev, sb.w1, sb.b1, sb.w2, sb.b2)
    created = not dist.is_initialized()
    if created:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29577")
        dist.init_process_group("nccl", rank=0, world_size=1, device_id=dev)
    try:
        ddp = torch.nn.parallel.DistributedDataParallel(ad, device_ids=[dev.index], find_unused_parameters=True)
        x = sb.x.to(dev)
        for _ in range(2):  # two iterations: the reducer must have been re-armed after the first
            ddp.zero_grad(set_to_none=True)
            y = ddp(x)
            (y.float() * torch.linspace(-1, 1, y.shape[-1], device=dev)).sum().backward()
        ref_ad.zero_grad(set_to_none=True)
        y0 = ref_ad(x)
        (y0.float() * torch.linspace(-1, 1, y0.shape[-1], device=dev)).sum().backward()
        assert torch.equal(y, y0)
        for n in ("fc1", "fc2"):
            assert torch.equal(getattr(ad, n).weight.grad, getattr(ref_ad, n).weight.grad)
            assert torch.equal(getattr(ad, n).bias.grad, getattr(ref_ad, n).bias.grad)
        assert ad.ln1.weight.grad is None and ad.ln2.bias.grad is None
    finally:
        if created:
            dist.destroy_process_group()
