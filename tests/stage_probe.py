"""Stage-by-stage parity report of the CUDA path against the oracle (prints, never asserts).

    python tests/stage_probe.py [config-name] [weight_gain]

Used during bring-up on the GPU box; the assertions proper live in tests/ (-m gpu).
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry
from oracle import restatement as R


def rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).abs().max() / (b.abs().max() + 1e-30)).item()


def cos(a, b):
    return torch.nn.functional.cosine_similarity(a.double().cpu().flatten(), b.double().cpu().flatten(), dim=0).item()


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "tiny"
    gain = float(sys.argv[2]) if len(sys.argv) > 2 else 6.0
    pkg = entry.load_package()
    core = sys.modules["p2t_b200._core"]
    synth = __import__("importlib").import_module("p2t_b200.synth")
    sb = synth.make_config_batch(name, weight_gain=gain)
    dev = torch.device("cuda:0")
    f = torch.float32
    B, L, d_in = sb.x.shape
    d_mid, d_out = sb.w1.shape[0], sb.w2.shape[0]
    print(f"== {name}: B={B} L={L} d_in={d_in} d_mid={d_mid} d_out={d_out} rows={int(sb.prot_lens.sum())} cta_group={core.default_cta_group()}")
    x, pm = sb.x.to(dev), sb.prot_mask.to(dev)
    w1, b1, w2, b2 = (t.to(dev) for t in (sb.w1, sb.b1, sb.w2, sb.b2))

    # oracle on bf16-rounded inputs, fp32 math
    st = R.step_forward(sb.x.to(f), sb.prot_mask, sb.w1.to(f), sb.b1.to(f), sb.w2.to(f), sb.b2.to(f), sb.text.to(f), sb.text_mask)
    gr = R.step_backward(st, sb.x.to(f), sb.prot_mask, sb.w1.to(f), sb.w2.to(f))
    tr = st.extras["adapter"]
    valid = sb.prot_mask.bool()

    # 1. plan + gather
    plan = core.plan_rows(pm)
    xp = core.gather_rows(x.view(B * L, d_in), plan)
    torch.cuda.synchronize()
    n = int(plan.n_rows.item())
    print(f"plan: n_rows={n} expected={int(valid.sum())} seq_off ok={torch.equal(plan.seq_off.cpu()[1:].long(), sb.prot_lens.cumsum(0))}")
    print(f"gather: exact={torch.equal(xp[:n].cpu(), sb.x[valid])} pad_zero={bool((xp[n:min(plan.rows_cap, (n + 255) // 256 * 256)] == 0).all())}")

    # 2. adapter forward
    acts = core.adapter_forward(xp, plan.rows_cap, plan.rows_cap, plan.n_rows, w1, b1, w2, b2, 0.0, 0, True)
    torch.cuda.synchronize()
    print(f"fc1: h1 rel={rel(acts.h1[:n], tr.h1[valid]):.3e}  g1 rel={rel(acts.g1[:n], R.gelu_erf_grad(tr.z1[valid])):.3e}")
    print(f"fc2: a rel={rel(acts.a[:n], tr.a[valid]):.3e}  g2 rel={rel(acts.g2[:n], R.gelu_erf_grad(tr.z2[valid])):.3e}")
    print(f"rowsq: rel={rel(acts.rowsq[:, :n].sum(0), tr.a[valid].pow(2).sum(-1)):.3e}")

    # 3/4. pool + normalise
    inv_norm = core.row_inv_norm(acts)
    stats = core.pool_forward(acts.a, plan, d_out, row_src=None, inv_norm=inv_norm)
    p_bf, p_f32, pn = core.l2norm_forward(stats)
    torch.cuda.synchronize()
    print(f"pool: mean rel={rel(stats[:, :d_out], st.e_prot[:, :d_out]):.3e} std rel={rel(stats[:, d_out:], st.e_prot[:, d_out:]):.3e} p rel={rel(p_f32, st.p):.3e}")

    # 5. text
    t_bf = pkg.text_embeddings(sb.text.to(dev), sb.text_mask.to(dev))
    torch.cuda.synchronize()
    print(f"text: t rel={rel(t_bf, st.t):.3e}")

    # 6. loss from ORACLE embeddings (isolates the loss kernels)
    labels = torch.arange(B, dtype=torch.int32, device=dev)
    po, to = st.p.to(torch.bfloat16), st.t.to(torch.bfloat16)
    for sym in (False, True):
        wr, wc = (0.5, 0.5) if sym else (1.0, 0.0)
        res = core.infonce_forward(po.to(dev), to.to(dev), labels, 0.05, w_row=wr, w_col=wc, want_col_argmax=True)
        dp, dt = core.infonce_backward(res, po.to(dev), to.to(dev), 0.05, need_dt=True)
        torch.cuda.synchronize()
        lab = torch.arange(B)
        lo = wr * R.infonce_rows(po.float(), to.float(), lab, 0.05) + (wc * R.infonce_cols(po.float(), to.float(), lab, 0.05) if wc else 0)
        _, dpo, dto = R.infonce_backward(po.float(), to.float(), lab, 0.05, wr, wc)
        am_r, am_c = R.retrieval_argmax(po.float(), to.float())
        print(f"infonce sym={sym}: loss {res.loss.item():.6f} vs {float(lo):.6f}  dp rel={rel(dp, dpo):.3e} dt rel={rel(dt, dto):.3e} "
              f"argmax_row ok={torch.equal(res.argmax_row.cpu().long(), am_r)} argmax_col ok={torch.equal(res.argmax_col.cpu().long(), am_c)}")

    # 7. full fused step
    cfg = pkg.ModalityAdapterConfig(input_dim=d_in, intermediate_dim=d_mid, output_dim=d_out)
    ad = pkg.ModalityAdapter(cfg).to(dev).to(torch.bfloat16).eval()
    with torch.no_grad():
        ad.fc1.weight.copy_(w1); ad.fc1.bias.copy_(b1); ad.fc2.weight.copy_(w2); ad.fc2.bias.copy_(b2)
    aux = pkg.StepAux()
    loss = pkg.contrastive_step(x, pm, ad, sb.text.to(dev), sb.text_mask.to(dev), aux=aux)
    loss.backward()
    torch.cuda.synchronize()
    print(f"step: loss {loss.item():.6f} oracle {st.loss.item():.6f} rel={abs(loss.item() - st.loss.item()) / abs(st.loss.item()):.3e}")
    for k, prm in (("fc1.weight", ad.fc1.weight), ("fc1.bias", ad.fc1.bias), ("fc2.weight", ad.fc2.weight), ("fc2.bias", ad.fc2.bias)):
        print(f"  grad {k}: cos={cos(prm.grad, gr[k]):.6f} maxrel={rel(prm.grad, gr[k]):.3e}")
    # same comparison against an oracle that rounds h1 exactly where the kernel does (bf16 operand of fc2)
    for hdt in (torch.bfloat16, torch.float16):
        w1f, b1f, w2f, b2f = (t.to(f).requires_grad_() for t in (sb.w1, sb.b1, sb.w2, sb.b2))
        z1 = sb.x.to(f) @ w1f.t() + b1f
        h1 = R.gelu_erf(z1)
        h1 = h1 + (h1.detach().to(hdt).to(f) - h1.detach())  # straight-through rounding
        a_ = R.gelu_erf(h1 @ w2f.t() + b2f)
        y_ = a_ / a_.pow(2).sum(-1, keepdim=True).sqrt().clamp_min(1e-12)
        p_, _ = R.l2_normalize(R.readout(y_, sb.prot_mask, "mix"))
        lo_ = R.infonce_rows(p_, st.t, torch.arange(B), 0.05)
        gs = torch.autograd.grad(lo_, (w1f, b1f, w2f, b2f))
        print(f"  vs oracle with h1 rounded to {str(hdt).split('.')[-1]}: loss rel={abs(loss.item() - lo_.item()) / abs(lo_.item()):.3e}")
        for (k, prm), g_ in zip((("fc1.weight", ad.fc1.weight), ("fc1.bias", ad.fc1.bias), ("fc2.weight", ad.fc2.weight), ("fc2.bias", ad.fc2.bias)), gs):
            print(f"    grad {k}: cos={cos(prm.grad, g_):.6f} maxrel={rel(prm.grad, g_):.3e}   [oracle-vs-oracle: cos={cos(g_, gr[k]):.6f} maxrel={rel(g_, gr[k]):.3e}]")
    am_r, am_c = R.retrieval_argmax(st.p, st.t)
    print(f"  argmax_row ok={torch.equal(aux.argmax_row.cpu().long(), am_r)} argmax_col ok={torch.equal(aux.argmax_col.cpu().long(), am_c)}")

    # 7b. the backward chain stage by stage (same kernels the fused step calls)
    labels_i = torch.arange(B, dtype=torch.int32, device=dev)
    t_f32 = pkg.text_embeddings(sb.text.to(dev), sb.text_mask.to(dev), dtype=torch.float32)
    res = core.infonce_forward(p_bf, t_bf, labels_i, 0.05, need_grad=True, p_f32=p_f32, t_f32=t_f32)
    dp, _ = core.infonce_backward(res, p_bf, t_bf, 0.05, p_f32=p_f32, t_f32=t_f32)
    de = core.l2norm_backward(dp, p_f32, pn)
    c1, c2 = core.pool_backward_coef(de, stats, plan, d_out, "mix")
    dz2, db2 = core.adapter_tail_backward(acts, inv_norm, plan, c1, c2)
    torch.cuda.synchronize()
    nseq = sb.prot_lens.to(f)[:, None]
    mu_o, sd_o = st.e_prot[:, :d_out], st.e_prot[:, d_out:]
    c2o = gr["de"][:, d_out:] / (nseq * sd_o)
    c1o = gr["de"][:, :d_out] / nseq - c2o * mu_o
    dz2o = gr["dz2"].view(B, L, d_out)[valid]
    print(f"bwd chain: dp rel={rel(dp, gr['dp']):.3e} cos={cos(dp, gr['dp']):.6f} | de rel={rel(de, gr['de']):.3e} cos={cos(de, gr['de']):.6f} | "
          f"c1 rel={rel(c1, c1o):.3e} c2 rel={rel(c2, c2o):.3e} cos={cos(c2, c2o):.6f}")
    print(f"           dz2 rel={rel(dz2[:n], dz2o):.3e} cos={cos(dz2[:n], dz2o):.6f} db2 cos={cos(db2, gr['fc2.bias']):.6f}")
    # dz2 recomputed by the oracle from OUR c1/c2/y (isolates the tail kernel itself)
    yv = tr.y[valid]
    seq_id = torch.repeat_interleave(torch.arange(B), sb.prot_lens)
    dy_ours = c1.cpu()[seq_id] + c2.cpu()[seq_id] * yv
    da = (dy_ours - yv * (yv * dy_ours).sum(-1, keepdim=True)) / tr.norm[valid]
    dz2_mix = da * R.gelu_erf_grad(tr.z2[valid])
    print(f"           dz2 vs oracle-tail(our c1,c2): rel={rel(dz2[:n], dz2_mix):.3e} cos={cos(dz2[:n], dz2_mix):.6f}")

    # 8. module API: y = adapter(x) and its backward
    ad.zero_grad()
    xg = x.clone().requires_grad_()
    y = ad(xg)
    gy = torch.randn(B, L, d_out, generator=torch.Generator().manual_seed(3)).to(torch.bfloat16)
    (y * gy.to(dev)).sum().backward()
    torch.cuda.synchronize()
    flat = R.adapter_rows(sb.x.to(f).reshape(-1, d_in), sb.w1.to(f), sb.b1.to(f), sb.w2.to(f), sb.b2.to(f))
    go = R.adapter_rows_backward(flat, gy.to(f).reshape(-1, d_out), sb.w1.to(f), sb.w2.to(f), need_dx=True)
    print(f"module: y rel={rel(y, flat.y.view(B, L, d_out)):.3e}")
    for k, prm in (("fc1.weight", ad.fc1.weight), ("fc1.bias", ad.fc1.bias), ("fc2.weight", ad.fc2.weight), ("fc2.bias", ad.fc2.bias)):
        print(f"  grad {k}: cos={cos(prm.grad, go[k]):.6f} maxrel={rel(prm.grad, go[k]):.3e}")
    print(f"  grad x: cos={cos(xg.grad, go['dx']):.6f} maxrel={rel(xg.grad, go['dx'].view(B, L, d_in)):.3e}")

    # 9. readout_embeddings
    emb = sb.text.to(dev).clone().requires_grad_()
    for fn in ("last", "mean", "std", "mix"):
        out = pkg.readout_embeddings(emb, sb.text_mask.to(dev), fn)
        gd = torch.randn(out.shape, generator=torch.Generator().manual_seed(4)).to(torch.bfloat16)
        (g,) = torch.autograd.grad((out * gd.to(dev)).sum(), emb)
        torch.cuda.synchronize()
        ro = R.readout(sb.text.to(f), sb.text_mask, fn)
        go_ = R.readout_backward(sb.text.to(f), sb.text_mask, fn, gd.to(f))
        print(f"readout {fn}: out rel={rel(out, ro):.3e} grad rel={rel(g, go_):.3e}")

    # 10. dropout: mask statistics and a step with the kernels' own masks fed to the oracle
    pdrop, seed = 0.3, 12345
    k1 = core.dropout_mask(plan.rows_cap, d_mid, pdrop, seed, 1, dev)[:n]
    k2 = core.dropout_mask(plan.rows_cap, d_out, pdrop, seed, 2, dev)[:n]
    print(f"dropout: keep frac layer1={float((k1 > 0).float().mean()):.4f} layer2={float((k2 > 0).float().mean()):.4f} (expect 0.7000) scale={float(k1.max()):.4f}")
    acts_d = core.adapter_forward(xp, plan.rows_cap, plan.rows_cap, plan.n_rows, w1, b1, w2, b2, pdrop, seed, True)
    torch.cuda.synchronize()
    trd = R.adapter_rows(sb.x[valid].to(f), sb.w1.to(f), sb.b1.to(f), sb.w2.to(f), sb.b2.to(f), k1.cpu(), k2.cpu())
    print(f"dropout fwd: h1 rel={rel(acts_d.h1[:n], trd.h1):.3e} a rel={rel(acts_d.a[:n], trd.a):.3e}")
    print(f"launches so far: {pkg._lib.launch_count()}")


if __name__ == "__main__":
    main()
