"""CPU: the oracle's restatements of the SURVEY.md §8f rows against golden vectors dumped from the real reference
(`prepare_decoder_inputs`) and from the third-party code its training loop calls (clip_grad_norm_ + AdamW), and the
host-side text hand-off (`llm_hidden_states_at`) against the stock HF decoders it truncates."""
import importlib
import os

import numpy as np
import pytest
import torch

from oracle import restatement as R


def _load(golden_dir, name):
    z = np.load(os.path.join(golden_dir, name))
    return {k: torch.from_numpy(np.asarray(z[k])) for k in z.files}


def test_adamw_and_clip_restatement_matches_torch_golden(golden_dir):
    g = _load(golden_dir, "next_adamw.npz")
    n, lr, max_norm = int(g["n_steps"]), float(g["lr"]), float(g["max_norm"])
    params = [g[f"p0.{i}"].double() for i in range(4)]
    m = [torch.zeros_like(p) for p in params]
    v = [torch.zeros_like(p) for p in params]
    clipped_steps = 0
    for step in range(n):
        grads = [g[f"g{step}.{i}"].double() for i in range(4)]
        norm, grads = R.clip_grad_norm(grads, max_norm)
        torch.testing.assert_close(norm.float(), g[f"norm{step}"].float(), rtol=1e-6, atol=0)
        clipped_steps += int(norm > max_norm)
        params, m, v = R.adamw_step(params, grads, m, v, step + 1, lr)
        for i in range(4):
            torch.testing.assert_close(params[i].float(), g[f"p{step + 1}.{i}"], rtol=2e-6, atol=2e-7)
    assert clipped_steps == 2  # the fixture exercises both branches of the clip


def test_adamw_restatement_matches_live_torch():
    torch.manual_seed(3)
    ps = [torch.nn.Parameter(torch.randn(5, 7, dtype=torch.float64)), torch.nn.Parameter(torch.randn(7, dtype=torch.float64))]
    opt = torch.optim.AdamW(ps, lr=1e-2, eps=1e-6, betas=(0.9, 0.999), weight_decay=0.05)
    mine = [p.detach().clone() for p in ps]
    m = [torch.zeros_like(p) for p in mine]
    v = [torch.zeros_like(p) for p in mine]
    for step in range(1, 6):
        grads = [torch.randn_like(p) for p in ps]
        for p, gr in zip(ps, grads):
            p.grad = gr.clone()
        tn = torch.nn.utils.clip_grad_norm_(ps, max_norm=1.0)
        opt.step()
        norm, cg = R.clip_grad_norm(grads, 1.0)
        torch.testing.assert_close(norm, tn.double())
        mine, m, v = R.adamw_step(mine, cg, m, v, step, 1e-2, weight_decay=0.05)
        for a, b in zip(mine, ps):
            torch.testing.assert_close(a, b.detach(), rtol=1e-12, atol=1e-14)


def test_placeholder_scatter_restatement_matches_reference_golden(golden_dir):
    g = _load(golden_dir, "next_scatter.npz")
    w1, b1, w2, b2 = (g["sd.fc1.weight"].requires_grad_(), g["sd.fc1.bias"].requires_grad_(),
                      g["sd.fc2.weight"].requires_grad_(), g["sd.fc2.bias"].requires_grad_())
    x, enc_mask, ids = g["x"], g["enc_mask"], g["input_ids"]
    B, L, _ = x.shape
    y = R.adapter_rows(x.reshape(B * L, -1), w1, b1, w2, b2).y.reshape(B, L, -1)
    ph_mask = ids == int(g["placeholder_id"])
    base = g["table"][ids]
    out = R.placeholder_scatter(base, ph_mask, y, enc_mask)
    torch.testing.assert_close(out, g["embeds"], rtol=2e-5, atol=2e-6)
    (out * g["gy"]).sum().backward()
    for k, prm in (("fc1.weight", w1), ("fc1.bias", b1), ("fc2.weight", w2), ("fc2.bias", b2)):
        torch.testing.assert_close(prm.grad, g["grad." + k], rtol=2e-4, atol=2e-6)
    # rows outside the placeholders are the token embeddings, untouched
    assert torch.equal(out[~ph_mask], base[~ph_mask])


def test_mean_allreduce_restatement():
    torch.manual_seed(0)
    per_rank = [torch.randn(1000).to(torch.bfloat16) for _ in range(4)]
    out = R.mean_allreduce_bf16(per_rank)
    exact = torch.stack([p.double() for p in per_rank]).mean(0)
    assert out.dtype == torch.bfloat16
    assert (out.double() - exact).abs().max() <= exact.abs().max() * 2 ** -8


@pytest.mark.parametrize("family", ["llama", "qwen2"])
def test_truncated_llm_forward_returns_the_same_hidden_state(p2t, family):
    """scripts/train_contrast.py:292-304 keeps hidden_states[16] of a full forward; the hand-off helper must return
    that very tensor from a forward over the first `layer` blocks only — bit for bit, on the stock HF decoder."""
    handoff = importlib.import_module("p2t_b200.handoff")
    import transformers
    torch.manual_seed(0)
    kw = dict(vocab_size=97, hidden_size=32, intermediate_size=64, num_hidden_layers=5, num_attention_heads=4,
              num_key_value_heads=2, max_position_embeddings=64)
    if family == "llama":
        model = transformers.LlamaModel(transformers.LlamaConfig(**kw)).eval()
    else:
        model = transformers.Qwen2Model(transformers.Qwen2Config(**kw)).eval()
    ids = torch.randint(0, 97, (3, 11))
    mask = torch.ones(3, 11, dtype=torch.long)
    mask[1, 7:] = 0
    mask[2, 4:] = 0
    with torch.no_grad():
        full = model(input_ids=ids, attention_mask=mask, use_cache=False, output_hidden_states=True, return_dict=True)
    for layer in (0, 2, 3, 5):
        got = handoff.llm_hidden_states_at(model, ids, mask, layer=layer)
        assert torch.equal(got, full.hidden_states[layer]), layer
    # the decoder is restored afterwards
    assert len(model.layers) == 5 and not isinstance(model.norm, torch.nn.Identity)
    with torch.no_grad():
        again = model(input_ids=ids, attention_mask=mask, use_cache=False, return_dict=True).last_hidden_state
    assert torch.equal(again, full.last_hidden_state)
    with pytest.raises(ValueError):
        handoff.llm_hidden_states_at(model, ids, mask, layer=6)


def test_next_row_entries_have_no_cpu_path(p2t):
    optim = importlib.import_module("p2t_b200.optim")
    peer = importlib.import_module("p2t_b200.peer")
    handoff = importlib.import_module("p2t_b200.handoff")
    w = torch.nn.Parameter(torch.zeros(8, 8, dtype=torch.bfloat16))
    w.grad = torch.ones_like(w)
    with pytest.raises(p2t.P2TError, match="CUDA"):
        optim.FusedAdamW([w]).step()
    if not torch.cuda.is_available():
        with pytest.raises(p2t.P2TError, match="CUDA"):
            peer.PeerBuffer(4096)
    ad = p2t.ModalityAdapter(p2t.ModalityAdapterConfig(input_dim=16, intermediate_dim=32, output_dim=24)).to(torch.bfloat16)
    with pytest.raises(p2t.P2TError, match="CUDA"):
        handoff.adapter_into_embeds(ad, torch.zeros(1, 2, 16, dtype=torch.bfloat16), None,
                                    torch.zeros(1, 4, 24, dtype=torch.bfloat16), torch.zeros(1, 4, dtype=torch.bool))


def test_fused_adamw_state_dict_keeps_fp32_state(p2t):
    """torch's Optimizer.load_state_dict casts floating state to the parameter dtype (bf16); FusedAdamW restores its fp32
    moments / master weights and the int64 step counter from the incoming dict (host logic, no kernel involved)."""
    optim = importlib.import_module("p2t_b200.optim")
    w = torch.nn.Parameter(torch.zeros(4, 8, dtype=torch.bfloat16))
    opt = optim.FusedAdamW([w], lr=1e-3)
    fine = torch.full((4, 8), 1.0 + 2.0 ** -12)  # not representable in bf16
    opt.state[w] = {"exp_avg": fine.clone(), "exp_avg_sq": fine.clone() * 2, "master": fine.clone() * 3,
                    "step": torch.tensor([7], dtype=torch.int64)}
    sd = opt.state_dict()
    w2 = torch.nn.Parameter(torch.zeros(4, 8, dtype=torch.bfloat16))
    opt2 = optim.FusedAdamW([w2], lr=5e-4)
    opt2.load_state_dict(sd)
    st = opt2.state[w2]
    assert st["exp_avg"].dtype == torch.float32 and torch.equal(st["exp_avg"], fine)
    assert torch.equal(st["exp_avg_sq"], fine * 2) and torch.equal(st["master"], fine * 3)
    assert st["step"].dtype == torch.int64 and int(st["step"]) == 7
    assert opt2.param_groups[0]["lr"] == 1e-3  # hyper-parameters come from the checkpoint, as in torch


def test_fused_adamw_pointer_tables_follow_the_state(p2t):
    """The per-call pointer tables are cached while parameters and state tensors stay the same objects and rebuilt when
    a load_state_dict (or anything else) swaps them — exercised on CPU tensors, no kernel involved."""
    optim = importlib.import_module("p2t_b200.optim")
    w = torch.nn.Parameter(torch.zeros(4, 8, dtype=torch.bfloat16))
    b = torch.nn.Parameter(torch.zeros(8, dtype=torch.bfloat16))
    opt = optim.FusedAdamW([w, b], lr=1e-3)
    st = opt._group_state(0, [w, b])
    assert st["m"][0] == opt.state[w]["exp_avg"].data_ptr() and st["w"][1] == opt.state[b]["master"].data_ptr()
    assert opt.state[w]["step"].dtype == torch.int64 and "step" not in opt.state[b]
    tables = (st["m"], st["step"])
    assert opt._group_state(0, [w, b])["m"] is tables[0]          # cached
    sd = opt.state_dict()
    opt.load_state_dict(sd)                                        # new state tensor objects
    st2 = opt._group_state(0, [w, b])
    assert st2["m"] is not tables[0] and st2["m"][0] == opt.state[w]["exp_avg"].data_ptr()
    assert st2["step"] is opt.state[w]["step"]
    nomaster = optim.FusedAdamW([w], master_weights=False)
    assert nomaster._group_state(0, [w])["w"][0] is None and "master" not in nomaster.state[w]
