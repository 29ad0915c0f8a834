"""CPU: host-side mirror of the reference interface — module structure, state-dict contract,
synthetic-input generator, workload definitions."""
import os

import numpy as np
import pytest
import torch


def test_adapter_state_dict_contract(p2t, golden_dir):
    """Keys exactly fc1/fc2/ln1/ln2 (reference checkpoints load strict, train_contrast.py:183)."""
    g = np.load(os.path.join(golden_dir, "adapter_eval_f32.npz"))
    sd = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd.")}
    cfg = p2t.ModalityAdapterConfig(input_dim=16, intermediate_dim=32, output_dim=24)
    ad = p2t.ModalityAdapter(cfg)
    assert sorted(ad.state_dict().keys()) == sorted(sd.keys()) == sorted(
        ["fc1.weight", "fc1.bias", "fc2.weight", "fc2.bias", "ln1.weight", "ln1.bias", "ln2.weight", "ln2.bias"])
    ad.load_state_dict(sd, strict=True)
    assert torch.equal(ad.fc1.weight, sd["fc1.weight"])
    # fc1/fc2 are addressable sub-modules (PEFT modules_to_save, train_instruct.py:177-181)
    assert isinstance(ad.fc1, torch.nn.Linear) and isinstance(ad.fc2, torch.nn.Linear)
    assert ad.fc1.weight.requires_grad
    assert ad.config_class is p2t.ModalityAdapterConfig and ad.config.dropout_rate == 0.3
    ad = ad.to(torch.bfloat16)
    assert ad.fc2.weight.dtype == torch.bfloat16
    ad.train(); assert ad.dropout_p() == pytest.approx(0.3)
    ad.eval(); assert ad.dropout_p() == 0.0


def test_adapter_is_a_pretrained_model(p2t, tmp_path):
    from transformers import PreTrainedModel
    cfg = p2t.ModalityAdapterConfig(input_dim=8, intermediate_dim=16, output_dim=8, dropout_rate=0.1)
    ad = p2t.ModalityAdapter(cfg)
    assert isinstance(ad, PreTrainedModel)
    ad.save_pretrained(tmp_path)
    again = p2t.ModalityAdapter.from_pretrained(tmp_path)
    assert again.config.dropout_rate == 0.1
    for (k1, v1), (k2, v2) in zip(ad.state_dict().items(), again.state_dict().items()):
        assert k1 == k2 and torch.equal(v1, v2)


def test_loss_modules_signature(p2t):
    assert p2t.BatchInfoNCELoss().temperature == 0.05
    assert p2t.SegmentedBatchInfoNCELoss(temperature=0.07).temperature == 0.07
    assert isinstance(p2t.BatchInfoNCELoss(), torch.nn.Module)


def test_synthetic_batches(p2t):
    import importlib
    synth = importlib.import_module("p2t_b200.synth")
    a = synth.make_config_batch("tiny")
    b = synth.make_config_batch("tiny")
    assert torch.equal(a.x, b.x) and torch.equal(a.w1, b.w1)  # seeded
    assert a.x.dtype == torch.bfloat16 and a.prot_mask.dtype == torch.int64
    assert torch.equal(a.prot_mask.sum(1), a.prot_lens)
    assert (a.x[a.prot_mask == 0] == 0).all()
    left = synth.make_config_batch("tiny", left_pad=True)
    assert left.prot_mask[:, -1].all() and torch.equal(left.prot_mask.sum(1), left.prot_lens)
    other_rank = synth.make_config_batch("tiny", rank=1)
    assert not torch.equal(a.x, other_rank.x) and torch.equal(a.w1, other_rank.w1)
    c2 = synth.CONFIGS["cfg2_esm2_3b_llama8b"]
    assert (c2["d_in"], c2["d_mid"], c2["d_out"], c2["batch"], c2["lmin"], c2["lmax"]) == (2560, 2048, 4096, 32, 50, 1024)
    c4 = synth.CONFIGS["cfg4_esmc600m_qwen7b"]
    assert (c4["d_in"], c4["d_out"]) == (1152, 3584)


def test_bench_reference_arm_runs_on_cpu(monkeypatch, capsys):
    """`bench.py --impl reference` needs no GPU and prints the contract's JSON line."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--workload", "tiny"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "pairs/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["e2e"]["h2d_bytes_per_step"] == 0


def test_balanced_shards_equalise_residue_rows(p2t):
    import importlib
    pdist = importlib.import_module("p2t_b200.dist")
    g = torch.Generator().manual_seed(1234)
    lengths = torch.randint(50, 1025, (256,), generator=g).tolist()
    shards = pdist.balanced_shards(lengths, 8)
    assert sorted(i for s in shards for i in s) == list(range(256)) and all(len(s) == 32 for s in shards)
    totals = [sum(lengths[i] for i in s) for s in shards]
    mean = sum(totals) / 8
    assert max(totals) <= 1.01 * mean, totals
    random_totals = [sum(lengths[r * 32:(r + 1) * 32]) for r in range(8)]
    assert max(random_totals) > max(totals)
    with pytest.raises(ValueError):
        pdist.balanced_shards(lengths[:10], 4)


def test_device_side_synthetic_batch_has_the_same_ragged_structure(p2t):
    import importlib
    synth = importlib.import_module("p2t_b200.synth")
    a = synth.make_config_batch("tiny", rank=3, same_lengths_as_rank0=True)
    b = synth.make_config_batch("tiny", rank=3, same_lengths_as_rank0=True, device=torch.device("cpu"))
    assert torch.equal(a.prot_lens, b.prot_lens) and torch.equal(a.prot_mask, b.prot_mask) and torch.equal(a.text_mask, b.text_mask)
    assert b.x.dtype == torch.bfloat16 and b.x.shape == a.x.shape and b.text.shape == a.text.shape
    assert not b.x[b.prot_mask == 0].any() and torch.equal(a.w1, b.w1)
    r0 = synth.make_config_batch("tiny", rank=0)
    assert sorted(a.prot_lens.tolist()) == sorted(r0.prot_lens.tolist())  # rank 0's multiset, permuted


def test_pull_table_describes_the_valid_rows_of_a_padded_batch(p2t):
    """host_io.pull_table (the segment table of the pull-mode staging kernel): replaying it on the CPU reproduces the
    packed valid rows for right and left padding, skips empty sequences and counts the 32 KB pieces."""
    import importlib
    host_io = importlib.import_module("p2t_b200.host_io")
    gen = torch.Generator().manual_seed(4)
    B, L, D = 6, 90, 24
    x = torch.randn(B, L, D, generator=gen).to(torch.bfloat16)
    counts = torch.tensor([90, 0, 1, 57, 89, 13], dtype=torch.int32)
    for starts in (torch.zeros(B, dtype=torch.int32), (L - counts).to(torch.int32)):
        table, prefix = host_io.pull_table(starts, counts, L, D * 2)
        assert table.shape == (5, 3) and prefix.shape == (6,) and prefix[0] == 0
        raw = x.contiguous().view(torch.int16).reshape(-1)  # 2-byte elements
        out = torch.zeros(int(counts.sum()) * D, dtype=torch.int16)
        pieces = 0
        for (src, dst, n), before in zip(table.tolist(), prefix.tolist()):
            assert src % 16 == 0 and dst % 16 == 0 and n % 16 == 0 and before == pieces
            out[dst // 2:(dst + n) // 2] = raw[src // 2:(src + n) // 2]
            pieces += -(-n // host_io.PULL_PIECE_BYTES)
        assert int(prefix[-1]) == pieces
        want = torch.cat([x[b, int(starts[b]):int(starts[b]) + int(counts[b])] for b in range(B)])
        assert torch.equal(out.view(torch.bfloat16).reshape(-1, D), want)
