"""CPU: the repo's ModalityAdapter inside the REAL reference model class.

north_star: "keeps the ModalityAdapter nn.Module ... so it drops into Esm2LlamaInstructForCausalLM ... unchanged".
`Esm2LlamaInstructForCausalLM(esm_encoder=, adapter=, llama_decoder=)` (models/modeling_esm2llama_instruct.py:80-106)
takes the adapter as a component, so the drop-in is literally passing the repo class; this test builds a tiny random
model that way through oracle/reference_loader.py (tests only; needs /root/reference, absent on the GPU box) and checks
everything the reference's training scripts do to the adapter without running it: state-dict keys and strict loading
in both directions (scripts/train_contrast.py:183,679-685), `.to(bfloat16)` (:159), `requires_grad_` freezing
(:186-187), PEFT `modules_to_save=["adapter.fc1", "adapter.fc2"]` addressing (scripts/train_instruct.py:177-181), the
optimizer seeing the never-applied ln1/ln2 (:611-626), save_pretrained / from_pretrained of the whole composite.
The ESMC/Qwen twin (models/esmc_qwen_arc.py:179-186) calls the adapter the same way (`self.adapter(protein_embeddings)`,
positional, one tensor) but imports the EvolutionaryScale `esm` package, which is not installed here.
"""
import importlib

import pytest
import torch

from oracle.reference_loader import load_reference, reference_available

pytestmark = pytest.mark.skipif(not reference_available(), reason="reference tree not mounted (GPU box)")

KEYS = ["fc1.weight", "fc1.bias", "fc2.weight", "fc2.bias", "ln1.weight", "ln1.bias", "ln2.weight", "ln2.bias"]


def _tiny_parts():
    from transformers import EsmConfig, EsmModel, LlamaConfig, LlamaForCausalLM
    torch.manual_seed(0)
    esm = EsmModel(EsmConfig(vocab_size=33, hidden_size=64, num_hidden_layers=1, num_attention_heads=4, intermediate_size=128,
                             pad_token_id=1, position_embedding_type="rotary"), add_pooling_layer=False)
    llama = LlamaForCausalLM(LlamaConfig(vocab_size=128, hidden_size=96, intermediate_size=192, num_hidden_layers=2,
                                         num_attention_heads=4, num_key_value_heads=2, max_position_embeddings=128))
    return esm, llama


def test_repo_adapter_drops_into_the_reference_model_class(p2t, tmp_path):
    ref = load_reference()
    modeling = importlib.import_module("models.modeling_esm2llama_instruct")
    esm, llama = _tiny_parts()
    ours = p2t.ModalityAdapter(p2t.ModalityAdapterConfig(input_dim=64, intermediate_dim=128, output_dim=96, dropout_rate=0.3))
    model = modeling.Esm2LlamaInstructForCausalLM(esm_encoder=esm, adapter=ours, llama_decoder=llama, placeholder_id=5)
    assert model.adapter is ours and model.config.adapter_config.output_dim == 96

    # the reference's own adapter with the same config: identical parameter names and shapes, strict loading both ways
    theirs = ref.ModalityAdapter(ref.ModalityAdapterConfig(input_dim=64, intermediate_dim=128, output_dim=96, dropout_rate=0.3))
    assert list(ours.state_dict().keys()) == KEYS == list(theirs.state_dict().keys())
    ours.load_state_dict(theirs.state_dict(), strict=True)
    theirs.load_state_dict(ours.state_dict(), strict=True)
    sd = model.state_dict()
    assert all(f"adapter.{k}" in sd for k in KEYS)
    twin = modeling.Esm2LlamaInstructForCausalLM(esm_encoder=esm, adapter=theirs, llama_decoder=llama, placeholder_id=5)
    assert sorted(twin.state_dict().keys()) == sorted(sd.keys())
    twin.load_state_dict(sd, strict=True)  # a checkpoint written with the repo adapter loads into the stock model

    # what train_contrast.py does to the model before wrapping it (:159, :186-187)
    model = model.to(torch.bfloat16)
    assert model.adapter.fc1.weight.dtype == torch.bfloat16 and model.adapter.ln2.bias.dtype == torch.bfloat16
    model.esm_encoder.requires_grad_(False)
    model.llama_decoder.requires_grad_(False)
    model.adapter.requires_grad_(True)
    trainable = [n for n, p in model.named_parameters() if p.requires_grad]
    assert sorted(trainable) == sorted(f"adapter.{k}" for k in KEYS)  # ln1/ln2 included: never applied, never get a grad
    model.train()
    assert model.adapter.training and model.adapter.dropout_p() == pytest.approx(0.3)
    model.eval()
    assert model.adapter.dropout_p() == 0.0

    # PEFT modules_to_save=["adapter.fc1", "adapter.fc2"] resolves real nn.Linear sub-modules by name
    for name in ("adapter.fc1", "adapter.fc2"):
        sub = model.get_submodule(name)
        assert isinstance(sub, torch.nn.Linear) and sub.weight.requires_grad
    # the optimizer of :621-626 is built over model.parameters(): the adapter's eight tensors are there once each
    opt = torch.optim.AdamW([p for p in model.parameters() if p.requires_grad], lr=2e-4, eps=1e-6)
    assert len(opt.param_groups[0]["params"]) == 8

    # the composite checkpoints through HF with the repo adapter inside
    model.adapter.save_pretrained(tmp_path / "adapter")
    again = p2t.ModalityAdapter.from_pretrained(tmp_path / "adapter")
    assert all(torch.equal(again.state_dict()[k], model.adapter.state_dict()[k]) for k in KEYS)

    # and there is no silent CPU path behind the reference's call site (:191): the forward fails loudly off the GPU
    if not torch.cuda.is_available():
        ids = torch.randint(4, 30, (2, 9))
        with pytest.raises(p2t.P2TError, match="no CPU path"):
            model(protein_input_ids=ids, protein_attention_mask=torch.ones_like(ids), return_adapter_outputs=True)


def test_reference_step_function_accepts_a_model_with_the_repo_adapter(p2t):
    """get_sequence_embeddings / teacher_forcing_forward_pass (scripts/train_contrast.py:251-379) only touch
    `model(...)`, `model.module` and the adapter's outputs: the same table-lookup model that generated
    tests/golden/grid_step.npz runs with the repo's adapter class on the CPU up to the adapter call, where the repo
    class refuses (no CPU fallback) — i.e. the call signature is what the reference passes."""
    ref = load_reference()
    ours = p2t.ModalityAdapter(p2t.ModalityAdapterConfig(input_dim=16, intermediate_dim=32, output_dim=24))
    x = torch.randn(2, 5, 16)
    want = ref.ModalityAdapter(ref.ModalityAdapterConfig(input_dim=16, intermediate_dim=32, output_dim=24))
    want.load_state_dict(ours.state_dict())
    assert want.eval()(x).shape == (2, 5, 24)  # the reference class runs on the CPU ...
    if not torch.cuda.is_available():
        with pytest.raises(p2t.P2TError, match="no CPU path"):  # ... the repo class only on the GPU, and says so
            ours(x.to(torch.bfloat16))
