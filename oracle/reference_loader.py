"""TEST INFRASTRUCTURE ONLY — loads the *real* reference implementation for pinning the oracle.

Only `tests/`, `oracle/make_golden.py` and nothing in the product package may import this.
It works only where `/root/reference` is mounted (the authoring container); the GPU box has
no such directory, so everything that must travel is dumped to `tests/golden/` by
`oracle/make_golden.py`.

The reference cannot be imported plainly: `models/__init__.py` pulls torch_geometric and the
EvolutionaryScale `esm` package, `scripts/train_contrast.py:37,44` pull `graphein` (through
`dataset`) and `esm`.  None of those is needed by the Stage-1 hot path, so the loader registers
empty stand-in modules for them and imports the two files that hold the hot path:

  * models/modeling_esm2llama_instruct.py  (ModalityAdapter, :45-68)
  * scripts/train_contrast.py              (BatchInfoNCELoss :72-91, SegmentedBatchInfoNCELoss
                                            :94-114, readout_embeddings :198-248,
                                            teacher_forcing_forward_pass :313-379)

No reference source is copied; the modules execute from where they lie.
"""
from __future__ import annotations

import importlib
import os
import sys
import types
from dataclasses import dataclass
from typing import Any

REFERENCE_ROOT = os.environ.get("P2T_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "scripts", "train_contrast.py"))


@dataclass
class ReferenceApi:
    ModalityAdapter: Any
    ModalityAdapterConfig: Any
    BatchInfoNCELoss: Any
    SegmentedBatchInfoNCELoss: Any
    readout_embeddings: Any
    teacher_forcing_forward_pass: Any
    get_sequence_embeddings: Any
    train_contrast: Any


_cached: ReferenceApi | None = None


def _stub(name: str, **attrs) -> types.ModuleType:
    mod = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(mod, k, v)
    sys.modules[name] = mod
    return mod


def load_reference() -> ReferenceApi:
    """Import the reference hot path with stand-ins for its unrelated heavy dependencies."""
    global _cached
    if _cached is not None:
        return _cached
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)

    # (1) a bare `models` package so models/__init__.py (torch_geometric, esm) never runs
    models_pkg = types.ModuleType("models")
    models_pkg.__path__ = [os.path.join(REFERENCE_ROOT, "models")]
    sys.modules["models"] = models_pkg
    modality_config = importlib.import_module("models.modality_config")
    modeling = importlib.import_module("models.modeling_esm2llama_instruct")

    # (2) stand-ins for what scripts/train_contrast.py imports but the hot path never touches
    class _Unavailable:  # pragma: no cover - never instantiated
        def __init__(self, *a, **k):
            raise RuntimeError("stand-in for a reference dependency outside the Stage-1 hot path")

    _stub("dataset", Prot2TextLightDataset=_Unavailable, Prot2TextLightCollater=_Unavailable)
    esm = _stub("esm")
    esm_models = _stub("esm.models")
    esm_esmc = _stub("esm.models.esmc", ESMC=_Unavailable)
    esm.models = esm_models
    esm_models.esmc = esm_esmc
    models_pkg.ModalityAdapter = modeling.ModalityAdapter
    models_pkg.ModalityAdapterConfig = modality_config.ModalityAdapterConfig
    models_pkg.ESMCConfig = _Unavailable
    models_pkg.ESMCQwen = _Unavailable

    # (3) the script module itself (argparse object is built at import, nothing is parsed)
    scripts_pkg = types.ModuleType("scripts")
    scripts_pkg.__path__ = [os.path.join(REFERENCE_ROOT, "scripts")]
    sys.modules["scripts"] = scripts_pkg
    tc = importlib.import_module("scripts.train_contrast")

    _cached = ReferenceApi(
        ModalityAdapter=modeling.ModalityAdapter,
        ModalityAdapterConfig=modality_config.ModalityAdapterConfig,
        BatchInfoNCELoss=tc.BatchInfoNCELoss,
        SegmentedBatchInfoNCELoss=tc.SegmentedBatchInfoNCELoss,
        readout_embeddings=tc.readout_embeddings,
        teacher_forcing_forward_pass=tc.teacher_forcing_forward_pass,
        get_sequence_embeddings=tc.get_sequence_embeddings,
        train_contrast=tc,
    )
    return _cached
