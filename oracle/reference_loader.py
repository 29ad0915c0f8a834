"""TEST INFRASTRUCTURE ONLY — loads the *real* reference implementation for pinning the oracle.

Only `tests/`, `oracle/make_golden.py` and nothing in the product package may import this.
It works where `/root/reference` is mounted (the authoring container) or where `oracle/build_ref.py`
has staged the reference's own hot-path files as `oracle/_ref/ref_hotpath.tgz` (the GPU box, which
has no /root/reference: the archive travels with the snapshot and is unpacked to a temporary
directory).  The golden vectors that pin the oracle are dumped to `tests/golden/` by
`oracle/make_golden.py` regardless.

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
_ARCHIVE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "ref_hotpath.tgz")


def _has_hot_path(root: str) -> bool:
    return os.path.isfile(os.path.join(root, "scripts", "train_contrast.py"))


def _staged_root() -> str | None:
    """Where /root/reference is absent (the GPU box): unpack oracle/_ref/ref_hotpath.tgz (packed from the reference's
    own files by oracle/build_ref.py, never committed) into a per-user temporary directory and use that as the root."""
    if not os.path.isfile(_ARCHIVE):
        return None
    import hashlib
    import tarfile
    import tempfile
    with open(_ARCHIVE, "rb") as fh:
        tag = hashlib.sha256(fh.read()).hexdigest()[:16]
    root = os.path.join(tempfile.gettempdir(), f"p2t_ref_{os.getuid()}_{tag}")
    if not _has_hot_path(root):
        os.makedirs(root, exist_ok=True)
        with tarfile.open(_ARCHIVE, "r:gz") as tar:
            tar.extractall(root, filter="data")
    return root


def reference_root() -> str | None:
    if _has_hot_path(REFERENCE_ROOT):
        return REFERENCE_ROOT
    return _staged_root()


def reference_available() -> bool:
    return reference_root() is not None


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
    root = reference_root()
    if root is None:
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT} and no staged archive at {_ARCHIVE}")
    if root not in sys.path:
        sys.path.insert(0, root)

    # (1) a bare `models` package so models/__init__.py (torch_geometric, esm) never runs
    models_pkg = types.ModuleType("models")
    models_pkg.__path__ = [os.path.join(root, "models")]
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
    scripts_pkg.__path__ = [os.path.join(root, "scripts")]
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
