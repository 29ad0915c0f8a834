"""B200-native Stage-1 contrastive hot path of Prot2Text-V2 (imported as `p2t_b200`).

Public surface = the reference's own call signatures for this path:
    ModalityAdapterConfig, ModalityAdapter           models/modeling_esm2llama_instruct.py:45-68
    readout_embeddings                               scripts/train_contrast.py:198-248
    BatchInfoNCELoss, SegmentedBatchInfoNCELoss      scripts/train_contrast.py:72-114
plus the fused entry `contrastive_step` (scripts/train_contrast.py:313-379 from trunk outputs on)
and its multi-GPU form in `dist` (exchange over NVLink peer memory: `peer`), and the callers either side of the path
(SURVEY.md §8f): `FusedAdamW` (clip + AdamW), `llm_hidden_states_at`, `adapter_into_embeds`.  All compute goes through the C-ABI CUDA library
(include/p2t_b200.h); there is no CPU fallback.
"""
from . import _lib
from ._lib import P2TError
from .adapter import ModalityAdapter, ModalityAdapterConfig
from .losses import BatchInfoNCELoss, SegmentedBatchInfoNCELoss, SymmetricInfoNCELoss
from .readout import readout_embeddings
from .graph import GraphedContrastiveStep
from .host_io import HostStager, StagedBatch
from .step import StepAux, contrastive_step, segment_pooling_mask, text_embeddings
from .optim import FusedAdamW
from .peer import OverlappedGradReduce, PeerAllGather, PeerBuffer, PeerGradAllReduce
from .handoff import adapter_into_embeds, llm_hidden_states_at

__all__ = [
    "ModalityAdapter", "ModalityAdapterConfig", "readout_embeddings", "BatchInfoNCELoss",
    "SegmentedBatchInfoNCELoss", "SymmetricInfoNCELoss", "contrastive_step", "text_embeddings", "segment_pooling_mask", "StepAux",
    "HostStager", "StagedBatch", "GraphedContrastiveStep",
    "FusedAdamW", "PeerAllGather", "PeerBuffer", "PeerGradAllReduce", "OverlappedGradReduce", "adapter_into_embeds", "llm_hidden_states_at",
    "P2TError",
]
