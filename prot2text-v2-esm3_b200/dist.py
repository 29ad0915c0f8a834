"""Multi-GPU form of the contrastive step: one process per GPU, batch sharded across ranks,
normalised text embeddings all-gathered over NCCL/NVLink to form global negatives (north_star;
the reference itself computes a local-only loss per rank and lets DDP average adapter grads,
scripts/train_contrast.py:551-556,611-614 — SURVEY.md D6).

Rank k holds pairs [k*B, (k+1)*B): it pools/projects its own rows, gathers t from all ranks and
computes the B x B_global block S_k = p_k T^T / tau.  The protein->text (row) term is then fully
local.  For the symmetric form only the per-column (max, sum-exp) statistics — 2*B_global floats —
cross ranks; because the text side carries no gradient (frozen LLM, :348-354) no E-wide gradient
has to be reduce-scattered.  When a caller does want d(loss)/d(text embeddings), `reduce_scatter_text_grad`
implements north_star's reduce-scatter of the gathered embeddings' gradient.

The returned loss is the LOCAL mean; averaging adapter grads across ranks (DDP) yields the
gradient of the global-batch mean loss.
"""
from __future__ import annotations

import os
from typing import Optional

import torch
import torch.distributed as dist

from . import step as _step
from .peer import PeerAllGather


class ShardedExchange:
    """The exchange channels of the sharded step over NVLink peer memory (csrc/peer.cu), built once per
    (pairs per rank, embedding width): `text` all-gathers the (B, E) fp32 unit-norm text embeddings, `stats` the
    (2, W*B) per-column (max, sum-exp) statistics of the symmetric loss.  With an exchange no NCCL call is on the data
    path and the sharded step is capturable in a CUDA graph (`graph.GraphedContrastiveStep(exchange=...)`)."""

    def __init__(self, pairs_per_rank: int, embed_dim: int, group=None, symmetric: bool = False, _buffers=None):
        self.group = group
        tb, sb = _buffers if _buffers is not None else (None, None)
        self.text = PeerAllGather(pairs_per_rank, embed_dim, torch.float32, group, _buffer=tb)
        self.rank, self.world = self.text.rank, self.text.world
        self.stats = PeerAllGather(2, self.world * pairs_per_rank, torch.float32, group, _buffer=sb) if symmetric else None

    @classmethod
    def virtual(cls, pairs_per_rank: int, embed_dim: int, world: int, symmetric: bool = False):
        """`world` exchanges in THIS process, one per simulated rank, over peer buffers that all live on the current
        device (single-GPU tests of the sharded step: the launches of all simulated ranks are issued phase by phase
        on one stream — every rank's push before any rank's arrival)."""
        from .peer import PeerBuffer
        tb = PeerBuffer.virtual(PeerAllGather.buffer_bytes(pairs_per_rank, embed_dim, torch.float32, world), world)
        sb = PeerBuffer.virtual(PeerAllGather.buffer_bytes(2, world * pairs_per_rank, torch.float32, world), world) \
            if symmetric else [None] * world
        return [cls(pairs_per_rank, embed_dim, symmetric=symmetric, _buffers=(tb[r], sb[r])) for r in range(world)]

    def push_column_stats(self, col_max: torch.Tensor, col_sum: torch.Tensor) -> None:
        if self.stats is None:
            raise ValueError("build the exchange with symmetric=True to merge column statistics")
        self.stats.push(torch.stack([col_max, col_sum]))

    def arrive_column_stats(self):
        gathered = self.stats.arrive().view(self.world, 2, -1)
        return _merge(gathered[:, 0], gathered[:, 1])

    def merge_column_stats(self, col_max: torch.Tensor, col_sum: torch.Tensor):
        self.push_column_stats(col_max, col_sum)
        return self.arrive_column_stats()

    def check(self) -> None:
        """Raise if a peer-memory wait timed out since the last check (synchronises the device)."""
        self.text.buffer.check()
        if self.stats is not None:
            self.stats.buffer.check()

    def close(self) -> None:
        self.text.close()
        if self.stats is not None:
            self.stats.close()


def balanced_shards(lengths, world: int):
    """Assign the pairs of a global batch to `world` ranks, the same number of pairs each, so that the ranks' residue
    row totals are as equal as a greedy longest-first pass gets them.  The step time of a rank is proportional to its
    valid residue rows (the adapter GEMMs), and every rank waits for the slowest at the exchange; the reference's
    DistributedSampler (scripts/train_contrast.py:551-556) shards at random, which at 32 pairs of 50..1024 residues
    leaves the heaviest of 8 ranks ~13 % above the mean.  The loss is invariant under the permutation (labels follow
    the pairs).  Returns a list of `world` index lists into `lengths`."""
    lengths = [int(v) for v in lengths]
    n = len(lengths)
    if world < 1 or n % world:
        raise ValueError("the global batch must split into equal shards")
    per = n // world
    order = sorted(range(n), key=lambda i: (-lengths[i], i))
    shards, totals = [[] for _ in range(world)], [0] * world
    for i in order:
        open_ranks = [r for r in range(world) if len(shards[r]) < per]
        r = min(open_ranks, key=lambda k: (totals[k], k))
        shards[r].append(i)
        totals[r] += lengths[i]
    return shards


def _merge(m_all: torch.Tensor, s_all: torch.Tensor):
    """M = max_k m_k, S = sum_k s_k exp(m_k - M) over the leading (rank) axis."""
    m = m_all.max(dim=0).values
    s = (s_all * torch.exp(m_all - m)).sum(dim=0)
    return m, s


def all_gather_embeddings(t_local: torch.Tensor, group=None, async_op: bool = False):
    """(B, E) per rank -> (W*B, E), rank-major.  One NCCL all-gather (B*E*elem bytes per rank).
    With async_op the collective stays on NCCL's stream and a zero-argument callable is returned that waits
    for it (on the then-current stream) and hands out the gathered tensor."""
    world = dist.get_world_size(group)
    out = torch.empty(world * t_local.shape[0], t_local.shape[1], dtype=t_local.dtype, device=t_local.device)
    work = dist.all_gather_into_tensor(out, t_local.contiguous(), group=group, async_op=async_op)
    if not async_op:
        return out

    def wait():
        work.wait()
        return out
    return wait


def merge_column_stats(col_max: torch.Tensor, col_sum: torch.Tensor, group=None):
    """Combine per-rank online-softmax column statistics: M = max_k m_k, S = sum_k s_k exp(m_k - M)."""
    world = dist.get_world_size(group)
    packed = torch.stack([col_max, col_sum])  # (2, C)
    gathered = torch.empty(world * 2, col_max.shape[0], dtype=packed.dtype, device=packed.device)
    dist.all_gather_into_tensor(gathered, packed, group=group)
    gathered = gathered.view(world, 2, -1)
    return _merge(gathered[:, 0], gathered[:, 1])


def reduce_scatter_text_grad(dt_full: torch.Tensor, group=None) -> torch.Tensor:
    """(W*B, E) local contribution to d(loss)/d(gathered text) -> (B, E) summed over ranks."""
    world = dist.get_world_size(group)
    out = torch.empty(dt_full.shape[0] // world, dt_full.shape[1], dtype=dt_full.dtype, device=dt_full.device)
    dist.reduce_scatter_tensor(out, dt_full.contiguous(), op=dist.ReduceOp.SUM, group=group)
    return out


def distributed_contrastive_step(residue_states, protein_mask, adapter, text_hidden, text_mask, *,
                                 residue_lengths=None, text_lengths=None,
                                 temperature: float = 0.05, symmetric: bool = False, group=None,
                                 aux: Optional[_step.StepAux] = None, exchange: Optional[ShardedExchange] = None,
                                 **step_kw) -> torch.Tensor:
    """Sharded-batch contrastive step with all-gathered negatives; returns the local mean loss.
    Accepts the padded (mask) or the packed (lengths) input form of `contrastive_step`.
    `exchange`: a ShardedExchange — the gather (and the column-statistics merge) then run as peer-memory kernels on
    the current stream instead of NCCL collectives."""
    if exchange is not None:
        rank, world = exchange.rank, exchange.world
    else:
        rank, world = dist.get_rank(group), dist.get_world_size(group)
    # the gather travels behind the protein side's kernels.  Peer-memory form: text branch and push run on the side
    # stream, which joins in front of the (persistent, one-CTA-per-SM) adapter GEMMs; the loss kernel waits for the
    # gathered rows itself and reads them in place (or, symmetric form, an arrive kernel right before the similarity).
    # NCCL form: the collective is waited for right before the adapter GEMMs.
    text_join = None
    if exchange is not None:
        if os.environ.get("P2T_TEXT_STREAM", "1") != "0":
            _, text_join = _step._text_branch(text_hidden, text_mask, text_lengths, after=exchange.text.push)
        else:
            exchange.text.push(_step.text_embeddings(text_hidden, text_mask, dtype=torch.float32, text_lengths=text_lengths))
        t_global = exchange.text
    else:
        t_local = _step.text_embeddings(text_hidden, text_mask, dtype=torch.float32, text_lengths=text_lengths)
        t_global = all_gather_embeddings(t_local, group, async_op=True) if world > 1 else t_local
    B = residue_lengths.shape[0] if residue_lengths is not None else residue_states.shape[0]
    labels = _step._rank_labels(rank, B, residue_states.device)
    if symmetric and world > 1:
        hook = exchange.merge_column_stats if (exchange is not None and exchange.stats is not None) \
            else (lambda m, s: merge_column_stats(m, s, group))
    else:
        hook = None
    return _step.contrastive_step(residue_states, protein_mask, adapter, text_embeds=t_global,
                                  residue_lengths=residue_lengths,
                                  temperature=temperature, symmetric=symmetric, labels=labels, aux=aux,
                                  col_stats_hook=hook, all_cols_labelled=symmetric and world > 1,
                                  late_text=exchange is not None, _text_join=text_join, **step_kw)
