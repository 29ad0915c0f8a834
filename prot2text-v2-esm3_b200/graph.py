"""CUDA-graph replay of the whole contrastive step (forward + backward).

One eager step launches ~45 small and large kernels from Python; between dependent kernels the GPU idles for a
launch latency each, and the host needs ~1 ms per step to issue them.  Every kernel of the step takes its ragged
extents from DEVICE memory (row counts, chunk tables, dropout seed), so the launch sequence depends only on the
padded shapes: it is captured once per (input buffers, shapes) and replayed with a single launch.

    step = GraphedContrastiveStep(adapter, residue_states, protein_mask, text_hidden, text_mask)
    for ...:                      # refill the SAME input tensors in place (or let the trunks write into them)
        loss = step.replay()      # fp32 0-dim tensor; adapter.fc1/fc2 .grad hold this step's gradients
        optimizer.step()

The gradients are written (not accumulated) into static tensors that `replay()` binds to `param.grad`.
Dropout (training mode) uses a device-side seed that every replay increments.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib
from .adapter import ModalityAdapter
from .step import StepAux, contrastive_step, step_backward


class GraphedContrastiveStep:
    def __init__(self, adapter: ModalityAdapter, residue_states: torch.Tensor, protein_mask: torch.Tensor,
                 text_hidden: torch.Tensor, text_mask: torch.Tensor, *, temperature: float = 0.05,
                 contrastive_num_segments: int = 1, symmetric: bool = False, seed: int = 0, warmup: int = 2,
                 exchange=None, grad_reducer=None, optimizer=None):
        if not residue_states.is_cuda:
            raise _lib.P2TError("GraphedContrastiveStep needs CUDA tensors: this package has no CPU path")
        self.adapter = adapter
        self.inputs = (residue_states, protein_mask, text_hidden, text_mask)
        # Sharded step: pass `exchange` (dist.ShardedExchange).  Its gather is a pair of peer-memory kernels, so the
        # exchange is captured with everything else.  (Capturing the NCCL all-gather instead hung on this stack —
        # torch 2.11, NCCL 2.28.9, async_op + wait inside capture.)  `grad_reducer` (peer.PeerGradAllReduce) appends
        # the mean all-reduce of the four weight gradients, DDP's job in the reference (scripts/train_contrast.py:611-614).
        # `optimizer` (optim.FusedAdamW over the adapter's four tensors) appends clip + AdamW: one replay is then a whole
        # training step — forward, backward, gradient mean over ranks, parameter update.  The warm-up passes below run
        # without it (they must not move the weights); its state is created by one dry step at learning rate 0 whose
        # moments and step count are zeroed again before the capture.
        self.exchange, self.grad_reducer, self.optimizer = exchange, grad_reducer, optimizer
        self._with_optimizer = False
        if exchange is not None and contrastive_num_segments != 1:
            raise ValueError("the sharded step averages over the whole local batch: contrastive_num_segments must be 1")
        self.kw = dict(temperature=temperature, contrastive_num_segments=contrastive_num_segments, symmetric=symmetric)
        self.params = [adapter.fc1.weight, adapter.fc1.bias, adapter.fc2.weight, adapter.fc2.bias]
        dev = residue_states.device
        self.seed = torch.full((1,), int(seed), dtype=torch.int64, device=dev)
        self.aux = StepAux()
        self.graph = torch.cuda.CUDAGraph()
        self.loss: Optional[torch.Tensor] = None
        self.launches_per_replay = 0
        # warm up on a side stream (lazy one-time setup: function attributes, driver entry points, allocator pools)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                self._eager_step()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        if optimizer is not None:
            self._prime_optimizer(dev)
        before = _lib.launch_count()
        self._with_optimizer = optimizer is not None
        with torch.cuda.graph(self.graph):
            self.loss, self.grads = self._eager_step()  # static tensors owned by the graph's memory pool
        self.launches_per_replay = _lib.launch_count() - before

    @torch.no_grad()
    def _eager_step(self):
        """One step without the autograd engine (its worker thread and AccumulateGrad stream bookkeeping do not mix
        with stream capture): forward, then the explicit backward for an upstream gradient of 1."""
        self.seed.add_(1)  # captured: each replay draws a fresh dropout mask
        x, pm, th, tm = self.inputs
        if self.exchange is None:
            loss, state = contrastive_step(x, pm, self.adapter, th, tm, aux=self.aux, seed_dev=self.seed, _raw=True, **self.kw)
        else:
            from .dist import distributed_contrastive_step
            loss, state = distributed_contrastive_step(x, pm, self.adapter, th, tm, aux=self.aux, seed_dev=self.seed,
                                                       _raw=True, exchange=self.exchange, temperature=self.kw["temperature"],
                                                       symmetric=self.kw["symmetric"])
        grads = list(step_backward(state, None))
        if self.grad_reducer is not None:
            self.grad_reducer.reduce_(grads)
        if self._with_optimizer:
            for p, g in zip(self.params, grads):
                p.grad = g
            self.optimizer.step()
        return loss, grads

    def _prime_optimizer(self, dev) -> None:
        """Allocate the optimizer's state outside the capture without moving weights, moments or the step count."""
        opt = self.optimizer
        if opt._dev and all("exp_avg" in opt.state[p] for p in self.params):
            return  # already stepped before: nothing would be allocated inside the capture
        lrs = [g["lr"] for g in opt.param_groups]
        for p in self.params:
            p.grad = torch.zeros_like(p)
        for g in opt.param_groups:
            g["lr"] = 0.0
        opt.step()  # fresh state, lr 0, zero gradients: parameters and (zero) moments keep their values; only the counter moves
        for p in self.params:
            st = opt.state[p]
            if "step" in st:
                st["step"].zero_()
        for g, lr in zip(opt.param_groups, lrs):
            g["lr"] = lr
        for gi, lr in enumerate(lrs):
            opt.set_lr(lr, gi)
        torch.cuda.synchronize(dev)

    def replay(self) -> torch.Tensor:
        """Run the captured step on the current contents of the input tensors; returns the (static) loss tensor."""
        self.graph.replay()
        for p, g in zip(self.params, self.grads):
            p.grad = g
        return self.loss
