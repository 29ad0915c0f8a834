"""CUDA-graph replay of the whole contrastive step (forward + backward [+ gradient mean over ranks + optimizer]).

One eager step launches ~20 kernels from Python; between dependent kernels the GPU idles for a launch latency each,
and the host needs ~1 ms per step to issue them.  Every kernel of the step takes its ragged extents from DEVICE memory
(row counts, chunk tables, dropout seed), so the launch sequence depends only on the padded shapes: it is captured once
per (input buffers, shapes) and replayed with a single launch.

    step = GraphedContrastiveStep(adapter, residue_states, protein_mask, text_hidden, text_mask)
    for ...:                      # refill the SAME input tensors in place (or let the trunks write into them)
        loss = step.replay()      # fp32 0-dim tensor; adapter.fc1/fc2 .grad hold this step's gradients
        optimizer.step()

Gradient accumulation (the reference accumulates `gradient_accumulation_steps` = 8 micro-batches before
`optimizer.step()`, scripts/train_contrast.py:57,432,448-465): with `accumulation_steps=k` every k-th replay is a
BOUNDARY step.  The first micro-step of a window overwrites the static gradient buffers, the others add to them (bf16
read-modify-write in the weight-gradient GEMMs' epilogues, fp32 for the biases), each with an upstream gradient of
1/k; only the boundary step runs the gradient mean over ranks (`grad_reducer`) and the optimizer (`optimizer`) and
binds `param.grad`.  That is DDP's `no_sync()` on the non-boundary micro-steps, which the reference lacks
(SURVEY.md §8f-1).  Dropout (training mode) uses a device-side seed that every replay increments.

With a `grad_reducer` (peer.PeerGradAllReduce.for_adapter: all four gradients in fp32) the backward kernels write
the unrounded gradients straight into the reducer's channel buffer (no staging copies), the mean over ranks is formed
in fp32 in rank order, and the only rounding is the final conversion to the bf16 tensors bound to `param.grad`.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _core, _lib
from .adapter import ModalityAdapter
from .step import StepAux, contrastive_step, step_backward


class GraphedContrastiveStep:
    def __init__(self, adapter: ModalityAdapter, residue_states: torch.Tensor, protein_mask: torch.Tensor,
                 text_hidden: torch.Tensor, text_mask: torch.Tensor, *, temperature: float = 0.05,
                 contrastive_num_segments: int = 1, symmetric: bool = False, seed: int = 0, warmup: int = 2,
                 exchange=None, grad_reducer=None, optimizer=None, accumulation_steps: int = 1, check_every: int = 0,
                 max_valid_rows: Optional[int] = None):
        if not residue_states.is_cuda:
            raise _lib.P2TError("GraphedContrastiveStep needs CUDA tensors: this package has no CPU path")
        self.adapter = adapter
        self.inputs = (residue_states, protein_mask, text_hidden, text_mask)
        # Sharded step: pass `exchange` (dist.ShardedExchange).  Its gather is peer-memory kernels, so the exchange is
        # captured with everything else.  (Capturing the NCCL all-gather instead hung on this stack — torch 2.11,
        # NCCL 2.28.9, async_op + wait inside capture.)  `grad_reducer` (peer.PeerGradAllReduce) appends the mean
        # all-reduce of the four gradients, DDP's job in the reference (scripts/train_contrast.py:611-614);
        # `optimizer` (optim.FusedAdamW over the adapter's four tensors) appends clip + AdamW: one boundary replay is
        # then a whole training step.  `check_every` = N > 0: every N-th replay synchronises and raises if an exchange
        # round timed out (a timed-out round already turns the loss / the gradients into NaN).
        self.exchange, self.grad_reducer, self.optimizer = exchange, grad_reducer, optimizer
        if exchange is not None and contrastive_num_segments != 1:
            raise ValueError("the sharded step averages over the whole local batch: contrastive_num_segments must be 1")
        if accumulation_steps < 1:
            raise ValueError("accumulation_steps must be >= 1")
        self.kw = dict(temperature=temperature, contrastive_num_segments=contrastive_num_segments, symmetric=symmetric,
                       max_valid_rows=max_valid_rows)
        self.params = [adapter.fc1.weight, adapter.fc1.bias, adapter.fc2.weight, adapter.fc2.bias]
        dev = residue_states.device
        self.device = dev
        self.k = int(accumulation_steps)
        self.check_every = int(check_every)
        self.seed = torch.full((1,), int(seed), dtype=torch.int64, device=dev)
        self.dloss = torch.full((), 1.0 / self.k, dtype=torch.float32, device=dev) if self.k > 1 else None
        self.aux = StepAux()
        self.loss: Optional[torch.Tensor] = None
        self.launches_per_replay = 0
        self._micro = 0
        self._replays = 0
        self._graphs = {}
        self._warmup = max(1, warmup)
        # static gradient buffers: local accumulators (the reducer's contribution area when there is one) ...
        d_mid, d_in = adapter.fc1.weight.shape
        d_out = adapter.fc2.weight.shape[0]
        bf, f32 = torch.bfloat16, torch.float32
        self._dw_f32 = grad_reducer is not None
        self._overlapped = grad_reducer is not None and hasattr(grad_reducer, "late")
        if grad_reducer is not None:
            want = [((d_mid, d_in), f32), ((d_mid,), f32), ((d_out, d_mid), f32), ((d_out,), f32)]
            if [(s, d) for s, d in zip(grad_reducer.shapes, grad_reducer.dtypes)] != want:
                raise _lib.P2TError("grad_reducer must be built over the four fp32 gradients [dW1, db1, dW2, db2] "
                                    "(peer.PeerGradAllReduce.for_adapter)")
            self._dw1, self._db1_f32, self._dw2, self._db2_f32 = grad_reducer.views_in()
            self._reduced = grad_reducer.views_out()
        else:
            self._dw1 = torch.zeros(d_mid, d_in, dtype=bf, device=dev)
            self._dw2 = torch.zeros(d_out, d_mid, dtype=bf, device=dev)
            self._db1_f32 = torch.zeros(d_mid, dtype=f32, device=dev)
            self._db2_f32 = torch.zeros(d_out, dtype=f32, device=dev)
        # ... and what param.grad is bound to on a boundary step
        self._db1 = torch.zeros(d_mid, dtype=bf, device=dev)
        self._db2 = torch.zeros(d_out, dtype=bf, device=dev)
        if grad_reducer is not None:
            self.grads = [torch.zeros(d_mid, d_in, dtype=bf, device=dev), self._db1,
                          torch.zeros(d_out, d_mid, dtype=bf, device=dev), self._db2]
        else:
            self.grads = [self._dw1, self._db1, self._dw2, self._db2]
        if optimizer is not None:
            self._prime_optimizer(dev)
        # capture every variant of the window now: the warm-up passes of a capture run real kernels on the static
        # gradient buffers, which is harmless only before the first window starts (its first micro-step overwrites)
        variants = [(True, True)] if self.k == 1 else [(True, False), (False, True)] + ([(False, False)] if self.k > 2 else [])
        for first, last in variants:
            self._graph_for(first, last)

    # ------------------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def _eager_step(self, first: bool, last: bool, with_tail: bool, aux: StepAux):
        """One step without the autograd engine (its worker thread and AccumulateGrad stream bookkeeping do not mix
        with stream capture): forward, then the explicit backward."""
        self.seed.add_(1)  # captured: each replay draws a fresh dropout mask
        x, pm, th, tm = self.inputs
        if self.exchange is None:
            loss, state = contrastive_step(x, pm, self.adapter, th, tm, aux=aux, seed_dev=self.seed, _raw=True,
                                           dloss_dev=self.dloss, **self.kw)
        else:
            from .dist import distributed_contrastive_step
            loss, state = distributed_contrastive_step(x, pm, self.adapter, th, tm, aux=aux, seed_dev=self.seed,
                                                       _raw=True, exchange=self.exchange, dloss_dev=self.dloss,
                                                       temperature=self.kw["temperature"], symmetric=self.kw["symmetric"],
                                                       max_valid_rows=self.kw["max_valid_rows"])
        # on a boundary step without a reducer the bf16 bias gradients land directly in what param.grad is bound to
        tail = last and with_tail
        # boundary step with an OverlappedGradReduce: the mean of dW2 / db2 rides inside the dW1 GEMM's launch
        ov = self.grad_reducer.late if (tail and self._overlapped) else None
        step_backward(state, None, accumulate=not first, dw_out=(self._dw1, self._dw2), dw_f32=self._dw_f32,
                      db_f32_out=(self._db1_f32, self._db2_f32), db_bf16_out=(self._db1, self._db2), overlap_reduce=ov)
        if tail:
            if self.grad_reducer is not None:
                if self._overlapped:
                    self.grad_reducer.finish()
                else:
                    self.grad_reducer.exchange()
                if self.optimizer is None or not hasattr(self.optimizer, "fp32_grad_sources"):
                    for mean_f32, grad_bf16 in zip(self._reduced, self.grads):  # the one rounding: mean (fp32) -> param.grad (bf16)
                        _lib.call("p2t_f32_to_bf16", mean_f32.data_ptr(), mean_f32.numel(), grad_bf16.data_ptr(), _core._stream())
            if self.optimizer is not None:
                for p, g in zip(self.params, self.grads):
                    p.grad = g
                fused_round = self.grad_reducer is not None and hasattr(self.optimizer, "fp32_grad_sources")
                if fused_round:  # FusedAdamW's norm pass does that rounding itself (no separate conversion launches)
                    self.optimizer.fp32_grad_sources = dict(zip(self.params, self._reduced))
                try:
                    self.optimizer.step()
                finally:
                    if fused_round:
                        self.optimizer.fp32_grad_sources = {}
        return loss

    def _graph_for(self, first: bool, last: bool):
        key = (first, last)
        if key in self._graphs:
            return self._graphs[key]
        dev = self.device
        # warm up on a side stream (lazy one-time setup: function attributes, driver entry points, allocator pools);
        # the warm-up passes run without the reducer / optimizer tail (they must not move the weights, and a
        # collective must run the same number of times on every rank: it does, every rank captures the same variants)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(self._warmup):
                self._eager_step(first, last, with_tail=False, aux=StepAux())
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        g = torch.cuda.CUDAGraph()
        aux = StepAux()
        before = _lib.launch_count()
        with torch.cuda.graph(g):
            loss = self._eager_step(first, last, with_tail=True, aux=aux)  # static tensors owned by the graph's memory pool
        launches = _lib.launch_count() - before
        self._graphs[key] = (g, loss, launches, aux)
        if not self.launches_per_replay:
            self.launches_per_replay = launches
        return self._graphs[key]

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

    # ------------------------------------------------------------------------------------------------------------
    @property
    def is_boundary(self) -> bool:
        """True when the NEXT replay closes an accumulation window (runs the reducer / optimizer, binds .grad)."""
        return self._micro == self.k - 1

    def replay(self) -> torch.Tensor:
        """Run the captured step on the current contents of the input tensors; returns the (static) loss tensor of this
        micro-step (not divided by `accumulation_steps`).  On a boundary step `param.grad` is bound to the window's
        gradients (their mean over ranks with a `grad_reducer`)."""
        first, last = self._micro == 0, self._micro == self.k - 1
        g, loss, _, aux = self._graph_for(first, last)
        g.replay()
        self.loss, self.aux = loss, aux
        self._micro = 0 if last else self._micro + 1
        if last:
            for p, gr in zip(self.params, self.grads):
                p.grad = gr
        self._replays += 1
        if self.check_every and self._replays % self.check_every == 0:
            if self.exchange is not None:
                self.exchange.check()
            if self.grad_reducer is not None:
                self.grad_reducer.buffer.check()
        return loss
