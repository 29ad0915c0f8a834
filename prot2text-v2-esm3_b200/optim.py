"""Gradient clipping + AdamW for the adapter parameters in three launches (csrc/optim.cu).

Replaces, in the reference's training loop (scripts/train_contrast.py:455-465, optimizer built at :621-626):

    gradnorm = torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm=...)
    optimizer.step()
    optimizer.zero_grad(set_to_none=True)

with

    optimizer = FusedAdamW(adapter.parameters(), lr=..., eps=1e-6, betas=(0.9, 0.999), max_grad_norm=...)
    optimizer.step()                    # clip + update (+ optional zero_grad) on the current stream, no host sync
    gradnorm = optimizer.grad_norm      # 0-dim device tensor, what clip_grad_norm_ would have returned

The update follows torch.optim.AdamW exactly (decoupled weight decay, bias corrections, eps outside the square
root); moments are fp32 and, by default, an fp32 master copy of each bf16 parameter carries the update (the bf16
parameter is its rounding).  `master_weights=False` keeps the bf16 parameter as the only weight state, which is what
the reference's bf16 `AdamW` does.  Step count, learning rate and clip coefficient live in device memory, so
`step()` can be captured in the same CUDA graph as the contrastive step; schedulers keep working through
`param_groups[i]["lr"]` (written to the device when it changes, outside capture) or `set_lr()`.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional

import torch

from . import _lib

MAX_TENSORS = 8


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 1e-2,
                 max_grad_norm: Optional[float] = None, master_weights: bool = True, zero_grad_in_step: bool = False):
        if lr < 0 or eps < 0 or not (0 <= betas[0] < 1) or not (0 <= betas[1] < 1) or weight_decay < 0:
            raise ValueError("invalid AdamW hyper-parameter")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.max_grad_norm = max_grad_norm
        self.master_weights = master_weights
        self.zero_grad_in_step = zero_grad_in_step
        self._dev = {}      # per group index: device-side scalars and workspaces
        self.grad_norm: Optional[torch.Tensor] = None

    # ------------------------------------------------------------------------------------------
    def _group_state(self, gi: int, tensors):
        st = self._dev.get(gi)
        key = tuple(t.data_ptr() for t in tensors)
        if st is not None and st["key"] == key:
            return st
        dev = tensors[0].device
        numel = (C.c_longlong * len(tensors))(*[t.numel() for t in tensors])
        nws = int(_lib.load().p2t_adamw_workspace_floats(len(tensors), numel))
        if st is None:
            st = dict(step=torch.zeros(1, dtype=torch.int64, device=dev),
                      lr=torch.zeros(1, dtype=torch.float32, device=dev), lr_host=None,
                      scal=torch.zeros(4, dtype=torch.float32, device=dev))
        st.update(key=key, numel=numel, partial=torch.empty(max(nws, 1), dtype=torch.float32, device=dev))
        for t in tensors:
            s = self.state[t]
            if "exp_avg" not in s:
                s["exp_avg"] = torch.zeros(t.shape, dtype=torch.float32, device=dev)
                s["exp_avg_sq"] = torch.zeros(t.shape, dtype=torch.float32, device=dev)
                if self.master_weights:
                    s["master"] = t.detach().to(torch.float32).clone()
        vp = C.c_void_p * len(tensors)
        st["params"] = vp(*[t.data_ptr() for t in tensors])
        st["m"] = vp(*[self.state[t]["exp_avg"].data_ptr() for t in tensors])
        st["v"] = vp(*[self.state[t]["exp_avg_sq"].data_ptr() for t in tensors])
        st["w"] = vp(*[self.state[t]["master"].data_ptr() if self.master_weights else None for t in tensors])
        self._dev[gi] = st
        return st

    def set_lr(self, lr: float, group: int = 0) -> None:
        """Write the learning rate of a group to the device (use this between CUDA-graph replays)."""
        self.param_groups[group]["lr"] = lr
        st = self._dev.get(group)
        if st is not None:
            st["lr"].fill_(float(lr))
            st["lr_host"] = float(lr)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        capturing = torch.cuda.is_current_stream_capturing() if torch.cuda.is_available() else False
        for gi, group in enumerate(self.param_groups):
            tensors = [p for p in group["params"] if p.grad is not None]
            if not tensors:
                continue
            if len(tensors) > MAX_TENSORS:
                raise _lib.P2TError(f"FusedAdamW handles up to {MAX_TENSORS} tensors per parameter group "
                                    f"(the adapter has 4); got {len(tensors)}")
            for p in tensors:
                if not p.is_cuda:
                    raise _lib.P2TError("FusedAdamW needs CUDA parameters: this package has no CPU path")
                if p.dtype != torch.bfloat16 or p.grad.dtype != torch.bfloat16:
                    raise _lib.P2TError("FusedAdamW updates bfloat16 parameters with bfloat16 gradients")
                if not p.is_contiguous() or not p.grad.is_contiguous():
                    raise _lib.P2TError("FusedAdamW needs contiguous parameters and gradients")
            st = self._group_state(gi, tensors)
            lr = float(group["lr"])
            if not capturing and st["lr_host"] != lr:
                st["lr"].fill_(lr)
                st["lr_host"] = lr
            grads = (C.c_void_p * len(tensors))(*[p.grad.data_ptr() for p in tensors])
            b1, b2 = group["betas"]
            mx = self.max_grad_norm
            mx = 0.0 if (mx is None or math.isinf(mx)) else float(mx)
            _lib.call("p2t_adamw_step", len(tensors), st["params"], grads, st["m"], st["v"], st["w"], st["numel"],
                      st["partial"].data_ptr(), st["scal"].data_ptr(), st["lr"].data_ptr(), st["step"].data_ptr(),
                      float(b1), float(b2), float(group["eps"]), float(group["weight_decay"]), mx,
                      int(self.zero_grad_in_step), torch.cuda.current_stream().cuda_stream)
            self.grad_norm = st["scal"][0]
        return loss
