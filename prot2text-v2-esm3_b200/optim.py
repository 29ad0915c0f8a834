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
        # {parameter: fp32 tensor}: gradients whose fp32 source (the mean over ranks left by peer.PeerGradAllReduce)
        # is rounded into the bf16 .grad by the norm pass of the next step() — set by graph.GraphedContrastiveStep
        self.fp32_grad_sources: dict = {}
        self.grad_norm: Optional[torch.Tensor] = None

    # ------------------------------------------------------------------------------------------
    def _group_state(self, gi: int, tensors):
        """Device-side scalars and workspace of a parameter group, plus this call's pointer tables.  Per-parameter
        state (`exp_avg`, `exp_avg_sq`, `master`: fp32; `step`: int64, one shared counter per group, kept in the
        first parameter's state so that it travels with state_dict) is created on first use and re-read on every
        call, so a `load_state_dict` in between is picked up."""
        dev = tensors[0].device
        st = self._dev.get(gi)
        # fast path: same parameter storage and same state tensor objects as last call -> the tables are still valid
        key = tuple((t.data_ptr(), id(s.get("exp_avg")), id(s.get("exp_avg_sq")), id(s.get("master")))
                    for t, s in ((t, self.state[t]) for t in tensors)) + (id(self.state[tensors[0]].get("step")),)
        if st is not None and st.get("key") == key:
            return st
        if st is None:
            st = self._dev[gi] = dict(lr=torch.zeros(1, dtype=torch.float32, device=dev), lr_host=None,
                                      scal=torch.zeros(4, dtype=torch.float32, device=dev), nws=-1, partial=None)
        numel = (C.c_longlong * len(tensors))(*[t.numel() for t in tensors])
        nws = int(_lib.load().p2t_adamw_workspace_floats(len(tensors), numel))
        if nws != st["nws"]:
            st["partial"], st["nws"] = torch.empty(max(nws, 1), dtype=torch.float32, device=dev), nws
        for t in tensors:
            s = self.state[t]
            if "exp_avg" not in s:
                s["exp_avg"] = torch.zeros(t.shape, dtype=torch.float32, device=dev)
                s["exp_avg_sq"] = torch.zeros(t.shape, dtype=torch.float32, device=dev)
            if self.master_weights and "master" not in s:
                s["master"] = t.detach().to(torch.float32).clone()
            for k in ("exp_avg", "exp_avg_sq", "master"):
                if k in s and (s[k].dtype != torch.float32 or s[k].device != dev or not s[k].is_contiguous()):
                    s[k] = s[k].to(device=dev, dtype=torch.float32).contiguous()
        # the shared step counter lives with the group's FIRST parameter (whether or not it has a gradient this step),
        # so the set of tensors with gradients may change between steps without moving the counter
        s0 = self.state[self.param_groups[gi]["params"][0]]
        step = s0.get("step")
        if step is None:
            s0["step"] = torch.zeros(1, dtype=torch.int64, device=dev)
        elif step.dtype != torch.int64 or step.device != dev or step.numel() != 1:
            s0["step"] = step.to(device=dev, dtype=torch.int64).reshape(1).clone()
        vp = C.c_void_p * len(tensors)
        st["numel"] = numel
        st["step"] = s0["step"]
        st["params"] = vp(*[t.data_ptr() for t in tensors])
        st["m"] = vp(*[self.state[t]["exp_avg"].data_ptr() for t in tensors])
        st["v"] = vp(*[self.state[t]["exp_avg_sq"].data_ptr() for t in tensors])
        st["w"] = vp(*[self.state[t]["master"].data_ptr() if self.master_weights else None for t in tensors])
        st["key"] = tuple((t.data_ptr(), id(s.get("exp_avg")), id(s.get("exp_avg_sq")), id(s.get("master")))
                          for t, s in ((t, self.state[t]) for t in tensors)) + (id(s0["step"]),)
        return st

    def load_state_dict(self, state_dict) -> None:
        """torch casts floating-point optimizer state to the parameter's dtype on load (bf16 here); the moments and
        master weights of this optimizer are fp32 whatever the parameter is, so they are restored from the incoming
        dict afterwards, uncast."""
        super().load_state_dict(state_dict)
        by_id = {}
        for saved, group in zip(state_dict["param_groups"], self.param_groups):
            for pid, p in zip(saved["params"], group["params"]):
                by_id[pid] = p
        for pid, saved in state_dict["state"].items():
            p = by_id.get(pid)
            if p is None:
                continue
            for k in ("exp_avg", "exp_avg_sq", "master"):
                if k in saved:
                    self.state[p][k] = saved[k].detach().to(device=p.device, dtype=torch.float32).clone()
            if "step" in saved:
                self.state[p]["step"] = torch.as_tensor(saved["step"]).detach().to(device=p.device, dtype=torch.int64).reshape(1).clone()
        for st in self._dev.values():
            st["lr_host"] = None  # re-send the (possibly restored) learning rate

    def set_lr(self, lr: float, group: int = 0) -> None:
        """Write the learning rate of a group to the device (use this between CUDA-graph replays)."""
        self.param_groups[group]["lr"] = lr
        st = self._dev.get(group)
        if st is not None:
            st["lr"].fill_(float(lr))
            st["lr_host"] = float(lr)

    def step_count(self, group: int = 0) -> int:
        """Optimizer steps taken by a group so far (reads the device counter: synchronises)."""
        st = self._dev.get(group)
        return 0 if st is None or "step" not in st else int(st["step"].item())

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        capturing = torch.cuda.is_current_stream_capturing() if torch.cuda.is_available() else False
        mxn = self.max_grad_norm
        if mxn is not None and not math.isinf(mxn) and sum(1 for g in self.param_groups if any(p.grad is not None for p in g["params"])) > 1:
            # torch.nn.utils.clip_grad_norm_(model.parameters()) of the reference (scripts/train_contrast.py:456-463)
            # clips by ONE norm over all parameters; the fused kernel forms the norm per launch, i.e. per group
            raise _lib.P2TError("FusedAdamW clips by the norm of ONE parameter group: with max_grad_norm set, put all "
                                "clipped parameters (the adapter's fc1/fc2) in a single group")
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
            srcs = None
            if self.fp32_grad_sources:
                for p in tensors:
                    f = self.fp32_grad_sources.get(p)
                    if f is not None and (f.dtype != torch.float32 or f.numel() != p.numel() or not f.is_contiguous() or f.device != p.device):
                        raise _lib.P2TError("FusedAdamW: an fp32 gradient source must be a contiguous fp32 tensor of the parameter's size")
                srcs = (C.c_void_p * len(tensors))(*[(self.fp32_grad_sources[p].data_ptr() if p in self.fp32_grad_sources else None)
                                                     for p in tensors])
            b1, b2 = group["betas"]
            mx = self.max_grad_norm
            mx = 0.0 if (mx is None or math.isinf(mx)) else float(mx)
            _lib.call("p2t_adamw_step", len(tensors), st["params"], grads, srcs, st["m"], st["v"], st["w"], st["numel"],
                      st["partial"].data_ptr(), st["scal"].data_ptr(), st["lr"].data_ptr(), st["step"].data_ptr(),
                      float(b1), float(b2), float(group["eps"]), float(group["weight_decay"]), mx,
                      int(self.zero_grad_in_step), torch.cuda.current_stream().cuda_stream)
            self.grad_norm = st["scal"][0]
        return loss
