"""Exchange steps of the sharded contrastive step over NVLink peer memory (csrc/peer.cu).

The sharded step has ONE data-path exchange: the all-gather of the unit-norm text embeddings that
form the global negatives; the training loop around it has a second one, the mean all-reduce of the adapter weight
gradients (DistributedDataParallel in the reference, scripts/train_contrast.py:611-614).  Both run here as plain
kernels that store into / load from the peers' buffers over NVLink, with arrival flags and epochs in device memory:
no NCCL call sits on the data path, nothing touches the host, and the whole sharded step is capturable in one CUDA
graph (`graph.GraphedContrastiveStep(..., exchange=...)`).  `torch.distributed` is used once, at construction, to
pass the 64-byte IPC handles around.

    gather = PeerAllGather(rows=B, cols=2 * H, dtype=torch.float32)          # once
    gather.push(t_local)                                                      # every step: returns immediately
    ... plan / pack kernels of the protein side overlap the transfer ...
    t_global = gather.arrive()                                                # (W*B, 2H), rank-major

    reducer = PeerGradAllReduce([p for p in adapter.parameters() if p.requires_grad])
    reducer.reduce_([p.grad for p in ...])                                    # grads := mean over ranks, in place
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import List, Optional, Sequence

import torch
import torch.distributed as dist

from . import _lib

_PUSH, _ARRIVE = 1, 2
_READY, _REDUCE, _WAIT = 1, 2, 4


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _rank_world(group) -> tuple:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


class _RawCuda:
    """Minimal __cuda_array_interface__ carrier: lets torch alias device memory it did not allocate."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2,
                                         "strides": None}


def raw_view(ptr: int, nbytes: int, dtype, shape, owner) -> torch.Tensor:
    """A torch tensor aliasing `nbytes` of device memory at `ptr` (a region of a peer buffer, which is cudaMalloc'ed by
    the library because it must be exportable over CUDA IPC); `owner` is kept alive by the tensor."""
    t = torch.as_tensor(_RawCuda(ptr, nbytes), device=torch.device("cuda", torch.cuda.current_device()))
    t = t.view(dtype).view(shape)
    t._p2t_owner = owner
    return t


class PeerBuffer:
    """One zero-initialised device buffer per rank, mapped into every rank of `group` (same node).

    `ptrs[r]` is this process's address of rank r's buffer; `table` is the same as a ctypes array for the C ABI.
    """

    def __init__(self, nbytes: int, group=None, _virtual: Optional[tuple] = None, _local: bool = False):
        if not torch.cuda.is_available():
            raise _lib.P2TError("PeerBuffer needs a CUDA device: this package has no CPU path")
        self.nbytes = int(nbytes)
        self.group = group
        self._own = None
        self._opened: List[int] = []
        if _virtual is not None:  # several ranks simulated in one process (tests): buffers handed in by the factory
            self.rank, self.world, self.ptrs = _virtual
            self._finish()
            return
        self.rank, self.world = (0, 1) if _local else _rank_world(group)
        torch.cuda.current_device()  # make sure the primary context exists before the runtime call below
        torch.zeros(1, device="cuda")
        own = C.c_void_p()
        handle = C.create_string_buffer(64)
        _lib.call("p2t_peer_alloc", self.nbytes, C.byref(own), handle)
        self._own = own.value
        self.ptrs = [None] * self.world
        self.ptrs[self.rank] = self._own
        if self.world > 1:
            handles = [None] * self.world
            dist.all_gather_object(handles, handle.raw, group=group)
            for r, h in enumerate(handles):
                if r == self.rank:
                    continue
                p = C.c_void_p()
                _lib.call("p2t_peer_open", bytes(h), C.byref(p))
                self.ptrs[r] = p.value
                self._opened.append(p.value)
            dist.barrier(group=group)
        self._finish()

    def _finish(self):
        self.table = (C.c_void_p * self.world)(*self.ptrs)
        self.base = self.ptrs[self.rank]

    @classmethod
    def virtual(cls, nbytes: int, world: int) -> List["PeerBuffer"]:
        """`world` buffers in THIS process, one PeerBuffer view per simulated rank (single-GPU tests of the
        exchange kernels: the launches of all simulated ranks are issued phase by phase on one stream)."""
        owners = [cls(nbytes, _local=True) for _ in range(world)]
        ptrs = [o.base for o in owners]
        views = [cls(nbytes, _virtual=(r, world, list(ptrs))) for r in range(world)]
        for v in views:
            v._keepalive = owners
        return views

    def status(self) -> int:
        """0, or 1 + the rank this rank timed out waiting for (synchronises the device)."""
        s = C.c_uint(0)
        _lib.call("p2t_peer_status", self.base, C.byref(s))
        return int(s.value)

    def check(self) -> None:
        s = self.status()
        if s:
            raise _lib.P2TError(f"rank {self.rank}: timed out waiting for the peer-memory flag of rank {s - 1}; the "
                                "round's results were poisoned with NaN.  Call reset() on every rank to resynchronise")

    def reset(self) -> None:
        """Collective: bring the channel's control block (epochs, counters, status, flags) back to zero on every rank
        after a timed-out round, so that pushes and arrivals pair up again."""
        torch.cuda.synchronize()
        if self.world > 1 and dist.is_initialized() and not hasattr(self, "_keepalive"):
            dist.barrier(group=self.group)
        _lib.call("p2t_peer_reset", self.base, _stream())
        torch.cuda.synchronize()
        if self.world > 1 and dist.is_initialized() and not hasattr(self, "_keepalive"):
            dist.barrier(group=self.group)

    def close(self) -> None:
        for p in self._opened:
            _lib.call("p2t_peer_close", p)
        self._opened = []
        if self.world > 1 and self._own is not None and dist.is_initialized():
            dist.barrier(group=self.group)
        if self._own is not None:
            _lib.call("p2t_peer_free", self._own)
            self._own = None


class PeerAllGather:
    """(rows, cols) per rank -> (world*rows, cols), rank-major, pushed over NVLink (SURVEY.md §8e exchange step)."""

    def __init__(self, rows: int, cols: int, dtype=torch.float32, group=None, _buffer: Optional[PeerBuffer] = None):
        self.rows, self.cols, self.dtype = int(rows), int(cols), dtype
        self.bytes_per_rank = self.rows * self.cols * torch.empty((), dtype=dtype).element_size()
        if self.bytes_per_rank % 16:
            raise ValueError("rows * cols * element size must be a multiple of 16 bytes")
        ctrl = int(_lib.load().p2t_peer_ctrl_bytes())
        rank, world = (_buffer.rank, _buffer.world) if _buffer is not None else _rank_world(group)
        self.buffer = _buffer if _buffer is not None else PeerBuffer(ctrl + 2 * world * self.bytes_per_rank, group)
        self.rank, self.world = rank, world
        self._pushed = False  # host-side pairing guard: every push is followed by exactly one arrival

    @staticmethod
    def buffer_bytes(rows: int, cols: int, dtype, world: int) -> int:
        return int(_lib.load().p2t_peer_ctrl_bytes()) + 2 * world * rows * cols * torch.empty((), dtype=dtype).element_size()

    def push(self, src: torch.Tensor) -> None:
        if src.shape != (self.rows, self.cols) or src.dtype != self.dtype or not src.is_cuda:
            raise ValueError(f"expected a CUDA {self.dtype} tensor of shape {(self.rows, self.cols)}, got {tuple(src.shape)} {src.dtype}")
        if self._pushed:
            raise _lib.P2TError("PeerAllGather.push: the previous round was never closed by arrive() — an exception "
                                "between push and arrive leaves the ranks out of step; call buffer.reset() on every rank")
        src = src.contiguous()
        _lib.call("p2t_peer_allgather", self.buffer.table, self.world, self.rank, src.data_ptr(), self.bytes_per_rank,
                  None, _PUSH, _stream())
        self._src_keepalive = src
        self._pushed = True

    def arrive(self, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        if not self._pushed:
            raise _lib.P2TError("PeerAllGather.arrive without a push in this round")
        if out is None:
            out = torch.empty(self.world * self.rows, self.cols, dtype=self.dtype, device="cuda")
        _lib.call("p2t_peer_allgather", self.buffer.table, self.world, self.rank, None, self.bytes_per_rank,
                  out.data_ptr(), _ARRIVE, _stream())
        self._pushed = False
        return out

    def note_arrived(self) -> None:
        """The round was closed by a kernel that waited for the flags itself (the fused loss kernel)."""
        if not self._pushed:
            raise _lib.P2TError("PeerAllGather: arrival without a push in this round")
        self._pushed = False

    def reset(self) -> None:
        self._pushed = False
        self.buffer.reset()

    def __call__(self, src: torch.Tensor) -> torch.Tensor:
        self.push(src)
        return self.arrive()

    def close(self) -> None:
        self.buffer.close()


class PeerGradAllReduce:
    """grads := mean over ranks, for a fixed list of tensors (the adapter's fc1/fc2 weights and biases).

    Two-shot over peer memory: rank k reduces slice k of the flattened gradients out of every peer's buffer in fp32,
    in rank order, and stores the mean into every peer's result area — all ranks end with identical bits.
    bfloat16 tensors travel and return as bf16; float32 tensors (the bias gradients of the fused step) travel and
    return in fp32, so they are rounded to bf16 once, after the mean.

    Zero-copy form: `views_in()` are tensors aliasing this rank's contribution area — let the backward kernels write
    the gradients there — then `exchange()` runs the three phases and `views_out()` alias the reduced result."""

    def __init__(self, like: Sequence[torch.Tensor], group=None, _buffer: Optional[PeerBuffer] = None):
        self.shapes = [tuple(t.shape) for t in like]
        self.dtypes = [t.dtype for t in like]
        for t in like:
            if t.dtype not in (torch.bfloat16, torch.float32):
                raise _lib.P2TError("PeerGradAllReduce reduces bfloat16 or float32 gradients")
        self.offsets = [0] * len(like)
        off = 0
        for want in (torch.bfloat16, torch.float32):  # all bf16 tensors first, then the fp32 ones
            if want == torch.float32:
                self.f32_from = off
            for i, t in enumerate(like):
                if t.dtype == want:
                    self.offsets[i] = off
                    off += (t.numel() * t.element_size() + 15) // 16 * 16
        self.n_bytes = off
        self.ctrl = int(_lib.load().p2t_peer_ctrl_bytes())
        rank, world = (_buffer.rank, _buffer.world) if _buffer is not None else _rank_world(group)
        self.buffer = _buffer if _buffer is not None else PeerBuffer(self.ctrl + 2 * self.n_bytes, group)
        self.rank, self.world = rank, world
        self._in = self._out = None

    @staticmethod
    def adapter_like(adapter) -> List[torch.Tensor]:
        """[dW1, db1, dW2, db2], all fp32 (meta tensors): the unrounded gradients of the fused step's backward."""
        w1, w2 = adapter.fc1.weight, adapter.fc2.weight
        f32 = torch.float32
        return [torch.empty(tuple(w1.shape), dtype=f32, device="meta"), torch.empty(w1.shape[0], dtype=f32, device="meta"),
                torch.empty(tuple(w2.shape), dtype=f32, device="meta"), torch.empty(w2.shape[0], dtype=f32, device="meta")]

    @classmethod
    def for_adapter(cls, adapter, group=None, _buffer: Optional[PeerBuffer] = None) -> "PeerGradAllReduce":
        """The reducer `graph.GraphedContrastiveStep(grad_reducer=...)` expects: all four gradients in fp32.  A rank's
        gradient can be several times larger than the mean over ranks (different batches pull in different
        directions), so rounding each rank's contribution to bf16 costs several times bf16's 2^-9 relative to the
        mean; in fp32 the only rounding is the final one, to the bf16 of param.grad."""
        return cls(cls.adapter_like(adapter), group, _buffer)

    @staticmethod
    def buffer_bytes(like: Sequence[torch.Tensor]) -> int:
        n = sum((t.numel() * t.element_size() + 15) // 16 * 16 for t in like)
        return int(_lib.load().p2t_peer_ctrl_bytes()) + 2 * n

    def _views(self, base_off: int):
        return [raw_view(self.buffer.base + base_off + off, math.prod(shape) * torch.empty((), dtype=dt).element_size(),
                         dt, shape, self.buffer)
                for shape, dt, off in zip(self.shapes, self.dtypes, self.offsets)]

    def views_in(self):
        if self._in is None:
            self._in = self._views(self.ctrl)
        return self._in

    def views_out(self):
        if self._out is None:
            self._out = self._views(self.ctrl + self.n_bytes)
        return self._out

    def _check(self, grads):
        if len(grads) != len(self.shapes):
            raise ValueError("gradient list does not match the tensors this reducer was built for")
        for g, s, dt in zip(grads, self.shapes, self.dtypes):
            if tuple(g.shape) != s or g.dtype != dt or not g.is_cuda or not g.is_contiguous():
                raise ValueError("gradients must be contiguous CUDA tensors of the registered shapes and dtypes")

    def _phase(self, bits: int) -> None:
        _lib.call("p2t_peer_allreduce_mean", self.buffer.table, self.world, self.rank, self.n_bytes, self.f32_from,
                  None, bits, _stream())

    def stage(self, grads: Sequence[torch.Tensor]) -> None:
        """Copy this rank's gradients into its channel buffer and announce them (phase 0)."""
        self._check(grads)
        st = _stream()
        for g, off in zip(grads, self.offsets):
            _lib.call("p2t_copy_d2d", self.buffer.base + self.ctrl + off, g.data_ptr(), g.numel() * g.element_size(), st)
        self._phase(_READY)

    def reduce(self) -> None:
        self._phase(_REDUCE)

    def finish(self, grads: Sequence[torch.Tensor]) -> None:
        """Wait for every slice of the mean and copy it over `grads` (phase 2)."""
        st = _stream()
        self._phase(_WAIT)
        for g, off in zip(grads, self.offsets):
            _lib.call("p2t_copy_d2d", g.data_ptr(), self.buffer.base + self.ctrl + self.n_bytes + off,
                      g.numel() * g.element_size(), st)

    def exchange(self) -> None:
        """Zero-copy form: the contribution is already in `views_in()`; after this call `views_out()` hold the mean."""
        self._phase(_READY | _REDUCE | _WAIT)

    def wait_only(self) -> None:
        """Close a round whose announce + reduce phases were serviced elsewhere (inside a GEMM launch)."""
        self._phase(_WAIT)

    def reduce_(self, grads: Sequence[torch.Tensor]) -> Sequence[torch.Tensor]:
        self.stage(grads)
        self.reduce()
        self.finish(grads)
        return grads

    def close(self) -> None:
        self._in = self._out = None
        self.buffer.close()


class OverlappedGradReduce:
    """The gradient mean of the sharded training step as TWO channels, so that most of it hides behind the backward:

      late : [dW2, db2] (fp32) — complete after the dW2 GEMM; its announce + reduce phases are serviced INSIDE the
             dW1 GEMM's launch by the epilogue warps of every CTA while their first accumulator is being computed
             (one kernel: tcgen05 GEMM + NVLink peer-memory reduce, no SM taken from the GEMM), DDP's bucket overlap
             (scripts/train_contrast.py:448 + :611-614) without a second kernel fighting the persistent GEMM for SMs;
      tail : [dW1, db1] (fp32) — reduced after the backward.

    `views_in()` / `views_out()` are in parameter order [dW1, db1, dW2, db2]; pass the object as
    `graph.GraphedContrastiveStep(grad_reducer=...)`."""

    def __init__(self, adapter, group=None, _buffers: Optional[tuple] = None):
        like = PeerGradAllReduce.adapter_like(adapter)
        lb, tb = _buffers if _buffers is not None else (None, None)
        self.late = PeerGradAllReduce([like[2], like[3]], group, _buffer=lb)
        self.tail = PeerGradAllReduce([like[0], like[1]], group, _buffer=tb)
        self.rank, self.world = self.late.rank, self.late.world
        self.shapes = [tuple(t.shape) for t in like]
        self.dtypes = [t.dtype for t in like]

    def views_in(self):
        (w2, b2), (w1, b1) = self.late.views_in(), self.tail.views_in()
        return [w1, b1, w2, b2]

    def views_out(self):
        (w2, b2), (w1, b1) = self.late.views_out(), self.tail.views_out()
        return [w1, b1, w2, b2]

    def finish(self) -> None:
        """After the backward: close the overlapped round, then reduce the rest."""
        self.late.wait_only()
        self.tail.exchange()

    class _Both:
        def __init__(self, a, b):
            self.a, self.b = a, b

        def check(self):
            self.a.check()
            self.b.check()

    @property
    def buffer(self):
        return OverlappedGradReduce._Both(self.late.buffer, self.tail.buffer)

    def close(self) -> None:
        self.late.close()
        self.tail.close()
