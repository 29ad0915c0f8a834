"""Host -> device staging of a padded batch as packed rows (the ragged hand-over format).

The reference moves whole padded tensors with `.to(rank)` (scripts/train_contrast.py:329-330).  When the
batch starts in host memory (data loader, CPU-side trunk, benchmark harness) only the VALID rows need to
cross PCIe: `HostStager` copies each sequence's valid range straight into a packed device buffer on a
dedicated copy stream, so the copy of batch i+1 overlaps the kernels of batch i, and the device-side
pack (gather) kernel disappears.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import Optional

import torch

from . import _lib


@dataclass
class StagedBatch:
    """Packed device tensors of one batch; feed to `contrastive_step(..., residue_lengths=, text_lengths=)`."""
    residue_rows: torch.Tensor     # (sum L_b, D_in) bf16
    residue_lengths: torch.Tensor  # (B,) int32, device
    text_rows: torch.Tensor        # (sum T_b, H) bf16
    text_lengths: torch.Tensor     # (B,) int32, device
    h2d_bytes: int
    _event: Optional[torch.cuda.Event] = None


def bind_host_thread_to_gpu(device) -> Optional[list]:
    """Pin the calling host thread (and the threads it spawns) to the CPUs of the NUMA node the GPU hangs off, so that
    pinned staging buffers allocated afterwards are NUMA-local to the GPU's PCIe root (first-touch placement) and the
    thread that issues the copies runs next to them.  On a two-socket 8-GPU box the default — every rank's buffers on
    node 0 — halves the host->device rate of the GPUs on the other socket.  Returns the CPU list, or None when the
    topology cannot be read (then nothing is changed)."""
    import os
    try:
        idx = torch.device(device).index
        idx = torch.cuda.current_device() if idx is None else idx
        pr = torch.cuda.get_device_properties(idx)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/local_cpulist") as fh:
            spec = fh.read().strip()
        cpus = []
        for part in spec.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.extend(range(int(a), int(b) + 1))
            elif part:
                cpus.append(int(part))
        allowed = sorted(set(cpus) & os.sched_getaffinity(0))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return allowed
    except Exception:  # noqa: BLE001 - topology files missing (containers): leave the affinity alone
        return None


def _valid_ranges(mask: torch.Tensor):
    """(starts, counts) as int32 CPU tensors if every row's valid positions are contiguous, else None."""
    m = mask != 0
    counts = m.sum(dim=1)
    starts = torch.argmax(m.to(torch.uint8), dim=1)
    L = mask.shape[1]
    ar = torch.arange(L)[None, :]
    expect = (ar >= starts[:, None]) & (ar < (starts + counts)[:, None])
    if not torch.equal(expect, m):
        return None
    return starts.to(torch.int32).contiguous(), counts.to(torch.int32).contiguous()


PULL_PIECE_BYTES = 256 * 8 * 16  # one sweep of the staging kernel's CTA (csrc/rows.cu: kPullPieceBytes)


def pull_table(starts: torch.Tensor, counts: torch.Tensor, L: int, row_bytes: int):
    """Segment table of `p2t_stage_rows_pull` for a padded host batch [B][L][row_bytes]: int64 [n_seg, 3] =
    (byte offset in the host batch, byte offset in the packed destination, byte length) for every sequence with at
    least one valid row, and int32 [n_seg + 1] = number of 32 KB pieces before each segment (last entry: total)."""
    B = counts.shape[0]
    seg_bytes = counts.to(torch.int64) * row_bytes
    table = torch.stack([(torch.arange(B, dtype=torch.int64) * L + starts.to(torch.int64)) * row_bytes,
                         torch.cumsum(seg_bytes, 0) - seg_bytes, seg_bytes], dim=1)[seg_bytes > 0].contiguous()
    pieces = (table[:, 2] + PULL_PIECE_BYTES - 1) // PULL_PIECE_BYTES
    prefix = torch.zeros(table.shape[0] + 1, dtype=torch.int32)
    prefix[1:] = torch.cumsum(pieces, 0).to(torch.int32)
    return table, prefix


class HostStager:
    """Stages padded HOST batches (ideally pinned) as packed device rows on a private copy stream.

        stager = HostStager(device)
        stager.submit(x, mask, text, text_mask)        # asynchronous
        batch = stager.take()                          # current stream now waits for the copies
        loss = contrastive_step(batch.residue_rows, None, adapter, batch.text_rows, None,
                                residue_lengths=batch.residue_lengths, text_lengths=batch.text_lengths)
    """

    def __init__(self, device, residue_streams: int = 3, mode: Optional[str] = None, pull_ctas: int = 32):
        """`mode`: "copy" = one copy-engine transfer per sequence (default); "pull" = ONE kernel per side that reads the
        valid rows out of the pinned host batch itself (no per-copy set-up; its `pull_ctas` CTAs share the SMs with
        the step's kernels).  Env P2T_STAGE_MODE overrides the default."""
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.P2TError("HostStager needs a CUDA device: this package has no CPU path")
        self.mode = mode or os.environ.get("P2T_STAGE_MODE", "copy")
        if self.mode not in ("copy", "pull"):
            raise _lib.P2TError(f"HostStager mode must be 'copy' or 'pull' (got {self.mode!r})")
        self.pull_ctas = int(os.environ.get("P2T_STAGE_PULL_CTAS", pull_ctas))
        # several copy streams: the residue rows of a batch are ~32-64 separate 1-5 MB copies (one per sequence); on ONE
        # stream every copy pays its DMA set-up behind the previous one (~5 us x 64 = 0.3 ms of a 2.6 ms step).  The
        # sequences are dealt round-robin to `residue_streams` streams so that set-up and transfer of neighbouring
        # copies overlap; the text rows travel on their own stream.
        self.nres = max(1, int(os.environ.get("P2T_STAGE_STREAMS", residue_streams)))
        # pull mode: high-priority streams, so that the staging kernels' few CTAs are placed as soon as an SM has room
        prio = -1 if self.mode == "pull" else 0
        self.streams = [torch.cuda.Stream(device=self.device, priority=prio) for _ in range(self.nres)]
        self.stream = self.streams[0]
        self.stream2 = torch.cuda.Stream(device=self.device, priority=prio)
        self._pending = []

    def _stage(self, src: torch.Tensor, mask: torch.Tensor, streams):
        if src.is_cuda or mask.is_cuda:
            raise _lib.P2TError("HostStager takes host tensors")
        if src.dtype != torch.bfloat16:
            raise _lib.P2TError(f"host states must be bfloat16 (got {src.dtype})")
        src = src.contiguous()
        B, L, D = src.shape
        ranges = _valid_ranges(mask)
        if ranges is None:
            raise _lib.P2TError("HostStager needs contiguous valid ranges (right or left padding); move masks with "
                                "holes to the device and use the mask form of contrastive_step")
        starts, counts = ranges
        total = int(counts.sum())
        lead = streams[0]
        with torch.cuda.stream(lead):
            rows = torch.empty(max(total, 1), D, dtype=torch.bfloat16, device=self.device)
            lens = counts.pin_memory().to(self.device, non_blocking=True)
        ready = torch.cuda.Event()
        ready.record(lead)  # the destination buffer exists (its allocation is stream-ordered on `lead`)
        for st in streams[1:]:
            st.wait_event(ready)
        if self.mode == "pull":
            if not src.is_pinned():
                raise _lib.P2TError("HostStager(mode='pull') reads the host batch from the device: it must be pinned")
            if (D * 2) % 16:
                raise _lib.P2TError("HostStager(mode='pull') needs rows of a multiple of 16 bytes")
            table, prefix = pull_table(starts, counts, L, D * 2)
            with torch.cuda.stream(lead):
                table_d = table.pin_memory().to(self.device, non_blocking=True)
                prefix_d = prefix.pin_memory().to(self.device, non_blocking=True)
                _lib.call("p2t_stage_rows_pull", src.data_ptr(), table_d.data_ptr(), prefix_d.data_ptr(), int(table.shape[0]),
                          rows.data_ptr(), self.pull_ctas, lead.cuda_stream)
            self._keep = getattr(self, "_keep", [])[-8:] + [(table_d, prefix_d, src)]  # alive until the kernel has run
            return rows[:total] if total else rows[:0], lens, total * D * 2 + B * 4
        handles = (C.c_void_p * len(streams))(*[st.cuda_stream for st in streams])
        _lib.call("p2t_stage_rows_h2d", src.data_ptr(), L * D * 2, D * 2, starts.data_ptr(), counts.data_ptr(), B,
                  rows.data_ptr(), handles, len(streams))
        for st in streams[1:]:
            ev = torch.cuda.Event()
            ev.record(st)
            lead.wait_event(ev)
        return rows[:total] if total else rows[:0], lens, total * D * 2 + B * 4

    def submit(self, x: torch.Tensor, mask: torch.Tensor, text: torch.Tensor, text_mask: torch.Tensor) -> None:
        tr, tl, b1 = self._stage(text, text_mask, [self.stream2])
        ev2 = torch.cuda.Event()
        ev2.record(self.stream2)
        xr, xl, b0 = self._stage(x, mask, self.streams)
        self.stream.wait_event(ev2)
        ev = torch.cuda.Event()
        ev.record(self.stream)
        self._pending.append(StagedBatch(xr, xl, tr, tl, b0 + b1, ev))

    def take(self) -> StagedBatch:
        if not self._pending:
            raise _lib.P2TError("HostStager.take() without a submitted batch")
        batch = self._pending.pop(0)
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(batch._event)
        for t in (batch.residue_rows, batch.residue_lengths, batch.text_rows, batch.text_lengths):
            t.record_stream(cur)  # allocated on the copy stream, consumed on the compute stream
        return batch
