#!/usr/bin/env python
"""Where the end-to-end path's host->device rate goes: the HostStager's per-sequence copies alone (no compute), the same
bytes as ONE copy per side from a packed pinned slab, and the 256 MiB probe bench.py quotes as the ceiling."""
import importlib
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import __graft_entry__ as entry
    pkg = entry.load_package()
    synth = importlib.import_module("p2t_b200.synth")
    dev = torch.device("cuda:0")
    batches = [synth.make_config_batch("cfg2_esm2_3b_llama8b", seed=1234 + 17 * i) for i in range(2)]
    host = [dict(x=b.x.pin_memory(), pm=b.prot_mask.pin_memory(), text=b.text.pin_memory(), tm=b.text_mask.pin_memory()) for b in batches]
    stager = pkg.HostStager(dev)
    steps = 30

    def run(fn):
        for _ in range(3):
            fn(0)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        nbytes = 0
        for i in range(steps):
            nbytes += fn(i)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        return nbytes / dt / 1e9, dt / steps * 1e3

    def staged(i):
        h = host[i % 2]
        stager.submit(h["x"], h["pm"], h["text"], h["tm"])
        b = stager.take()
        return b.h2d_bytes

    print("HostStager copies alone (one per sequence): %.2f GB/s, %.3f ms per batch" % run(staged))
    for ctas in (8, 16, 32, 64):
        stager = pkg.HostStager(dev, mode="pull", pull_ctas=ctas)
        print("HostStager pull kernels alone, %3d CTAs:     %.2f GB/s, %.3f ms per batch" % ((ctas,) + run(staged)))
    # the same valid rows packed on the host beforehand: one copy per side
    packed = []
    for b in batches:
        xr = torch.cat([b.x[j, :int(n)] for j, n in enumerate(b.prot_lens)]).pin_memory()
        tr = torch.cat([b.text[j, :int(n)] for j, n in enumerate(b.text_mask.sum(1))]).pin_memory()
        packed.append((xr, tr))
    dx = [torch.empty_like(p[0], device=dev) for p in packed]
    dt_ = [torch.empty_like(p[1], device=dev) for p in packed]

    def one_copy(i):
        xr, tr = packed[i % 2]
        dx[i % 2].copy_(xr, non_blocking=True)
        dt_[i % 2].copy_(tr, non_blocking=True)
        return xr.numel() * 2 + tr.numel() * 2

    print("packed slabs, one copy per side:            %.2f GB/s, %.3f ms per batch" % run(one_copy))
    ph = torch.empty(256 << 20, dtype=torch.uint8).pin_memory()
    pd = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def probe(i):
        pd.copy_(ph, non_blocking=True)
        return 256 << 20

    print("256 MiB probe:                              %.2f GB/s, %.3f ms per copy" % run(probe))


if __name__ == "__main__":
    main()
