"""Bring-up probe for the tcgen05 GEMM: one variant per process (so a hang or fault is contained).

    python tools/gemm_probe.py <cta_group> <a_mn> <b_mn> <m> <n> <k> [f32]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry


def main():
    cg, a_mn, b_mn, m, n, k = (int(v) for v in sys.argv[1:7])
    f32 = len(sys.argv) > 7 and "f32" in sys.argv[7:]
    pkg = entry.load_package()
    core = sys.modules["p2t_b200._core"]
    torch.manual_seed(0)
    dev = torch.device("cuda:0")
    A = torch.randn(m, k, device=dev).to(torch.bfloat16)
    B = torch.randn(n, k, device=dev).to(torch.bfloat16)
    a_store = A.t().contiguous() if a_mn else A
    b_store = B.t().contiguous() if b_mn else B
    out = core.gemm(a_store, b_store, m, n, k, a_mn=bool(a_mn), b_mn=bool(b_mn),
                    out_dtype=torch.float32 if f32 else torch.bfloat16, alpha=0.5, cta_group=cg)
    torch.cuda.synchronize()
    ref = 0.5 * (A.float() @ B.float().t())
    err = (out.float() - ref).abs().max().item()
    scale = ref.abs().max().item()
    tol = (2e-3 if f32 else 1.2e-2) * scale
    # timing
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        core.gemm(a_store, b_store, m, n, k, a_mn=bool(a_mn), b_mn=bool(b_mn), out_dtype=torch.float32 if f32 else torch.bfloat16, cta_group=cg)
    ev0.record()
    iters = 10
    for _ in range(iters):
        core.gemm(a_store, b_store, m, n, k, a_mn=bool(a_mn), b_mn=bool(b_mn), out_dtype=torch.float32 if f32 else torch.bfloat16, cta_group=cg)
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / iters
    tf = 2.0 * m * n * k / ms / 1e9
    status = "OK" if err <= tol else "MISMATCH"
    print(f"{status} cg={cg} a_mn={a_mn} b_mn={b_mn} m={m} n={n} k={k} f32={int(f32)} max_err={err:.4g} ref_max={scale:.4g} "
          f"{ms*1000:.1f}us {tf:.1f} TFLOP/s", flush=True)
    if err > tol:
        bad = ((out.float() - ref).abs() > tol).nonzero()
        print("  first bad:", bad[:5].tolist(), "count", bad.shape[0], "of", m * n, flush=True)
        sys.exit(1)


if __name__ == "__main__":
    main()
