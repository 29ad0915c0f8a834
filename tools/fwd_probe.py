"""Times the two forward adapter GEMMs (fc1, fc2 with fused epilogues) in isolation at a config's shape,
with / without the derivative outputs and dropout, using the library's per-launch GEMM timing.

    python tools/fwd_probe.py [config-name]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "cfg2_esm2_3b_llama8b"
    pkg = entry.load_package()
    core = sys.modules["p2t_b200._core"]
    synth = __import__("importlib").import_module("p2t_b200.synth")
    lib = pkg._lib
    sb = synth.make_config_batch(name)
    dev = torch.device("cuda:0")
    B, L, d_in = sb.x.shape
    x, pm = sb.x.to(dev), sb.prot_mask.to(dev)
    w1, b1, w2, b2 = (t.to(dev) for t in (sb.w1, sb.b1, sb.w2, sb.b2))
    plan = core.plan_rows(pm)
    xp = core.gather_rows(x.view(B * L, d_in), plan)
    for need_grad in (True, False):
        for p in (0.3, 0.0):
            for _ in range(3):
                core.adapter_forward(xp, plan.rows_cap, plan.rows_cap, plan.n_rows, w1, b1, w2, b2, p, 1, need_grad)
            torch.cuda.synchronize()
            lib.gemm_timing_enable(True)
            for _ in range(10):
                core.adapter_forward(xp, plan.rows_cap, plan.rows_cap, plan.n_rows, w1, b1, w2, b2, p, 1, need_grad)
            torch.cuda.synchronize()
            ms, n, each = lib.gemm_timing_collect()
            lib.gemm_timing_enable(False)
            fc1 = 1e3 * sum(each[0::2]) / 10
            fc2 = 1e3 * sum(each[1::2]) / 10
            print(f"{name}: derivative outputs={need_grad} dropout={p}: fc1 {fc1:.1f} us  fc2 {fc2:.1f} us", flush=True)


if __name__ == "__main__":
    main()
