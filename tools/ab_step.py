#!/usr/bin/env python
"""A/B timing of step variants inside ONE process, interleaved, so that clock / power-cap drift hits every variant alike.

    python tools/ab_step.py [--workload cfg2_esm2_3b_llama8b] [--reps 6] [--steps 40] VARIANT [VARIANT ...]

A VARIANT is a comma-separated list of ENV=VALUE settings applied while that variant's CUDA graphs are captured (the
package reads its switches at call time), e.g.  base  P2T_TEXT_STREAM=0  P2T_FUSED_LOSS=0,P2T_TEXT_STREAM=0 .
Prints min / median ms per step of every variant over the repetitions.
"""
import argparse
import importlib
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg2_esm2_3b_llama8b")
    ap.add_argument("--reps", type=int, default=6)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--eval-mode", action="store_true")
    ap.add_argument("variants", nargs="+")
    args = ap.parse_args()
    import torch
    import __graft_entry__ as entry
    pkg = entry.load_package()
    synth = importlib.import_module("p2t_b200.synth")
    dev = torch.device("cuda:0")
    cfg = synth.CONFIGS[args.workload]
    batches = [synth.make_config_batch(args.workload, seed=1234 + 17 * i) for i in range(2)]
    acfg = pkg.ModalityAdapterConfig(input_dim=cfg["d_in"], intermediate_dim=cfg["d_mid"], output_dim=cfg["d_out"], dropout_rate=0.3)
    ad = pkg.ModalityAdapter(acfg).to(dev).to(torch.bfloat16)
    with torch.no_grad():
        b0 = batches[0]
        ad.fc1.weight.copy_(b0.w1); ad.fc1.bias.copy_(b0.b1); ad.fc2.weight.copy_(b0.w2); ad.fc2.bias.copy_(b0.b2)
    ad.eval() if args.eval_mode else ad.train()
    res = [dict(x=b.x.to(dev), pm=b.prot_mask.to(dev), text=b.text.to(dev), tm=b.text_mask.to(dev)) for b in batches]
    graphs = {}
    for v in args.variants:
        saved = {}
        if v != "base":
            for kv in v.split(","):
                k, val = kv.split("=")
                saved[k] = os.environ.get(k)
                os.environ[k] = val
        graphs[v] = [pkg.GraphedContrastiveStep(ad, r["x"], r["pm"], r["text"], r["tm"], seed=7) for r in res]
        for k, old in saved.items():
            if old is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = old
    times = {v: [] for v in args.variants}
    for rep in range(args.reps + 1):
        for v in args.variants:
            g = graphs[v]
            for i in range(5):
                g[i % 2].replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(args.steps):
                g[i % 2].replay()
            e1.record()
            torch.cuda.synchronize()
            if rep:
                times[v].append(e0.elapsed_time(e1) / args.steps)
    for v in args.variants:
        t = times[v]
        print(f"{v:50s} min {min(t):.4f} ms  median {statistics.median(t):.4f} ms  launches {graphs[v][0].launches_per_replay}  loss {graphs[v][0].loss.item():.5f}")


if __name__ == "__main__":
    main()
