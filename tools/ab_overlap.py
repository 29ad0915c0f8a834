#!/usr/bin/env python
"""A/B timing of the sharded TRAINING step's gradient-mean forms on N GPUs, interleaved inside one job.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/ab_overlap.py [--workload cfg2_esm2_3b_llama8b] [--reps 8] [--steps 20]

Forms: the step alone; step + optimizer with the plain peer reducer (one channel after the backward); step + optimizer
with the fused form (dW2 / db2 mean by the idle epilogue warps inside the dW1 GEMM launch).  Every form is one CUDA graph
per step; times are device times, max over ranks, each taken right after a pass of the bare step (paired difference).
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
    ap.add_argument("--reps", type=int, default=8)
    ap.add_argument("--steps", type=int, default=20)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import __graft_entry__ as entry
    pkg = entry.load_package()
    synth = importlib.import_module("p2t_b200.synth")
    pdist = importlib.import_module("p2t_b200.dist")
    peer = importlib.import_module("p2t_b200.peer")
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = synth.CONFIGS[args.workload]
    b = synth.make_config_batch(args.workload, seed=1234, rank=rank, same_lengths_as_rank0=True)
    acfg = pkg.ModalityAdapterConfig(input_dim=cfg["d_in"], intermediate_dim=cfg["d_mid"], output_dim=cfg["d_out"], dropout_rate=0.3)
    ad = pkg.ModalityAdapter(acfg).to(dev).to(torch.bfloat16)
    with torch.no_grad():
        ad.fc1.weight.copy_(b.w1); ad.fc1.bias.copy_(b.b1); ad.fc2.weight.copy_(b.w2); ad.fc2.bias.copy_(b.b2)
    ad.train()
    params = [ad.fc1.weight, ad.fc1.bias, ad.fc2.weight, ad.fc2.bias]
    r = dict(x=b.x.to(dev), pm=b.prot_mask.to(dev), text=b.text.to(dev), tm=b.text_mask.to(dev))
    rows_bound = int(b.prot_lens.sum()) + 256
    exchange = pdist.ShardedExchange(cfg["batch"], 2 * cfg["d_out"]) if world > 1 else None
    opt = pkg.FusedAdamW(params, lr=1e-6, eps=1e-6, betas=(0.9, 0.999), max_grad_norm=1.0)

    def graph(reducer, optimizer):
        return pkg.GraphedContrastiveStep(ad, r["x"], r["pm"], r["text"], r["tm"], seed=7, exchange=exchange,
                                          grad_reducer=reducer, optimizer=optimizer, max_valid_rows=rows_bound)

    forms, reducers = {}, []
    forms["step only"] = graph(None, None)
    red = peer.PeerGradAllReduce.for_adapter(ad)
    reducers.append(red)
    forms["plain reducer, no optimizer"] = graph(red, None)
    forms["plain reducer + optimizer"] = graph(red, opt)
    fr = peer.OverlappedGradReduce(ad)
    reducers.append(fr)
    forms["fused GEMM + reduce, no optimizer"] = graph(fr, None)
    forms["fused GEMM + reduce + optimizer"] = graph(fr, opt)
    # per repetition every form is timed right after its own pass of the bare step, and the DIFFERENCE is what is
    # kept: power-cap drift between repetitions (several percent on a loaded box) then cancels to first order
    def timed(g):
        for _ in range(3):
            g.replay()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            g.replay()
        e1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / args.steps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    times = {k: [] for k in forms}
    extra = {k: [] for k in forms}
    for rep in range(args.reps + 1):
        for k, g in forms.items():
            base = timed(forms["step only"])
            t = timed(g)
            if rep:
                times[k].append(t)
                extra[k].append(t - base)
    for red in reducers:
        red.buffer.check()
    if rank == 0:
        for k, t in times.items():
            print(f"world {world} {k:42s} median {statistics.median(t):.4f} ms  over the bare step: median +{statistics.median(extra[k]):.4f}"
                  f"  min +{min(extra[k]):.4f}  max +{max(extra[k]):.4f} ms  launches {forms[k].launches_per_replay}", flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
