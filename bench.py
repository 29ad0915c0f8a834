#!/usr/bin/env python
"""Benchmark of the Stage-1 contrastive step (BASELINE.json metric: pairs/sec + roofline fraction).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

One "step" = one fused contrastive step, forward AND backward, over one synthetic batch of the
BASELINE config: residue states -> ModalityAdapter (fc1/GELU/dropout/fc2/GELU/dropout/L2-norm) ->
'mix' readout -> L2 normalise; text hidden states -> 'mix' readout -> normalise; similarity / tau;
InfoNCE; gradients of fc1/fc2 weights and biases.  Training mode (dropout p = 0.3) as in the
reference's train_epoch.  N > 1 (under torchrun): every rank holds B pairs (weak scaling: each rank gets rank 0's
multiset of sequence lengths with its own data); the unit-norm text embeddings are all-gathered by peer-memory
kernels over NVLink (csrc/peer.cu) to form global negatives, inside the same CUDA graph as the rest of the step.

Prints ONE JSON line (rank 0).  `value` = pairs/s with inputs resident in HBM; `e2e` = the same
through the public API with inputs in pinned host memory (H2D of the step's inputs and D2H of the
loss inside the timed region); `roofline` = the tcgen05 GEMM kernel against the measured bf16
peak; `cpu_baseline` = the CPU oracle timed on this box's host cores on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "contrastive_step_pairs_per_sec"
UNIT = "pairs/s"
DEFAULT_WORKLOAD = "cfg2_esm2_3b_llama8b"
CPU_SAMPLE_PAIRS = 32      # per step of the CPU arms: one whole config-2 step (the same 32 pairs the GPU arm steps)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-optimizer", action="store_true", help="skip the (separately reported) optimizer leg")
    ap.add_argument("--no-exchange-probe", action="store_true", help="skip timing the exchange kernels alone (N > 1)")
    ap.add_argument("--pairs-per-gpu", type=int, default=0,
                    help="override the workload's pairs per GPU (config 5: global batch 4096 -> 2048 / 1024 / 512 at 2 / 4 / 8 GPUs)")
    ap.add_argument("--no-stages", action="store_true", help="skip the per-kernel event-stamp pass")
    ap.add_argument("--no-imbalance-leg", action="store_true",
                    help="N > 1: skip the second weak-scaling leg with independent per-rank sequence lengths")
    ap.add_argument("--stages", action="store_true",
                    help="extra eager pass with a CUDA event after EVERY library launch: per-kernel share of the step and "
                         "achieved fraction of the HBM / tensor roofline (adds a `stages` list to the JSON line)")
    ap.add_argument("--eval-mode", action="store_true", help="dropout off (parity runs); default is training mode")
    ap.add_argument("--no-graph", action="store_true", help="time eager launches instead of CUDA-graph replays")
    ap.add_argument("--device-synth", action="store_true",
                    help="draw the synthetic residue/text states on the GPU (big sweep configs: no multi-GB host "
                         "buffers); implies --no-e2e and --no-cpu-baseline, which need host copies")
    return ap.parse_args()


# --------------------------------------------------------------------------------------------------
# CPU oracle timing (cpu_baseline leg and --impl reference)
# --------------------------------------------------------------------------------------------------
def cpu_step_time(sb, pairs: int, iters: int, warmup: int):
    """Time one CPU step (fwd + bwd, fp32, all host threads, dropout off) on the first `pairs` pairs of the batch.

    kind "reference": the reference's OWN ModalityAdapter, readout_embeddings(..., "mix"), F.normalize and
    SegmentedBatchInfoNCELoss (2 segments, its default) composed as teacher_forcing_forward_pass composes them
    (scripts/train_contrast.py:345-379, trunk outputs given), imported from /root/reference or from the archive
    oracle/build_ref.py staged under oracle/_ref/.  kind "port": oracle/restatement.py (same ATen CPU kernels) when
    neither is present.  Returns (times, threads, kind)."""
    import torch
    from oracle import reference_loader as RL
    torch.set_num_threads(os.cpu_count() or 1)
    f = torch.float32
    n = pairs
    lmax = int(sb.prot_lens[:n].max())
    x = sb.x[:n, :lmax].to(f)
    pm = sb.prot_mask[:n, :lmax]
    text, tm = sb.text[:n].to(f), sb.text_mask[:n]
    times = []
    if RL.reference_available():
        ref = RL.load_reference()
        d_mid, d_in = sb.w1.shape
        adapter = ref.ModalityAdapter(ref.ModalityAdapterConfig(input_dim=d_in, intermediate_dim=d_mid,
                                                               output_dim=sb.w2.shape[0], dropout_rate=0.3)).to(f).eval()
        with torch.no_grad():
            adapter.fc1.weight.copy_(sb.w1); adapter.fc1.bias.copy_(sb.b1)
            adapter.fc2.weight.copy_(sb.w2); adapter.fc2.bias.copy_(sb.b2)
        loss_fn = ref.SegmentedBatchInfoNCELoss()
        nseg = 2 if n % 2 == 0 else 1
        seg = n // nseg
        for it in range(warmup + iters):
            adapter.zero_grad(set_to_none=True)
            t0 = time.perf_counter()
            with torch.no_grad():
                t_emb = torch.nn.functional.normalize(ref.readout_embeddings(text, tm, "mix"), p=2, dim=-1)
            p_emb = torch.nn.functional.normalize(ref.readout_embeddings(adapter(x), pm, "mix"), p=2, dim=-1)
            loss = 0.0
            for s_id in range(nseg):
                labels = torch.arange(s_id * seg, (s_id + 1) * seg)
                loss = loss + loss_fn(p_emb[s_id * seg:(s_id + 1) * seg], t_emb, labels)
            (loss / nseg).backward()
            dt = time.perf_counter() - t0
            if it >= warmup:
                times.append(dt)
        return times, torch.get_num_threads(), "reference"
    from oracle import restatement as R
    params = [t.to(f).requires_grad_() for t in (sb.w1, sb.b1, sb.w2, sb.b2)]
    for it in range(warmup + iters):
        for p in params:
            p.grad = None
        t0 = time.perf_counter()
        st = R.step_forward(x, pm, params[0], params[1], params[2], params[3], text, tm, 0.05, 1)
        st.loss.backward()
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    return times, torch.get_num_threads(), "port"


def run_reference_arm(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path on the host cores (rank 0 only): its own
    files when /root/reference or the staged archive oracle/_ref/ref_hotpath.tgz is present (kind "reference"), else
    the oracle port (same ATen CPU kernels, kind "port").  Each step is one whole 32-pair config-2 step."""
    if rank != 0:
        return
    import __graft_entry__ as entry
    entry.load_package()
    import importlib
    synth = importlib.import_module("p2t_b200.synth")
    sb = synth.make_config_batch(args.workload)
    cfg = synth.CONFIGS[args.workload]
    n_cpu = min(cfg["batch"], CPU_SAMPLE_PAIRS)
    times, cores, kind = cpu_step_time(sb, n_cpu, args.steps, args.warmup)
    total = sum(times)
    value = n_cpu * len(times) / total
    sample = (f"{'the whole' if n_cpu == cfg['batch'] else 'the first ' + str(n_cpu) + ' pairs of the'} {args.workload} batch "
              f"({n_cpu} pairs) per step, fp32, fwd+bwd, eval-mode dropout, "
              + ("the reference's own ModalityAdapter / readout_embeddings / SegmentedBatchInfoNCELoss (2 segments)"
                 if kind == "reference" else "CPU restatement (oracle/restatement.py)"))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * total / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "d_in": cfg["d_in"], "d_mid": cfg["d_mid"], "d_out": cfg["d_out"],
                   "pairs_per_gpu": cfg["batch"], "pairs_per_step": n_cpu},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock, power and clock-event reasons of one GPU DURING the timed region: an NVML polling thread
    (every ~10 ms, cheap enough not to disturb the launching thread; `nvidia-smi -lms` is too coarse for a ~0.1 s region), with
    `nvidia-smi` as the fallback when NVML cannot be loaded."""
    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu_index = gpu_index
        self.proc = None
        self.thread = None
        self.stop_flag = False
        self.sm, self.power, self.reasons = [], [], set()
        self.sm_max = None

    def _physical_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if self.gpu_index < len(ids) and ids[self.gpu_index].isdigit():
                return int(ids[self.gpu_index])
        return self.gpu_index

    def _poll(self, nv, handle):
        names = {nv.nvmlClocksEventReasonHwSlowdown: "hw_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksEventReasonSwThermalSlowdown: "sw_thermal_slowdown", nv.nvmlClocksEventReasonSwPowerCap: "sw_power_cap",
                 nv.nvmlClocksEventReasonHwPowerBrakeSlowdown: "hw_power_brake_slowdown"}
        while not self.stop_flag:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(handle, nv.NVML_CLOCK_SM)))
                self.power.append(nv.nvmlDeviceGetPowerUsage(handle) / 1e3)
                mask = nv.nvmlDeviceGetCurrentClocksEventReasons(handle)
                for bit, nm in names.items():
                    if mask & bit:
                        self.reasons.add(nm)
            except Exception:  # noqa: BLE001 - sampling must never break the benchmark
                pass
            time.sleep(0.01)

    def start(self):
        try:
            import threading
            import pynvml as nv
            nv.nvmlInit()
            handle = nv.nvmlDeviceGetHandleByIndex(self._physical_index())
            self.sm_max = float(nv.nvmlDeviceGetMaxClockInfo(handle, nv.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._poll, args=(nv, handle), daemon=True)
            self.thread.start()
            return
        except Exception:  # noqa: BLE001
            self.thread = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu_index), f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                 "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if self.thread is not None:
            self.stop_flag = True
            self.thread.join(timeout=2)
            return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": self.sm_max,
                    "reasons": sorted(self.reasons), "samples": len(self.sm), "source": "nvml",
                    "power_w_max": max(self.power) if self.power else None}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in out.strip().splitlines():
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); smax.append(float(parts[1]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi"}


def gemm_source_digest() -> str:
    """sha256 (16 hex digits) of the GEMM kernel's sources: ties profiles/roofline_traffic.json to the code it measured."""
    import hashlib
    h = hashlib.sha256()
    for f in ("gemm_sm100.cuh", "gemm_host.cu", "ptx.cuh", "mathfn.cuh"):
        with open(os.path.join(ROOT, "prot2text-v2-esm3_b200", "csrc", f), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


# --------------------------------------------------------------------------------------------------
def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun the way the driver does
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29511"), os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import __graft_entry__ as entry
    pkg = entry.load_package()
    import importlib
    synth = importlib.import_module("p2t_b200.synth")
    pdist = importlib.import_module("p2t_b200.dist")
    lib = pkg._lib
    lib.load()  # fail loudly if the CUDA extension is missing

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    cfg = dict(synth.CONFIGS[args.workload])
    if args.pairs_per_gpu > 0:
        cfg["batch"] = args.pairs_per_gpu
    nbatches = 2
    # saved activations of one step (packed rows x (x, h1, g1, a, g2, dz1, dz2) + padded inputs): above ~25 GB keep ONE
    # resident batch (its inputs alone exceed the L2 a hundred times) and time eager launches — a captured graph would
    # pin a second copy of every intermediate in its private pool, and at config 5 / 2 GPUs that no longer fits 180 GB
    est_rows = cfg["batch"] * (cfg["lmin"] + cfg["lmax"]) / 2
    est_bytes = est_rows * 2 * (2 * cfg["d_in"] + 3 * cfg["d_mid"] + 3 * cfg["d_out"]) + cfg["batch"] * cfg["lmax"] * cfg["d_in"] * 2
    big_workload = est_bytes > 25e9
    if big_workload:
        nbatches = 1
        args.no_graph = True
        args.no_stages = True
    host_io = importlib.import_module("p2t_b200.host_io")
    numa_cpus = host_io.bind_host_thread_to_gpu(dev) if world > 1 else None  # before any pinned buffer exists
    if args.device_synth:
        args.no_e2e = args.no_cpu_baseline = True
    batches = [synth.make_config_batch(args.workload, seed=1234 + 17 * i, rank=rank, same_lengths_as_rank0=True,
                                       device=dev if args.device_synth else None, batch=cfg["batch"])
               for i in range(nbatches)]
    B = cfg["batch"]
    acfg = pkg.ModalityAdapterConfig(input_dim=cfg["d_in"], intermediate_dim=cfg["d_mid"], output_dim=cfg["d_out"], dropout_rate=0.3)
    adapter = pkg.ModalityAdapter(acfg).to(dev).to(torch.bfloat16)
    with torch.no_grad():
        sb0 = batches[0]
        adapter.fc1.weight.copy_(sb0.w1); adapter.fc1.bias.copy_(sb0.b1)
        adapter.fc2.weight.copy_(sb0.w2); adapter.fc2.bias.copy_(sb0.b2)
    adapter.eval() if args.eval_mode else adapter.train()
    params = [adapter.fc1.weight, adapter.fc1.bias, adapter.fc2.weight, adapter.fc2.bias]

    if args.device_synth:
        host = None
        resident = [dict(x=b.x, pm=b.prot_mask.to(dev), text=b.text, tm=b.text_mask.to(dev)) for b in batches]
    else:
        host = [dict(x=b.x.pin_memory(), pm=b.prot_mask.pin_memory(), text=b.text.pin_memory(), tm=b.text_mask.pin_memory())
                for b in batches]
        resident = [{k: v.to(dev) for k, v in h.items()} for h in host]
    valid_rows = [int(b.prot_lens.sum()) for b in batches]

    # N > 1: the exchange step of the sharded batch runs as peer-memory kernels (no NCCL on the data path)
    exchange = pdist.ShardedExchange(B, 2 * cfg["d_out"]) if world > 1 else None

    # the collater knows the sequence lengths: packed activations are sized for the valid rows, not for B * L_max
    rows_bound = max(valid_rows) + 256

    def step(inp):
        for p in params:
            p.grad = None
        if world > 1:
            loss = pdist.distributed_contrastive_step(inp["x"], inp["pm"], adapter, inp["text"], inp["tm"], exchange=exchange,
                                                      max_valid_rows=rows_bound)
        else:
            loss = pkg.contrastive_step(inp["x"], inp["pm"], adapter, inp["text"], inp["tm"], max_valid_rows=rows_bound)
        loss.backward()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms: float) -> float:
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ------------------------------ device-resident timing ------------------------------
    # Pass A (eager launches): the library brackets every GEMM launch with CUDA events on its stream -> roofline.
    # Pass B (CUDA-graph replay of the same step through GraphedContrastiveStep; for N > 1 the sharded step including
    # its peer-memory exchange): the headline `value`.
    for i in range(args.warmup):
        step(resident[i % nbatches])
    barrier()
    lib.gemm_timing_enable(True)
    lib.reset_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    t_host0 = time.perf_counter()
    for i in range(args.steps):
        loss = step(resident[i % nbatches])
    host_issue_ms = (time.perf_counter() - t_host0) * 1e3 / args.steps  # CPU time to enqueue one eager step
    ev1.record()
    barrier()
    launches = lib.launch_count()
    gemm_ms, gemm_launches, gemm_each = lib.gemm_timing_collect()
    lib.gemm_timing_enable(False)
    eager_ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    last_loss = float(loss.item())
    use_graph = not args.no_graph
    sampler = ClockSampler(local_rank)
    if use_graph:
        graphs = [pkg.GraphedContrastiveStep(adapter, r["x"], r["pm"], r["text"], r["tm"], seed=1000 * (i + 1),
                                             exchange=exchange, max_valid_rows=rows_bound)
                  for i, r in enumerate(resident)]
        for i in range(args.warmup):
            graphs[i % nbatches].replay()
        barrier()
        if rank == 0:
            sampler.start()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        g0.record()
        for i in range(args.steps):
            loss = graphs[i % nbatches].replay()
        g1.record()
        barrier()
        clocks = sampler.stop() if rank == 0 else None
        ms_total = max_over_ranks(g0.elapsed_time(g1))
        launches = sum(graphs[i % nbatches].launches_per_replay for i in range(args.steps))
        last_loss = float(loss.item())
    else:
        # eager numbers are the headline; sample the clocks over a second identical eager pass
        if rank == 0:
            sampler.start()
        barrier()
        ev0.record()
        for i in range(args.steps):
            loss = step(resident[i % nbatches])
        ev1.record()
        barrier()
        clocks = sampler.stop() if rank == 0 else None
        ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    ms_per_step = ms_total / args.steps
    value = world * B * args.steps / (ms_total / 1e3)

    # ------------------------------ N > 1: the same step with independent per-rank lengths ------------------------------
    # `value` above is the idealised weak-scaling workload (every rank holds rank 0's multiset of lengths: exactly fixed
    # per-GPU work).  A real DistributedSampler split (scripts/train_contrast.py:551-556) gives every rank its own draw:
    # the ranks' residue-row totals differ and everybody waits for the heaviest at the exchange.  Two more timings of
    # the captured step: random shards (each rank its own draw) and dist.balanced_shards over the same global batch.
    imbalance = None
    if world > 1 and use_graph and not args.no_imbalance_leg and not args.device_synth:
        imbalance = {}
        draws = [[synth.draw_lengths(args.workload, seed=1234 + 17 * i, rank=r) for r in range(world)] for i in range(nbatches)]
        for leg_name in ("random_shards", "balanced_shards"):
            leg_batches, ratios = [], []
            for i in range(nbatches):
                all_l = torch.cat([d[0] for d in draws[i]])
                all_t = torch.cat([d[1] for d in draws[i]])
                if leg_name == "balanced_shards":
                    mine = torch.tensor(pdist.balanced_shards(all_l.tolist(), world)[rank])
                    shards_rows = [int(all_l[torch.tensor(sh)].sum()) for sh in pdist.balanced_shards(all_l.tolist(), world)]
                else:
                    mine = torch.arange(rank * B, (rank + 1) * B)
                    shards_rows = [int(all_l[r * B:(r + 1) * B].sum()) for r in range(world)]
                ratios.append(max(shards_rows) / (sum(shards_rows) / world))
                leg_batches.append(synth.make_config_batch(args.workload, seed=1234 + 17 * i, rank=rank,
                                                           lens=all_l[mine].tolist(), tlens=all_t[mine].tolist()))
            leg_res = [dict(x=b.x.to(dev), pm=b.prot_mask.to(dev), text=b.text.to(dev), tm=b.text_mask.to(dev)) for b in leg_batches]
            leg_graphs = [pkg.GraphedContrastiveStep(adapter, r["x"], r["pm"], r["text"], r["tm"], seed=77 * (i + 1), exchange=exchange)
                          for i, r in enumerate(leg_res)]
            for i in range(args.warmup):
                leg_graphs[i % nbatches].replay()
            barrier()
            l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            l0.record()
            for i in range(args.steps):
                leg_graphs[i % nbatches].replay()
            l1.record()
            barrier()
            leg_ms = max_over_ranks(l0.elapsed_time(l1))
            imbalance[leg_name] = {"value": world * B * args.steps / (leg_ms / 1e3), "unit": UNIT, "ms_per_step": leg_ms / args.steps,
                                   "heaviest_rank_rows_over_mean": round(sum(ratios) / len(ratios), 4)}
            del leg_graphs, leg_res, leg_batches
        torch.cuda.empty_cache()

    # ------------------------------ per-kernel breakdown ------------------------------
    # One more eager pass with a CUDA event after EVERY library launch.  The text branch is kept on the main stream
    # (one in-order stream) and every step starts behind a ~2 ms device-side sleep, so the host has enqueued the whole
    # step before its first kernel runs.  Every kernel is bracketed by a "begin" stamp right before its launch and an
    # end stamp right after: the difference is the kernel on the stream (plus one event record), not the Python launch
    # rate and not the gap to its predecessor.
    stages = None
    if not args.no_stages and world == 1:
        os.environ["P2T_TEXT_STREAM"] = "0"
        for i in range(2):
            step(resident[i % nbatches])
        barrier()
        lib.launch_timing_enable(True)
        cur = torch.cuda.current_stream().cuda_stream
        n_st = min(args.steps, 10)
        for i in range(n_st):
            torch.cuda._sleep(4_000_000)
            lib.launch_timing_mark(cur)
            step(resident[i % nbatches])
        barrier()
        stamps = lib.launch_timing_collect()
        lib.launch_timing_enable(False)
        os.environ.pop("P2T_TEXT_STREAM", None)
        per, order, seen, first_after_mark = {}, [], {}, False
        for name, ms in stamps:
            if name == "mark":
                seen, first_after_mark = {}, True
                continue
            if name == "begin":  # opens a bracket: the next stamp's time is the kernel alone
                first_after_mark = False
                continue
            k = seen[name] = seen.get(name, 0) + 1
            key = f"{name}#{k}"
            if key not in per:
                per[key] = []
                order.append(key)
            if not first_after_mark:  # the first kernel after a mark is timed from the mark, i.e. through the sleep
                per[key].append(ms)
            first_after_mark = False
        rows = sum(valid_rows[i % nbatches] for i in range(n_st)) / n_st
        trows = sum(int(batches[i % nbatches].text_mask.sum()) for i in range(n_st)) / n_st
        di, dm, do = cfg["d_in"], cfg["d_mid"], cfg["d_out"]
        pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
        hbm_peak = float(pk.get("hbm_gbs", 6545.0))
        tf_burst = float(pk.get("bf16_tflops", 1672.7))
        algo = {  # algorithmic bytes (HBM kernels) or flops (GEMMs) per step, SURVEY.md §8d
            "pool_partial_kernel#1": ("hbm", 2.0 * trows * do), "pool_partial_kernel#2": ("hbm", 2.0 * rows * do),
            "gather_rows_kernel#1": ("hbm", 4.0 * rows * di), "adapter_tail_bwd_kernel#1": ("hbm", 6.0 * rows * do),
            "gemm_fc1#1": ("tensor", 2.0 * rows * di * dm), "gemm_fc2#1": ("tensor", 2.0 * rows * dm * do),
            "gemm_dgrad#1": ("tensor", 2.0 * rows * dm * do), "gemm_wgrad#1": ("tensor", 2.0 * rows * dm * do),
            "gemm_wgrad#2": ("tensor", 2.0 * rows * di * dm)}
        total_us = sum(statistics.median(v) for v in per.values() if v) * 1e3
        stages = []
        for key in order:
            if not per[key]:
                continue
            us = statistics.median(per[key]) * 1e3
            ent = {"kernel": key, "us": round(us, 2), "share": round(us / total_us, 4)}
            if key in algo and us > 0:
                bound, work = algo[key]
                if bound == "hbm":
                    ent.update(bound="hbm", achieved=round(work / 1e9 / (us / 1e6), 1), unit="GB/s", peak=hbm_peak,
                               frac=round(work / 1e9 / (us / 1e6) / hbm_peak, 3))
                else:
                    ent.update(bound="tensor", achieved=round(work / 1e12 / (us / 1e6), 1), unit="TFLOP/s", peak=tf_burst,
                               frac=round(work / 1e12 / (us / 1e6) / tf_burst, 3))
            stages.append(ent)

    # ------------------------------ end to end (host buffers) ------------------------------
    # Inputs start in pinned HOST memory.  The public HostStager copies only the valid rows of the step's residue
    # states and text hidden states to the device (packed) on a copy stream; the copy of step i+1 is issued before
    # step i's kernels so it overlaps them.  Every step's H2D copy and the D2H read of its loss are inside the
    # timed region.
    e2e = None
    if not args.no_e2e:
        stager = pkg.HostStager(dev)

        def submit(h):
            stager.submit(h["x"], h["pm"], h["text"], h["tm"])

        def e2e_step(i, last):
            batch = stager.take()
            if not last:
                submit(host[(i + 1) % nbatches])
            for p in params:
                p.grad = None
            kw = dict(residue_lengths=batch.residue_lengths, text_lengths=batch.text_lengths)
            if world > 1:
                l = pdist.distributed_contrastive_step(batch.residue_rows, None, adapter, batch.text_rows, None,
                                                       exchange=exchange, **kw)
            else:
                l = pkg.contrastive_step(batch.residue_rows, None, adapter, batch.text_rows, None, **kw)
            l.backward()
            return float(l.item()), batch.h2d_bytes  # device -> host read of the step's result

        nwarm = min(3, args.warmup)
        submit(host[0])
        for i in range(nwarm):
            e2e_step(i, last=(i == nwarm - 1))
        barrier()
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        submit(host[0])
        h2d = 0
        for i in range(args.steps):
            _, nbytes = e2e_step(i, last=(i == args.steps - 1))
            h2d += nbytes
        e1.record()
        barrier()
        e2e_ms = max_over_ranks(max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3))
        # what bounds it: the step's inputs cross PCIe once (packed valid rows); compare with one large pinned copy
        probe_h = torch.empty(256 << 20, dtype=torch.uint8).pin_memory()
        probe_d = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        probe_d.copy_(probe_h, non_blocking=True)
        barrier()  # N > 1: every rank probes at the same time -> the per-rank figure is the CONCURRENT ceiling
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for _ in range(4):
            probe_d.copy_(probe_h, non_blocking=True)
        c1.record()
        torch.cuda.synchronize()
        pcie_peak = 4 * (256 << 20) / 1e9 / (c0.elapsed_time(c1) / 1e3)
        pcie_aggregate = pcie_peak
        if world > 1:
            tp = torch.tensor([pcie_peak], dtype=torch.float64, device=dev)
            dist.all_reduce(tp, op=dist.ReduceOp.SUM)
            pcie_aggregate = float(tp.item())
        del probe_h, probe_d
        e2e = {"value": world * B * args.steps / (e2e_ms / 1e3), "unit": UNIT, "h2d_bytes_per_step": h2d // args.steps,
               "d2h_bytes_per_step": 4, "ms_per_step": e2e_ms / args.steps,
               "bound": "pcie", "h2d_gb_per_s": (h2d / args.steps) / 1e9 / (e2e_ms / args.steps / 1e3),
               "h2d_peak_gb_per_s": pcie_peak, "h2d_peak_aggregate_gb_per_s": pcie_aggregate,
               "h2d_aggregate_gb_per_s": world * (h2d / args.steps) / 1e9 / (e2e_ms / args.steps / 1e3),
               "h2d_peak_how": "4 x 256 MiB pinned cudaMemcpyAsync, all ranks probing at the same time (rank 0's figure; "
                               "aggregate = sum over ranks): the box's concurrent host->device ceiling",
               "host_numa_binding": ("rank bound to the CPUs local to its GPU before pinning its buffers" if numa_cpus
                                     else ("not needed (one GPU)" if world == 1 else "topology unreadable: default placement")),
               "how": "pinned host batch -> HostStager (valid rows only, copy stream, next step's copy overlaps this "
                      "step's kernels) -> contrastive_step(packed rows + lengths) -> backward -> loss.item()"}

    # ------------------------------ outside the metric: gradient exchange + optimizer ------------------------------
    # (the DDP weight-gradient all-reduce and the optimizer step are reported separately.)
    # mean all-reduce of the four weight gradients over peer memory (N > 1) + clip_grad_norm_ + AdamW in 3 launches
    optim_leg = None
    if not args.no_optimizer:
        peer = importlib.import_module("p2t_b200.peer")
        if exchange is not None:
            exchange.check()
        opt = pkg.FusedAdamW(params, lr=1e-4, eps=1e-6, betas=(0.9, 0.999), max_grad_norm=1.0)
        reducer = peer.PeerGradAllReduce(params) if world > 1 else None  # eager form: stage, reduce, copy back
        step(resident[0])
        grads = [p.grad for p in params]

        def opt_step():
            if reducer is not None:
                reducer.reduce_(grads)
            opt.step()

        for _ in range(3):
            opt_step()
        barrier()
        o0, o1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n_opt = 20
        o0.record()
        for _ in range(n_opt):
            opt_step()
        o1.record()
        barrier()
        opt_ms = max_over_ranks(o0.elapsed_time(o1)) / n_opt
        n_el = sum(p.numel() for p in params)
        opt_bytes = n_el * (2 + 2 + 4 * 3 * 2 + 2)  # norm pass reads g; update reads g, m, v, master and writes m, v, master, p
        optim_leg = {"ms_per_step": opt_ms, "parameters": n_el, "algorithmic_bytes": opt_bytes,
                     "gb_per_s": (opt_bytes / 1e9) / (opt_ms / 1e3) if reducer is None else None,
                     "includes": ("peer-memory mean all-reduce of the weight gradients + " if reducer is not None else "") +
                                 "clip_grad_norm_ + AdamW (FusedAdamW, 3 launches); outside `value`"}
        if reducer is not None:
            reducer.buffer.check()

    # ------------------------------ the exchange step alone (N > 1): NVLink roofline ------------------------------
    exchange_leg = None
    if exchange is not None and not args.no_exchange_probe:
        try:
            E = 2 * cfg["d_out"]
            t_probe = torch.randn(B, E, device=dev, dtype=torch.float32)
            for _ in range(3):
                exchange.text(t_probe)
            barrier()
            x0, x1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n_x = 20
            x0.record()
            for _ in range(n_x):
                exchange.text(t_probe)  # push kernel + arrive kernel, the pair the step launches
            x1.record()
            barrier()
            us = max_over_ranks(x0.elapsed_time(x1)) / n_x * 1e3
            sent = (world - 1) * B * E * 4
            exchange_leg = {"us_per_exchange": us, "bytes_sent_per_rank": sent,
                            "gb_per_s_sent_per_rank": sent / 1e9 / (us / 1e6), "peak_gb_per_s": 900.0,
                            "peak_source": "nominal NVLink 5 per direction per GPU (not measured by the driver)",
                            "note": "push + arrive kernels back to back, eager launches, both launch latencies included; "
                                    "at B*E*4 bytes per peer the pair is latency-bound, not link-bound"}
            exchange.check()
        except Exception as exc:  # noqa: BLE001 - a probe must never cost the benchmark its line
            exchange_leg = {"error": f"{type(exc).__name__}: {exc}"}

    # ------------------------------ one whole training step in one graph (outside the metric) ------------------------------
    # forward + backward + gradient mean over ranks (fp32, zero-copy through the reducer's channel buffer) + clip +
    # AdamW captured together: what the optimizer side costs ON TOP of `value`'s step when nothing is staged or relaunched
    train_leg = None
    if use_graph and not args.no_optimizer and not big_workload:
        try:
            peer = importlib.import_module("p2t_b200.peer")
            opt2 = pkg.FusedAdamW(params, lr=1e-5, eps=1e-6, betas=(0.9, 0.999), max_grad_norm=1.0)
            r0 = resident[0]
            variants = [("step_only", graphs[0], None)]
            if world > 1:
                red_plain = peer.PeerGradAllReduce.for_adapter(adapter)
                red_fused = peer.OverlappedGradReduce(adapter)
                variants.append(("step_reduce_optimizer_unfused", pkg.GraphedContrastiveStep(
                    adapter, r0["x"], r0["pm"], r0["text"], r0["tm"], seed=4242, exchange=exchange, grad_reducer=red_plain,
                    optimizer=opt2, max_valid_rows=rows_bound), red_plain))
                variants.append(("step_reduce_optimizer", pkg.GraphedContrastiveStep(
                    adapter, r0["x"], r0["pm"], r0["text"], r0["tm"], seed=4242, exchange=exchange, grad_reducer=red_fused,
                    optimizer=opt2, max_valid_rows=rows_bound), red_fused))
            else:
                variants.append(("step_reduce_optimizer", pkg.GraphedContrastiveStep(
                    adapter, r0["x"], r0["pm"], r0["text"], r0["tm"], seed=4242, optimizer=opt2, max_valid_rows=rows_bound), None))
            res_ms = {}
            for name, g, _ in variants:
                for _ in range(3):
                    g.replay()
                barrier()
                t0e, t1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                t0e.record()
                for _ in range(args.steps):
                    g.replay()
                t1e.record()
                barrier()
                res_ms[name] = max_over_ranks(t0e.elapsed_time(t1e)) / args.steps
            full = variants[-1][1]
            train_leg = {"ms_per_step": res_ms["step_reduce_optimizer"], "step_only_ms": res_ms["step_only"],
                         "optimizer_side_ms": res_ms["step_reduce_optimizer"] - res_ms["step_only"],
                         "launches_per_replay": full.launches_per_replay,
                         "includes": "fused step + " + ("fp32 gradient mean over ranks — dW2/db2 by the idle epilogue warps inside the dW1 GEMM's "
                                                        "launch (fused tcgen05 GEMM + NVLink peer-memory reduce), dW1/db1 after it + "
                                                        if world > 1 else "")
                                     + "clip_grad_norm_ + AdamW, one CUDA graph per step (same batch replayed)"}
            if world > 1:
                train_leg["optimizer_side_ms_unfused"] = res_ms["step_reduce_optimizer_unfused"] - res_ms["step_only"]
            for _, _, red in variants:
                if red is not None:
                    red.buffer.check()
        except Exception as exc:  # noqa: BLE001 - an extra leg must never cost the benchmark its line
            train_leg = {"error": f"{type(exc).__name__}: {exc}"}

    # ------------------------------ roofline of the dominant kernel ------------------------------
    d_in, d_mid, d_out = cfg["d_in"], cfg["d_mid"], cfg["d_out"]
    flops_per_row = 2.0 * d_in * d_mid + 2.0 * d_mid * d_out + 2.0 * d_in * d_mid + 4.0 * d_mid * d_out
    rows_timed = sum(valid_rows[i % nbatches] for i in range(args.steps))
    gemm_flops = flops_per_row * rows_timed
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peaks = json.load(open(peaks_path))
        peak_burst = float(peaks.get("bf16_tflops", 1672.7))
        peak_sustained = float(peaks.get("bf16_tflops_sustained", peak_burst))
        peak_origin = "measured (MEASURED_PEAKS.json)"
    else:
        peak_burst, peak_sustained, peak_origin = 1672.7, 1400.0, "fallback (B200_PROFILING.md)"
    # which denominator: the burst figure for a timed window well under a second (the GPU never reaches the power /
    # thermal steady state MEASURED_PEAKS' sustained figure was taken in), the sustained one for longer windows
    window_s = eager_ms_total / 1e3
    use_burst = window_s < 1.0
    peak_tf = peak_burst if use_burst else peak_sustained
    peak_src = f"{peak_origin}: {'bf16_tflops (burst)' if use_burst else 'bf16_tflops_sustained'}, timed window {window_s:.3f} s"
    achieved_tf = gemm_flops / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else 0.0
    # mean duration of each of the step's GEMM launches, in launch order: fc1, fc2, fc2-dgrad, dW2, dW1
    per_step = gemm_launches // args.steps if args.steps else 0
    per_gemm_us = None
    if per_step and per_step * args.steps == gemm_launches and len(gemm_each) == gemm_launches:
        per_gemm_us = [round(1e3 * sum(gemm_each[i::per_step]) / args.steps, 1) for i in range(per_step)]
    # DRAM bytes per GEMM launch from the last ncu --set full capture, valid only for the kernel source it was taken
    # with (tools/traffic_from_ncu.py stamps the capture with a digest of csrc/gemm_*): otherwise null, never stale
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        if tj.get("gemm_source_sha16") == gemm_source_digest():
            traffic = tj.get("dram_bytes_per_launch")
    algo_bytes_gemm = None
    if world == 1 or True:
        r_ = rows_timed / args.steps
        # operands + outputs of the five launches moved once: x, W1, h1, g1 | h1, W2, a, g2 | dz2, W2, g1, dz1 | dz2, h1, dW2 | dz1, x, dW1
        algo_bytes_gemm = 2.0 * (r_ * (d_in + 2 * d_mid) + d_in * d_mid + r_ * (d_mid + 2 * d_out) + d_mid * d_out
                                 + r_ * (d_out + 2 * d_mid) + d_mid * d_out + r_ * (d_out + d_mid) + d_mid * d_out
                                 + r_ * (d_mid + d_in) + d_in * d_mid) / 5.0
    roofline = {"bound": "tensor", "kernel": "gemm_bf16_tcgen05_kernel", "achieved": achieved_tf, "peak": peak_tf,
                "unit": "TFLOP/s", "frac": achieved_tf / peak_tf, "traffic": traffic,
                "algorithmic_bytes_per_launch": algo_bytes_gemm, "peak_source": peak_src,
                "frac_of_burst": achieved_tf / peak_burst, "frac_of_sustained": achieved_tf / peak_sustained,
                "launches_timed": gemm_launches, "kernel_ms_per_step": gemm_ms / args.steps,
                "kernel_share_of_step": gemm_ms / eager_ms_total if world == 1 else None,
                "timed_in": "eager pass of the same K steps (per-launch CUDA events on the launching stream)",
                "algorithmic_flops_per_step": gemm_flops / args.steps,
                "per_gemm_us": per_gemm_us,
                "whole_step_frac": (gemm_flops / (ms_total / 1e3) / 1e12) / peak_tf,
                "whole_step_frac_of_burst": (gemm_flops / (ms_total / 1e3) / 1e12) / peak_burst,
                "whole_step_frac_of_sustained": (gemm_flops / (ms_total / 1e3) / 1e12) / peak_sustained}

    if rank == 0:
        cpu_baseline = None
        if world == 1 and not args.no_cpu_baseline:
            # ~10 s of CPU work: the whole batch when it is config 2's 32 pairs, else its first 32 pairs
            n_cpu = min(B, CPU_SAMPLE_PAIRS)
            times, cores, kind = cpu_step_time(batches[0], n_cpu, iters=5, warmup=1)
            cpu_baseline = {"value": n_cpu / min(times), "unit": UNIT, "cores": cores, "kind": kind,
                            "seconds_of_cpu_work": round(sum(times), 2),
                            "sample": f"first {n_cpu} pairs of the {args.workload} batch ({'the whole step' if n_cpu == B else 'a bounded sample'}), "
                                      "fp32 fwd+bwd, eval-mode dropout, best of 5 after 1 warm-up; "
                                      + ("the reference's own ModalityAdapter / readout_embeddings / SegmentedBatchInfoNCELoss"
                                         if kind == "reference" else "CPU restatement oracle/restatement.py")}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic (drawn on the device)" if args.device_synth else "synthetic",
            "config": {"workload": args.workload, "d_in": d_in, "d_mid": d_mid, "d_out": d_out, "pairs_per_gpu": B,
                       "global_pairs": world * B, "residue_len": [cfg["lmin"], cfg["lmax"]],
                       "valid_rows_per_step": valid_rows[0], "dropout_p": 0.0 if args.eval_mode else 0.3,
                       "parallelism": f"dp{world}",
                       "exchange": None if world == 1 else "text embeddings all-gathered by peer-memory kernels over NVLink "
                                                           "(csrc/peer.cu), inside the step's CUDA graph; no NCCL on the data path",
                       "rank_lengths": "rank 0's multiset of sequence lengths on every rank, own data (fixed per-GPU work)",
                       "l2": ("inputs+activations per step exceed the 126 MB L2; 2 batches alternate" if nbatches == 2 else
                              "one resident batch whose inputs+activations exceed the 126 MB L2 a hundred times; eager launches"),
                       "cta_group": int(os.environ.get("P2T_CTA_GROUP", "2"))},
            "loss": last_loss, "gpu_launches": int(launches), "cuda_graph": bool(use_graph),
            "eager_ms_per_step": eager_ms_total / args.steps, "host_issue_ms_per_step_eager": host_issue_ms, "clocks": clocks, "e2e": e2e, "roofline": roofline,
            "cpu_baseline": cpu_baseline, "optimizer": optim_leg, "training_step": train_leg, "exchange": exchange_leg,
            "imbalance": imbalance,
            "stages": stages,
        }
        print(json.dumps(line), flush=True)
    if exchange is not None:
        exchange.check()
        torch.cuda.synchronize()
        dist.barrier()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
