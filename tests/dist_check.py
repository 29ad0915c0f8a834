"""Multi-GPU parity check of the sharded contrastive step (run under torchrun, one rank per GPU).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
        tests/dist_check.py [workload] [weight_gain]

Every rank runs `dist.distributed_contrastive_step` on its own shard, with and without the symmetric column term, in
three forms: NCCL all-gather (eager), the peer-memory exchange (eager), and the peer-memory exchange captured in a
CUDA graph together with the peer-memory mean all-reduce of the weight gradients.  Gradients of the first two are
averaged across ranks with NCCL the way DDP does.  Rank 0 rebuilds the
GLOBAL batch on the CPU and runs the oracle on it in one process: the mean of the local losses must
equal the global-batch loss, the averaged gradients the global-batch gradients.  Exit code 0 = parity.
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry
from oracle import restatement as R

PARAMS = ("fc1.weight", "fc1.bias", "fc2.weight", "fc2.bias")


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "cfg1_esm2_t6_llama1b"
    gain = float(sys.argv[2]) if len(sys.argv) > 2 else 2.5
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    pkg = entry.load_package()
    import importlib
    synth = importlib.import_module("p2t_b200.synth")
    pdist = importlib.import_module("p2t_b200.dist")
    shards = [synth.make_config_batch(name, weight_gain=gain, rank=r) for r in range(world)]
    sb = shards[rank]
    cfg = pkg.ModalityAdapterConfig(input_dim=sb.w1.shape[1], intermediate_dim=sb.w1.shape[0], output_dim=sb.w2.shape[0])
    ad = pkg.ModalityAdapter(cfg).to(dev).to(torch.bfloat16).eval()
    with torch.no_grad():
        ad.fc1.weight.copy_(sb.w1); ad.fc1.bias.copy_(sb.b1); ad.fc2.weight.copy_(sb.w2); ad.fc2.bias.copy_(sb.b2)
    ok = True
    peer = importlib.import_module("p2t_b200.peer")
    prms = (ad.fc1.weight, ad.fc1.bias, ad.fc2.weight, ad.fc2.bias)
    exchange = pdist.ShardedExchange(sb.x.shape[0], 2 * sb.w2.shape[0], symmetric=True)
    reducer = peer.PeerGradAllReduce.for_adapter(ad)  # graph form: all four gradients cross the ranks in fp32
    f32m = lambda n: torch.empty(n, dtype=torch.float32, device="meta")
    reducer_mixed = peer.PeerGradAllReduce([prms[0], f32m(prms[1].numel()), prms[2], f32m(prms[3].numel())])  # eager form
    dx, dpm, dtx, dtm = sb.x.to(dev), sb.prot_mask.to(dev), sb.text.to(dev), sb.text_mask.to(dev)
    overlapped = peer.OverlappedGradReduce(ad)  # graph form with the dW2 / db2 mean inside the dW1 GEMM's launch
    for sym, form in ((False, "nccl"), (True, "nccl"), (False, "peer"), (True, "peer"), (False, "peer+graph"), (True, "peer+graph"),
                      (False, "peer+graph+overlap"), (True, "peer+graph+overlap")):
        ad.zero_grad(set_to_none=True)
        aux = pkg.StepAux()
        torch.cuda.synchronize()
        dist.barrier()  # rank 0 has finished the CPU oracle of the previous form: peer-memory waits give up after 20 s
        if form.startswith("peer+graph"):
            gstep = pkg.GraphedContrastiveStep(ad, dx, dpm, dtx, dtm, symmetric=sym, exchange=exchange,
                                               grad_reducer=overlapped if form.endswith("overlap") else reducer)
            for _ in range(3):
                loss = gstep.replay()
            aux = gstep.aux
            reduced = [g.float().cpu() for g in gstep.grads]  # already the mean over ranks, identical bits everywhere
            same = torch.tensor([1], device=dev)
            for g in gstep.grads:
                g0 = g.clone()
                dist.broadcast(g0, 0)
                same &= torch.equal(g0, g)
            dist.all_reduce(same, op=dist.ReduceOp.MIN)
            ok &= bool(same.item())
        else:
            loss = pdist.distributed_contrastive_step(dx, dpm, ad, dtx, dtm, symmetric=sym, aux=aux,
                                                      exchange=exchange if form == "peer" else None)
            loss.backward()
        lt = loss.detach().clone()
        dist.all_reduce(lt)
        grads = {}
        if form.startswith("peer+graph"):
            grads = {k: reduced[i] for i, k in enumerate(PARAMS)}
        elif form == "peer":
            # eager form of the same gradient mean: the mixed bf16 / fp32 reducer over NVLink peer memory (stage, reduce,
            # copy back).  The biases travel in fp32 (aux.bias_grads_f32) and are NOT rounded per rank first.
            db1_f32, db2_f32 = aux.bias_grads_f32
            mixed = [prms[0].grad.clone(), db1_f32.clone(), prms[2].grad.clone(), db2_f32.clone()]
            reducer_mixed.reduce_(mixed)
            grads = {k: g.float().cpu() for k, g in zip(PARAMS, mixed)}
        else:
            # what DDP does (scripts/train_contrast.py:611-614), with the biases taken in fp32 as above
            db1_f32, db2_f32 = aux.bias_grads_f32
            grads = {}
            for k, g in zip(PARAMS, (prms[0].grad.float(), db1_f32, prms[2].grad.float(), db2_f32)):
                g = g.clone()
                dist.all_reduce(g)
                grads[k] = (g / world).cpu()
        exchange.check()
        reducer.buffer.check()
        reducer_mixed.buffer.check()
        overlapped.buffer.check()
        am_row = [torch.empty_like(aux.argmax_row) for _ in range(world)]
        dist.all_gather(am_row, aux.argmax_row)
        if rank == 0:
            f = torch.float32
            B = sb.x.shape[0]
            Lmax = max(s.x.shape[1] for s in shards)
            Tmax = max(s.text.shape[1] for s in shards)
            X = torch.zeros(world * B, Lmax, sb.x.shape[2])
            PM = torch.zeros(world * B, Lmax, dtype=torch.long)
            TX = torch.zeros(world * B, Tmax, sb.text.shape[2])
            TM = torch.zeros(world * B, Tmax, dtype=torch.long)
            for r, s in enumerate(shards):
                X[r * B:(r + 1) * B, :s.x.shape[1]] = s.x.to(f)
                PM[r * B:(r + 1) * B, :s.x.shape[1]] = s.prot_mask
                TX[r * B:(r + 1) * B, :s.text.shape[1]] = s.text.to(f)
                TM[r * B:(r + 1) * B, :s.text.shape[1]] = s.text_mask
            st = R.step_forward(X, PM, sb.w1.to(f), sb.b1.to(f), sb.w2.to(f), sb.b2.to(f), TX, TM, 0.05, 1, sym)
            ref = R.step_backward(st, X, PM, sb.w1.to(f), sb.w2.to(f), 0.05, 1, sym)
            lrel = abs(lt.item() / world - st.loss.item()) / abs(st.loss.item())
            line = f"world={world} {name} symmetric={sym} exchange={form}: loss {lt.item() / world:.6f} vs global oracle {st.loss.item():.6f} rel={lrel:.2e}"
            ok &= lrel <= 1e-3
            for k in PARAMS:
                a, b = grads[k].double().flatten(), ref[k].double().flatten()
                c = torch.nn.functional.cosine_similarity(a, b, dim=0).item()
                m = ((a - b).abs().max() / b.abs().max()).item()
                line += f" | {k} cos={c:.6f} maxrel={m:.2e}"
                ok &= c >= 0.999 and m <= 1e-2  # north_star's tolerance, row term and symmetric extension alike
            am_ref, _ = R.retrieval_argmax(st.p, st.t)
            am = torch.cat([a.cpu().long() for a in am_row])
            line += f" | argmax_row exact={torch.equal(am, am_ref)}"
            print(line, flush=True)
    # north_star's reduce-scatter form of the gathered embeddings' gradient (SURVEY.md §8e "Backward exchange"): every
    # rank forms its B x B_global block's contribution to d(loss)/d(text embeddings) for ALL gathered rows, the
    # contributions are reduce-scattered (NCCL over NVLink) so that rank k ends with the gradient of ITS text rows.
    # (The Stage-1 step itself never needs it — the text side is frozen — which is why it lives here as a checked
    # building block.)  Oracle: d(global mean loss)/dT on the concatenated batch.
    core = sys.modules["p2t_b200._core"]
    step_mod = sys.modules["p2t_b200.step"]
    B = dx.shape[0]
    t_local = pkg.text_embeddings(dtx, dtm, dtype=torch.float32)
    t_global = pdist.all_gather_embeddings(t_local)
    p_f32 = aux.protein_embeddings.float().contiguous()   # unit-norm protein embeddings of the last step (bf16 values)
    labels = step_mod._rank_labels(rank, B, dev)
    res = core.infonce_forward(aux.protein_embeddings, t_global, labels, 0.05, need_grad=True, p_f32=p_f32, t_f32=t_global)
    _, dt_full = core.infonce_backward(res, aux.protein_embeddings, None, 0.05, need_dp=False, need_dt=True, p_f32=p_f32, t_f32=t_global)
    dt_mine = pdist.reduce_scatter_text_grad(dt_full / world)
    p_all = pdist.all_gather_embeddings(p_f32)
    dt_all = pdist.all_gather_embeddings(dt_mine)
    torch.cuda.synchronize()
    if rank == 0:
        _, _, dto = R.infonce_backward(p_all.double().cpu(), t_global.double().cpu(), torch.arange(world * B), 0.05)
        a, b = dt_all.double().cpu().flatten(), dto.flatten()
        c = torch.nn.functional.cosine_similarity(a, b, dim=0).item()
        m = ((a - b).abs().max() / b.abs().max()).item()
        print(f"world={world} {name} reduce-scattered dT: cos={c:.6f} maxrel={m:.2e}", flush=True)
        ok &= c >= 0.9999 and m <= 1e-3
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    torch.cuda.synchronize()
    exchange.close()
    reducer.close()
    reducer_mixed.close()
    overlapped.close()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
