"""One eager pass over the late-round kernels (for ncu): the sharded step through a world-of-one peer exchange, the
gradient mean all-reduce, clip + AdamW, and the Stage-2 scatter hand-off — config 2 shapes.

    ncu --set full --clock-control none --import-source on -k regex:"adamw|grad_sqnorm|sim_tile|loss_bwd_dp|row_inv_norm|peer_|scale_rows" \
        -o gpurun_out/next_r01 python tools/next_probe.py
"""
import importlib
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry


def main():
    pkg = entry.load_package()
    synth = importlib.import_module("p2t_b200.synth")
    pdist = importlib.import_module("p2t_b200.dist")
    handoff = importlib.import_module("p2t_b200.handoff")
    dev = torch.device("cuda:0")
    name = sys.argv[1] if len(sys.argv) > 1 else "cfg2_esm2_3b_llama8b"
    sb = synth.make_config_batch(name)
    cfg = pkg.ModalityAdapterConfig(input_dim=sb.w1.shape[1], intermediate_dim=sb.w1.shape[0], output_dim=sb.w2.shape[0])
    ad = pkg.ModalityAdapter(cfg).to(dev).to(torch.bfloat16).train()
    with torch.no_grad():
        ad.fc1.weight.copy_(sb.w1); ad.fc1.bias.copy_(sb.b1); ad.fc2.weight.copy_(sb.w2); ad.fc2.bias.copy_(sb.b2)
    params = [ad.fc1.weight, ad.fc1.bias, ad.fc2.weight, ad.fc2.bias]
    x, pm, th, tm = (t.to(dev) for t in (sb.x, sb.prot_mask, sb.text, sb.text_mask))
    ex = pdist.ShardedExchange(x.shape[0], 2 * sb.w2.shape[0])
    red = pkg.PeerGradAllReduce(params)  # all-bf16 eager form
    opt = pkg.FusedAdamW(params, lr=1e-4, eps=1e-6, max_grad_norm=1.0)
    for _ in range(3):
        for p in params:
            p.grad = None
        loss = pdist.distributed_contrastive_step(x, pm, ad, th, tm, exchange=ex)
        loss.backward()
        red.reduce_([p.grad for p in params])
        opt.step()
    # Stage-2 hand-off at the same shapes: residues into placeholder slots of a (B, L + 16, D_out) embedding tensor
    B, L, _ = x.shape
    embeds = torch.zeros(B, L + 16, sb.w2.shape[0], dtype=torch.bfloat16, device=dev)
    ph = torch.zeros(B, L + 16, dtype=torch.bool, device=dev)
    ph[:, 8:8 + L] = pm.bool()
    with torch.no_grad():
        handoff.adapter_into_embeds(ad.eval(), x, pm, embeds, ph)
    torch.cuda.synchronize()
    ex.check()
    print("loss", float(loss), "grad_norm", float(opt.grad_norm), "embeds norm", float(embeds.float().norm()))


if __name__ == "__main__":
    main()
