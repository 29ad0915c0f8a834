"""CPU, world_size 2 over gloo: the sharded-batch logic of prot2text-v2-esm3_b200/dist.py.

The kernels cannot run here, so the per-rank block computation is done with the oracle; what is
under test is the exchange logic (all-gather of embeddings, merge of column statistics,
reduce-scatter of the gathered embeddings' gradient) and the claim that local-mean losses whose
gradients are averaged across ranks reproduce the single-process global-batch loss."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, results):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import importlib
    import __graft_entry__ as entry
    entry.load_package()
    pdist = importlib.import_module("p2t_b200.dist")
    from oracle import restatement as R
    torch.manual_seed(0)
    B, E, tau = 5, 24, 0.05
    P = torch.nn.functional.normalize(torch.randn(world * B, E, dtype=torch.float64), dim=-1)
    T = torch.nn.functional.normalize(torch.randn(world * B, E, dtype=torch.float64), dim=-1)
    p_loc = P[rank * B:(rank + 1) * B].clone().requires_grad_()
    t_loc = T[rank * B:(rank + 1) * B].clone()
    t_glob = pdist.all_gather_embeddings(t_loc)
    assert torch.equal(t_glob, T)
    labels = torch.arange(rank * B, (rank + 1) * B)
    # row term: fully local
    loss_row = R.infonce_rows(p_loc, t_glob, labels, tau)
    # column term: local stats merged across ranks
    s = R.similarity(p_loc.detach(), t_glob, tau)
    m_loc = s.max(dim=0).values
    sum_loc = torch.exp(s - m_loc).sum(dim=0)
    m, ssum = pdist.merge_column_stats(m_loc, sum_loc)
    lse_col = m + torch.log(ssum)
    full_lse_col = torch.logsumexp(R.similarity(P, T, tau), dim=0)
    torch.testing.assert_close(lse_col, full_lse_col)
    # north_star reduce-scatter of d(loss)/d(gathered t): each rank contributes dS_k^T p_k / tau
    ds, dp, dt_full = R.infonce_backward(p_loc.detach(), t_glob, labels, tau)
    dt_mine = pdist.reduce_scatter_text_grad(dt_full / world)
    # mean over ranks of local losses == global loss; averaged grads == global grads
    loss_t = loss_row.detach().clone()
    dist.all_reduce(loss_t)
    (g_loc,) = torch.autograd.grad(loss_row, p_loc)
    if rank == 0:
        Pg = P.clone().requires_grad_()
        Tg = T.clone().requires_grad_()
        glob = R.infonce_rows(Pg, Tg, torch.arange(world * B), tau)
        gP, gT = torch.autograd.grad(glob, (Pg, Tg))
        results["loss_ok"] = bool(torch.allclose(loss_t / world, glob.detach()))
        results["gp_ok"] = bool(torch.allclose(g_loc / world, gP[:B]))
        results["gt_ok"] = bool(torch.allclose(dt_mine, gT[:B]))
    dist.destroy_process_group()


def test_sharded_batch_matches_global_batch():
    mgr = mp.Manager()
    results = mgr.dict()
    port = 29600 + os.getpid() % 300
    mp.spawn(_worker, args=(2, port, results), nprocs=2, join=True)
    assert results["loss_ok"] and results["gp_ok"] and results["gt_ok"], dict(results)
