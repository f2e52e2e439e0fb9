"""World-size-2 gloo tests of the multi-GPU host logic (sharding, diagnostic reductions, gathers)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as td
import torch.multiprocessing as mp

from golden_util import ROOT


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from quinn_b200 import dist
    dist.init(backend='gloo')
    rs = np.random.RandomState(0)
    K, n = 11, 50
    draws = rs.randn(K, n, 3) + rs.randn(K, 1, 3) * 0.3       # [chain, step, monitored scalar]
    lo, hi = dist.shard_range(K, rank, world)
    mine = torch.as_tensor(draws[lo:hi])
    r = dist.rhat(mine.mean(1), mine.var(1, unbiased=True), n)
    y = torch.as_tensor(rs.randn(K, 6))
    m, v = dist.reduce_predictive_moments(y[lo:hi].sum(0), (y[lo:hi] ** 2).sum(0), hi - lo)
    g = dist.gather_to_rank0(mine[:, 0, :])
    mx = dist.max_over_ranks(float(rank + 1))
    # running diagnostics hook (CPU tensors: no side stream) and the sharded-data sum
    class Rec:
        pass
    rec = Rec()
    rec.logpost = mine[:, :, 0]
    rec.accepted = (mine[:, :, 1] > 0).to(torch.uint8)
    diag = dist.RunningDiagnostics(None)
    diag(None, rec, n)
    hist = diag.finish()
    part = torch.as_tensor(draws[:, lo * 4:hi * 4 if rank == world - 1 else hi * 4, 0].sum(1))     # per-rank partial sums over "points"
    tot = dist.allreduce_sum_(part.clone())
    b = dist.broadcast_from_rank0(torch.full((3,), float(rank + 5), dtype=torch.float64))
    dist.barrier()
    q.put((rank, r.numpy(), m.numpy(), v.numpy(), None if g is None else g.numpy(), mx, hist, tot.numpy(), b.numpy()))
    td.destroy_process_group()


def test_two_rank_reductions_match_single_process():
    world, port = 2, 29500 + os.getpid() % 2000
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rs = np.random.RandomState(0)
    K, n = 11, 50
    draws = rs.randn(K, n, 3) + rs.randn(K, 1, 3) * 0.3
    cm, cv = draws.mean(1), draws.var(1, ddof=1)
    W = cv.mean(0)
    var_plus = (n - 1.0) / n * W + cm.var(0, ddof=1)
    rhat = np.sqrt(var_plus / W)
    y = rs.randn(K, 6)
    lo0, hi0 = 0, 6
    for rank, r, m, v, g, mx, hist, tot, b in res:
        assert len(hist) == 1 and hist[0]['step'] == n
        np.testing.assert_allclose(hist[0]['rhat_logpost'], rhat[0], rtol=1e-12)
        np.testing.assert_allclose(hist[0]['accept_rate'], (draws[:, :, 1] > 0).mean(), rtol=1e-12)
        np.testing.assert_allclose(tot, draws[:, :44, 0].sum(1), rtol=1e-12)      # ranks own point slices [0,24) and [24,44)
        np.testing.assert_array_equal(b, np.full(3, 5.0))
        np.testing.assert_allclose(r, rhat, rtol=1e-12)
        np.testing.assert_allclose(m, y.mean(0), rtol=1e-12, atol=1e-15)
        np.testing.assert_allclose(v, y.var(0, ddof=1), rtol=1e-10)
        assert mx == 2.0
        if rank == 0:
            np.testing.assert_allclose(g, draws[:, 0, :])
        else:
            assert g is None
