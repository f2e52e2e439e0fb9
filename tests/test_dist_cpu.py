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
    dist.barrier()
    q.put((rank, r.numpy(), m.numpy(), v.numpy(), None if g is None else g.numpy(), mx))
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
    for rank, r, m, v, g, mx in res:
        np.testing.assert_allclose(r, rhat, rtol=1e-12)
        np.testing.assert_allclose(m, y.mean(0), rtol=1e-12, atol=1e-15)
        np.testing.assert_allclose(v, y.var(0, ddof=1), rtol=1e-10)
        assert mx == 2.0
        if rank == 0:
            np.testing.assert_allclose(g, draws[:, 0, :])
        else:
            assert g is None
