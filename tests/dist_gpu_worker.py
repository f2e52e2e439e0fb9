"""Worker of tests/test_gpu_dist.py, launched with torchrun on 2 GPUs (one process per GPU, NCCL):
   python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port P tests/dist_gpu_worker.py OUT.npz
Runs NN_MCMC.fit(distributed=True) and the N-sharded mode and writes what rank 0 has to OUT.npz; the parent test
compares with the same runs done in ONE process."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))


def problem():
    from quinn_b200.nns import MLP
    np.random.seed(3)
    torch.manual_seed(3)
    net = MLP(2, 1, (16, 16), activ='tanh')
    x = np.random.rand(200, 2) * 2 - 1
    y = np.sin(2 * x[:, :1]) * x[:, 1:] + 0.05 * np.random.randn(200, 1)
    th0 = 0.3 * np.random.randn(10, sum(p.numel() for p in net.parameters()))
    return net, x, y, th0


def main():
    from quinn_b200 import dist
    from quinn_b200.solvers import NN_MCMC
    rank, world, local = dist.init()
    net, x, y, th0 = problem()
    out = {}
    # (1) chains sharded over the ranks, diagnostics every 20 steps, gathered on rank 0
    uq = NN_MCMC(net, verbose=False, dtype=torch.float64, device=f'cuda:{local}')
    res = uq.fit(x, y, zflag=False, datanoise=0.1, nmcmc=60, param_ini=th0, sampler='amcmc', sampler_params={'gamma': 0.1, 'adapt': 'diag', 't0': 10, 'tadapt': 20},
                 seed=11, distributed=True, diag_every=20)
    if rank == 0:
        out.update(chain=res['chain'], logpost=res['logpost'], accepted=res['accepted'], accrate=res['accrate'],
                   rhat=np.array([d['rhat_logpost'] for d in uq.diagnostics]), acc_hist=np.array([d['accept_rate'] for d in uq.diagnostics]))
    # (2) same with HMC and the predictive moments reduced over the ranks (no gather)
    uq2 = NN_MCMC(net, verbose=False, dtype=torch.float64, device=f'cuda:{local}')
    uq2.fit(x, y, zflag=False, datanoise=0.1, nmcmc=40, param_ini=th0, sampler='hmc', sampler_params={'epsilon': 2e-3, 'L': 3}, seed=12,
            distributed=True, gather=False)
    xt = np.linspace(-1, 1, 14).reshape(7, 2)
    m, v = uq2.predict_moments_distributed(xt, nens=4, nburn=8)
    if rank == 0:
        out.update(pm=m, pv=v)
    # (3) N-sharded data: this rank holds half of the points, all chains
    lo, hi = dist.shard_range(x.shape[0], rank, world)
    uq3 = NN_MCMC(net, verbose=False, dtype=torch.float64, device=f'cuda:{local}')
    res3 = uq3.fit(x[lo:hi], y[lo:hi], datanoise=0.1, nmcmc=25, param_ini=th0[:4], sampler='hmc', sampler_params={'epsilon': 2e-3, 'L': 2}, seed=13,
                   data_sharded=True)
    from quinn_b200.mcmc.mcmc import ShardedDataLogPost
    from quinn_b200 import ops
    prob = ops.Problem(uq3.desc, x[lo:hi], y[lo:hi], 0.1, dtype=torch.float64, device=f'cuda:{local}')
    slp = ShardedDataLogPost(prob, x.shape[0])
    lp_sh = slp(th0).cpu().numpy()
    g_sh = slp.grad(th0).cpu().numpy()
    if rank == 0:
        out.update(ns_chain=res3['chain'], ns_logpost=res3['logpost'], ns_accepted=res3['accepted'], lp_sh=lp_sh, g_sh=g_sh)
        np.savez(sys.argv[1], **out)
    dist.barrier()
    import torch.distributed as td
    td.destroy_process_group()


if __name__ == '__main__':
    main()
