"""Multi-GPU product API on 2 GPUs over NCCL (skipped on a 1-GPU box; run with `gpurun --gpus 2`):
NN_MCMC.fit(distributed=True) -- chains sharded over ranks, running R-hat / acceptance reduced on a side stream, result
gathered on rank 0 -- must reproduce the single-process run of the same chains exactly (Philox is keyed by the global
chain index); predictive moments reduced over ranks; and the N-sharded mode (each rank a slice of the data, all-reduced
log-likelihood partial sums) must reproduce the unsharded log-posterior, gradient and chain."""
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest
import torch

from golden_util import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs')
def test_two_rank_nccl_fit_equals_single_process():
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    from dist_gpu_worker import problem
    from quinn_b200 import ops, dist
    from quinn_b200.solvers import NN_MCMC
    out = os.path.join(tempfile.mkdtemp(), 'dist.npz')
    port = 29600 + os.getpid() % 1000
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
           '--master-port', str(port), os.path.join(ROOT, 'tests', 'dist_gpu_worker.py'), out]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    got = np.load(out)
    net, x, y, th0 = problem()
    # (1) single process, all chains
    uq = NN_MCMC(net, verbose=False, dtype=torch.float64)
    res = uq.fit(x, y, zflag=False, datanoise=0.1, nmcmc=60, param_ini=th0, sampler='amcmc',
                 sampler_params={'gamma': 0.1, 'adapt': 'diag', 't0': 10, 'tadapt': 20}, seed=11, distributed=True, diag_every=20)
    np.testing.assert_array_equal(got['accepted'], res['accepted'])
    np.testing.assert_allclose(got['chain'], res['chain'], rtol=0, atol=0)
    np.testing.assert_allclose(got['logpost'], res['logpost'], rtol=0, atol=0)
    np.testing.assert_allclose(got['rhat'], [d['rhat_logpost'] for d in uq.diagnostics], rtol=1e-10)
    np.testing.assert_allclose(got['acc_hist'], [d['accept_rate'] for d in uq.diagnostics], rtol=1e-12)
    assert len(got['rhat']) == 3 and np.isfinite(got['rhat']).all()
    # (2) predictive moments over all chains
    uq2 = NN_MCMC(net, verbose=False, dtype=torch.float64)
    uq2.fit(x, y, zflag=False, datanoise=0.1, nmcmc=40, param_ini=th0, sampler='hmc', sampler_params={'epsilon': 2e-3, 'L': 3}, seed=12)
    xt = np.linspace(-1, 1, 14).reshape(7, 2)
    yens = uq2.predict_ens(xt, nens=4, nburn=8)
    np.testing.assert_allclose(got['pm'], yens.mean(0), rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(got['pv'], yens.var(0, ddof=1), rtol=1e-8, atol=1e-14)
    # (3) N-sharded data against the unsharded problem
    prob = ops.Problem(uq.desc, x, y, 0.1, dtype=torch.float64)
    lp, g = ops.logpost_grad(prob, th0)
    np.testing.assert_allclose(got['lp_sh'], lp.cpu().numpy(), rtol=1e-12)
    np.testing.assert_allclose(got['g_sh'], g.cpu().numpy(), rtol=1e-9, atol=1e-9)
    from quinn_b200.mcmc import HMC, DeviceLogPost
    sam = HMC(epsilon=2e-3, L=2)
    dl = DeviceLogPost(prob)

    class Batched:          # the unsharded log-posterior through the same generic driver (same torch generator stream)
        batched = True

        def __call__(self, th, **_):
            return ops.logpost(prob, th).clone()

        def grad(self, th, **_):
            return ops.logpost_grad(prob, th)[1].double().clone()
    b = Batched()
    Batched.grad.batched = True
    sam.setLogPost(b, b.grad)
    ref = sam.run(25, th0[:4], seed=13, verbose=False)
    np.testing.assert_array_equal(got['ns_accepted'], ref['accepted'])
    np.testing.assert_allclose(got['ns_chain'], ref['chain'], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(got['ns_logpost'], ref['logpost'], rtol=1e-9)
