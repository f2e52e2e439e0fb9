"""BASELINE.json-size checks through size-independent properties (the oracle is too slow at these sizes):
fp32 vs fp64 kernels, invariance under data permutation / chain sharding, fused moments vs explicit moments,
energy conservation of the leapfrog, finite differences of the ELBO.  A few oracle spot checks anchor them."""
import math

import numpy as np
import pytest
import torch

from golden_util import netdesc_from_layers
from oracle import quinn_oracle as qo

pytestmark = pytest.mark.gpu


def _net(d, hls):
    layers, P = qo.mlp_layers(d, 1, hls, True, 'tanh')
    return layers, netdesc_from_layers(layers, P)


def test_config5_logpost_fp32_vs_fp64_vs_oracle_and_permutation():
    from quinn_b200 import ops
    rs = np.random.RandomState(0)
    layers, desc = _net(3, (64, 64))
    N, K = 10_000, 96
    x = rs.rand(N, 3) * 2 * math.pi - math.pi
    y = np.sin(x).sum(1, keepdims=True) + 0.05 * rs.randn(N, 1)
    th = rs.rand(K, desc.n_params)                                   # the chains' starting points (nn_mcmc.py:124)
    th[K // 2:] = 0.2 * rs.randn(K - K // 2, desc.n_params)            # and some near-zero states
    p64 = ops.Problem(desc, x, y, 0.05, dtype=torch.float64)
    p32 = ops.Problem(desc, x, y, 0.05, dtype=torch.float32)
    lp64 = ops.logpost(p64, th).cpu().numpy()
    lp32 = ops.logpost(p32, th).cpu().numpy()
    assert np.all(np.abs(lp32 - lp64) <= 1e-4 * np.abs(lp64)), np.abs(lp32 / lp64 - 1).max()
    for k in (0, K - 1):
        ref = qo.logpost(layers, th[k], x, y, 0.05)
        assert abs(lp64[k] - ref) <= 1e-10 * abs(ref)
    perm = rs.permutation(N)
    lpp = ops.logpost(ops.Problem(desc, x[perm], y[perm], 0.05, dtype=torch.float64), th).cpu().numpy()
    np.testing.assert_allclose(lpp, lp64, rtol=1e-11)
    # gradient: fp32 vs fp64 at full size, fp64 vs oracle for one state
    l64, g64 = ops.logpost_grad(p64, th[K // 2:K // 2 + 8])
    l32, g32 = ops.logpost_grad(p32, th[K // 2:K // 2 + 8])
    g64, g32 = g64.cpu().numpy(), g32.double().cpu().numpy()
    for k in range(8):
        assert np.abs(g32[k] - g64[k]).max() <= 2e-3 * np.abs(g64[k]).max()
    _, gref = qo.logpost_grad(layers, th[K // 2], x, y, 0.05)
    assert np.abs(g64[0] - gref).max() <= 1e-9 * np.abs(gref).max()


def test_config5_chains_do_not_depend_on_batching_or_sharding():
    """The same chains run as one launch of 300, or as 3 shards with chain_offset, give identical states."""
    from quinn_b200 import ops
    rs = np.random.RandomState(1)
    layers, desc = _net(3, (64, 64))
    N, K, steps = 10_000, 300, 4
    x = rs.rand(N, 3) * 2 * math.pi - math.pi
    y = np.sin(x).sum(1, keepdims=True) + 0.05 * rs.randn(N, 1)
    th0 = rs.rand(K, desc.n_params).astype(np.float32)
    prob = ops.Problem(desc, x, y, 0.05, dtype=torch.float32)

    def run(lo, hi):
        st = ops.ChainState(prob, th0[lo:hi])
        am = ops.AmcmcState(st, gamma=0.01, adapt='diag')
        ops.amcmc_run(st, am, steps, None, seed=77, chain_offset=lo)
        return st.theta.cpu().numpy(), st.lp.cpu().numpy(), st.naccept.cpu().numpy()
    full = run(0, K)
    parts = [run(0, 100), run(100, 200), run(200, 300)]
    np.testing.assert_array_equal(full[0], np.concatenate([p[0] for p in parts]))
    np.testing.assert_array_equal(full[1], np.concatenate([p[1] for p in parts]))
    assert 0 < full[2].sum() < K * steps
    # the recorded log-posterior of the final states equals a fresh evaluation (kernel 3 == kernel 1)
    np.testing.assert_allclose(ops.logpost(prob, full[0]).cpu().numpy(), full[1], rtol=1e-6)


def test_config3_fused_moments_equal_moments_of_the_full_array():
    from quinn_b200 import ops
    rs = np.random.RandomState(2)
    layers, desc = _net(10, (128, 128))
    M, N = 256, 100_000
    th = ((2 * rs.rand(M, desc.n_params) - 1) / math.sqrt(10)).astype(np.float32)
    x = rs.rand(N, 10).astype(np.float32)
    out, mean, var = ops.predict(desc, th, x, dtype=torch.float32, want_out=True, want_moments=True)
    o64 = out.double()
    np.testing.assert_allclose(mean.double().cpu().numpy(), o64.mean(0).cpu().numpy(), rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(var.double().cpu().numpy(), o64.var(0, unbiased=True).cpu().numpy(), rtol=2e-3, atol=1e-8)
    _, mean2, var2 = ops.predict(desc, th, x, dtype=torch.float32, want_out=False, want_moments=True)
    np.testing.assert_allclose(mean2.cpu().numpy(), mean.cpu().numpy(), rtol=1e-6, atol=1e-7)
    sel = rs.randint(0, N, 64)
    ref = qo.predict_ens(layers, th[:3].astype(np.float64), x[sel].astype(np.float64))
    np.testing.assert_allclose(out[:3].double().cpu().numpy()[:, sel], ref, rtol=1e-4, atol=2e-5)


@pytest.mark.parametrize('method', ['hmc', 'mala'])
def test_config2_leapfrog_conserves_energy_for_small_steps(method):
    """1,024 chains, MLP 2-32-32-1, N=1000: with a tiny step the Hamiltonian error is O(eps^2), so every proposal is
    accepted and the recorded MH ratios are ~1; with a huge step they are not."""
    from quinn_b200 import ops
    rs = np.random.RandomState(3)
    layers, desc = _net(2, (32, 32))
    N, K = 1_000, 1_024
    x = rs.rand(N, 2) * 3 - 1.5
    y = np.cos(2 * x[:, :1]) + np.sin(2 * x[:, 1:]) + 0.02 * rs.randn(N, 1)
    th0 = 0.1 * rs.randn(K, desc.n_params)
    prob = ops.Problem(desc, x, y, 0.02, dtype=torch.float64)
    st = ops.ChainState(prob, th0)
    hm = ops.HmcState(st, epsilon=1e-7, L=3, method=method)
    rec = ops.Recorder(st, 5, store_every=5)
    ops.hmc_run(st, hm, 5, rec, seed=5)
    al = rec.alpha.cpu().numpy()
    assert np.all(rec.accepted.cpu().numpy() == 1) or np.mean(rec.accepted.cpu().numpy()) > 0.97
    assert np.all(np.abs(np.log(al)) < 0.1), np.abs(np.log(al)).max()
    lp, g = ops.logpost_grad(prob, st.theta)
    np.testing.assert_allclose(lp.cpu().numpy(), st.lp.cpu().numpy(), rtol=1e-12)
    np.testing.assert_allclose(g.cpu().numpy(), hm.grad_cur.cpu().numpy(), rtol=1e-9, atol=1e-6)   # cached gradient is current
    for k in (0, K - 1):
        rl, rg = qo.logpost_grad(layers, st.theta[k].cpu().numpy(), x, y, 0.02)
        assert abs(lp[k].item() - rl) <= 1e-10 * abs(rl)
        assert np.abs(g[k].cpu().numpy() - rg).max() <= 1e-9 * np.abs(rg).max()


def test_config4_elbo_gradient_matches_finite_differences():
    """NN_VI shape (MLP 10-128-128-1, 16 MC samples, N=20,000): d viloss / d mu, d rho against central differences
    with the same eps draws (the ELBO is deterministic given eps)."""
    from quinn_b200.nns import MLP
    from quinn_b200.vi import BNet
    torch.manual_seed(0)
    rs = np.random.RandomState(4)
    net = MLP(10, 1, (128, 128), activ='tanh')
    b = BNet(net, pi=0.5, sigma1=1.0, sigma2=0.5, seed=9)
    N, nsam = 20_000, 16
    x = torch.as_tensor(rs.rand(N, 10), device='cuda')
    y = torch.as_tensor(np.sin(rs.rand(N, 1) * 3), device='cuda')
    b.loss_params = [0.05, nsam, 1]
    eps = rs.randn(nsam, b.desc.n_params)
    loss = b.viloss(x, y, eps=eps)
    loss.backward()
    for tensor_idx, coord in [(0, (3, 2)), (2, (17, 100)), (4, (0, 5)), (1, (7,))]:
        for which in (0, 1):                                   # mu, rho
            par = b.params[2 * tensor_idx + which]
            g = par.grad[coord].item()
            h = 1e-5 if which == 0 else 1e-3
            with torch.no_grad():
                old = par[coord].item()
                par[coord] = old + h
                lp_ = b.viloss(x, y, eps=eps).item()
                par[coord] = old - h
                lm_ = b.viloss(x, y, eps=eps).item()
                par[coord] = old
            fd = (lp_ - lm_) / (2 * h)
            assert abs(fd - g) <= 2e-4 * max(abs(g), abs(fd)) + 1e-6 * abs(loss.item()) / max(h, 1e-12) * 1e-9, (tensor_idx, which, fd, g)


def test_degenerate_sizes_are_refused_or_handled():
    from quinn_b200 import ops
    layers, desc = _net(2, (4,))
    with pytest.raises((RuntimeError, ValueError)):
        p = ops.Problem(desc, np.zeros((0, 2)), np.zeros((0, 1)), 0.1, dtype=torch.float64)
        ops.logpost(p, np.zeros((1, desc.n_params)))
    with pytest.raises(ValueError):
        ops.Problem(desc, np.zeros((3, 5)), np.zeros((3, 1)), 0.1)
    p = ops.Problem(desc, np.ones((1, 2)), np.ones((1, 1)), 0.1, dtype=torch.float64)     # a single point, single chain
    with pytest.raises(ValueError):
        ops.logpost(p, np.zeros((1, desc.n_params + 1)))
    lp = ops.logpost(p, np.zeros((1, desc.n_params)))
    assert abs(lp.item() - qo.logpost(layers, np.zeros(desc.n_params), np.ones((1, 2)), np.ones((1, 1)), 0.1)) < 1e-12
    th_nan = np.zeros((2, desc.n_params))
    th_nan[1, 0] = np.nan                                        # NaN state: lp is NaN and every proposal from it rejects
    assert np.isnan(ops.logpost(p, th_nan)[1].item())


@pytest.mark.parametrize('shape', ['c5_hot', 'c3_wide'])
def test_tensor_core_value_path_against_the_oracle_at_bench_size(shape):
    """Direct fp64-oracle comparison of the tensor-core value kernels at N >= 10^4: the 79-tile loop (odd count, ragged
    last tile) of the half-K pipeline of the config-5 shape, and the 128-wide (4 column groups) pipe of the config-3/4
    net at 80 tiles + 1 point, in the log-posterior kernel, the fused chain kernel and the predictive kernel."""
    from quinn_b200 import ops
    if shape == 'c5_hot':
        d, hls, N, sigma, tscale = 3, (64, 64), 10_000, 0.05, None
    else:
        d, hls, N, sigma, tscale = 10, (128, 128), 10_241, 0.05, 0.1
    rs = np.random.RandomState(len(shape))
    layers, desc = _net(d, hls)
    x = (rs.rand(N, d) * 2 * math.pi - math.pi).astype(np.float32).astype(np.float64)
    y = (np.sin(x).sum(1, keepdims=True) + sigma * rs.randn(N, 1)).astype(np.float32).astype(np.float64)
    K = 160
    th = (rs.rand(K, desc.n_params) if tscale is None else tscale * rs.randn(K, desc.n_params)).astype(np.float32).astype(np.float64)
    prob = ops.Problem(desc, x, y, sigma, dtype=torch.float32)
    assert prob.plan_info(K)['tensor_core'] == 2
    lp = ops.logpost(prob, th).cpu().numpy()
    picks = (0, K // 2, K - 1)
    refs = {k: qo.logpost(layers, th[k], x, y, sigma) for k in picks}
    for k in picks:
        assert abs(lp[k] - refs[k]) <= 1e-5 * abs(refs[k]), (k, lp[k], refs[k])
    # fused chain kernel: the log-posterior it records for the initial state and for an accepted step
    st = ops.ChainState(prob, th)
    rec = ops.Recorder(st, 2, store_every=1)
    ops.amcmc_run(st, ops.AmcmcState(st, gamma=0.01), 2, rec, seed=4)
    lp0 = rec.logpost0.cpu().numpy()
    for k in picks:
        assert abs(lp0[k] - refs[k]) <= 1e-5 * abs(refs[k])
        s1 = rec.samples[k, 1].double().cpu().numpy()
        r1 = qo.logpost(layers, s1, x, y, sigma)
        assert abs(rec.logpost[k, 1].item() - r1) <= 1e-5 * abs(r1)
    # predictive kernel
    out, _, _ = ops.predict(desc, th[:3], x, dtype=torch.float32)
    for k in range(3):
        ref = qo.forward(layers, th[k], x)
        assert np.abs(out[k].double().cpu().numpy() - ref).max() <= 2e-5 * max(1.0, np.abs(ref).max())
