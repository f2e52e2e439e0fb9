"""SURVEY.md 8f ranks 2 and 4 on the device: RNets whose weights are polynomials in depth (Lin / Quad / Cubic / Poly(n),
rnet.py:244-347) against fixtures recorded from the reference; the diagonal Fisher of NNWrap.calc_hess_diag
(nnwrap.py:204-229) against the reference's values; quantiles / get_stats (utils/stats.py:8-32) against numpy; per-chain
moments, R-hat and effective sample size; chains that stay on the device from fit() to the predictive."""
import numpy as np
import pytest
import torch

from golden_util import NET_CASES_R2, make_inputs, make_thetas, load, netdesc_from_spec
from oracle import quinn_oracle as qo

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.mark.parametrize('name', list(NET_CASES_R2))
@pytest.mark.parametrize('dtype', ['f64', 'f32'])
def test_polynomial_rnet_logpost_grad_golden(name, dtype):
    from quinn_b200 import ops
    spec = NET_CASES_R2[name]
    g = load(f'logpost_{name}.npz')
    desc = netdesc_from_spec(spec)
    assert desc.n_params == int(g['pdim'])
    x, y = make_inputs(spec)
    thetas = make_thetas(spec, desc.n_params)
    td = torch.float64 if dtype == 'f64' else torch.float32
    tol_lp, tol_g = (1e-10, 1e-9) if dtype == 'f64' else (1e-4, 2e-3)
    prob = ops.Problem(desc, x, y, spec['sigma'], dtype=td)
    lp = ops.logpost(prob, thetas).cpu().numpy()
    assert np.all(np.abs(lp - g['lp']) <= tol_lp * np.abs(g['lp'])), (lp, g['lp'])
    lp2, gr = ops.logpost_grad(prob, thetas)
    assert np.all(np.abs(lp2.cpu().numpy() - g['lp']) <= tol_lp * np.abs(g['lp']))
    gr = gr.double().cpu().numpy()
    for i in range(len(thetas)):
        assert _rel(gr[i], g['grad'][i]) <= tol_g, (i, _rel(gr[i], g['grad'][i]))
    probp = ops.Problem(desc, x, y, spec['sigma'], dtype=td, prior_sigma=float(g['sigma_prior']), prior_anchor=g['anchor'],
                        fulldatasize=int(g['nfull']))
    lpp, grp = ops.logpost_grad(probp, thetas)
    assert np.all(np.abs(lpp.cpu().numpy() - g['lp_prior']) <= tol_lp * np.abs(g['lp_prior']))
    for i in range(len(thetas)):
        assert _rel(grp[i].double().cpu().numpy(), g['grad_prior'][i]) <= tol_g
    out, _, _ = ops.predict(desc, thetas[:2], x, dtype=td)
    assert _rel(out.double().cpu().numpy(), g['pred']) <= (1e-11 if dtype == 'f64' else 2e-5)


def test_polynomial_rnet_modules_and_sampling():
    """The torch modules (Lin / Quad / Cubic / Poly) describe themselves to the kernels; a short HMC run on a Quad RNet
    through NN_MCMC replays against the oracle chain."""
    from quinn_b200.nns import RNet, Lin, Quad, Cubic, Poly
    from quinn_b200.netdesc import netdesc_from_module, flatten_module
    from quinn_b200.solvers import NN_MCMC
    from quinn_b200 import ops
    rs = np.random.RandomState(4)
    for wp, npar in ((Lin(), 2), (Quad(), 3), (Cubic(), 4), (Poly(4), 5)):
        torch.manual_seed(1)
        net = RNet(4, 2, wp_function=wp, indim=2, outdim=1, layer_pre=True, layer_post=True)
        desc = netdesc_from_module(net)
        assert desc.n_params == sum(p.numel() for p in net.parameters())
        x = rs.rand(9, 2)
        ref = net(torch.as_tensor(x)).detach().numpy()
        out, _, _ = ops.predict(desc, flatten_module(net)[None], x, dtype=torch.float64)
        np.testing.assert_allclose(out[0].cpu().numpy(), ref, rtol=1e-12, atol=1e-14)
    torch.manual_seed(2)
    net = RNet(3, 3, wp_function=Quad(), indim=1, outdim=1, layer_pre=True, layer_post=True)
    x = rs.rand(20, 1) * 2 - 1
    y = np.sin(3 * x) + 0.05 * rs.randn(20, 1)
    uq = NN_MCMC(net, verbose=False)
    th0 = 0.3 * rs.randn(uq.pdim)
    steps = 30
    mom, u = rs.randn(steps, uq.pdim), rs.rand(steps)
    res = uq.fit(x, y, zflag=False, datanoise=0.1, nmcmc=steps, param_ini=th0, sampler='hmc', sampler_params={'epsilon': 3e-3, 'L': 3},
                 replay=dict(incr=mom, unif=u))
    layers = uq.desc.as_oracle_layers()
    ref = qo.run_chain(lambda th: qo.logpost(layers, th, x, y, 0.1), th0, steps, 'hmc', dict(p=mom, u=u),
                       grad_fn=lambda th: qo.logpost_grad(layers, th, x, y, 0.1)[1], epsilon=3e-3, L=3)
    np.testing.assert_allclose(res['chain'], ref['chain'], rtol=1e-8, atol=1e-11)
    np.testing.assert_allclose(res['logpost'], ref['logpost'], rtol=1e-9)
    assert 0.1 < res['accrate'] <= 1.0


@pytest.mark.parametrize('dtype', [torch.float64, torch.float32])
def test_diag_fisher_matches_the_reference(dtype):
    from quinn_b200.nns import MLP, NNWrap, NegLogPost
    from quinn_b200.nns.tchutils import tch
    from quinn_b200.netdesc import unflatten_module
    g = load('hessdiag_mlp.npz')
    torch.manual_seed(31)
    net = MLP(2, 1, (6, 5), activ='tanh')
    if dtype == torch.float32:
        net = net.float()
    wrap = NNWrap(net)
    tol = 1e-9 if dtype == torch.float64 else 2e-4
    for tag, prior in (('noprior', None), ('prior', {'sigma': float(g['sigma_prior']), 'anchor': tch(g['anchor'])})):
        loss = NegLogPost(wrap.nnmodel, int(g['nfull']), float(g['sigma']), prior)
        f = wrap.calc_fisher_diag(g['theta'], loss, g['x'], g['y'])
        np.testing.assert_allclose(f, g['fisher_' + tag], rtol=tol)
        f2 = wrap.calc_fisher_diag(g['theta'], loss, g['x'], g['y'], chunk=5)          # chunked over the data points
        np.testing.assert_allclose(f2, f, rtol=1e-12)
    h = wrap.calc_hess_diag(g['theta'], loss, g['x'], g['y'])
    assert h.shape == (g['theta'].size, g['theta'].size) and np.count_nonzero(h - np.diag(np.diag(h))) == 0
    np.testing.assert_allclose(np.diag(h), g['fisher_prior'], rtol=tol)


@pytest.mark.parametrize('M,n', [(1, 5), (2, 7), (7, 33), (256, 1000), (1000, 129), (5000, 17)])
@pytest.mark.parametrize('dtype', [torch.float64, torch.float32])
def test_quantiles_and_get_stats_against_numpy(M, n, dtype):
    from quinn_b200 import post
    rs = np.random.RandomState(M * 7 + n)
    y = rs.randn(M, n).astype(np.float32 if dtype == torch.float32 else np.float64)
    y[:, 0] = 3.0                                                  # a constant column
    if M > 4:
        y[:3, 1] = y[3, 1]                                         # ties
    yt = torch.as_tensor(y, device='cuda')
    qs = (0.0, 0.05, 0.25, 0.5, 0.75, 0.9, 1.0)
    got = post.quantiles(yt, qs).cpu().numpy()
    want = np.quantile(y.astype(np.float64), qs, axis=0)
    np.testing.assert_allclose(got, want, rtol=1e-12 if dtype == torch.float64 else 1e-6, atol=1e-12 if dtype == torch.float64 else 1e-6)
    # get_stats (utils/stats.py:8-32), on a (M, N, o) array
    y3 = yt.reshape(M, n, 1)
    mb, lb, ub = post.get_stats(y3, True)
    q = np.quantile(y.astype(np.float64), [0.25, 0.5, 0.75], axis=0)
    tol = 1e-12 if dtype == torch.float64 else 1e-5
    np.testing.assert_allclose(mb.cpu().numpy()[:, 0], q[1], rtol=tol, atol=tol)
    np.testing.assert_allclose(lb.cpu().numpy()[:, 0], q[1] - q[0], rtol=tol, atol=tol)
    np.testing.assert_allclose(ub.cpu().numpy()[:, 0], q[2] - q[1], rtol=tol, atol=tol)
    mean, sd, sd2 = post.get_stats(y3, False)
    np.testing.assert_allclose(mean.cpu().numpy()[:, 0], y.astype(np.float64).mean(0), rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(sd.cpu().numpy()[:, 0], y.astype(np.float64).std(0), rtol=1e-5, atol=1e-6)


def _ess_numpy(x, max_lag=None):
    n = len(x)
    d = x - x.mean()
    c0 = np.sum(d * d)
    max_lag = n - 1 if max_lag is None else max_lag
    if max_lag % 2 == 0:
        max_lag -= 1
    tau, prev = -1.0, 1.0
    for t in range(1, max_lag + 1):
        rho = np.sum(d[:n - t] * d[t:]) / c0
        if t % 2 == 1:
            pair = prev + rho
            if not pair > 0:
                break
            tau += 2 * pair
        else:
            prev = rho
    return n / tau


def test_row_moments_rhat_and_ess():
    from quinn_b200 import post
    rs = np.random.RandomState(9)
    K, n, phi = 12, 4000, 0.8
    x = np.zeros((K, n))
    e = rs.randn(K, n)
    for t in range(1, n):
        x[:, t] = phi * x[:, t - 1] + e[:, t]                      # AR(1): ESS ~ n (1 - phi) / (1 + phi)
    x += rs.randn(K, 1) * 0.1
    m, v = post.row_moments(x)
    np.testing.assert_allclose(m.cpu().numpy(), x.mean(1), rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(v.cpu().numpy(), x.var(1, ddof=1), rtol=1e-10)
    r = post.rhat(x).item()
    W = x.var(1, ddof=1).mean()
    want = np.sqrt(((n - 1.0) / n * W + x.mean(1).var(ddof=1)) / W)
    assert abs(r - want) <= 1e-10 * want
    ess = post.ess(x, max_lag=400).cpu().numpy()
    for k in (0, K - 1):
        assert abs(ess[k] - _ess_numpy(x[k], 400)) <= 1e-8 * ess[k]
    theory = n * (1 - phi) / (1 + phi)
    assert 0.5 * theory < np.median(ess) < 2.0 * theory
    white = post.ess(rs.randn(3, 2000)).cpu().numpy()
    assert (white > 1000).all()


def test_chain_stays_on_the_device_from_fit_to_predictive():
    """fit(keep_on_device=True): samples / MAP are CUDA tensors; thinning (nn_mcmc.py:194-196), the ensemble forward, its
    moments, get_stats and the diagnostics all run without the chain visiting the host; results equal the host path."""
    from quinn_b200.nns import MLP
    from quinn_b200.solvers import NN_MCMC
    np.random.seed(2)
    torch.manual_seed(2)
    net = MLP(1, 1, (8,), activ='tanh')
    x = np.random.rand(25, 1) * 2 - 1
    y = np.sin(3 * x) + 0.1 * np.random.randn(25, 1)
    th0 = 0.3 * np.random.randn(6, sum(p.numel() for p in net.parameters()))
    kw = dict(zflag=False, datanoise=0.1, nmcmc=1400, param_ini=th0, sampler='amcmc', sampler_params={'gamma': 0.1}, seed=3)
    a = NN_MCMC(net, verbose=False)
    a.fit(x, y, **kw)
    b = NN_MCMC(net, verbose=False)
    b.fit(x, y, keep_on_device=True, **kw)
    assert torch.is_tensor(b.samples) and b.samples.is_cuda and torch.is_tensor(b.cmode) and b.cmode.is_cuda
    np.testing.assert_array_equal(b.samples.cpu().numpy(), a.samples)
    xt = np.linspace(-1, 1, 7)[:, None]
    np.testing.assert_allclose(b.predict_ens(xt, nens=5, nburn=50), a.predict_ens(xt, nens=5, nburn=50), rtol=0, atol=0)
    np.testing.assert_allclose(b.predict_MAP(xt), a.predict_MAP(xt), rtol=0, atol=0)
    mb, lb, ub = b.predict_stats(xt, nsam=8, qt=True)              # thinned with the reference's nburn = 1000 rule (quinn.py:84)
    ye = a.predict_ens(xt, nens=8, nburn=1000)
    q = np.quantile(ye, [0.25, 0.5, 0.75], axis=0)
    np.testing.assert_allclose(mb, q[1], rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(lb, q[1] - q[0], rtol=1e-10, atol=1e-13)
    np.testing.assert_allclose(ub, q[2] - q[1], rtol=1e-10, atol=1e-13)
    mean, sd, _ = b.predict_stats(xt, nsam=8, qt=False)
    np.testing.assert_allclose(mean, ye.mean(0), rtol=1e-10, atol=1e-13)
    np.testing.assert_allclose(sd, ye.std(0), rtol=1e-8, atol=1e-12)
    d = b.diagnose()
    assert np.isfinite(d['rhat_logpost']) and d['ess_logpost'].shape == (6,) and (d['ess_logpost'] > 0).all()
    da = a.diagnose()
    assert abs(da['rhat_logpost'] - d['rhat_logpost']) <= 1e-12 * abs(d['rhat_logpost'])
