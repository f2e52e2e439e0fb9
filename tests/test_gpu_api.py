"""The reference-facing Python API on the GPU: same calls and assertions as the reference's own tests for the
hot-path rows (tests/test_mcmc.py, test_nnwrap.py, test_solvers.py, test_ensemble.py, test_vi.py there), plus
parity with the golden fixtures through that API."""
import numpy as np
import pytest
import torch

from golden_util import load, NET_CASES, make_inputs, make_thetas
from oracle import quinn_oracle as qo

pytestmark = pytest.mark.gpu


def make_gaussian_logpost(mean, cov):
    cov_inv = np.linalg.inv(cov)
    return (lambda x: -0.5 * (x - mean) @ cov_inv @ (x - mean)), (lambda x: -cov_inv @ (x - mean))


# ---------------------------------------------------------------- samplers on a generic callable (test_mcmc.py)
def test_amcmc_generic_callable_gaussian():
    from quinn_b200.mcmc import AMCMC
    np.random.seed(42)
    mean, cov = np.array([1.0, 2.0]), np.array([[1.0, 0.3], [0.3, 1.0]])
    lp, _ = make_gaussian_logpost(mean, cov)
    s = AMCMC(gamma=0.5, t0=50, tadapt=100)
    assert (s.gamma, s.t0, s.tadapt) == (0.5, 50, 100)
    s.setLogPost(lp, None)
    res = s.run(1500, np.zeros(2), verbose=False)
    for key in ('chain', 'mapparams', 'maxpost', 'accrate', 'logpost', 'alphas'):
        assert key in res
    assert res['chain'].shape == (1501, 2) and res['logpost'].shape == (1501,) and res['alphas'][0] == 0.0
    assert np.allclose(res['mapparams'], mean, atol=0.5)
    assert 0.05 < res['accrate'] < 0.95
    assert res['maxpost'] >= res['logpost'].max() - 1e-12


@pytest.mark.parametrize('cls,kw', [('HMC', dict(epsilon=0.1, L=5)), ('MALA', dict(epsilon=0.3))])
def test_gradient_samplers_generic_callable(cls, kw):
    import quinn_b200.mcmc as M
    np.random.seed(42)
    mean, cov = np.array([1.0, -1.0]), np.eye(2)
    lp, lg = make_gaussian_logpost(mean, cov)
    s = getattr(M, cls)(**kw)
    s.setLogPost(lp, lg)
    res = s.run(800, np.zeros(2), verbose=False)
    assert res['chain'].shape == (801, 2)
    assert np.allclose(res['chain'][200:].mean(0), mean, atol=0.5)
    assert res['accrate'] > 0.3


def test_defaults_match_reference():
    from quinn_b200.mcmc import AMCMC, HMC, MALA
    a, h, m = AMCMC(), HMC(), MALA()
    assert (a.gamma, a.t0, a.tadapt, a.cov_ini) == (0.1, 100, 1000, None)
    assert (h.epsilon, h.L, m.epsilon) == (0.05, 3, 0.05)
    with pytest.raises(AssertionError):
        a.run(10, np.zeros(2))


# ---------------------------------------------------------------- NNWrap (test_nnwrap.py)
def test_nnwrap_calls():
    from quinn_b200.nns import MLP, NNWrap, NegLogPost, nn_p, nnwrapper
    torch.manual_seed(0)
    net = MLP(2, 1, (5,), activ='tanh')
    wrap = NNWrap(net)
    flat = wrap.p_flatten().detach().numpy().flatten()
    assert len(flat) == net.numpar() == 21
    x, y = np.random.rand(10, 2), np.random.rand(10, 1)
    out = wrap(x)
    assert isinstance(out, np.ndarray) and out.shape == (10, 1)
    np.testing.assert_allclose(out, net(torch.as_tensor(x)).detach().numpy(), rtol=1e-12, atol=1e-14)
    wrap.p_unflatten(flat)
    np.testing.assert_allclose(wrap(x), out)
    assert wrap.predict(x, flat).shape == (10, 1)
    loss = NegLogPost(net, 10, 0.1, None)
    val = wrap.calc_loss(flat, loss, x, y)
    assert isinstance(val, float)
    ref = loss(torch.as_tensor(x), torch.as_tensor(y)).item()
    assert abs(val - ref) <= 1e-10 * abs(ref)
    g = wrap.calc_lossgrad(flat, loss, x, y)
    assert g.shape == flat.shape
    ref_l = loss(torch.as_tensor(x), torch.as_tensor(y))
    ref_l.backward()
    ref_g = np.concatenate([p.grad.numpy().ravel() for p in net.parameters()])
    np.testing.assert_allclose(g, ref_g, rtol=1e-9, atol=1e-9 * np.abs(ref_g).max())
    y1, y2 = nn_p(flat, x, net), nn_p(flat + 0.1, x, net)
    assert y1.shape == (10, 1) and not np.allclose(y1, y2)
    assert nnwrapper(x, net).shape == (10, 1)
    with pytest.raises(NotImplementedError):
        wrap.calc_loss(flat, torch.nn.MSELoss(), x, y)


def test_nnwrap_golden_with_prior():
    """NNWrap.calc_loss / calc_lossgrad with NegLogPost + prior reproduce the reference's numbers."""
    from quinn_b200.nns import MLP, NNWrap, NegLogPost, tch
    spec = NET_CASES['mlp_tanh_o2']
    g = load('logpost_mlp_tanh_o2.npz')
    net = MLP(spec['indim'], spec['outdim'], spec['hls'], activ=spec['activ'])
    x, y = make_inputs(spec)
    th = make_thetas(spec, net.numpar())
    wrap = NNWrap(net)
    loss = NegLogPost(net, int(g['nfull']), spec['sigma'], {'sigma': float(g['sigma_prior']), 'anchor': tch(g['anchor'])})
    for i in range(3):
        assert abs(-wrap.calc_loss(th[i], loss, x, y) - g['lp_prior'][i]) <= 1e-10 * abs(g['lp_prior'][i])
        gr = -wrap.calc_lossgrad(th[i], loss, x, y)
        assert np.abs(gr - g['grad_prior'][i]).max() <= 1e-9 * np.abs(g['grad_prior'][i]).max()


# ---------------------------------------------------------------- NN_MCMC (test_solvers.py + golden replay)
def _c1_net():
    from quinn_b200.nns import RNet, Poly
    return RNet(3, 3, wp_function=Poly(0), indim=1, outdim=1, layer_pre=True, layer_post=True, biasorno=True,
                nonlin=True, mlp=False, final_layer=None)


@pytest.mark.parametrize('fname,sampler,sp', [
    ('chain_c1_amcmc.npz', 'amcmc', {'gamma': 0.01}),
    ('chain_c1_hmc.npz', 'hmc', {'L': 3, 'epsilon': 0.0025})])
def test_nn_mcmc_fit_replays_reference_config1(fname, sampler, sp):
    """examples/ex_ufit.py (config 1) through NN_MCMC.fit with the reference's recorded draws: same
    accept/reject sequence, chain, MAP and predictive ensemble for the first 1,000 steps."""
    from quinn_b200.solvers import NN_MCMC
    g = load(fname)
    uq = NN_MCMC(_c1_net(), verbose=False)
    assert uq.pdim == 22
    n = len(g['u'])
    incr = g['xi'] if sampler == 'amcmc' else g['p']
    res = uq.fit(g['x'], g['y'], zflag=False, datanoise=float(g['sigma']), nmcmc=n, param_ini=g['theta0'], sampler=sampler,
                 sampler_params=sp, replay=dict(incr=incr, unif=g['u']))
    assert uq.samples.shape == (n + 1, 22)
    ref_acc = np.any(np.diff(g['chain'], axis=0) != 0, axis=1)
    assert np.array_equal(res['accepted'], ref_acc)
    np.testing.assert_allclose(uq.samples, g['chain'], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(res['logpost'], g['logpost'], rtol=1e-10)
    np.testing.assert_allclose(uq.cmode, g['mapparams'], rtol=1e-9, atol=1e-12)
    assert abs(res['accrate'] - float(g['accrate'])) < 1e-12
    np.testing.assert_allclose(uq.predict_ens(g['xg'], nens=5, nburn=100), g['pred_ens'], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(uq.predict_MAP(g['xg']), g['pred_map'], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(uq.predict_sample(g['xg'], uq.cmode), g['pred_map'], rtol=1e-9, atol=1e-12)


def test_nn_mcmc_solver_flow_like_reference_tests():
    from quinn_b200.nns import MLP
    from quinn_b200.solvers import NN_MCMC
    np.random.seed(42)
    torch.manual_seed(42)
    net = MLP(1, 1, (5,), activ='tanh')
    x = np.random.rand(20, 1) * 2 - 1
    y = np.sin(3 * x) + 0.1 * np.random.randn(20, 1)
    uq = NN_MCMC(net, verbose=False)
    uq.fit(x, y, nmcmc=300, sampler='amcmc', zflag=False, sampler_params={})
    assert uq.samples.shape == (301, uq.pdim) and uq.cmode.shape == (uq.pdim,)
    xt = np.linspace(-1, 1, 10).reshape(-1, 1)
    assert uq.predict_sample(xt, uq.cmode).shape == (10, 1)
    ye = uq.predict_ens(xt, nens=5, nburn=100)
    assert ye.shape == (5, 10, 1) and np.all(np.isfinite(ye))
    assert uq.predict_MAP(xt).shape == (10, 1)
    lp = uq.logpost(uq.cmode, uq.lpinfo)
    assert isinstance(lp, float) and np.isfinite(lp)
    assert uq.logpostgrad(uq.cmode, uq.lpinfo).shape == (uq.pdim,)
    with pytest.raises(IndexError):                             # nburn=1000 > chain length: the reference raises too
        uq.predict_mom_sample(xt, msc=1, nsam=20)               # (quinn.py:84 -> nn_mcmc.py:194-196)
    uq.fit(x, y, nmcmc=1400, sampler='amcmc', zflag=False, sampler_params={})
    m, v, c = uq.predict_mom_sample(xt, msc=1, nsam=20)
    ref = uq.predict_ens(xt, nens=20, nburn=1000)
    np.testing.assert_allclose(m, ref.mean(0), rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(v, ref.var(0, ddof=1), rtol=1e-7, atol=1e-14)
    # MALA is reachable (the reference's docstring promises it, nn_mcmc.py:110) and zflag uses the analytic gradient
    uq.fit(x, y, nmcmc=50, sampler='mala', zflag=True, sampler_params={'epsilon': 0.002})
    assert uq.samples.shape == (51, uq.pdim)
    with pytest.raises(ValueError):
        uq.fit(x, y, nmcmc=5, sampler='nope', zflag=False, sampler_params={})


def test_nn_mcmc_many_chains_extension():
    from quinn_b200.nns import MLP
    from quinn_b200.solvers import NN_MCMC
    np.random.seed(1)
    torch.manual_seed(1)
    net = MLP(2, 1, (8, 8), activ='tanh')
    x = np.random.rand(64, 2)
    y = np.sin(x.sum(1, keepdims=True))
    uq = NN_MCMC(net, verbose=False, dtype=torch.float32)
    res = uq.fit(x, y, zflag=False, datanoise=0.1, nmcmc=40, sampler='hmc', sampler_params={'epsilon': 1e-3, 'L': 2},
                 nchains=16, seed=3, store_every=10)
    assert uq.samples.shape == (16, 5, uq.pdim) and res['logpost'].shape == (16, 41)
    assert res['accrate'].shape == (16,) and np.all(res['maxpost'] >= res['logpost'][:, 0] - 1e-9)
    lps = uq.logpost(uq.samples[:, -1, :], uq.lpinfo)
    np.testing.assert_allclose(lps, res['logpost'][:, -1], rtol=2e-4)
    assert uq.predict_ens(x[:7], nens=2, nburn=1).shape == (32, 7, 1)


# ---------------------------------------------------------------- NN_Ens (test_ensemble.py + golden)
def test_nn_ens_predictive():
    from quinn_b200.nns import MLP
    from quinn_b200.solvers import NN_Ens
    np.random.seed(11)
    torch.manual_seed(11)
    net = MLP(2, 2, (7, 4), activ='tanh')
    x = np.random.rand(30, 2)
    y = np.stack([np.sin(x.sum(1)), np.cos(x[:, 0])], axis=1)
    ens = NN_Ens(net, nens=4, dfrac=0.8)
    assert len(ens.learners) == 4
    ens.fit(x, y, nepochs=3, lrate=0.01, batch_size=10, freq_out=1000)
    xt = np.random.rand(9, 2)
    ye = ens.predict_ens(xt)
    assert ye.shape == (4, 9, 2)
    # every member's kernel-4 prediction equals its torch module
    ref = np.stack([l.best_model(torch.as_tensor(xt)).detach().numpy() for l in ens.learners])
    np.testing.assert_allclose(np.sort(ye, axis=0), np.sort(ref, axis=0), rtol=1e-11, atol=1e-13)
    assert ens.predict_sample(xt).shape == (9, 2)
    assert ens.predict_ens_fromsamples(xt, nens=3).shape == (3, 9, 2)
    m, v, c = ens.predict_mom_sample(xt, msc=2, nsam=4)
    np.testing.assert_allclose(m, ref.mean(0), rtol=1e-11, atol=1e-13)
    np.testing.assert_allclose(v, ref.var(0, ddof=1), rtol=1e-8, atol=1e-16)
    assert c.shape == (9, 9, 2)
    m1, v1, _ = ens.predict_mom_sample(xt, msc=1, nsam=1000)     # warns + clips to 4 members (nn_ens.py:99-101)
    np.testing.assert_allclose(m1, ref.mean(0), rtol=1e-11, atol=1e-13)
    np.testing.assert_allclose(v1, ref.var(0, ddof=1), rtol=1e-8, atol=1e-16)
    np.testing.assert_allclose(ens.predict(xt), ref.mean(0), rtol=1e-11, atol=1e-13)


def test_nn_ens_golden_permutation_rule():
    from quinn_b200 import ops
    from golden_util import netdesc_from_layers
    g = load('predict_ens.npz')
    layers, P = qo.mlp_layers(2, 2, (7, 4), True, 'tanh')
    out, _, _ = ops.predict(netdesc_from_layers(layers, P), g['thetas'][g['perm']], g['x'], dtype=torch.float64)
    np.testing.assert_allclose(out.cpu().numpy(), g['yens'], rtol=1e-11, atol=1e-13)


# ---------------------------------------------------------------- NN_VI (test_vi.py + golden)
@pytest.mark.parametrize('name,net', [('mlp', (2, 1, (6,), 'tanh')), ('mlp2', (3, 2, (8, 5), 'relu'))])
def test_bnet_viloss_golden(name, net):
    from quinn_b200.nns import MLP
    from quinn_b200.vi import BNet
    g = load(f'vi_{name}.npz')
    m = MLP(net[0], net[1], net[2], activ=net[3])
    b = BNet(m, pi=float(g['pi']), sigma1=float(g['sigma1']), sigma2=float(g['sigma2']))
    # load the reference's mu / rho into the per-tensor parameters (flat layout order)
    off = 0
    with torch.no_grad():
        for i in range(b.nparams):
            n = b.params[2 * i].numel()
            b.params[2 * i].copy_(torch.as_tensor(g['mu'][off:off + n]).view_as(b.params[2 * i]))
            b.params[2 * i + 1].copy_(torch.as_tensor(g['rho'][off:off + n]).view_as(b.params[2 * i + 1]))
            off += n
    b.loss_params = [float(g['datanoise']), int(g['nsam']), int(g['num_batches'])]
    loss = b.viloss(torch.as_tensor(g['x'], device='cuda'), torch.as_tensor(g['y'], device='cuda'), eps=g['eps'])
    assert abs(loss.item() - float(g['loss'])) <= 1e-10 * abs(float(g['loss']))
    loss.backward()
    gmu = np.concatenate([b.params[2 * i].grad.cpu().numpy().ravel() for i in range(b.nparams)])
    grho = np.concatenate([b.params[2 * i + 1].grad.cpu().numpy().ravel() for i in range(b.nparams)])
    np.testing.assert_allclose(gmu, g['gmu'], rtol=1e-8, atol=1e-9 * np.abs(g['gmu']).max())
    np.testing.assert_allclose(grho, g['grho'], rtol=1e-8, atol=1e-9 * np.abs(g['grho']).max())


def test_nn_vi_fit_and_predict():
    from quinn_b200.nns import MLP
    from quinn_b200.solvers import NN_VI
    np.random.seed(0)
    torch.manual_seed(0)
    net = MLP(1, 1, (8,), activ='tanh')
    x = np.random.rand(40, 1) * 2 - 1
    y = np.sin(2 * x) + 0.05 * np.random.randn(40, 1)
    vi = NN_VI(net, verbose=False, seed=4)
    vi.fit(x, y, nepochs=60, lrate=0.02, nsam=4, datanoise=0.1, freq_out=1000)
    assert vi.trained
    hist = np.array(vi.fit_info['history'])
    assert hist[-10:, 3].mean() < hist[:5, 3].mean()              # the ELBO loss goes down
    xt = np.linspace(-1, 1, 11).reshape(-1, 1)
    s1, s2 = vi.predict_sample(xt), vi.predict_sample(xt)
    assert s1.shape == (11, 1) and not np.allclose(s1, s2)
    ye = vi.predict_ens(xt, nens=6)
    assert ye.shape == (6, 11, 1)
    m, v, _ = vi.predict_mom_sample(xt, msc=1, nsam=200)
    assert m.shape == (11, 1) and np.all(v > 0)


def test_sharded_host_pipeline_equals_single_launch():
    """run() splits big host batches into shards on two streams; results must not change (Philox keyed by chain id)."""
    from quinn_b200 import ops
    from quinn_b200.mcmc import AMCMC, HMC, DeviceLogPost
    from golden_util import netdesc_from_layers
    rs = np.random.RandomState(21)
    layers, P = qo.mlp_layers(2, 1, (8,), True, 'tanh')
    desc = netdesc_from_layers(layers, P)
    x, y = rs.rand(40, 2), rs.randn(40, 1)
    th0 = (0.3 * rs.randn(37, P)).astype(np.float32)
    prob = ops.Problem(desc, x, y, 0.3, dtype=torch.float32)
    for make in (lambda: AMCMC(gamma=0.3, t0=4, tadapt=4, adapt='diag'), lambda: HMC(epsilon=0.01, L=2)):
        a = make()
        a.setLogPost(DeviceLogPost(prob), None)
        ref = a.run(12, th0, seed=9, store_every=4, verbose=False)
        b = make()
        b.setLogPost(DeviceLogPost(prob), None)
        b._auto_shards = lambda *args: 4
        out = b.run(12, torch.from_numpy(th0), seed=9, store_every=4, verbose=False)
        for k in ('chain', 'mapparams', 'maxpost', 'accrate', 'logpost', 'alphas', 'accepted'):
            np.testing.assert_array_equal(np.asarray(out[k], dtype=np.float64), np.asarray(ref[k], dtype=np.float64), err_msg=k)
        assert out['chain'].shape == (37, 4, P)


def test_nn_mcmc_batched_map_start():
    """zflag=True with many chains (extension): all start points are pre-conditioned together by Adam ascent on the
    log-posterior (kernel 2 + qb_adam_step); every chain starts at least as high as its random draw."""
    from quinn_b200.nns import MLP
    from quinn_b200.solvers import NN_MCMC
    np.random.seed(5)
    torch.manual_seed(5)
    net = MLP(2, 1, (16, 16), activ='tanh')
    x = np.random.rand(60, 2) * 2 - 1
    y = np.sin(2 * x[:, :1]) * x[:, 1:] + 0.05 * np.random.randn(60, 1)
    uq = NN_MCMC(net, verbose=False)
    np.random.seed(6)
    start = np.random.rand(32, uq.pdim)                       # what fit() draws for param_ini (nn_mcmc.py:124)
    np.random.seed(6)
    uq.fit(x, y, zflag=True, datanoise=0.05, nmcmc=20, sampler='amcmc', sampler_params={'gamma': 0.1}, nchains=32, seed=1)
    lp_start = uq.logpost(start, uq.lpinfo)
    assert uq.map_start_logpost.shape == (32,) and np.isfinite(uq.map_start_logpost).all()
    assert (uq.map_start_logpost >= lp_start).all() and uq.map_start_logpost.mean() > lp_start.mean() + 10.0
    np.testing.assert_allclose(uq.mcmc_results['logpost'][:, 0], uq.map_start_logpost, rtol=1e-9)
    assert uq.samples.shape == (32, 21, uq.pdim)


def test_more_than_65535_chains_through_the_solver_api():
    """The advertised scale is 1e5 chains: row-copy, member-parallel predictive and the chain kernels put chains on
    gridDim.x, so K > 65535 works through NN_MCMC.fit(zflag=True) (batched MAP start), predict_MAP and predict_ens."""
    from quinn_b200.nns import MLP
    from quinn_b200.solvers import NN_MCMC
    np.random.seed(8)
    torch.manual_seed(8)
    net = MLP(1, 1, (4,), activ='tanh')
    x = np.random.rand(16, 1) * 2 - 1
    y = np.sin(2 * x) + 0.05 * np.random.randn(16, 1)
    uq = NN_MCMC(net, verbose=False, dtype=torch.float32)
    uq.map_start_steps = 20
    K = 70_000
    res = uq.fit(x, y, zflag=True, datanoise=0.1, nmcmc=6, sampler='amcmc', sampler_params={'gamma': 0.1}, nchains=K, seed=2)
    assert uq.samples.shape == (K, 7, uq.pdim) and np.isfinite(uq.samples).all()
    assert uq.map_start_logpost.shape == (K,)
    xt = np.linspace(-1, 1, 5)[:, None]
    pm = uq.predict_MAP(xt)
    assert pm.shape == (K, 5, 1) and np.isfinite(pm).all()
    # spot check of the last chains (beyond 65535) against the oracle forward
    from oracle import quinn_oracle as qo
    layers, P = qo.mlp_layers(1, 1, (4,), True, 'tanh')
    for k in (65_535, 65_536, K - 1):
        ref = qo.forward(layers, np.asarray(uq.cmode[k], dtype=np.float64), xt)
        np.testing.assert_allclose(pm[k], ref, rtol=1e-4, atol=1e-5)
    pe = uq.predict_ens(xt, nens=2, nburn=2)
    assert pe.shape == (2 * K, 5, 1) and np.isfinite(pe).all()
    assert 0.0 < np.asarray(res['accrate']).mean() <= 1.0


def test_batched_map_start_against_per_chain_bfgs():
    """the batched MAP pre-conditioning (Adam ascent on kernel 2 for all starts at once) against what
    the reference's zflag=True does per chain (scipy BFGS on -logpost, nn_mcmc.py:125-127; here with the analytic
    gradient): from the same random starts it recovers most of the log-posterior improvement BFGS finds."""
    from quinn_b200.nns import MLP
    from quinn_b200.solvers import NN_MCMC
    np.random.seed(11)
    torch.manual_seed(11)
    net = MLP(1, 1, (6,), activ='tanh')
    x = np.random.rand(30, 1) * 2 - 1
    y = np.sin(3 * x) + 0.05 * np.random.randn(30, 1)
    uq = NN_MCMC(net, verbose=False)
    uq.lpinfo = {'model': None, 'xd': x, 'yd': [yy for yy in y], 'ltype': 'classical', 'lparams': {'sigma': 0.1}}
    starts = np.random.rand(12, uq.pdim)
    lp0 = uq.logpost(starts, uq.lpinfo)
    uq.map_batched_above = 10 ** 9
    bfgs = uq._map_start(starts)                                  # one BFGS run per start
    lp_bfgs = uq.logpost(bfgs, uq.lpinfo)
    uq.map_batched_above, uq.map_start_steps = 8, 3000
    batched = uq._map_start(starts)                               # all starts together
    lp_b = uq.logpost(batched, uq.lpinfo)
    assert (lp_b >= lp0).all() and (lp_bfgs >= lp0 - 1e-9).all()
    gain_b, gain_bfgs = lp_b - lp0, lp_bfgs - lp0
    assert np.median(gain_b) >= 0.9 * np.median(gain_bfgs), (np.median(gain_b), np.median(gain_bfgs))
    assert (gain_b >= 0.5 * gain_bfgs).mean() >= 0.75
