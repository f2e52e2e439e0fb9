"""Pin the CPU oracle (oracle/quinn_oracle.py) against fixtures produced by the reference itself
(tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from golden_util import NET_CASES, NET_CASES_R2, make_inputs, make_thetas, oracle_layers, load
from oracle import quinn_oracle as qo

RTOL = 1e-12


@pytest.mark.parametrize('name', list(NET_CASES) + list(NET_CASES_R2))
def test_logpost_and_grad_match_reference(name):
    spec = NET_CASES[name] if name in NET_CASES else NET_CASES_R2[name]
    g = load(f'logpost_{name}.npz')
    layers, P = oracle_layers(spec)
    assert P == int(g['pdim'])
    x, y = make_inputs(spec)
    thetas = make_thetas(spec, P)
    prior = dict(sigma=float(g['sigma_prior']), anchor=g['anchor'])
    for i, th in enumerate(thetas):
        lp = qo.logpost(layers, th, x, y, spec['sigma'])
        assert abs(lp - g['lp'][i]) <= RTOL * abs(g['lp'][i])
        lp2, gr = qo.logpost_grad(layers, th, x, y, spec['sigma'])
        assert abs(lp2 - g['lp'][i]) <= RTOL * abs(g['lp'][i])
        scale = np.abs(g['grad'][i]).max()
        assert np.abs(gr - g['grad'][i]).max() <= 1e-11 * scale
        lpp, grp = qo.logpost_grad(layers, th, x, y, spec['sigma'], fulldatasize=int(g['nfull']), prior=prior)
        assert abs(lpp - g['lp_prior'][i]) <= RTOL * abs(g['lp_prior'][i])
        assert np.abs(grp - g['grad_prior'][i]).max() <= 1e-11 * np.abs(g['grad_prior'][i]).max()
    for i in range(2):
        np.testing.assert_allclose(qo.forward(layers, thetas[i], x), g['pred'][i], rtol=1e-12, atol=1e-14)


def _c1_layers():
    return qo.rnet_layers(3, 3, 1, 1, biasorno=True, nonlin=True, mlp=False, shared=True)


def _check_chain(res, g, nsteps):
    # accept/reject sequence = positions where the chain state changed
    np.testing.assert_allclose(res['chain'], g['chain'], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(res['logpost'], g['logpost'], rtol=1e-9)
    ref_acc = np.any(np.diff(g['chain'], axis=0) != 0, axis=1)
    assert np.array_equal(res['accepted'], ref_acc)
    fin = np.isfinite(g['alphas'])
    np.testing.assert_allclose(res['alphas'][fin], g['alphas'][fin], rtol=1e-6, atol=1e-300)
    assert abs(res['accrate'] - float(g['accrate'])) < 1e-12
    np.testing.assert_allclose(res['mapparams'], g['mapparams'], rtol=1e-9, atol=1e-12)
    assert abs(res['maxpost'] - float(g['maxpost'])) <= 1e-9 * abs(float(g['maxpost']))


@pytest.mark.parametrize('fname', ['chain_c1_amcmc.npz', 'chain_c1_amcmc_adapt.npz'])
def test_amcmc_replay_config1(fname):
    g = load(fname)
    layers, P = _c1_layers()
    x, y, s = g['x'], g['y'], float(g['sigma'])
    kw = {}
    if 'tadapt' in g.files:
        kw = dict(gamma=float(g['gamma']), t0=int(g['t0']), tadapt=int(g['tadapt']))
    else:
        kw = dict(gamma=0.01)
    n = len(g['u'])
    res = qo.run_chain(lambda th: qo.logpost(layers, th, x, y, s), g['theta0'], n, 'amcmc',
                       dict(xi=g['xi'], u=g['u']), **kw)
    _check_chain(res, g, n)
    np.testing.assert_allclose(res['_Xm'], g['Xm'], rtol=1e-10)
    np.testing.assert_allclose(res['_cov'], g['cov'], rtol=1e-8, atol=1e-14)
    np.testing.assert_allclose(res['_propcov'], g['propcov'], rtol=1e-8, atol=1e-14)
    # thinning rule + predictive
    rows = qo.mcmc_thinning_rows(res['chain'].shape[0], 5, 100)
    ye = qo.predict_ens(layers, res['chain'][rows], g['xg'])
    np.testing.assert_allclose(ye, g['pred_ens'], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(qo.forward(layers, res['mapparams'], g['xg']), g['pred_map'], rtol=1e-9, atol=1e-12)


def test_hmc_replay_config1():
    g = load('chain_c1_hmc.npz')
    layers, P = _c1_layers()
    x, y, s = g['x'], g['y'], float(g['sigma'])
    n = len(g['u'])
    res = qo.run_chain(lambda th: qo.logpost(layers, th, x, y, s), g['theta0'], n, 'hmc',
                       dict(p=g['p'], u=g['u']), grad_fn=lambda th: qo.logpost_grad(layers, th, x, y, s)[1],
                       epsilon=float(g['epsilon']), L=int(g['L']))
    _check_chain(res, g, n)


def test_mala_replay_mlp():
    g = load('chain_mlp_mala.npz')
    layers, P = qo.mlp_layers(1, 1, (5,), True, 'tanh')
    x, y, s = g['x'], g['y'], float(g['sigma'])
    n = len(g['u'])
    res = qo.run_chain(lambda th: qo.logpost(layers, th, x, y, s), g['theta0'], n, 'mala',
                       dict(p=g['p'], u=g['u']), grad_fn=lambda th: qo.logpost_grad(layers, th, x, y, s)[1],
                       epsilon=float(g['epsilon']))
    _check_chain(res, g, n)


def test_amcmc_replay_mlp():
    g = load('chain_mlp_amcmc.npz')
    layers, P = qo.mlp_layers(1, 1, (5,), True, 'tanh')
    x, y, s = g['x'], g['y'], float(g['sigma'])
    n = len(g['u'])
    res = qo.run_chain(lambda th: qo.logpost(layers, th, x, y, s), g['theta0'], n, 'amcmc',
                       dict(xi=g['xi'], u=g['u']), gamma=float(g['gamma']))
    _check_chain(res, g, n)


@pytest.mark.parametrize('name,net', [('mlp', (2, 1, (6,), 'tanh')), ('mlp2', (3, 2, (8, 5), 'relu'))])
def test_vi_loss_and_grads(name, net):
    g = load(f'vi_{name}.npz')
    layers, P = qo.mlp_layers(net[0], net[1], net[2], True, net[3])
    assert P == g['mu'].size
    loss, gmu, grho = qo.vi_loss(layers, g['mu'], g['rho'], g['eps'], g['x'], g['y'], float(g['datanoise']),
                                 int(g['num_batches']), float(g['pi']), float(g['sigma1']), float(g['sigma2']),
                                 want_grad=True)
    assert abs(loss - float(g['loss'])) <= 1e-12 * abs(float(g['loss']))
    np.testing.assert_allclose(gmu, g['gmu'], rtol=1e-9, atol=1e-9 * np.abs(g['gmu']).max())
    np.testing.assert_allclose(grho, g['grho'], rtol=1e-9, atol=1e-9 * np.abs(g['grho']).max())


def test_ensemble_predictive():
    g = load('predict_ens.npz')
    layers, P = qo.mlp_layers(2, 2, (7, 4), True, 'tanh')
    ye = qo.predict_ens(layers, g['thetas'][g['perm']], g['x'])
    np.testing.assert_allclose(ye, g['yens'], rtol=1e-12, atol=1e-14)
    m, v, c = qo.predict_moments(ye, msc=2)
    np.testing.assert_allclose(m, g['ymean'], rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(v, g['yvar'], rtol=1e-10, atol=1e-16)
    np.testing.assert_allclose(c, g['ycov'], rtol=1e-10, atol=1e-16)
    m1, v1, _ = qo.predict_moments(ye, msc=1)
    np.testing.assert_allclose(v1, g['yvar'], rtol=1e-10, atol=1e-16)


def test_torch_port_matches_golden():
    """oracle/torch_port.py (the CPU-baseline flow) computes the same numbers as the reference."""
    from oracle.torch_port import RefPort
    for name in ('mlp_c2', 'mlp_tanh_o2'):
        spec = NET_CASES[name]
        g = load(f'logpost_{name}.npz')
        port = RefPort(spec['indim'], spec['outdim'], spec['hls'], spec['activ'])
        x, y = make_inputs(spec)
        th = make_thetas(spec, port.pdim)
        lp = port.logpost(th[0], x, [r for r in y], spec['sigma'])
        assert abs(lp - g['lp'][0]) <= 1e-12 * abs(g['lp'][0])
        gr = port.logpostgrad(th[0], x, [r for r in y], spec['sigma'])
        assert np.abs(gr - g['grad'][0]).max() <= 1e-11 * np.abs(g['grad'][0]).max()


def test_diag_fisher_matches_reference():
    """oracle.diag_fisher against NNWrap.calc_hess_diag (nnwrap.py:204-229) recorded from the reference."""
    g = load('hessdiag_mlp.npz')
    layers, P = qo.mlp_layers(2, 1, (6, 5), True, 'tanh')
    f0 = qo.diag_fisher(layers, g['theta'], g['x'], g['y'], float(g['sigma']))
    np.testing.assert_allclose(f0, g['fisher_noprior'], rtol=1e-10)
    f1 = qo.diag_fisher(layers, g['theta'], g['x'], g['y'], float(g['sigma']), fulldatasize=int(g['nfull']),
                        prior=dict(sigma=float(g['sigma_prior']), anchor=g['anchor']))
    np.testing.assert_allclose(f1, g['fisher_prior'], rtol=1e-10)


def test_philox_known_answers():
    """oracle/philox.py (the counter-based streams of the chain kernels) against the published Philox4x32-10 vectors."""
    from oracle import philox
    assert [int(v) for v in philox.philox4x32_10((0, 0, 0, 0), (0, 0))] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert [int(v) for v in philox.philox4x32_10((0xffffffff,) * 4, (0xffffffff,) * 2)] == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert [int(v) for v in philox.philox4x32_10((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0))] == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]
    z = philox.normals(5, 3, 7, philox.STREAM_INCR, 40001)
    assert abs(z.mean()) < 0.02 and abs(z.std() - 1.0) < 0.02
