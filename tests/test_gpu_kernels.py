"""Parity of the CUDA kernels (through the C ABI, quinn_b200.ops) against the oracle and the golden
fixtures the reference produced.  fp64: 1e-10 relative (north_star); fp32: 1e-4 relative on log-posteriors."""
import numpy as np
import pytest
import torch

from golden_util import NET_CASES, make_inputs, make_thetas, oracle_layers, load, netdesc_from_spec, netdesc_from_layers
from oracle import quinn_oracle as qo

pytestmark = pytest.mark.gpu


def _ops():
    from quinn_b200 import ops
    return ops


def _rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.mark.parametrize('name', list(NET_CASES))
@pytest.mark.parametrize('dtype', ['f64', 'f32'])
def test_logpost_grad_golden(name, dtype):
    ops = _ops()
    spec = NET_CASES[name]
    g = load(f'logpost_{name}.npz')
    desc = netdesc_from_spec(spec)
    x, y = make_inputs(spec)
    thetas = make_thetas(spec, desc.n_params)
    td = torch.float64 if dtype == 'f64' else torch.float32
    tol_lp, tol_g = (1e-10, 1e-9) if dtype == 'f64' else (1e-4, 2e-3)
    prob = ops.Problem(desc, x, y, spec['sigma'], dtype=td)
    lp = ops.logpost(prob, thetas).cpu().numpy()
    assert np.all(np.abs(lp - g['lp']) <= tol_lp * np.abs(g['lp'])), (lp, g['lp'])
    lp2, gr = ops.logpost_grad(prob, thetas)
    lp2, gr = lp2.cpu().numpy(), gr.double().cpu().numpy()
    assert np.all(np.abs(lp2 - g['lp']) <= tol_lp * np.abs(g['lp']))
    for i in range(len(thetas)):
        assert _rel(gr[i], g['grad'][i]) <= tol_g, (i, _rel(gr[i], g['grad'][i]))
    # with the Gaussian prior (losses.py:202-204)
    probp = ops.Problem(desc, x, y, spec['sigma'], dtype=td, prior_sigma=float(g['sigma_prior']),
                        prior_anchor=g['anchor'], fulldatasize=int(g['nfull']))
    lpp, grp = ops.logpost_grad(probp, thetas)
    lpp, grp = lpp.cpu().numpy(), grp.double().cpu().numpy()
    assert np.all(np.abs(lpp - g['lp_prior']) <= tol_lp * np.abs(g['lp_prior']))
    for i in range(len(thetas)):
        assert _rel(grp[i], g['grad_prior'][i]) <= tol_g
    lpv = ops.logpost(probp, thetas).cpu().numpy()
    assert np.all(np.abs(lpv - g['lp_prior']) <= tol_lp * np.abs(g['lp_prior']))
    # forward values (nnwrap.py:49-62)
    out, _, _ = ops.predict(desc, thetas[:2], x, dtype=td)
    assert _rel(out.double().cpu().numpy(), g['pred']) <= (1e-11 if dtype == 'f64' else 2e-5)


@pytest.mark.parametrize('K,N', [(1, 1), (2, 3000), (3, 257), (700, 33), (5, 1024)])
@pytest.mark.parametrize('dtype', ['f64', 'f32'])
def test_logpost_shapes_vs_oracle(K, N, dtype):
    """Ragged / split / many-chain launches (N not a multiple of the tile, N-split when K is small)."""
    ops = _ops()
    rs = np.random.RandomState(K * 1000 + N)
    layers, P = qo.mlp_layers(3, 2, (16, 12), True, 'tanh')
    desc = netdesc_from_layers(layers, P)
    x = rs.rand(N, 3) * 2 - 1
    y = rs.randn(N, 2)
    th = 0.5 * rs.randn(K, P)
    td = torch.float64 if dtype == 'f64' else torch.float32
    prob = ops.Problem(desc, x, y, 0.3, dtype=td)
    lp = ops.logpost(prob, th).cpu().numpy()
    lp2, gr = ops.logpost_grad(prob, th)
    lp2, gr = lp2.cpu().numpy(), gr.double().cpu().numpy()
    idx = list(range(K)) if K <= 5 else [0, K // 2, K - 1]
    for k in idx:
        rl, rg = qo.logpost_grad(layers, th[k], x, y, 0.3)
        tol_lp, tol_g = (1e-10, 1e-9) if dtype == 'f64' else (1e-4, 2e-3)
        assert abs(lp[k] - rl) <= tol_lp * abs(rl)
        assert abs(lp2[k] - rl) <= tol_lp * abs(rl)
        assert _rel(gr[k], rg) <= tol_g


def test_final_exp_and_wide_outputs():
    ops = _ops()
    rs = np.random.RandomState(5)
    layers, P = qo.mlp_layers(2, 6, (9,), True, 'relu')
    desc = netdesc_from_layers(layers, P, final_exp=True)
    x, y, th = rs.rand(40, 2), rs.rand(40, 6) + 0.5, 0.3 * rs.randn(3, P)
    prob = ops.Problem(desc, x, y, 0.5, dtype=torch.float64)
    lp, gr = ops.logpost_grad(prob, th)
    for k in range(3):
        rl, rg = qo.logpost_grad(layers, th[k], x, y, 0.5, final='exp')
        assert abs(lp[k].item() - rl) <= 1e-10 * abs(rl)
        assert _rel(gr[k].cpu().numpy(), rg) <= 1e-9


def _c1_desc():
    layers, P = qo.rnet_layers(3, 3, 1, 1, biasorno=True, nonlin=True, mlp=False, shared=True)
    return layers, netdesc_from_layers(layers, P)


def _replay(kind, g, desc, extra=None, track=0, adapt='none'):
    ops = _ops()
    n = len(g['u'])
    prob = ops.Problem(desc, g['x'], g['y'], float(g['sigma']), dtype=torch.float64)
    st = ops.ChainState(prob, g['theta0'])
    rec = ops.Recorder(st, n, store_every=1)
    unif = torch.as_tensor(g['u'], dtype=torch.float64, device='cuda').reshape(n, 1).contiguous()
    if kind == 'amcmc':
        incr = torch.as_tensor(g['xi'], dtype=torch.float64, device='cuda').reshape(n, 1, -1).contiguous()
        am = ops.AmcmcState(st, adapt=adapt, track=track, **(extra or {}))
        ops.amcmc_run(st, am, n, rec, incr=incr, unif=unif)
        aux = am
    else:
        incr = torch.as_tensor(g['p'], dtype=torch.float64, device='cuda').reshape(n, 1, -1).contiguous()
        hm = ops.HmcState(st, method=kind, **(extra or {}))
        ops.hmc_run(st, hm, n, rec, incr=incr, unif=unif)
        aux = hm
    torch.cuda.synchronize()
    return st, rec, aux


def _check_replay(st, rec, g):
    chain = rec.samples[0].cpu().numpy()
    ref_acc = np.any(np.diff(g['chain'], axis=0) != 0, axis=1)
    acc = rec.accepted[0].cpu().numpy().astype(bool)
    assert np.array_equal(acc, ref_acc), f'accept/reject sequence differs at steps {np.nonzero(acc != ref_acc)[0][:5]}'
    np.testing.assert_allclose(chain, g['chain'][1:], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(rec.logpost[0].cpu().numpy(), g['logpost'][1:], rtol=1e-10)
    al, ral = rec.alpha[0].cpu().numpy(), g['alphas'][1:]
    fin = np.isfinite(ral) & (ral > 1e-290)
    np.testing.assert_allclose(al[fin], ral[fin], rtol=1e-5)
    assert st.naccept[0].item() == int(round(float(g['accrate']) * len(ral)))
    np.testing.assert_allclose(st.map_theta[0].cpu().numpy(), g['mapparams'], rtol=1e-9, atol=1e-12)
    assert abs(st.map_lp[0].item() - float(g['maxpost'])) <= 1e-10 * abs(float(g['maxpost']))


def test_amcmc_replay_config1_first_1000_steps():
    g = load('chain_c1_amcmc.npz')
    _, desc = _c1_desc()
    st, rec, am = _replay('amcmc', g, desc, dict(gamma=0.01), track=2)
    _check_replay(st, rec, g)
    np.testing.assert_allclose(am.xm[0].cpu().numpy(), g['Xm'], rtol=1e-10)
    np.testing.assert_allclose(am.cov[0].cpu().numpy(), g['cov'], rtol=1e-8, atol=1e-14)


def test_amcmc_replay_with_adaptation_tracks_reference_covariance():
    g = load('chain_c1_amcmc_adapt.npz')
    _, desc = _c1_desc()
    st, rec, am = _replay('amcmc', g, desc, dict(gamma=float(g['gamma']), t0=int(g['t0']), tadapt=int(g['tadapt'])),
                          adapt='full')
    _check_replay(st, rec, g)
    np.testing.assert_allclose(am.xm[0].cpu().numpy(), g['Xm'], rtol=1e-10)
    np.testing.assert_allclose(am.cov[0].cpu().numpy(), g['cov'], rtol=1e-8, atol=1e-14)
    # the in-kernel Cholesky factor reproduces the reference's final proposal covariance
    Lf = am.chol[0].cpu().numpy()
    np.testing.assert_allclose(Lf @ Lf.T, g['propcov'], rtol=1e-7, atol=1e-13)


def test_hmc_replay_config1():
    g = load('chain_c1_hmc.npz')
    _, desc = _c1_desc()
    st, rec, _ = _replay('hmc', g, desc, dict(epsilon=float(g['epsilon']), L=int(g['L'])))
    _check_replay(st, rec, g)


def test_mala_replay_mlp():
    g = load('chain_mlp_mala.npz')
    layers, P = qo.mlp_layers(1, 1, (5,), True, 'tanh')
    st, rec, _ = _replay('mala', g, netdesc_from_layers(layers, P), dict(epsilon=float(g['epsilon'])))
    _check_replay(st, rec, g)


def test_amcmc_replay_mlp():
    g = load('chain_mlp_amcmc.npz')
    layers, P = qo.mlp_layers(1, 1, (5,), True, 'tanh')
    st, rec, _ = _replay('amcmc', g, netdesc_from_layers(layers, P), dict(gamma=float(g['gamma'])))
    _check_replay(st, rec, g)


def test_segmented_run_equals_single_run():
    """Two 150-step calls == one 300-step call (state carried in qb_chain_t / qb_amcmc_t)."""
    ops = _ops()
    g = load('chain_mlp_amcmc.npz')
    layers, P = qo.mlp_layers(1, 1, (5,), True, 'tanh')
    desc = netdesc_from_layers(layers, P)
    n = len(g['u'])
    prob = ops.Problem(desc, g['x'], g['y'], float(g['sigma']), dtype=torch.float64)
    unif = torch.as_tensor(g['u'], dtype=torch.float64, device='cuda').reshape(n, 1).contiguous()
    incr = torch.as_tensor(g['xi'], dtype=torch.float64, device='cuda').reshape(n, 1, -1).contiguous()
    st = ops.ChainState(prob, g['theta0'])
    am = ops.AmcmcState(st, gamma=float(g['gamma']))
    h = n // 2
    r1, r2 = ops.Recorder(st, h), ops.Recorder(st, n - h)
    ops.amcmc_run(st, am, h, r1, incr=incr[:h].contiguous(), unif=unif[:h].contiguous())
    ops.amcmc_run(st, am, n - h, r2, incr=incr[h:].contiguous(), unif=unif[h:].contiguous())
    chain = torch.cat([r1.samples[0], r2.samples[0]]).cpu().numpy()
    np.testing.assert_allclose(chain, g['chain'][1:], rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize('sampler', ['amcmc', 'amcmc_diag', 'hmc', 'mala'])
@pytest.mark.parametrize('dtype', ['f64', 'f32'])
def test_native_philox_chains_sample_a_gaussian_posterior(sampler, dtype):
    """Linear model y = w x + b with known Gaussian posterior: many independent Philox-driven chains must
    reproduce its mean and standard deviation (statistical check of propose/accept + RNG)."""
    ops = _ops()
    rs = np.random.RandomState(0)
    N, sig = 40, 0.5
    x = rs.rand(N, 1) * 2 - 1
    y = 1.5 * x - 0.7 + sig * rs.randn(N, 1)
    layers, P = qo.mlp_layers(1, 1, (), True, 'tanh') if False else ([dict(n_in=1, n_out=1, w_off=0, b_off=1, act='identity', res_step=0.0)], 2)
    desc = netdesc_from_layers(layers, P)
    A = np.hstack([x, np.ones_like(x)])
    cov = np.linalg.inv(A.T @ A) * sig ** 2
    mean = cov @ A.T @ y[:, 0] / sig ** 2
    td = torch.float64 if dtype == 'f64' else torch.float32
    K, nsteps = 2048, 600
    prob = ops.Problem(desc, x, y, sig, dtype=td)
    st = ops.ChainState(prob, np.tile(mean, (K, 1)) + 0.05 * rs.randn(K, 2))
    if sampler.startswith('amcmc'):
        am = ops.AmcmcState(st, gamma=1.0, t0=50, tadapt=50, adapt='diag' if sampler == 'amcmc_diag' else 'full')
        ops.amcmc_run(st, am, nsteps, None, seed=1234)
    else:
        hm = ops.HmcState(st, epsilon=0.08 if sampler == 'hmc' else 0.1, L=5, method=sampler)
        ops.hmc_run(st, hm, nsteps, None, seed=99)
    torch.cuda.synchronize()
    th = st.theta.double().cpu().numpy()
    acc = st.naccept.cpu().numpy() / nsteps
    assert 0.05 < acc.mean() < 0.999, acc.mean()
    se = np.sqrt(np.diag(cov) / K)
    assert np.all(np.abs(th.mean(0) - mean) < 6 * se + 1e-3), (th.mean(0), mean)
    assert np.all(np.abs(th.std(0) / np.sqrt(np.diag(cov)) - 1) < 0.12), (th.std(0), np.sqrt(np.diag(cov)))


def test_philox_results_do_not_depend_on_sharding():
    ops = _ops()
    rs = np.random.RandomState(3)
    layers, P = qo.mlp_layers(2, 1, (4,), True, 'tanh')
    desc = netdesc_from_layers(layers, P)
    x, y = rs.rand(30, 2), rs.randn(30, 1)
    th0 = 0.3 * rs.randn(8, P)
    prob = ops.Problem(desc, x, y, 0.4, dtype=torch.float64)

    def run(sl, off):
        st = ops.ChainState(prob, th0[sl])
        am = ops.AmcmcState(st, gamma=0.5, t0=10, tadapt=10, adapt='diag')
        ops.amcmc_run(st, am, 40, None, seed=7, chain_offset=off)
        return st.theta.cpu().numpy()
    full = run(slice(0, 8), 0)
    np.testing.assert_array_equal(full[:4], run(slice(0, 4), 0))
    np.testing.assert_array_equal(full[4:], run(slice(4, 8), 4))


def test_predictive_ensemble_golden_and_moments():
    ops = _ops()
    g = load('predict_ens.npz')
    layers, P = qo.mlp_layers(2, 2, (7, 4), True, 'tanh')
    desc = netdesc_from_layers(layers, P)
    th = g['thetas'][g['perm']]
    out, mean, var = ops.predict(desc, th, g['x'], dtype=torch.float64, want_out=True, want_moments=True)
    np.testing.assert_allclose(out.cpu().numpy(), g['yens'], rtol=1e-11, atol=1e-13)
    np.testing.assert_allclose(mean.cpu().numpy(), g['ymean'], rtol=1e-11, atol=1e-13)
    np.testing.assert_allclose(var.cpu().numpy(), g['yvar'], rtol=1e-8, atol=1e-16)
    _, mean2, var2 = ops.predict(desc, th, g['x'], dtype=torch.float64, want_out=False, want_moments=True)
    np.testing.assert_allclose(mean2.cpu().numpy(), g['ymean'], rtol=1e-11, atol=1e-13)
    np.testing.assert_allclose(var2.cpu().numpy(), g['yvar'], rtol=1e-8, atol=1e-16)


@pytest.mark.parametrize('dtype', ['f64', 'f32'])
def test_predictive_fused_moments_large_n(dtype):
    """Point-parallel fused mean/var path (enough tiles to fill the GPU) against the oracle on a sample of points."""
    ops = _ops()
    rs = np.random.RandomState(8)
    layers, P = qo.mlp_layers(3, 1, (16, 16), True, 'tanh')
    desc = netdesc_from_layers(layers, P)
    M, N = 7, 90001
    th, x = 0.6 * rs.randn(M, P), rs.rand(N, 3)
    td = torch.float64 if dtype == 'f64' else torch.float32
    out, mean, var = ops.predict(desc, th, x, dtype=td, want_out=True, want_moments=True)
    sel = np.r_[0:50, N - 50:N, rs.randint(0, N, 200)]
    ref = qo.predict_ens(layers, th, x[sel])
    tol = 1e-11 if dtype == 'f64' else 2e-5
    np.testing.assert_allclose(out.double().cpu().numpy()[:, sel], ref, rtol=tol, atol=tol)
    np.testing.assert_allclose(mean.double().cpu().numpy()[sel], ref.mean(0), rtol=tol, atol=tol)
    np.testing.assert_allclose(var.double().cpu().numpy()[sel], ref.var(0, ddof=1), rtol=1e3 * tol, atol=tol)


@pytest.mark.parametrize('name,net', [('mlp', (2, 1, (6,), 'tanh')), ('mlp2', (3, 2, (8, 5), 'relu'))])
def test_vi_loss_and_grads_golden(name, net):
    """BNet.viloss + autograd (bnet.py:181-232) rebuilt from qb_vi_sample + qb_logpost_grad + qb_vi_backward."""
    ops = _ops()
    g = load(f'vi_{name}.npz')
    layers, P = qo.mlp_layers(net[0], net[1], net[2], True, net[3])
    desc = netdesc_from_layers(layers, P)
    nsam, nb, dn = int(g['nsam']), int(g['num_batches']), float(g['datanoise'])
    pi, s1, s2 = float(g['pi']), float(g['sigma1']), float(g['sigma2'])
    mu = torch.as_tensor(g['mu'], device='cuda')
    rho = torch.as_tensor(g['rho'], device='cuda')
    w, eps, logq, logp = ops.vi_sample(mu, rho, nsam, pi, s1, s2, eps=g['eps'])
    B, o = g['y'].shape
    prob = ops.Problem(desc, g['x'], g['y'], 1.0, dtype=torch.float64)
    lp, glp = ops.logpost_grad(prob, w)
    ssq = -2.0 * (lp + 0.5 * B * np.log(2 * np.pi))
    c_nll = 0.5 * B / (nsam * B * o) / dn ** 2
    loss = (logq.mean() - logp.mean()) / nb + B * np.log(dn) + 0.5 * B * np.log(2 * np.pi) + c_nll * ssq.sum()
    assert abs(loss.item() - float(g['loss'])) <= 1e-10 * abs(float(g['loss']))
    gmu, grho = ops.vi_backward(mu, rho, eps, w, glp, pi, s1, s2, c_nll, -1.0 / (nsam * nb), 1.0 / (nsam * nb))
    np.testing.assert_allclose(gmu.cpu().numpy(), g['gmu'], rtol=1e-8, atol=1e-9 * np.abs(g['gmu']).max())
    np.testing.assert_allclose(grho.cpu().numpy(), g['grho'], rtol=1e-8, atol=1e-9 * np.abs(g['grho']).max())


def test_vi_philox_normals_are_standard():
    ops = _ops()
    mu = torch.zeros(4096, device='cuda', dtype=torch.float32)
    rho = torch.zeros(4096, device='cuda', dtype=torch.float32)
    w, eps, _, _ = ops.vi_sample(mu, rho, 64, 0.5, 1.0, 1.0, seed=5, step=2)
    e = eps.double().cpu().numpy()
    assert abs(e.mean()) < 0.01 and abs(e.std() - 1) < 0.01
    assert abs(np.mean(e ** 4) - 3) < 0.1
    np.testing.assert_allclose(w.cpu().numpy(), eps.cpu().numpy(), rtol=1e-6)
    w2, eps2, _, _ = ops.vi_sample(mu, rho, 64, 0.5, 1.0, 1.0, seed=5, step=3)
    assert not np.array_equal(eps2.cpu().numpy(), eps.cpu().numpy())


@pytest.mark.parametrize('kind', ['amcmc', 'hmc'])
def test_checkpoint_resume_is_exact(kind, tmp_path):
    """state_dict -> torch.save -> fresh objects -> load_state_dict continues the Philox-driven chains bit for bit."""
    ops = _ops()
    rs = np.random.RandomState(12)
    layers, P = qo.mlp_layers(2, 1, (6,), True, 'tanh')
    desc = netdesc_from_layers(layers, P)
    x, y, th0 = rs.rand(50, 2), rs.randn(50, 1), 0.3 * rs.randn(16, P)
    prob = ops.Problem(desc, x, y, 0.3, dtype=torch.float32)

    def make():
        st = ops.ChainState(prob, th0)
        sm = (ops.AmcmcState(st, gamma=0.5, t0=5, tadapt=5, adapt='full') if kind == 'amcmc'
              else ops.HmcState(st, epsilon=0.01, L=2))
        return st, sm
    adv = (lambda st, sm, n: ops.amcmc_run(st, sm, n, None, seed=3)) if kind == 'amcmc' else \
          (lambda st, sm, n: ops.hmc_run(st, sm, n, None, seed=3))
    st, sm = make()
    adv(st, sm, 30)
    ref = st.theta.cpu().numpy().copy()
    st2, sm2 = make()
    adv(st2, sm2, 12)
    torch.save(dict(chain=st2.state_dict(), sampler=sm2.state_dict()), tmp_path / 'ck.pt')
    ck = torch.load(tmp_path / 'ck.pt')
    st3, sm3 = make()
    st3.load_state_dict(ck['chain'])
    sm3.load_state_dict(ck['sampler'])
    adv(st3, sm3, 18)
    np.testing.assert_array_equal(st3.theta.cpu().numpy(), ref)
    assert st3.t == 30 and np.array_equal(st3.naccept.cpu().numpy(), st.naccept.cpu().numpy())
