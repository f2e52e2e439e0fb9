#!/usr/bin/env python
"""Generate the golden fixtures in this directory by RUNNING THE REFERENCE.

Run only in the build container (needs /root/reference; matplotlib is absent
there, so it is stubbed exactly as SURVEY.md section 8c describes):

    python tests/golden/make_golden.py

Outputs (committed): logpost_*.npz, chain_*.npz, vi_*.npz, predict_*.npz.
Large random inputs are NOT stored; they are re-derived from the recorded seed
with ``np.random.RandomState(seed)`` by tests/golden_util.py.
"""
import os
import sys
from unittest.mock import MagicMock

for _n in ['matplotlib', 'matplotlib.pyplot', 'matplotlib.colors', 'matplotlib.lines', 'matplotlib.cm']:
    sys.modules[_n] = MagicMock()
sys.path.insert(0, os.environ.get('QUINN_REF', '/root/reference'))

import numpy as np
import torch

from quinn.nns.mlp import MLP
from quinn.nns.rnet import RNet, Poly, NonPar
from quinn.nns.nnwrap import NNWrap
from quinn.nns.losses import NegLogPost
from quinn.nns.tchutils import tch
from quinn.solvers.nn_mcmc import NN_MCMC
from quinn.solvers.nn_ens import NN_Ens
from quinn.solvers.nn_vi import NN_VI
from quinn.mcmc.admcmc import AMCMC
from quinn.mcmc.hmc import HMC
from quinn.mcmc.mala import MALA
from quinn.vi.bnet import BNet

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from golden_util import NET_CASES, make_inputs, make_thetas   # noqa: E402


def build_net(spec):
    kind = spec['kind']
    if kind == 'mlp':
        return MLP(spec['indim'], spec['outdim'], spec['hls'], biasorno=spec['bias'], activ=spec['activ'])
    if kind == 'rnet':
        wp = Poly(0) if spec['shared'] else NonPar(spec['nlayers'] + 1)
        return RNet(spec['rdim'], spec['nlayers'], wp_function=wp, indim=spec['indim'], outdim=spec['outdim'],
                    layer_pre=True, layer_post=True, biasorno=spec['bias'], nonlin=spec['nonlin'], mlp=spec['mlp'])
    raise ValueError(kind)


# ---------------------------------------------------------------- A: log-posterior and gradient
def gen_logpost():
    for name, spec in NET_CASES.items():
        torch.manual_seed(0)
        net = build_net(spec)
        uq = NN_MCMC(net, verbose=False)
        P = uq.pdim
        x, y = make_inputs(spec)
        thetas = make_thetas(spec, P)
        lpinfo = {'model': None, 'xd': x, 'yd': [yy for yy in y], 'ltype': 'classical',
                  'lparams': {'sigma': spec['sigma']}}
        lps = np.array([uq.logpost(th, lpinfo) for th in thetas])
        grads = np.array([uq.logpostgrad(th, lpinfo) for th in thetas])
        # the same with a Gaussian prior through the generic NNWrap seam (nnwrap.py:109-150, losses.py:202-204)
        rs = np.random.RandomState(spec['seed'] + 77)
        anchor = 0.3 * rs.randn(P)
        nfull = 3 * x.shape[0] + 1
        sp = 0.7
        wrap = NNWrap(uq.nnmodel)
        loss = NegLogPost(uq.nnmodel, nfull, spec['sigma'], {'sigma': sp, 'anchor': tch(anchor)})
        lps_prior = np.array([-wrap.calc_loss(th, loss, x, y) for th in thetas])
        grads_prior = np.array([-wrap.calc_lossgrad(th, loss, x, y) for th in thetas])
        preds = np.array([wrap.predict(x, th) for th in thetas[:2]])
        np.savez_compressed(os.path.join(HERE, f'logpost_{name}.npz'), pdim=P, lp=lps, grad=grads,
                            lp_prior=lps_prior, grad_prior=grads_prior, anchor=anchor, nfull=nfull,
                            sigma_prior=sp, pred=preds)
        print(name, 'P', P, 'lp', lps[:2])


# ---------------------------------------------------------------- B: chains with recorded draws
class Recorder:
    """Wraps the three np.random entry points the samplers use and logs what they returned."""

    def __init__(self):
        self.xi, self.p, self.u = [], [], []
        self._mvn, self._randn, self._rs = np.random.multivariate_normal, np.random.randn, np.random.random_sample

    def __enter__(self):
        def mvn(mean, cov, *a, **k):
            v = self._mvn(mean, cov, *a, **k)
            self.xi.append(np.array(v))
            return v

        def randn(*a):
            v = self._randn(*a)
            self.p.append(np.array(v))
            return v

        def rs(*a, **k):
            v = self._rs(*a, **k)
            self.u.append(float(v))
            return v
        np.random.multivariate_normal, np.random.randn, np.random.random_sample = mvn, randn, rs
        return self

    def __exit__(self, *exc):
        np.random.multivariate_normal, np.random.randn, np.random.random_sample = self._mvn, self._randn, self._rs


def config1_data():
    """examples/ex_ufit.py:39-93 with np.random.seed(0); torch.manual_seed(0) first."""
    from quinn.utils.maps import scale01ToDom
    from quinn.func.funcs import Sine
    np.random.seed(0)
    torch.manual_seed(0)
    nall, ndim, datanoise = 15, 1, 0.02
    domain = np.tile(np.array([-np.pi, np.pi]), (ndim, 1))
    xall = scale01ToDom(np.random.rand(nall, ndim), domain)
    yall = Sine(xall, datanoise=datanoise)
    ntrn = int(0.9 * nall)
    net = RNet(3, 3, wp_function=Poly(0), indim=ndim, outdim=1, layer_pre=True, layer_post=True,
               biasorno=True, nonlin=True, mlp=False, final_layer=None)
    return net, xall[:ntrn], yall[:ntrn], datanoise


def gen_chains():
    # B1: config 1, AMCMC through NN_MCMC.fit, 1000 steps (proposal covariance never adapts: t0=100, tadapt=1000)
    net, x, y, dn = config1_data()
    uq = NN_MCMC(net, verbose=False)
    theta0 = np.random.rand(uq.pdim)
    with Recorder() as rec:
        uq_res = _fit_capture(uq, x, y, dn, 1000, theta0, 'amcmc', {'gamma': 0.01})
    np.savez_compressed(os.path.join(HERE, 'chain_c1_amcmc.npz'), x=x, y=y, sigma=dn, theta0=theta0,
                        xi=np.array(rec.xi), u=np.array(rec.u), **uq_res)
    print('c1 amcmc accrate', uq_res['accrate'])

    # B2: config 1 shape with adaptation exercised (t0=50, tadapt=200, 1200 steps)
    net, x, y, dn = config1_data()
    uq = NN_MCMC(net, verbose=False)
    theta0 = np.random.rand(uq.pdim)
    with Recorder() as rec:
        uq_res = _fit_capture(uq, x, y, dn, 1200, theta0, 'amcmc', {'gamma': 0.05, 't0': 50, 'tadapt': 200})
    np.savez_compressed(os.path.join(HERE, 'chain_c1_amcmc_adapt.npz'), x=x, y=y, sigma=dn, theta0=theta0,
                        xi=np.array(rec.xi), u=np.array(rec.u), gamma=0.05, t0=50, tadapt=200, **uq_res)
    print('c1 amcmc adapt accrate', uq_res['accrate'])

    # B3: config 1, HMC L=3 eps=0.0025 (examples/ex_ufit.py:103-107), 1000 steps
    net, x, y, dn = config1_data()
    uq = NN_MCMC(net, verbose=False)
    theta0 = np.random.rand(uq.pdim)
    with Recorder() as rec:
        uq_res = _fit_capture(uq, x, y, dn, 1000, theta0, 'hmc', {'L': 3, 'epsilon': 0.0025})
    np.savez_compressed(os.path.join(HERE, 'chain_c1_hmc.npz'), x=x, y=y, sigma=dn, theta0=theta0,
                        p=np.array(rec.p), u=np.array(rec.u), epsilon=0.0025, L=3, **uq_res)
    print('c1 hmc accrate', uq_res['accrate'])

    # B4: MALA (not reachable through NN_MCMC.fit, nn_mcmc.py:130-135) on MLP(1,1,(5,),tanh), N=20
    np.random.seed(3)
    torch.manual_seed(3)
    net = MLP(1, 1, (5,), activ='tanh')
    x = np.random.rand(20, 1) * 2 - 1
    y = np.sin(3 * x) + 0.1 * np.random.randn(20, 1)
    uq = NN_MCMC(net, verbose=False)
    lpinfo = {'model': None, 'xd': x, 'yd': [yy for yy in y], 'ltype': 'classical', 'lparams': {'sigma': 0.1}}
    theta0 = 0.5 * np.random.randn(uq.pdim)
    sam = MALA(epsilon=0.004)
    sam.setLogPost(uq.logpost, uq.logpostgrad, lpinfo=lpinfo)
    with Recorder() as rec:
        res = sam.run(1000, theta0)
    np.savez_compressed(os.path.join(HERE, 'chain_mlp_mala.npz'), x=x, y=y, sigma=0.1, theta0=theta0,
                        p=np.array(rec.p), u=np.array(rec.u), epsilon=0.004,
                        chain=res['chain'], logpost=res['logpost'], alphas=res['alphas'],
                        mapparams=res['mapparams'], maxpost=res['maxpost'], accrate=res['accrate'])
    print('mlp mala accrate', res['accrate'])

    # B5: the reference solver-test shape: MLP(1,1,(5,),tanh), N=20, AMCMC 300 steps (tests/test_solvers.py:16-100)
    sam = AMCMC(gamma=0.1)
    sam.setLogPost(uq.logpost, None, lpinfo=lpinfo)
    with Recorder() as rec:
        res = sam.run(300, theta0)
    np.savez_compressed(os.path.join(HERE, 'chain_mlp_amcmc.npz'), x=x, y=y, sigma=0.1, theta0=theta0,
                        xi=np.array(rec.xi), u=np.array(rec.u), gamma=0.1,
                        chain=res['chain'], logpost=res['logpost'], alphas=res['alphas'],
                        mapparams=res['mapparams'], maxpost=res['maxpost'], accrate=res['accrate'],
                        Xm=sam._Xm, cov=sam._cov, propcov=sam._propcov)
    print('mlp amcmc accrate', res['accrate'])


def _fit_capture(uq, x, y, dn, nmcmc, theta0, sampler, sp):
    """NN_MCMC.fit but keeping the sampler object's result dict and final AMCMC state."""
    captured = {}
    import quinn.solvers.nn_mcmc as mod
    klass = {'amcmc': 'AMCMC', 'hmc': 'HMC'}[sampler]
    orig = getattr(mod, klass)

    class Spy(orig):
        def run(self, **kw):
            res = super().run(**kw)
            captured['res'] = res
            captured['obj'] = self
            return res
    setattr(mod, klass, Spy)
    try:
        uq.fit(x, y, zflag=False, datanoise=dn, nmcmc=nmcmc, param_ini=theta0.copy(), sampler=sampler,
               sampler_params=sp)
    finally:
        setattr(mod, klass, orig)
    res = captured['res']
    out = dict(chain=res['chain'], logpost=res['logpost'], alphas=res['alphas'], mapparams=res['mapparams'],
               maxpost=res['maxpost'], accrate=res['accrate'])
    if sampler == 'amcmc':
        o = captured['obj']
        out.update(Xm=o._Xm, cov=o._cov, propcov=o._propcov)
    # predictive thinning rule on the fitted chain (nn_mcmc.py:180-200)
    xg = np.linspace(-3, 3, 7)[:, None]
    out['xg'] = xg
    out['pred_ens'] = uq.predict_ens(xg, nens=5, nburn=100)
    out['pred_map'] = uq.predict_MAP(xg)
    return out


# ---------------------------------------------------------------- C: variational loss with recorded eps
def gen_vi():
    for name, mk, nsam, N in [('mlp', lambda: MLP(2, 1, (6,), activ='tanh'), 3, 15),
                              ('mlp2', lambda: MLP(3, 2, (8, 5), activ='relu'), 4, 21)]:
        torch.manual_seed(5)
        np.random.seed(5)
        net = mk()
        d = net.indim
        o = net.outdim
        x = np.random.rand(N, d)
        y = np.random.randn(N, o)
        bnet = BNet(net, pi=0.4, sigma1=1.3, sigma2=0.2)
        datanoise, num_batches = 0.07, 3
        bnet.loss_params = [datanoise, nsam, num_batches]
        eps_log = []
        orig = torch.distributions.Normal.sample

        def rec_sample(self, shape=torch.Size()):
            v = orig(self, shape)
            eps_log.append(v.detach().numpy().ravel().copy())
            return v
        torch.distributions.Normal.sample = rec_sample
        try:
            loss = bnet.viloss(tch(x), tch(y))
        finally:
            torch.distributions.Normal.sample = orig
        loss.backward()
        mus = [p for n, p in bnet.named_parameters() if n.endswith('_mu') and not n.startswith('params')]
        rhos = [p for n, p in bnet.named_parameters() if n.endswith('_rho') and not n.startswith('params')]
        # named_parameters() de-duplicates; the registered names come first (bnet.py:70-74)
        nten = len(bnet.rparams)
        assert len(mus) == nten and len(rhos) == nten, (len(mus), len(rhos), nten)
        mu = np.concatenate([p.detach().numpy().ravel() for p in mus])
        rho = np.concatenate([p.detach().numpy().ravel() for p in rhos])
        gmu = np.concatenate([p.grad.numpy().ravel() for p in mus])
        grho = np.concatenate([p.grad.numpy().ravel() for p in rhos])
        eps = np.concatenate(eps_log).reshape(nsam, -1)
        assert eps.shape[1] == mu.size
        np.savez_compressed(os.path.join(HERE, f'vi_{name}.npz'), x=x, y=y, mu=mu, rho=rho, eps=eps,
                            loss=float(loss.item()), gmu=gmu, grho=grho, datanoise=datanoise,
                            num_batches=num_batches, pi=0.4, sigma1=1.3, sigma2=0.2, nsam=nsam)
        print('vi', name, float(loss.item()))


# ---------------------------------------------------------------- D: ensemble predictive
def gen_predict():
    np.random.seed(11)
    torch.manual_seed(11)
    net = MLP(2, 2, (7, 4), activ='tanh')
    x = np.random.rand(30, 2)
    y = np.stack([np.sin(x.sum(1)), np.cos(x[:, 0])], axis=1)
    ens = NN_Ens(net, nens=4, dfrac=0.8)
    ens.fit(x, y, nepochs=3, lrate=0.01, batch_size=10, freq_out=1000, freq_plot=10 ** 9)
    flat = np.array([np.concatenate([p.detach().numpy().ravel() for p in l.best_model.parameters()])
                     for l in ens.learners])
    xt = np.random.rand(9, 2)
    perm_log = []
    orig = np.random.permutation

    def rec_perm(n):
        v = orig(n)
        perm_log.append(np.array(v))
        return v
    np.random.permutation = rec_perm
    try:
        yens = ens.predict_ens(xt)
    finally:
        np.random.permutation = orig
    ymean, yvar, ycov = None, None, None
    np.random.permutation = lambda n: perm_log[0]
    try:
        ymean, yvar, ycov = ens.predict_mom_sample(xt, msc=2, nsam=4)
    finally:
        np.random.permutation = orig
    np.savez_compressed(os.path.join(HERE, 'predict_ens.npz'), x=xt, thetas=flat, perm=perm_log[0], yens=yens,
                        ymean=ymean, yvar=yvar, ycov=ycov)
    print('predict ens', yens.shape)


if __name__ == '__main__':
    gen_logpost()
    gen_chains()
    gen_vi()
    gen_predict()
