"""NN_MCMC with the reference's interface (quinn/solvers/nn_mcmc.py:15-200).

``logpost`` / ``logpostgrad`` evaluate in CUDA kernels 1 / 2; ``fit`` runs the fused chain-step kernel 3
for one chain (reference behaviour) or for K chains when ``param_ini`` is (K,P) / ``nchains`` is given;
the predictive calls run kernel 4."""
import copy

import numpy as np
import torch

from .. import ops
from ..mcmc.admcmc import AMCMC
from ..mcmc.hmc import HMC
from ..mcmc.mala import MALA
from ..mcmc.mcmc import DeviceLogPost
from ..netdesc import netdesc_from_module
from ..nns.nnwrap import NNWrap, nn_p
from .quinn import QUiNNBase


class NN_MCMC(QUiNNBase):
    def __init__(self, nnmodel, verbose=True, dtype=torch.float64, device='cuda'):
        """``dtype`` (extension): arithmetic type of the kernels; float64 reproduces the reference
        (tchutils.py:9), float32 is the throughput mode."""
        super().__init__(nnmodel)
        self.verbose = verbose
        self.pdim = sum(p.numel() for p in self.nnmodel.parameters())
        print("Number of parameters:", self.pdim)
        if self.verbose:
            self.print_params(names_only=True)
        self.samples = None
        self.cmode = None
        self.lpinfo = {}
        self.dtype, self.device = dtype, device
        self.desc = netdesc_from_module(self.nnmodel)
        self._problems = {}

    # ---- device handle
    def device_logpost(self, lpinfo):
        """GPU-resident log-posterior for this lpinfo dict (built once, unlike nn_mcmc.py:55-66)."""
        if lpinfo['ltype'] != 'classical':
            raise ValueError('Likelihood type is not recognized.')
        key = id(lpinfo)
        h = self._problems.get(key)
        if h is None:
            prob = ops.Problem(self.desc, np.asarray(lpinfo['xd']), np.asarray(lpinfo['yd']),
                               lpinfo['lparams']['sigma'], dtype=self.dtype, device=self.device)
            h = DeviceLogPost(prob)
            self._problems = {key: h}
            self._problems_ref = lpinfo      # keep the dict alive so its id stays unique
        return h

    def logpost(self, modelpars, lpinfo):
        """log p(theta | D) as a float for theta (P,), or an array (K,) for theta (K,P) (nn_mcmc.py:45-71)."""
        modelpars = np.asarray(modelpars, dtype=np.float64)
        lp = ops.logpost(self.device_logpost(lpinfo).problem, modelpars)
        return float(lp[0].item()) if modelpars.ndim == 1 else lp.cpu().numpy()

    def logpostgrad(self, modelpars, lpinfo):
        """Gradient of the log-posterior, (P,) or (K,P) (nn_mcmc.py:73-98)."""
        modelpars = np.asarray(modelpars, dtype=np.float64)
        _, g = ops.logpost_grad(self.device_logpost(lpinfo).problem, modelpars)
        g = g.double().cpu().numpy()
        return g[0] if modelpars.ndim == 1 else g

    def fit(self, xtrn, ytrn, zflag=True, datanoise=0.05, nmcmc=6000, param_ini=None, sampler='amcmc',
            sampler_params=None, *, nchains=None, seed=None, store_every=1, replay=None, chain_offset=0):
        """Sample the posterior of the flat parameters (nn_mcmc.py:100-139).

        Extensions (keyword-only): ``nchains`` / a (K,P) ``param_ini`` run K chains at once (``samples``
        becomes (K, M+1, P)); ``seed``, ``store_every``, ``replay``, ``chain_offset`` are passed to the sampler.
        With ``zflag`` the start point is refined by BFGS on -logpost using the analytic gradient of kernel 2
        (the reference uses finite differences, nn_mcmc.py:126)."""
        assert xtrn.shape[0] == ytrn.shape[0]
        self.lpinfo = {'model': nn_p, 'xd': xtrn, 'yd': [y for y in ytrn], 'ltype': 'classical',
                       'lparams': {'sigma': datanoise}}
        if param_ini is None:
            shape = (self.pdim,) if nchains is None else (nchains, self.pdim)
            param_ini = np.random.rand(*shape)
            if zflag:
                param_ini = self._map_start(param_ini)
        sampler_params = {} if sampler_params is None else sampler_params
        if sampler == 'amcmc':
            mymcmc = AMCMC(**sampler_params)
            mymcmc.setLogPost(self.logpost, None, lpinfo=self.lpinfo)
        elif sampler == 'hmc':
            mymcmc = HMC(**sampler_params)
            mymcmc.setLogPost(self.logpost, self.logpostgrad, lpinfo=self.lpinfo)
        elif sampler == 'mala':                     # promised by the docstring at nn_mcmc.py:110
            mymcmc = MALA(**sampler_params)
            mymcmc.setLogPost(self.logpost, self.logpostgrad, lpinfo=self.lpinfo)
        else:
            raise ValueError(f"sampler {sampler!r} is not one of 'amcmc', 'hmc', 'mala'")
        res = mymcmc.run(param_ini=param_ini, nmcmc=nmcmc, seed=seed, store_every=store_every, replay=replay,
                         chain_offset=chain_offset, verbose=self.verbose)
        self.sampler_obj, self.mcmc_results = mymcmc, res
        self.samples, self.cmode = res['chain'], res['mapparams']
        return res

    def _map_start(self, param_ini):
        from scipy.optimize import minimize

        def fg(th):
            lp, g = ops.logpost_grad(self.device_logpost(self.lpinfo).problem, th)
            return -float(lp[0].item()), -g[0].double().cpu().numpy()
        rows = np.atleast_2d(param_ini)
        if rows.shape[0] > self.map_batched_above:
            return self._map_start_batched(rows)
        out = np.array([minimize(fg, r, jac=True, method='BFGS', options={'gtol': 1e-13}).x for r in rows])
        return out[0] if np.ndim(param_ini) == 1 else out

    map_batched_above = 8          # more start points than this: one batched optimisation instead of K BFGS runs
    map_start_steps = 300          # Adam ascent steps of the batched MAP start

    def _map_start_batched(self, rows, nsteps=None, lrate=0.01):
        """MAP pre-conditioning of MANY chains at once (the reference has one chain and BFGS with
        finite differences, nn_mcmc.py:125-127): Adam ascent on log p(theta|D) for all K starts together - kernel 2,
        qb_adam_step and qb_copy_rows_where, nothing leaves the device - returning the best state each chain visited."""
        import ctypes as C
        from .. import _lib
        from ..ops import _ptr, _stream, qb_dtype
        nsteps = self.map_start_steps if nsteps is None else nsteps
        prob = self.device_logpost(self.lpinfo).problem
        lib = _lib.load()
        theta = prob.theta(rows).clone()
        K, P = theta.shape
        m, v, best = torch.zeros_like(theta), torch.zeros_like(theta), theta.clone()
        best_lp = torch.full((K,), -float('inf'), dtype=torch.float64, device=theta.device)
        lp = torch.empty(K, dtype=torch.float64, device=theta.device)
        g = torch.empty_like(theta)
        qdt = qb_dtype(prob.dtype)
        for step in range(1, nsteps + 2):
            ops.logpost_grad(prob, theta, lp, g)
            better = lp > best_lp
            mask = better.to(torch.uint8)
            with torch.cuda.device(theta.device):
                _lib.check(lib.qb_copy_rows_where(qdt, _ptr(best), _ptr(theta), _ptr(mask), K, P, _stream()), 'qb_copy_rows_where')
            best_lp = torch.where(better, lp, best_lp)
            if step > nsteps:
                break
            with torch.cuda.device(theta.device):
                _lib.check(lib.qb_adam_step(qdt, _ptr(theta), _ptr(g), _ptr(m), _ptr(v), K * P, lrate, 0.9, 0.999, 1e-8, 0.0,
                                            step, -1.0, _stream()), 'qb_adam_step')
        self.map_start_logpost = best_lp.cpu().numpy()
        return best.double().cpu().numpy()

    # ---- predictive (nn_mcmc.py:142-200)
    def get_best_model(self, param):
        nnw = NNWrap(self.nnmodel)
        nnw.p_unflatten(param)
        return copy.deepcopy(nnw.nnmodel)

    def _forward(self, thetas, x):
        out, _, _ = ops.predict(self.desc, np.asarray(thetas, dtype=np.float64), np.asarray(x), dtype=self.dtype,
                                device=self.device)
        return out.double().cpu().numpy()

    def predict_MAP(self, x):
        cm = np.asarray(self.cmode)
        return self._forward(cm, x)[0] if cm.ndim == 1 else self._forward(cm, x)

    def predict_sample(self, x, param):
        return self._forward(param, x)[0]

    def _thinned(self, nens, nburn):
        """Rows samples[nburn + j*nevery], nevery = int((M'-nburn)/nens) (nn_mcmc.py:194-196); for a
        multi-chain run the same rows of every chain, chain-major."""
        s = np.asarray(self.samples)
        nevery = int((s.shape[-2] - nburn) / nens)
        rows = [nburn + j * nevery for j in range(nens)]
        return s[..., rows, :].reshape(-1, s.shape[-1])

    def predict_ens(self, x, nens=10, nburn=1000):
        return self._forward(self._thinned(nens, nburn), x)

    def _ens_thetas(self, nens):
        # predict_mom_sample -> base predict_ens(x, nens=nsam) in the reference calls the override with
        # nburn=1000 (quinn.py:84 -> nn_mcmc.py:180)
        return self.desc, self._thinned(nens, 1000), self.dtype
