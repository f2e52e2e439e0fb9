"""Hamiltonian Monte Carlo with the reference's constructor (quinn/mcmc/hmc.py:16-25)."""
import torch

from .. import ops
from .mcmc import MCMCBase


class HMC(MCMCBase):
    def __init__(self, epsilon=0.05, L=3):
        super().__init__()
        self.epsilon, self.L = epsilon, L

    def _device_sampler_state(self, st):
        return ops.HmcState(st, epsilon=self.epsilon, L=self.L, method='hmc')

    def _device_advance(self, st, samp, nsteps, rec, kw):
        ops.hmc_run(st, samp, nsteps, rec, **kw)

    def sampler(self, current, imcmc):
        """Leapfrog of hmc.py:43-68, batched over chains (generic-callable adapter)."""
        assert self.logPostGrad is not None
        grad = lambda th: self._eval_generic(self.logPostGrad, th)     # noqa: E731
        p = torch.randn(current.shape, dtype=current.dtype, device=current.device, generator=self._gen)
        K_cur = p.square().sum(1) / 2
        prop = current.clone()
        p = p + self.epsilon * grad(prop) / 2
        for jj in range(self.L):
            prop = prop + self.epsilon * p
            if jj != self.L - 1:
                p = p + self.epsilon * grad(prop)
        p = p + self.epsilon * grad(prop) / 2
        return prop, K_cur, p.square().sum(1) / 2
