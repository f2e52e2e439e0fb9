"""Metropolis-adjusted Langevin with the reference's constructor (quinn/mcmc/mala.py:15-22)."""
import torch

from .. import ops
from .mcmc import MCMCBase


class MALA(MCMCBase):
    def __init__(self, epsilon=0.05):
        super().__init__()
        self.epsilon = epsilon

    def _device_sampler_state(self, st):
        return ops.HmcState(st, epsilon=self.epsilon, L=1, method='mala')

    def _device_advance(self, st, samp, nsteps, rec, kw):
        ops.hmc_run(st, samp, nsteps, rec, **kw)

    def sampler(self, current, imcmc):
        """mala.py:42-51, batched over chains (generic-callable adapter)."""
        assert self.logPostGrad is not None
        grad = lambda th: self._eval_generic(self.logPostGrad, th)     # noqa: E731
        p = torch.randn(current.shape, dtype=current.dtype, device=current.device, generator=self._gen)
        g0 = grad(current)
        prop = current + 0.5 * self.epsilon ** 2 * g0 + self.epsilon * p
        g1 = grad(prop)
        K_cur = p.square().sum(1) / 2
        p = p + self.epsilon * (g0 + g1) / 2
        return prop, K_cur, p.square().sum(1) / 2
