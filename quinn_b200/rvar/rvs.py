"""Random-variable descriptions used by the variational model (quinn/rvar/rvs.py:51-173).  As torch
modules they evaluate the reference's formulas with torch ops (API conformance); inside BNet the same
formulas run fused in qb_vi_sample / qb_vi_backward."""
import math

import torch


class RV(torch.nn.Module):
    def sample(self, num_samples=1):
        raise NotImplementedError

    def log_prob(self, x):
        raise NotImplementedError


class Gaussian_1d(RV):
    """N(mu, sigma) with sigma = log(1+exp(rho)) or sigma = exp(logsigma) (rvs.py:66-127)."""

    def __init__(self, mu, rho=None, logsigma=None):
        super().__init__()
        assert (rho is None) != (logsigma is None)
        self.mu, self.rho, self.logsigma = mu, rho, logsigma
        assert (rho if rho is not None else logsigma).shape == mu.shape

    def _sigma(self):
        return torch.log1p(torch.exp(self.rho)) if self.rho is not None else torch.exp(self.logsigma)

    def sample(self):
        sigma = self._sigma()
        return self.mu + sigma * torch.randn(sigma.shape, dtype=sigma.dtype, device=sigma.device)

    def log_prob(self, x):
        sigma = self._sigma()
        return (-math.log(math.sqrt(2 * math.pi)) - torch.log(sigma) - ((x - self.mu) ** 2) / (2 * sigma ** 2)).sum()


class GMM2_1d(RV):
    """pi N(0,sigma1) + (1-pi) N(0,sigma2); log_prob is exp-then-log as in rvs.py:169-171."""

    def __init__(self, pi, sigma1, sigma2):
        super().__init__()
        self.pi, self.sigma1, self.sigma2 = pi, sigma1, sigma2

    @staticmethod
    def _pdf(x, s):
        return torch.exp(-x * x / (2 * s * s) - math.log(s) - 0.5 * math.log(2 * math.pi))

    def log_prob(self, x):
        return torch.log(self.pi * self._pdf(x, self.sigma1) + (1 - self.pi) * self._pdf(x, self.sigma2)).sum()
