"""BNet: Bayes-by-backprop variational model with the reference's interface (quinn/vi/bnet.py:10-232).

Per parameter tensor it owns ``<name>_mu`` and ``<name>_rho`` (sigma = exp(rho): bnet.py:80 passes rho as
``logsigma``).  The reference samples, rebinds and forwards one Monte-Carlo weight sample at a time
(bnet.py:202-205); here all ``nsam`` samples are drawn by qb_vi_sample, pushed through kernels 1/2
(K = nsam parameter vectors) and folded back onto (mu, rho) by qb_vi_backward, wrapped in a
torch.autograd.Function so ``nnfit`` and torch optimisers work unchanged."""
import copy
import math

import numpy as np
import torch

from .. import ops
from ..netdesc import netdesc_from_module
from ..rvar.rvs import Gaussian_1d, GMM2_1d


class _ElboTerms(torch.autograd.Function):
    """(mu_flat, rho_flat) -> (mean_s log q, mean_s log p, sum_s ssq_s); everything on the GPU."""

    @staticmethod
    def forward(ctx, mu, rho, bnet, x, y, nsam, eps):
        prob = ops.Problem(bnet.desc, x, y, 1.0, dtype=mu.dtype, device=mu.device)
        bnet._step += 1
        w, eps, logq, logp = ops.vi_sample(mu.detach(), rho.detach(), nsam, bnet.pi, bnet.sigma1, bnet.sigma2, eps=eps,
                                           seed=0 if eps is not None else bnet.seed, step=bnet._step)
        B = x.shape[0]
        need_grad = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]     # False under torch.no_grad()
        if need_grad:
            lp, glp = ops.logpost_grad(prob, w)
            ctx.save_for_backward(mu.detach(), rho.detach(), eps, w, glp)
            ctx.bnet, ctx.nsam = bnet, nsam
        else:
            lp = ops.logpost(prob, w)
        ssq = -2.0 * (lp + 0.5 * B * math.log(2.0 * math.pi))      # lp(sigma=1) = -ssq/2 - B/2 log 2pi
        bnet.last_eps = eps
        return logq.mean().to(mu.dtype), logp.mean().to(mu.dtype), ssq.sum().to(mu.dtype)

    @staticmethod
    def backward(ctx, g_q, g_p, g_s):
        mu, rho, eps, w, glp = ctx.saved_tensors
        b, nsam = ctx.bnet, ctx.nsam
        gmu, grho = ops.vi_backward(mu, rho, eps, w, glp, b.pi, b.sigma1, b.sigma2, float(g_s), float(g_p) / nsam,
                                    float(g_q) / nsam)
        return gmu, grho, None, None, None, None, None


class BNet(torch.nn.Module):
    def __init__(self, nnmodel, pi=0.5, sigma1=1.0, sigma2=1.0, mu_init_lower=-0.2, mu_init_upper=0.2,
                 rho_init_lower=-5.0, rho_init_upper=-4.0, device='cuda', seed=None):
        super().__init__()
        assert isinstance(nnmodel, torch.nn.Module)
        self.nnmodel_ref = [copy.deepcopy(nnmodel)]          # kept out of the module tree (no extra parameters)
        self.desc = netdesc_from_module(nnmodel)
        self.device = device
        self.pi, self.sigma1, self.sigma2 = float(pi), float(sigma1), float(sigma2)
        self.param_names, self.rparams, self.param_priors, self.shapes = [], [], [], []
        plist = []
        for name, param in nnmodel.named_parameters():
            if not param.requires_grad:
                raise NotImplementedError('BNet on the fused path needs every parameter to be variational')
            mu = torch.nn.Parameter(torch.empty(param.shape).uniform_(mu_init_lower, mu_init_upper))
            rho = torch.nn.Parameter(torch.empty(param.shape).uniform_(rho_init_lower, rho_init_upper))
            self.register_parameter(name.replace('.', '_') + '_mu', mu)
            self.register_parameter(name.replace('.', '_') + '_rho', rho)
            plist += [mu, rho]
            self.rparams.append(Gaussian_1d(mu, logsigma=rho))
            self.param_priors.append(GMM2_1d(pi, sigma1, sigma2))
            self.param_names.append(name)
            self.shapes.append(tuple(param.shape))
        self.params = torch.nn.ParameterList(plist)
        self.nparams = len(self.rparams)
        self.log_prior = 0.0
        self.log_variational_posterior = 0.0
        self.loss_params = None
        self.seed = int(np.random.randint(1, 2 ** 31 - 1)) if seed is None else int(seed)
        self._step = 0
        self.last_eps = None
        self.to(device)

    # ---- flat views of the variational parameters (flat layout of nnwrap.py:64-106)
    def flat_mu(self):
        return torch.cat([self.params[2 * i].flatten() for i in range(self.nparams)])

    def flat_rho(self):
        return torch.cat([self.params[2 * i + 1].flatten() for i in range(self.nparams)])

    def sample_weights(self, nsam, eps=None):
        """(nsam, P) weight samples mu + exp(rho)*eps drawn on the device."""
        self._step += 1
        w, eps, logq, logp = ops.vi_sample(self.flat_mu().detach(), self.flat_rho().detach(), nsam, self.pi, self.sigma1,
                                           self.sigma2, eps=eps, seed=0 if eps is not None else self.seed, step=self._step)
        return w, logq, logp

    def forward(self, x, sample=False, par_samples=None):
        """Network output (N,o) at a sampled weight vector (training / sample=True) or at mu (bnet.py:131-178)."""
        if self.training or sample:
            assert par_samples is None
            w, logq, logp = self.sample_weights(1)
            theta = w[0]
            if self.training:
                self.log_prior, self.log_variational_posterior = logp[0], logq[0]
        else:
            theta = self.flat_mu().detach() if par_samples is None else torch.cat([p.flatten() for p in par_samples])
            self.log_prior, self.log_variational_posterior = 0, 0
        out, _, _ = ops.predict(self.desc, theta, x, dtype=theta.dtype, device=theta.device)
        return out[0]

    def sample_elbo(self, x, target, nsam, likparams=None, eps=None):
        """(log_prior, log_variational_posterior, negative_log_likelihood), differentiable wrt mu/rho (bnet.py:181-217)."""
        B, o = target.shape
        assert x.shape[0] == B
        mu, rho = self.flat_mu(), self.flat_rho()
        x = ops.as_device(x, mu.dtype, mu.device)
        target = ops.as_device(target, mu.dtype, mu.device)
        logq, logp, ssq = _ElboTerms.apply(mu, rho, self, x, target, int(nsam), eps)
        sd = float(likparams[0])
        nll = B * math.log(sd) + 0.5 * B * math.log(2.0 * math.pi) + (0.5 * B / (nsam * B * o) / sd ** 2) * ssq
        return logp, logq, nll

    def viloss(self, data, target, eps=None):
        """(log q - log p)/num_batches + NLL (bnet.py:219-232)."""
        datanoise, nsam, num_batches = self.loss_params
        log_prior, log_q, nll = self.sample_elbo(data, target, nsam, likparams=[datanoise], eps=eps)
        return (log_q - log_prior) / num_batches + nll
