"""NegLogPost / NegLogPrior with the reference's constructors (quinn/nns/losses.py:152-256).

Called as torch modules they evaluate the formula with torch ops on whatever device the tensors live
(needed by nnfit(loss_fn='logpost') training, which is outside the hot path); handed to
NNWrap.calc_loss / calc_lossgrad or used by NN_MCMC they are only *descriptions* of the likelihood and
prior, and the arithmetic runs in CUDA kernels 1 and 2."""
import numpy as np
import torch

from .tchutils import tch


class NegLogPrior(torch.nn.Module):
    def __init__(self, sigma, anchor):
        super().__init__()
        self.sigma = tch(float(sigma), rgrad=False)
        self.pi = tch(np.pi, rgrad=False)
        self.anchor = anchor

    def forward(self, model):
        flat = torch.cat([p.flatten() for p in model.parameters()])
        anchor = self.anchor.to(flat.device)
        sigma = self.sigma.to(flat.device)
        return torch.sum((flat - anchor) ** 2) / 2 / sigma ** 2 + (flat.numel() / 2) * torch.log(2 * self.pi.to(flat.device) * sigma ** 2)


class NegLogPost(torch.nn.Module):
    def __init__(self, nnmodel, fulldatasize, sigma, priorparams):
        super().__init__()
        self.nnmodel = nnmodel
        self.sigma = tch(float(sigma), rgrad=False)
        self.priorparams = priorparams
        self.pi = tch(np.pi, rgrad=False)
        self.fulldatasize = fulldatasize

    def forward(self, inputs, targets):
        pred = self.nnmodel(inputs)
        sigma, pi = self.sigma.to(pred.device), self.pi.to(pred.device)
        n = len(pred)                                   # N, not N*o (losses.py:199-200)
        val = 0.5 * torch.sum((targets - pred) ** 2) / sigma ** 2 + (n / 2) * torch.log(2 * pi) + n * torch.log(sigma)
        if self.priorparams is not None:
            prior = NegLogPrior(self.priorparams['sigma'], self.priorparams['anchor'])
            val = val + n * prior(self.nnmodel) / self.fulldatasize
        return val
