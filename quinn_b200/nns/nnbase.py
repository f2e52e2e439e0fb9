"""MLPBase: the torch.nn.Module base of the reference's networks (quinn/nns/nnbase.py:19-115), with
``predict`` routed through posterior-predictive kernel 4.  Plot helpers are out of scope."""
import numpy as np
import torch

from .tchutils import torch as _t  # noqa: F401  (sets the double default dtype like the reference)


class MLPBase(torch.nn.Module):
    def __init__(self, indim, outdim, device='cpu'):
        super().__init__()
        self.indim = indim
        self.outdim = outdim
        self.best_model = None
        self.trained = False
        self.history = None
        self.device = device

    def forward(self, x):
        raise NotImplementedError

    def predict(self, x):
        """numpy (N,d) -> numpy (N,o); uses best_model once trained (nnbase.py:59-84).  Runs kernel 4."""
        from .nnwrap import device_forward
        model = self.best_model if self.trained else self
        return device_forward(model, np.asarray(x))

    def numpar(self):
        return sum(p.numel() for p in self.parameters())

    def fit(self, xtrn, ytrn, **kwargs):
        """nnbase.py:95-115: train with nnfit, remember the best model."""
        from .nnfit import nnfit
        fit_info = nnfit(self, xtrn, ytrn, **kwargs)
        object.__setattr__(self, 'best_model', fit_info['best_nnmodel'])
        self.history = fit_info['history']
        self.trained = True
        return self.best_model

    def printParams(self):
        for name, param in self.named_parameters():
            print(name, param.data)
