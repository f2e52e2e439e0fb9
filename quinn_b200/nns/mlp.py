"""MLP with the reference's constructor (quinn/nns/mlp.py:22-88).  The layer stack it builds
(Linear, activation, ..., Linear[, Expon]) is what quinn_b200.netdesc turns into a kernel descriptor."""
import torch

from .nnbase import MLPBase


class Expon(torch.nn.Module):
    """exp(x) final transform (quinn/nns/nns.py Expon)."""

    def forward(self, x):
        return torch.exp(x)


class Sine(torch.nn.Module):
    def forward(self, x):
        return torch.sin(x)


_ACTIVATIONS = {'tanh': torch.nn.Tanh, 'relu': torch.nn.ReLU, 'sin': Sine}


class MLP(MLPBase):
    def __init__(self, indim, outdim, hls, biasorno=True, activ='relu', bnorm=False, bnlearn=True, dropout=0.0,
                 final_transform=None, device='cpu'):
        super().__init__(indim, outdim, device=device)
        assert len(hls) > 0
        self.nlayers = len(hls)
        self.hls, self.biasorno, self.dropout = hls, biasorno, dropout
        self.bnorm, self.bnlearn, self.final_transform = bnorm, bnlearn, final_transform
        act = _ACTIVATIONS.get(activ, torch.nn.Identity)()
        widths = [indim] + list(hls) + [outdim]
        mods = []
        for l in range(len(widths) - 1):
            if l > 0:
                mods.append(act)                      # one shared activation module, as in mlp.py:65,76
            mods.append(torch.nn.Linear(widths[l], widths[l + 1], bias=biasorno))
            if dropout > 0.0:
                mods.append(torch.nn.Dropout(p=dropout))
            if bnorm:
                mods.append(torch.nn.BatchNorm1d(widths[l + 1], affine=bnlearn))
        if final_transform == 'exp':
            mods.append(Expon())
        self.nnmodel = torch.nn.Sequential(*mods)
        self.to(device)

    def forward(self, x):
        return self.nnmodel(x)
