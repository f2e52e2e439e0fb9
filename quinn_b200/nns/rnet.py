"""ResNet with the reference's constructor and parameter order (quinn/nns/rnet.py:39-164).
Poly(n) (Poly(0) = shared weights, Lin, Quad, Cubic) and NonPar weight parameterisations are all on the fused path."""
import math

import torch
import torch.nn.functional as F

from .nnbase import MLPBase


class LayerFcn:
    npar = None

    def __call__(self, pars, t):
        raise NotImplementedError


class Poly(LayerFcn):
    """Polynomial-in-depth weights sum_i pars[i] t^i (rnet.py:330-347)."""

    def __init__(self, order):
        self.npar = order + 1

    def __call__(self, pars, t):
        assert len(pars) == self.npar
        return sum(p * t ** i for i, p in enumerate(pars))


class Lin(Poly):
    """pars[0] + pars[1] t (rnet.py:244-264)."""

    def __init__(self):
        super().__init__(1)


class Quad(Poly):
    """pars[0] + pars[1] t + pars[2] t^2 (rnet.py:267-288)."""

    def __init__(self):
        super().__init__(2)


class Cubic(Poly):
    """pars[0] + ... + pars[3] t^3 (rnet.py:290-311)."""

    def __init__(self):
        super().__init__(3)


class NonPar(LayerFcn):
    """A separate parameter per layer (rnet.py:349-377)."""

    def __init__(self, npar):
        self.npar = npar

    def __call__(self, pars, t):
        assert len(pars) == self.npar
        return pars[int(t * self.npar)]


class RNet(MLPBase):
    def __init__(self, rdim, nlayers, wp_function=None, indim=None, outdim=None, biasorno=True, nonlin=True,
                 mlp=False, layer_pre=False, layer_post=False, final_layer=None, device='cpu', init_factor=1.0,
                 sum_dim=1):
        super().__init__(indim, outdim, device=device)
        self.indim = rdim if indim is None else indim
        self.outdim = rdim if outdim is None else outdim
        self.rdim, self.nlayers, self.biasorno = rdim, nlayers, biasorno
        self.wp_function = NonPar(nlayers + 1) if wp_function is None else wp_function
        assert isinstance(self.wp_function, LayerFcn)
        self.step_size = 1.0 / (nlayers + 1.0)
        self.mlp, self.layer_pre, self.layer_post = mlp, layer_pre, layer_post
        self.final_layer, self.init_factor, self.sum_dim = final_layer, init_factor, sum_dim
        assert self.indim == rdim or layer_pre
        assert self.outdim == rdim or layer_post

        def uniform(*shape, fan):
            return torch.nn.Parameter(init_factor * (2.0 * torch.rand(*shape) - 1.0) / math.sqrt(fan))
        # registration order fixes the flat layout: pre, post, ww_*, bb_*  (rnet.py:90-111)
        if layer_pre:
            self.weight_pre = uniform(rdim, self.indim, fan=self.indim)
            self.bias_pre = uniform(rdim, fan=self.indim)
        if layer_post:
            self.weight_post = uniform(self.outdim, rdim, fan=rdim)
            self.bias_post = uniform(self.outdim, fan=rdim)
        for ip in range(self.wp_function.npar):
            self.register_parameter(f'ww_{ip}', uniform(rdim, rdim, fan=rdim))
        if biasorno:
            for ip in range(self.wp_function.npar):
                self.register_parameter(f'bb_{ip}', uniform(rdim, fan=rdim))
        self.activ = torch.nn.Tanh() if nonlin else torch.nn.Identity()
        self.to(device)

    def forward(self, x):
        out = x + 0.0
        if self.layer_pre:
            out = self.activ(F.linear(out, self.weight_pre, self.bias_pre))
        ws = [getattr(self, f'ww_{ip}') for ip in range(self.wp_function.npar)]
        bs = [getattr(self, f'bb_{ip}') for ip in range(self.wp_function.npar)] if self.biasorno else None
        for i in range(self.nlayers + 1):
            t = self.step_size * i
            w = self.wp_function(ws, t)
            b = self.wp_function(bs, t) if self.biasorno else None
            z = self.activ(F.linear(out, w, b))
            out = z if self.mlp else out + self.step_size * z
        if self.layer_post:
            out = F.linear(out, self.weight_post, self.bias_post)
        if self.final_layer == 'exp':
            out = torch.exp(out)
        elif self.final_layer == 'logabs':
            out = torch.log(torch.abs(out))
        elif self.final_layer == 'sum':
            out = torch.sum(out, dim=self.sum_dim)
        return out
