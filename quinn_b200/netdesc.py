"""Flat-layout network descriptor: torch module -> qb_net_t.

The descriptor pins the flat parameter contract of the reference (quinn/nns/nnwrap.py:64-106):
``nnmodel.parameters()`` order, each tensor flattened C-order, Linear weights (n_out, n_in) row-major.
Supported modules are the ones on the hot path: ``MLP`` (mlp.py:59-86),
``RNet`` with Poly(n) / Lin / Quad / Cubic / NonPar weight parameterisation (rnet.py:124-164, 244-377), ``torch.nn.Linear`` and plain
``torch.nn.Sequential`` stacks of Linear / Tanh / ReLU / Identity.  Anything else (batch norm, dropout,
higher-order Poly, 'sin') raises -- there is no slow path.
"""
from dataclasses import dataclass, field
from typing import List

import numpy as np
import torch

from . import _lib

_ACT = {'identity': _lib.QB_ACT_IDENTITY, 'tanh': _lib.QB_ACT_TANH, 'relu': _lib.QB_ACT_RELU}


@dataclass
class Layer:
    n_in: int
    n_out: int
    w_off: int
    b_off: int = -1
    act: str = 'identity'
    res_step: float = 0.0
    terms: list = None      # polynomial-in-depth weights (rnet.py:244-347): [(coef, w_off_m, b_off_m), ...]; W = sum coef*ww_m


@dataclass
class NetDesc:
    in_dim: int
    out_dim: int
    n_params: int
    layers: List[Layer] = field(default_factory=list)
    final_exp: bool = False

    def to_c(self):
        if len(self.layers) > _lib.QB_MAX_LAYERS:
            raise ValueError(f'quinn_b200 supports at most {_lib.QB_MAX_LAYERS} layers')
        net = _lib.qb_net_t()
        net.n_layers, net.in_dim, net.out_dim, net.n_params = len(self.layers), self.in_dim, self.out_dim, self.n_params
        net.final_exp = 1 if self.final_exp else 0
        for i, L in enumerate(self.layers):
            c = net.layers[i]
            c.n_in, c.n_out, c.w_off, c.b_off = L.n_in, L.n_out, L.w_off, L.b_off
            c.act, c.res_step = _ACT[L.act], float(L.res_step)
            if L.terms:
                if len(L.terms) > _lib.QB_MAX_TERMS:
                    raise NotImplementedError(f'at most {_lib.QB_MAX_TERMS} polynomial terms per layer')
                ws = (L.terms[1][1] - L.terms[0][1]) if len(L.terms) > 1 else 0
                bs = (L.terms[1][2] - L.terms[0][2]) if (len(L.terms) > 1 and L.b_off >= 0) else 0
                for m, (_, wo, bo) in enumerate(L.terms):
                    if wo != L.w_off + m * ws or (L.b_off >= 0 and bo != L.b_off + m * bs):
                        raise ValueError('polynomial terms must be equally spaced tensors starting at w_off / b_off')
                c.n_terms, c.w_stride, c.b_stride = len(L.terms), ws, bs
                for m, t in enumerate(L.terms):
                    c.coef[m] = float(t[0])
        return net

    def as_oracle_layers(self):
        """Same description in the dict form oracle/quinn_oracle.py takes (tests only)."""
        out = []
        for L in self.layers:
            d = dict(n_in=L.n_in, n_out=L.n_out, w_off=L.w_off, b_off=L.b_off, act=L.act, res_step=L.res_step)
            if L.terms:
                d['terms'] = list(L.terms)
            out.append(d)
        return out

    def macs_per_point(self):
        """S of SURVEY.md section 8: sum of n_in*n_out over the executed layers."""
        return sum(L.n_in * L.n_out for L in self.layers)


def _param_offsets(module):
    offs, off = {}, 0
    for name, p in module.named_parameters():
        offs[name] = off
        off += p.numel()
    return offs, off


def _act_of(m):
    if isinstance(m, torch.nn.Tanh):
        return 'tanh'
    if isinstance(m, torch.nn.ReLU):
        return 'relu'
    if isinstance(m, torch.nn.Identity):
        return 'identity'
    return None


def _from_sequential(seq, prefix, offs, layers):
    """Linear [act] Linear [act] ... ; an activation directly follows the Linear it applies to."""
    final_exp = False
    mods = list(seq._modules.items())      # NOT named_children(): MLP reuses one activation module (mlp.py:65,76)
    for name, m in mods:
        if isinstance(m, torch.nn.Linear):
            w = offs[f'{prefix}{name}.weight']
            b = offs.get(f'{prefix}{name}.bias', -1) if m.bias is not None else -1
            layers.append(Layer(m.in_features, m.out_features, w, b, 'identity', 0.0))
        elif _act_of(m) is not None:
            if not layers:
                if _act_of(m) != 'identity':
                    raise NotImplementedError('activation before the first Linear is not supported')
                continue
            if layers[-1].act != 'identity':
                raise NotImplementedError('two activations in a row are not supported')
            layers[-1].act = _act_of(m)
        elif type(m).__name__ == 'Expon':
            if name != mods[-1][0]:
                raise NotImplementedError('Expon is only supported as the final transform')
            final_exp = True
        else:
            raise NotImplementedError(
                f'quinn_b200: module {type(m).__name__} is outside the fused path (only Linear/Tanh/ReLU/Identity/'
                'final Expon; batch norm and dropout carry training-mode state, SURVEY.md section 2 row 4)')
    return final_exp


def netdesc_from_module(module) -> NetDesc:
    """Walk a torch module and describe it for the kernels; raises NotImplementedError outside the path."""
    offs, total = _param_offsets(module)
    layers: List[Layer] = []
    final_exp = False
    if isinstance(module, torch.nn.Linear):
        layers.append(Layer(module.in_features, module.out_features, offs['weight'],
                            offs['bias'] if module.bias is not None else -1))
    elif isinstance(module, torch.nn.Sequential):
        final_exp = _from_sequential(module, '', offs, layers)
    elif hasattr(module, 'wp_function') and hasattr(module, 'rdim'):
        layers, final_exp = _from_rnet(module, offs)
    elif hasattr(module, 'nnmodel') and isinstance(module.nnmodel, torch.nn.Sequential):
        final_exp = _from_sequential(module.nnmodel, 'nnmodel.', offs, layers)
    else:
        raise NotImplementedError(f'quinn_b200: cannot describe module of type {type(module).__name__}')
    if not layers:
        raise NotImplementedError('module has no Linear layers')
    return NetDesc(layers[0].n_in, layers[-1].n_out, total, layers, final_exp)


def _from_rnet(m, offs):
    """RNet.forward (rnet.py:124-164): pre layer (with activation), nlayers+1 residual steps, post layer."""
    wp = m.wp_function
    kind = type(wp).__name__
    poly = False
    if kind == 'Poly' and wp.npar == 1:
        pick = lambda i: 0                                   # noqa: E731  (Poly(0): pars[0]*t**0)
    elif kind == 'NonPar':
        pick = lambda i: int((m.step_size * i) * wp.npar)    # noqa: E731  (rnet.py:377)
    elif kind in ('Poly', 'Lin', 'Quad', 'Cubic') and wp.npar <= _lib.QB_MAX_TERMS:
        # polynomial in the depth variable t_i = i/(nlayers+1) (rnet.py:244-347): the kernels stage W_i = sum_m t_i^m ww_m
        # and scatter dW_i back to every term (chain rule)
        poly, pick = True, (lambda i: 0)                     # noqa: E731
    else:
        raise NotImplementedError(f'RNet weight parameterisation {kind}(npar={wp.npar}) is outside the fused path '
                                  '(SURVEY.md section 8f rank 4)')
    if m.final_layer not in (None, 'exp'):
        raise NotImplementedError(f"RNet final_layer={m.final_layer!r} is outside the fused path")
    act = 'tanh' if isinstance(m.activ, torch.nn.Tanh) else 'identity'
    layers = []
    if m.layer_pre:
        layers.append(Layer(m.indim, m.rdim, offs['weight_pre'], offs['bias_pre'], act, 0.0))
    for i in range(m.nlayers + 1):
        ip = pick(i)
        terms = None
        if poly:
            t = m.step_size * i
            terms = [(t ** k, offs[f'ww_{k}'], offs[f'bb_{k}'] if m.biasorno else -1) for k in range(wp.npar)]
        layers.append(Layer(m.rdim, m.rdim, offs[f'ww_{ip}'], offs[f'bb_{ip}'] if m.biasorno else -1, act,
                            0.0 if m.mlp else float(m.step_size), terms=terms))
    if m.layer_post:
        layers.append(Layer(m.rdim, m.outdim, offs['weight_post'], offs['bias_post'], 'identity', 0.0))
    return layers, m.final_layer == 'exp'


def flatten_module(module) -> np.ndarray:
    """p_flatten (nnwrap.py:64-79) as a float64 numpy vector."""
    return np.concatenate([p.detach().cpu().double().numpy().ravel() for p in module.parameters()])


def unflatten_module(module, flat) -> None:
    """p_unflatten (nnwrap.py:81-106): fill the module's parameters, in parameters() order, from a flat vector."""
    s = 0
    flat = np.asarray(flat)
    for p in module.parameters():
        n = p.numel()
        p.data = torch.as_tensor(flat[s:s + n], dtype=p.dtype, device=p.device).view(p.shape).clone()
        s += n
    if s != flat.size:
        raise ValueError(f'flat vector has {flat.size} entries, the module {s} parameters')
