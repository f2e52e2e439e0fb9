"""Device-side post-processing of chains and predictive ensembles: the arithmetic runs in the
kernels of quinn_b200/csrc/qb_post.cu through the C ABI; tensors stay on the GPU.

  get_stats(yy, qt)          quinn/utils/stats.py:8-32 on a CUDA (M, ...) array: (mean, std, std) or (median, q50-q25, q75-q50)
  quantiles(y, q)            quantiles over the leading axis (numpy's default linear interpolation)
  row_moments(x)             per-chain mean / variance of a recorded scalar [K, n]
  rhat(x)                    Gelman-Rubin R-hat over chains (all ranks when torch.distributed is initialised)
  ess(x)                     effective sample size per chain (Geyer's initial positive sequence)
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from .ops import _ptr, _stream, qb_dtype


def _dev(t, dtype=None):
    t = torch.as_tensor(t)
    if not t.is_cuda:
        t = t.cuda()
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous()


def row_moments(x):
    """x: [K, n] -> (mean[K], var[K] ddof=1), float64 CUDA tensors."""
    x = _dev(x, torch.float64)
    K, n = x.shape
    mean = torch.empty(K, dtype=torch.float64, device=x.device)
    var = torch.empty(K, dtype=torch.float64, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().qb_row_moments(_ptr(x), K, n, _ptr(mean), _ptr(var), _stream()), 'qb_row_moments')
    return mean, var


def rhat(x):
    """R-hat of a scalar recorded in K chains for n steps, x: [K, n] (this rank's chains); cross-rank sums are all-reduced."""
    from . import dist
    x = _dev(x, torch.float64)
    m, v = row_moments(x)
    return dist.rhat(m, v, x.shape[1])


def ess(x, max_lag=0, return_tau=False):
    """Effective sample size of every row of x[K, n] (float64 CUDA tensor [K]); max_lag <= 0: up to n - 1."""
    x = _dev(x, torch.float64)
    K, n = x.shape
    out = torch.empty(K, dtype=torch.float64, device=x.device)
    tau = torch.empty(K, dtype=torch.float64, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().qb_ess(_ptr(x), K, n, int(max_lag), _ptr(out), _ptr(tau), _stream()), 'qb_ess')
    return (out, tau) if return_tau else out


def quantiles(y, q):
    """Quantiles q (sequence, at most 8 values) over the leading axis of a CUDA array y[M, ...] -> [len(q), ...]."""
    y = _dev(y)
    if y.dtype not in (torch.float32, torch.float64):
        y = y.double()
    M = y.shape[0]
    flat = y.reshape(M, -1).contiguous()
    n = flat.shape[1]
    qs = (C.c_double * len(q))(*[float(v) for v in q])
    out = torch.empty((len(q), n), dtype=y.dtype, device=y.device)
    with torch.cuda.device(y.device):
        _lib.check(_lib.load().qb_quantiles(qb_dtype(y.dtype), _ptr(flat), M, n, qs, len(q), _ptr(out), _stream()), 'qb_quantiles')
    return out.reshape((len(q),) + tuple(y.shape[1:]))


def get_stats(yy, qt):
    """quinn/utils/stats.py:8-32 for a CUDA array of sampled predictions yy[M, ...]: (median, q50-q25, q75-q50) when qt,
    else (mean, std, std) with numpy's population std (ddof = 0).  Returns CUDA tensors."""
    yy = _dev(yy)
    if qt:
        q = quantiles(yy, (0.25, 0.5, 0.75))
        return q[1], q[1] - q[0], q[2] - q[1]
    M = yy.shape[0]
    flat = yy.reshape(M, -1).double().t().contiguous()          # rows = points: per-row moments kernel
    mean, var1 = row_moments(flat)
    std = torch.sqrt(var1 * ((M - 1.0) / M)) if M > 1 else torch.zeros_like(mean)
    shape = tuple(yy.shape[1:])
    return mean.reshape(shape), std.reshape(shape), std.reshape(shape)


def fisher_diag(g):
    """out[p] = mean_k g[k, p]^2 (float64): per-point loss gradients -> diagonal Fisher (nnwrap.py:204-229)."""
    g = _dev(g)
    K, P = g.shape
    out = torch.empty(P, dtype=torch.float64, device=g.device)
    with torch.cuda.device(g.device):
        _lib.check(_lib.load().qb_colsq_mean(qb_dtype(g.dtype), _ptr(g), K, P, _ptr(out), _stream()), 'qb_colsq_mean')
    return out
