"""Batched training of all ensemble members at once.

The reference trains the members of NN_Ens one after the other (quinn/solvers/nn_ens.py:51-69 ->
quinn/ens/learner.py:59-73 -> quinn/nns/nnfit.py:125-166): per member an Adam loop over random minibatches of its own
data subset, keeping the parameters with the smallest validation loss seen *before* each update.  Here the flat
parameters of all K members live in one [K,P] device array and every iteration is

    kernel 2 with per-member data (qb_logpost_members)  ->  gradient of the MSE of every member's minibatch
    [kernel 1 with per-member / shared data             ->  validation loss, when it is not the minibatch loss itself]
    qb_copy_rows_where                                   ->  best-model bookkeeping (nnfit.py:147-152)
    qb_adam_step                                         ->  torch.optim.Adam update (nnfit.py:92-93)

MSE (nnfit.py:70, torch.nn.MSELoss mean over N*o) is obtained from the Gaussian log-posterior kernels with sigma = 1:
lp = -SSE/2 - N/2 log(2 pi)  =>  MSE = -(2 lp + N log 2 pi)/(N o),  grad MSE = -2 grad lp / (N o).
Index bookkeeping (subsets, minibatch gathers) is torch plumbing on the device; no arithmetic of the path runs on the
host and there is no CPU fallback.
"""
import ctypes as C
import math

import numpy as np
import torch

from .. import _lib
from ..ops import _ptr, _stream, as_device, qb_dtype

LOG_2PI = math.log(2.0 * math.pi)


class _MemberEval:
    """qb_logpost_members on [K, n, d] / [K, n, o] (per-member) or [n, d] / [n, o] (shared) data."""

    def __init__(self, desc, K, dtype, device):
        self.desc, self.K, self.dtype, self.device = desc, K, dtype, device
        self.cnet = desc.to_c()
        self.qdt = qb_dtype(dtype)
        self.lik = _lib.qb_lik_t(1.0, 0.0, 1.0, None, 0, 0)
        self.lib = _lib.load()
        self.lp = torch.empty(K, dtype=torch.float64, device=device)
        self.grad = torch.empty((K, desc.n_params), dtype=dtype, device=device)
        self.ws = None

    def __call__(self, theta, x, y, want_grad):
        n = x.shape[-2]
        per_member = x.dim() == 3
        xs = n * x.shape[-1] if per_member else 0
        ys = n * y.shape[-1] if per_member else 0
        need = self.lib.qb_eval_workspace_bytes(C.byref(self.cnet), self.qdt, self.K, n, 1 if want_grad else 0)
        if self.ws is None or self.ws.numel() < need:
            self.ws = torch.empty(max(int(need), 256), dtype=torch.uint8, device=self.device)
        data = _lib.qb_data_t(_ptr(x), _ptr(y), n)
        with torch.cuda.device(self.device):
            rc = self.lib.qb_logpost_members(C.byref(self.cnet), self.qdt, _ptr(theta), self.K, C.byref(data), xs, ys,
                                             C.byref(self.lik), _ptr(self.lp), _ptr(self.grad) if want_grad else None,
                                             _ptr(self.ws), self.ws.numel(), _stream())
        _lib.check(rc, 'qb_logpost_members')
        # MSE of every member from the sigma=1 log-posterior
        return -(2.0 * self.lp + n * LOG_2PI) / (n * self.desc.out_dim)


def fit_members(desc, theta0, xtrn, ytrn, subsets, val=None, nepochs=5000, lrate=0.1, wd=0.0, batch_size=None,
                perms=None, dtype=torch.float64, device='cuda', freq_out=100, verbose=True):
    """Train K members together.  theta0: [K,P] (or [P], replicated); subsets: [K, nsub] integer rows of xtrn that
    member k trains on (nn_ens.py:62-63); val: None (a member's own subset, nnfit.py:108-109) or (xval, yval) shared by
    all members; perms: optional [K, nepochs, nsub] minibatch orders (otherwise drawn with torch.randperm in the
    reference's member-major order when that array is small, else epoch by epoch).
    Returns dict(best_theta [K,P], best_loss [K], best_epoch [K], theta [K,P], history [niter, K] of validation MSE)."""
    if not torch.cuda.is_available():
        raise RuntimeError('batched ensemble training needs a CUDA device (there is no CPU fallback)')
    device = torch.device(device)
    lib = _lib.load()
    subsets = torch.as_tensor(np.asarray(subsets), dtype=torch.long, device=device)
    K, nsub = subsets.shape
    P = desc.n_params
    th0 = as_device(theta0, dtype, device)
    theta = (th0[None, :].repeat(K, 1) if th0.dim() == 1 else th0.clone()).contiguous()
    if theta.shape != (K, P):
        raise ValueError(f'theta0 must be [P] or [K,P] with K={K}, P={P}')
    x_all = as_device(xtrn, dtype, device)
    y_all = as_device(np.asarray(ytrn).reshape(len(ytrn), -1), dtype, device)
    xs = x_all[subsets].contiguous()                  # [K, nsub, d]: every member's own training subset
    ys = y_all[subsets].contiguous()
    if val is None:
        xv, yv = xs, ys
    else:
        xv = as_device(np.asarray(val[0]), dtype, device).contiguous()
        yv = as_device(np.asarray(val[1]).reshape(len(val[1]), -1), dtype, device).contiguous()
    if batch_size is None or batch_size > nsub:
        batch_size = nsub
    full_batch = batch_size == nsub
    starts = list(range(0, nsub, batch_size))
    if perms is not None:
        perms = torch.as_tensor(np.asarray(perms), dtype=torch.long, device=device)
    elif not full_batch and K * nepochs * nsub <= 5 * 10 ** 7:
        # the reference's order of torch.randperm calls: all epochs of member 0, then member 1, ... (nn_ens.py:56-66)
        perms = torch.stack([torch.stack([torch.randperm(nsub) for _ in range(nepochs)]) for _ in range(K)]).to(device)
    ev = _MemberEval(desc, K, dtype, device)
    qdt = qb_dtype(dtype)
    m = torch.zeros_like(theta)
    v = torch.zeros_like(theta)
    best_theta = theta.clone()
    best_loss = torch.full((K,), 1.0e100, dtype=torch.float64, device=device)
    best_epoch = torch.zeros(K, dtype=torch.long, device=device)
    history = []
    karange = torch.arange(K, device=device)[:, None]
    step = 0
    for t in range(nepochs):
        if full_batch:
            perm_t = None
        elif perms is not None:
            perm_t = perms[:, t]
        else:
            perm_t = torch.stack([torch.randperm(nsub) for _ in range(K)]).to(device)
        for i in starts:
            if full_batch:
                xb, yb = xs, ys
            else:
                idx = perm_t[:, i:i + batch_size]                     # [K, b] positions inside each member's subset
                xb = xs[karange, idx].contiguous()
                yb = ys[karange, idx].contiguous()
            loss_trn = ev(theta, xb, yb, True)                        # fills ev.grad = d lp / d theta
            nb = xb.shape[1]
            if val is None and full_batch:
                crit = loss_trn                                       # validation data == the minibatch
            else:
                grad_keep = ev.grad
                crit = ev(theta, xv, yv, False).clone()
                assert ev.grad is grad_keep
            better = crit < best_loss                                 # before the update (nnfit.py:143-152)
            mask = better.to(torch.uint8)
            with torch.cuda.device(device):
                _lib.check(lib.qb_copy_rows_where(qdt, _ptr(best_theta), _ptr(theta), _ptr(mask), K, P, _stream()),
                           'qb_copy_rows_where')
            best_loss = torch.where(better, crit, best_loss)
            best_epoch = torch.where(better, torch.full_like(best_epoch, t), best_epoch)
            history.append(crit.clone())
            step += 1
            with torch.cuda.device(device):
                _lib.check(lib.qb_adam_step(qdt, _ptr(theta), _ptr(ev.grad), _ptr(m), _ptr(v), K * P, float(lrate), 0.9,
                                            0.999, 1e-8, float(wd), step, -2.0 / (nb * desc.out_dim), _stream()),
                           'qb_adam_step')
        if verbose and freq_out and ((t + 1) % freq_out == 0 or t == 0 or t == nepochs - 1):
            print(f'{t + 1:>10}{step:>10}   validation MSE: mean {history[-1].mean().item():.6f} '
                  f'best {best_loss.mean().item():.6f}', flush=True)
    return dict(best_theta=best_theta, best_loss=best_loss, best_epoch=best_epoch, theta=theta,
                history=torch.stack(history) if history else torch.empty((0, K), device=device))
