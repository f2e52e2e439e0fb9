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

    def __init__(self, desc, K, dtype, device, logpost=None):
        """logpost = None: MSE loss (from the sigma = 1 log-posterior); dict(sigma=, prior_sigma=, anchor=[K,P] tensor,
        nfull=): the loss is NegLogPost itself (losses.py:186-206; nnfit's 'logpost' option, nnfit.py:64-65)."""
        self.desc, self.K, self.dtype, self.device = desc, K, dtype, device
        self.cnet = desc.to_c()
        self.qdt = qb_dtype(dtype)
        self.logpost = logpost
        if logpost is None:
            self.lik = _lib.qb_lik_t(1.0, 0.0, 1.0, None, 0, 0)
        else:
            self.anchor = logpost['anchor']
            self.lik = _lib.qb_lik_t(float(logpost['sigma']), float(logpost['prior_sigma']), 1.0, self.anchor.data_ptr(), 1, 0)
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
        if self.logpost is not None:
            self.lik.prior_scale = float(n) / float(self.logpost['nfull'])          # losses.py:204
        with torch.cuda.device(self.device):
            rc = self.lib.qb_logpost_members(C.byref(self.cnet), self.qdt, _ptr(theta), self.K, C.byref(data), xs, ys,
                                             C.byref(self.lik), _ptr(self.lp), _ptr(self.grad) if want_grad else None,
                                             _ptr(self.ws), self.ws.numel(), _stream())
        _lib.check(rc, 'qb_logpost_members')
        if self.logpost is not None:
            return -self.lp
        # MSE of every member from the sigma=1 log-posterior
        return -(2.0 * self.lp + n * LOG_2PI) / (n * self.desc.out_dim)


class MemberTrainer:
    """State of a batched ensemble fit; `epoch(t, perm_t)` advances every member by one epoch (all its minibatches)."""

    def __init__(self, desc, theta0, xtrn, ytrn, subsets, val=None, lrate=0.1, wd=0.0, batch_size=None,
                 dtype=torch.float64, device='cuda', logpost=None):
        if not torch.cuda.is_available():
            raise RuntimeError('batched ensemble training needs a CUDA device (there is no CPU fallback)')
        self.device = device = torch.device(device)
        self.desc, self.dtype, self.lrate, self.wd = desc, dtype, float(lrate), float(wd)
        self.lib = _lib.load()
        subsets = torch.as_tensor(np.asarray(subsets), dtype=torch.long, device=device)
        self.K, self.nsub = K, nsub = subsets.shape
        self.P = P = desc.n_params
        th0 = as_device(theta0, dtype, device)
        self.theta = (th0[None, :].repeat(K, 1) if th0.dim() == 1 else th0.clone()).contiguous()
        if self.theta.shape != (K, P):
            raise ValueError(f'theta0 must be [P] or [K,P] with K={K}, P={P}')
        x_all = as_device(xtrn, dtype, device)
        y_all = as_device(np.asarray(ytrn).reshape(len(ytrn), -1), dtype, device)
        self.xs = x_all[subsets].contiguous()             # [K, nsub, d]: every member's own training subset
        self.ys = y_all[subsets].contiguous()
        self.own_val = val is None
        if val is None:
            self.xv, self.yv = self.xs, self.ys
        else:
            self.xv = as_device(np.asarray(val[0]), dtype, device).contiguous()
            self.yv = as_device(np.asarray(val[1]).reshape(len(val[1]), -1), dtype, device).contiguous()
        if batch_size is None or batch_size > nsub:
            batch_size = nsub
        self.batch_size = int(batch_size)
        self.full_batch = self.batch_size == nsub
        if logpost is not None:
            # NN_RMS (nn_rms.py:50-54): NegLogPost with a per-member anchor; fulldatasize = the member's subset size
            logpost = dict(logpost, anchor=as_device(logpost['anchor'], dtype, device).contiguous(), nfull=nsub)
            if logpost['anchor'].shape != (K, P):
                raise ValueError('logpost anchor must be [K,P]')
        self.ev = _MemberEval(desc, K, dtype, device, logpost)
        self.grad_scale = None if logpost is None else -1.0
        self.qdt = qb_dtype(dtype)
        self.m = torch.zeros_like(self.theta)
        self.v = torch.zeros_like(self.theta)
        self.best_theta = self.theta.clone()
        self.best_loss = torch.full((K,), 1.0e100, dtype=torch.float64, device=device)
        self.best_epoch = torch.zeros(K, dtype=torch.long, device=device)
        self.karange = torch.arange(K, device=device)[:, None]
        self.step = 0
        self.last_crit = None

    def epoch(self, t, perm_t=None, history=None):
        ev, K, P, lib = self.ev, self.K, self.P, self.lib
        for i in range(0, self.nsub, self.batch_size):
            if self.full_batch:
                xb, yb = self.xs, self.ys
            else:
                idx = perm_t[:, i:i + self.batch_size]                # [K, b] positions inside each member's subset
                xb = self.xs[self.karange, idx].contiguous()
                yb = self.ys[self.karange, idx].contiguous()
            loss_trn = ev(self.theta, xb, yb, True)                   # fills ev.grad = d lp / d theta
            nb = xb.shape[1]
            if self.own_val and self.full_batch:
                crit = loss_trn                                       # validation data == the minibatch
            else:
                crit = ev(self.theta, self.xv, self.yv, False)
            better = crit < self.best_loss                            # before the update (nnfit.py:143-152)
            mask = better.to(torch.uint8)
            with torch.cuda.device(self.device):
                _lib.check(lib.qb_copy_rows_where(self.qdt, _ptr(self.best_theta), _ptr(self.theta), _ptr(mask), K, P,
                                                  _stream()), 'qb_copy_rows_where')
            self.best_loss = torch.where(better, crit, self.best_loss)
            self.best_epoch = torch.where(better, torch.full_like(self.best_epoch, t), self.best_epoch)
            if history is not None:
                history.append(crit.clone())
            self.last_crit = crit
            self.step += 1
            with torch.cuda.device(self.device):
                _lib.check(lib.qb_adam_step(self.qdt, _ptr(self.theta), _ptr(ev.grad), _ptr(self.m), _ptr(self.v), K * P,
                                            self.lrate, 0.9, 0.999, 1e-8, self.wd, self.step,
                                            self.grad_scale if self.grad_scale is not None else -2.0 / (nb * self.desc.out_dim),
                                            _stream()), 'qb_adam_step')


def fit_members(desc, theta0, xtrn, ytrn, subsets, val=None, nepochs=5000, lrate=0.1, wd=0.0, batch_size=None,
                perms=None, dtype=torch.float64, device='cuda', freq_out=100, verbose=True, logpost=None):
    """Train K members together.  theta0: [K,P] (or [P], replicated); subsets: [K, nsub] integer rows of xtrn that
    member k trains on (nn_ens.py:62-63); val: None (a member's own subset, nnfit.py:108-109) or (xval, yval) shared by
    all members; logpost: None (MSE) or dict(sigma=, prior_sigma=, anchor=[K,P]) for nnfit's 'logpost' loss (NN_RMS);
    perms: optional [K, nepochs, nsub] minibatch orders (otherwise drawn with torch.randperm in the
    reference's member-major order when that array is small, else epoch by epoch).
    Returns dict(best_theta [K,P], best_loss [K], best_epoch [K], theta [K,P], history [niter, K] of validation MSE)."""
    tr = MemberTrainer(desc, theta0, xtrn, ytrn, subsets, val=val, lrate=lrate, wd=wd, batch_size=batch_size, dtype=dtype,
                       device=device, logpost=logpost)
    K, nsub, device = tr.K, tr.nsub, tr.device
    if perms is not None:
        perms = torch.as_tensor(np.asarray(perms), dtype=torch.long, device=device)
    elif not tr.full_batch and K * nepochs * nsub <= 5 * 10 ** 7:
        # the reference's order of torch.randperm calls: all epochs of member 0, then member 1, ... (nn_ens.py:56-66)
        perms = torch.stack([torch.stack([torch.randperm(nsub) for _ in range(nepochs)]) for _ in range(K)]).to(device)
    history = []
    for t in range(nepochs):
        if tr.full_batch:
            perm_t = None
        elif perms is not None:
            perm_t = perms[:, t]
        else:
            perm_t = torch.stack([torch.randperm(nsub) for _ in range(K)]).to(device)
        tr.epoch(t, perm_t, history)
        if verbose and freq_out and ((t + 1) % freq_out == 0 or t == 0 or t == nepochs - 1):
            print(f'{t + 1:>10}{tr.step:>10}   validation MSE: mean {history[-1].mean().item():.6f} '
                  f'best {tr.best_loss.mean().item():.6f}', flush=True)
    return dict(best_theta=tr.best_theta, best_loss=tr.best_loss, best_epoch=tr.best_epoch, theta=tr.theta,
                history=torch.stack(history) if history else torch.empty((0, K), device=device))
