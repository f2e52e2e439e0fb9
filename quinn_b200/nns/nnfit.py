"""Host-side training loop used by the deterministic fits, the ensemble learners and VI
(the role of quinn/nns/nnfit.py:15-218).  It only *calls* the hot path:
optimisers stay torch's, plotting is out of scope.  Runs on the model's device."""
import copy

import numpy as np
import torch

from .losses import NegLogPost
from .tchutils import tch


def nnfit(nnmodel, xtrn, ytrn, val=None, loss_fn='mse', loss_xy=None, datanoise=None, wd=0.0, priorparams=None,
          lossparams=None, optimizer='adam', lrate=0.1, lmbd=None, scheduler_lr=None, nepochs=5000, batch_size=None,
          gradcheck=False, cooldown=100, factor=0.95, freq_out=100, freq_plot=1000, lhist_suffix=''):
    ntrn = xtrn.shape[0]
    if loss_xy is None:
        if loss_fn == 'mse':
            mse = torch.nn.MSELoss(reduction='mean')
            loss_xy = lambda x, y: mse(nnmodel(x), y)       # noqa: E731
        elif loss_fn == 'logpost':
            loss_xy = NegLogPost(nnmodel, ntrn, datanoise, priorparams)
        else:
            raise ValueError(f'Loss function {loss_fn} is unknown.')
    params = list(nnmodel.parameters())
    if optimizer == 'adam':
        opt = torch.optim.Adam(params, lr=lrate, weight_decay=wd)
    elif optimizer == 'sgd':
        opt = torch.optim.SGD(params, lr=lrate, weight_decay=wd)
    else:
        raise ValueError(f'Optimizer {optimizer} is unknown.')
    if scheduler_lr == 'ReduceLROnPlateau':
        if lmbd is not None:
            raise ValueError('Trying to use two schedulers.')
        sched = torch.optim.lr_scheduler.ReduceLROnPlateau(opt, mode='min', cooldown=cooldown, factor=factor)
    else:
        sched = torch.optim.lr_scheduler.LambdaLR(opt, lr_lambda=lmbd if lmbd is not None else (lambda e: 1.0))
    if batch_size is None or batch_size > ntrn:
        batch_size = ntrn
    device = getattr(nnmodel, 'device', 'cpu')
    x_, y_ = tch(xtrn, device=device), tch(ytrn, device=device)
    xv, yv = (xtrn, ytrn) if val is None else val
    xv_, yv_ = tch(np.asarray(xv), device=device), tch(np.asarray(yv), device=device)
    info = {'best_fepoch': 0, 'best_epoch': 0, 'best_loss': 1.e+100, 'best_nnmodel': nnmodel, 'history': []}
    fepoch = 0.0
    nsub = len(range(0, ntrn, batch_size))
    for t in range(nepochs):
        perm = torch.randperm(ntrn, device='cpu').to(x_.device)
        for i in range(0, ntrn, batch_size):
            idx = perm[i:i + batch_size]
            loss_trn = loss_xy(x_[idx, :], y_[idx, :])
            with torch.no_grad():
                loss_val = loss_xy(xv_, yv_)
                if i == 0:
                    loss_full = loss_xy(x_, y_)
            fepoch += 1.0 / nsub
            crit = loss_val.item()
            info['history'].append([fepoch, loss_trn.item(), loss_full.item(), crit])
            if crit < info['best_loss']:
                info.update(best_loss=crit, best_nnmodel=copy.deepcopy(nnmodel), best_fepoch=fepoch, best_epoch=t)
            opt.zero_grad()
            loss_trn.backward()
            opt.step()
        if scheduler_lr == 'ReduceLROnPlateau':
            sched.step(info['history'][-1][3])
        else:
            sched.step()
        if freq_out and ((t + 1) % freq_out == 0 or t == 0 or t == nepochs - 1):
            h = info['history'][-1]
            print(f"{t + 1:>10}{len(info['history']):>10}{h[1]:>14.6f}{h[2]:>13.6f}{h[3]:>13.6f}"
                  f"{info['best_loss']:>14.6f} ({info['best_epoch']}){opt.param_groups[0]['lr']:>10}", flush=True)
    return info
