"""NN_Ens: deep ensemble with the reference's interface (quinn/solvers/nn_ens.py:9-127).  The default MSE / Adam
training runs for all members at once on the device (quinn_b200/ens/batched.py, section 8f rank 1), other nnfit
options fall back to the member-by-member torch loop; every predictive call is one kernel-4 launch over the stacked
member weights."""
import numpy as np
import torch

from ..ens.learner import Learner
from ..netdesc import netdesc_from_module, flatten_module, unflatten_module
from .quinn import QUiNNBase


class NN_Ens(QUiNNBase):
    def __init__(self, nnmodel, nens=1, dfrac=1.0, verbose=False, dtype=torch.float64):
        super().__init__(nnmodel)
        self.verbose, self.nens, self.dfrac, self.dtype = verbose, nens, dfrac, dtype
        self.learners = [Learner(nnmodel) for _ in range(nens)]
        if self.verbose:
            self.print_params(names_only=True)

    def print_params(self, names_only=False):
        for i, learner in enumerate(self.learners):
            print(f"==========  Learner {i + 1}/{self.nens}  ============")
            learner.print_params(names_only=names_only)

    # nnfit options the batched device trainer reproduces (quinn/nns/nnfit.py:15-21); anything else -> member by member
    _BATCHED_OK = {'val', 'loss_fn', 'wd', 'optimizer', 'lrate', 'nepochs', 'batch_size', 'freq_out', 'freq_plot',
                   'lhist_suffix', 'gradcheck', 'scheduler_lr', 'lmbd', 'loss_xy'}

    def _can_batch(self, kwargs):
        import os
        if os.environ.get('QB_ENS_SEQUENTIAL') or not torch.cuda.is_available():
            return False
        if set(kwargs) - self._BATCHED_OK:
            return False
        if kwargs.get('loss_fn', 'mse') != 'mse' or kwargs.get('optimizer', 'adam') != 'adam':
            return False
        if kwargs.get('loss_xy') is not None or kwargs.get('lmbd') is not None or kwargs.get('scheduler_lr') is not None:
            return False
        if kwargs.get('gradcheck', False):
            return False
        from ..nns.nnbase import MLPBase
        for l in self.learners:       # a model with its own training procedure keeps it (learner.py:66-70); MLPBase.fit IS nnfit
            f = getattr(type(l.nnmodel), 'fit', None)
            if f is not None and f is not MLPBase.fit:
                return False
        try:
            netdesc_from_module(self.learners[0].nnmodel)
        except NotImplementedError:
            return False
        return True

    def fit(self, xtrn, ytrn, **kwargs):
        """Same call as the reference (nn_ens.py:51-69).  When the options are the default MSE / Adam training, all
        members are trained TOGETHER on the device (quinn_b200/ens/batched.py): one kernel-2 launch per iteration for
        the whole ensemble; the member subsets are drawn with the reference's np.random.permutation calls."""
        ntrn = ytrn.shape[0]
        self.batched_fit = False
        if not self._can_batch(kwargs):
            for jens, learner in enumerate(self.learners):
                print(f"======== Fitting Learner {jens + 1}/{self.nens} =======")
                ind = np.random.permutation(ntrn)[:int(ntrn * self.dfrac)]
                kwargs['lhist_suffix'] = f'_e{jens}'
                learner.fit(xtrn[ind], ytrn[ind], **kwargs)
            return
        import copy
        from ..ens.batched import fit_members
        print(f"======== Fitting {self.nens} learners together =======")
        subsets = np.stack([np.random.permutation(ntrn)[:int(ntrn * self.dfrac)] for _ in range(self.nens)])
        desc = netdesc_from_module(self.learners[0].nnmodel)
        theta0 = np.stack([flatten_module(l.nnmodel) for l in self.learners])
        res = fit_members(desc, theta0, np.asarray(xtrn), np.asarray(ytrn), subsets, val=kwargs.get('val'),
                          nepochs=kwargs.get('nepochs', 5000), lrate=kwargs.get('lrate', 0.1), wd=kwargs.get('wd', 0.0),
                          batch_size=kwargs.get('batch_size'), dtype=self.dtype, freq_out=kwargs.get('freq_out', 100))
        best = res['best_theta'].double().cpu().numpy()
        self.fit_info = {k: v.cpu().numpy() for k, v in res.items() if k in ('best_loss', 'best_epoch')}
        self.batched_fit = True
        for k, learner in enumerate(self.learners):
            learner.best_model = copy.deepcopy(learner.nnmodel)
            unflatten_module(learner.best_model, best[k])
            learner.trained = True

    def member_thetas(self, order=None):
        """Flat weights of the trained members, (nens, P) float64, optionally re-ordered."""
        assert all(l.trained for l in self.learners)
        th = np.stack([flatten_module(l.best_model) for l in self.learners])
        return th if order is None else th[np.asarray(order)]

    def _desc(self):
        return netdesc_from_module(self.learners[0].best_model)

    def predict_sample(self, x):
        return self.learners[np.random.randint(0, self.nens)].predict(x)

    def _ens_thetas(self, nens):
        if nens is None:
            nens = self.nens
        if nens > self.nens:
            print(f"Warning: Requested {nens} but only {self.nens} ensemble members available.")
            nens = self.nens
        return self._desc(), self.member_thetas(np.random.permutation(nens)), self.dtype     # nn_ens.py:103

    def predict_ens_fromsamples(self, x, nens=1):
        return np.array([self.predict_sample(x) for _ in range(nens)])
