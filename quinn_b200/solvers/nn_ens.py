"""NN_Ens: deep ensemble with the reference's interface (quinn/solvers/nn_ens.py:9-127).  Training the
members stays a host-side torch loop (section 8f rank 1); every predictive call is one kernel-4 launch
over the stacked member weights."""
import numpy as np
import torch

from ..ens.learner import Learner
from ..netdesc import netdesc_from_module, flatten_module
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

    def fit(self, xtrn, ytrn, **kwargs):
        for jens, learner in enumerate(self.learners):
            print(f"======== Fitting Learner {jens + 1}/{self.nens} =======")
            ntrn = ytrn.shape[0]
            ind = np.random.permutation(ntrn)[:int(ntrn * self.dfrac)]
            kwargs['lhist_suffix'] = f'_e{jens}'
            learner.fit(xtrn[ind], ytrn[ind], **kwargs)

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
