"""Learner: one ensemble member (quinn/ens/learner.py:10-93); ``predict`` runs kernel 4."""
import copy

import numpy as np

from ..nns.nnfit import nnfit
from ..nns.nnwrap import device_forward
from ..nns.tchutils import print_nnparams


class Learner:
    def __init__(self, nnmodel, verbose=False):
        self.nnmodel = copy.deepcopy(nnmodel)     # every learner starts from the SAME weights (nn_ens.py:33-34)
        self.trained = False
        self.verbose = verbose
        self.best_model = None
        if self.verbose:
            self.print_params(names_only=True)

    def print_params(self, names_only=False):
        print_nnparams(self.best_model if self.trained else self.nnmodel, names_only=names_only)

    def fit(self, xtrn, ytrn, **kwargs):
        if callable(getattr(self.nnmodel, 'fit', None)):
            self.best_model = self.nnmodel.fit(xtrn, ytrn, **kwargs)
        else:
            self.best_model = nnfit(self.nnmodel, xtrn, ytrn, **kwargs)['best_nnmodel']
        self.trained = True

    def predict(self, x):
        assert self.trained
        return device_forward(self.best_model, np.asarray(x))
