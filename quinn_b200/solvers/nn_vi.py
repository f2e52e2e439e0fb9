"""NN_VI: Bayes-by-backprop wrapper with the reference's interface (quinn/solvers/nn_vi.py:13-132)."""
import numpy as np
import torch

from .. import ops
from ..nns.nnfit import nnfit
from ..nns.tchutils import print_nnparams
from ..vi.bnet import BNet
from .quinn import QUiNNBase


class NN_VI(QUiNNBase):
    def __init__(self, nnmodel, verbose=False, pi=0.5, sigma1=1.0, sigma2=1.0, mu_init_lower=-0.2, mu_init_upper=0.2,
                 rho_init_lower=-5.0, rho_init_upper=-4.0, dtype=torch.float64, seed=None):
        super().__init__(nnmodel)
        self.device = 'cuda'
        self.bmodel = BNet(nnmodel, pi=pi, sigma1=sigma1, sigma2=sigma2, mu_init_lower=mu_init_lower,
                           mu_init_upper=mu_init_upper, rho_init_lower=rho_init_lower, rho_init_upper=rho_init_upper,
                           device=self.device, seed=seed)
        self.bmodel.to(dtype)
        self.verbose, self.trained, self.best_model = verbose, False, None
        if self.verbose:
            print("=========== Deterministic model parameters ================")
            self.print_params(names_only=True)
            print("=========== Variational model parameters ==================")
            print_nnparams(self.bmodel, names_only=True)
            print("===========================================================")

    def fit(self, xtrn, ytrn, val=None, nepochs=600, lrate=0.01, batch_size=None, freq_out=100, freq_plot=1000, wd=0,
            cooldown=100, factor=0.95, nsam=1, scheduler_lr=None, datanoise=0.05):
        ntrn = xtrn.shape[0]
        assert ntrn == ytrn.shape[0]
        if batch_size is None or batch_size > ntrn:
            batch_size = ntrn
        num_batches = ntrn if batch_size == 1 else (ntrn + 1) // batch_size        # nn_vi.py:97-100
        self.bmodel.loss_params = [datanoise, nsam, num_batches]
        fit_info = nnfit(self.bmodel, xtrn, ytrn, val=val, loss_xy=self.bmodel.viloss, lrate=lrate, batch_size=batch_size,
                         nepochs=nepochs, wd=wd, cooldown=cooldown, factor=factor, freq_plot=freq_plot,
                         scheduler_lr=scheduler_lr, freq_out=freq_out)
        self.best_model = fit_info['best_nnmodel']
        self.fit_info = fit_info
        self.trained = True

    def predict_sample(self, x):
        assert self.trained
        y = self.best_model(ops.as_device(np.asarray(x), self.best_model.flat_mu().dtype, self.device), sample=True)
        return y.double().cpu().numpy()

    def _ens_thetas(self, nens):
        assert self.trained
        w, _, _ = self.best_model.sample_weights(int(nens))
        return self.best_model.desc, w, w.dtype
