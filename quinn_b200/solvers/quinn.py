"""QUiNNBase: posterior-predictive front end shared by the solvers (quinn/solvers/quinn.py:15-104).
Plotting helpers are out of scope."""
import copy

import numpy as np
import torch

from .. import ops
from ..netdesc import netdesc_from_module
from ..nns.tchutils import print_nnparams


class QUiNNBase:
    def __init__(self, nnmodel):
        self.nnmodel = copy.deepcopy(nnmodel)
        self.nens = None

    def print_params(self, names_only=False):
        print_nnparams(self.nnmodel, names_only=names_only)

    def predict_sample(self, x):
        raise NotImplementedError

    # ---- weight-matrix hook: solvers that can name their ensemble as theta[M,P] get the batched kernel
    def _ens_thetas(self, nens):
        """Return (desc, theta[M,P] tensor/array, dtype) for an nens-member predictive, or None."""
        return None

    def predict_ens(self, x, nens=None):
        """(M, N, o) array of sampled predictions (quinn.py:51-70); one kernel-4 launch when the solver
        exposes its ensemble as a weight matrix, else a loop over predict_sample."""
        if nens is None:
            nens = self.nens
        ens = self._ens_thetas(nens)
        if ens is None:
            return np.array([self.predict_sample(x) for _ in range(nens)])
        desc, thetas, dtype = ens
        out, _, _ = ops.predict(desc, thetas, np.asarray(x), dtype=dtype)
        return out.double().cpu().numpy()

    def predict(self, x):
        return self.predict_mom_sample(x)[0]

    def predict_stats(self, x, nsam=1000, qt=True):
        """get_stats (quinn/utils/stats.py:8-32) of nsam sampled predictions without the (M, N, o) array visiting the host:
        (median, q50-q25, q75-q50) when qt, else (mean, std, std); kernel 4 + the quantile / moment kernels of qb_post.cu."""
        from .. import post
        ens = self._ens_thetas(nsam)
        if ens is None:
            y = torch.as_tensor(self.predict_ens(x, nens=nsam)).cuda()
        else:
            desc, thetas, dtype = ens
            y, _, _ = ops.predict(desc, thetas, np.asarray(x), dtype=dtype)
        return tuple(t.double().cpu().numpy() for t in post.get_stats(y, qt))

    def predict_mom_sample(self, x, msc=0, nsam=1000):
        """mean / variance (ddof=1) / covariance of nsam sampled predictions (quinn.py:75-104).

        msc 0/1 never materialise the (M,N,o) array when the solver exposes a weight matrix: mean and
        variance are accumulated inside kernel 4.  msc=2 (N x N covariance per output) is host numpy on
        the sampled array and only sensible for small N."""
        if msc not in (0, 1, 2):
            raise ValueError(f"msc={msc}, but needs to be 0,1, or 2.")
        ens = self._ens_thetas(nsam) if msc < 2 else None
        if ens is not None and len(ens[1]) > 1:
            desc, thetas, dtype = ens
            _, mean, var = ops.predict(desc, thetas, np.asarray(x), dtype=dtype, want_out=False, want_moments=True)
            ymean = mean.double().cpu().numpy()
            yvar = var.double().cpu().numpy() if msc == 1 else None
            return ymean, yvar, None
        y = self.predict_ens(x, nens=nsam)
        _, nx, nout = y.shape
        ymean = np.mean(y, axis=0)
        yvar = ycov = None
        if msc == 2:
            ycov = np.empty((nx, nx, nout))
            yvar = np.empty((nx, nout))
            for io in range(nout):
                ycov[:, :, io] = np.cov(y[:, :, io], rowvar=False, ddof=1)
                yvar[:, io] = np.diag(ycov[:, :, io])
        elif msc == 1:
            yvar = np.var(y, axis=0, ddof=1)
        return ymean, yvar, ycov
