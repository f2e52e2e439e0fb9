"""Adaptive Metropolis (Haario 2001) with the reference's constructor (quinn/mcmc/admcmc.py:17-36)."""
import numpy as np
import torch

from .. import ops
from .mcmc import MCMCBase


class AMCMC(MCMCBase):
    def __init__(self, cov_ini=None, gamma=0.1, t0=100, tadapt=1000, adapt='auto'):
        """``adapt`` (extension): 'full' keeps the reference's dense PxP covariance per chain (in-kernel
        Cholesky at every adaptation), 'diag' only its diagonal (the scalable deviation, SURVEY.md
        section 7 hard part 1), 'none' never adapts; 'auto' = 'full' while K*P*P <= 2**27 entries, else 'diag'."""
        super().__init__()
        self.cov_ini, self.gamma, self.t0, self.tadapt, self.adapt = cov_ini, gamma, t0, tadapt, adapt
        self._Xm = self._cov = self._propcov = None

    # ---- fused path
    def _device_sampler_state(self, st):
        adapt = self.adapt
        if adapt == 'auto':
            adapt = 'full' if st.K * st.P * st.P <= 2 ** 27 else 'diag'
            if adapt == 'diag' and self.tadapt < 10 ** 9:
                print(f'AMCMC: K*P*P = {st.K * st.P * st.P:.3g} covariance entries do not fit; adapting the DIAGONAL only '
                      "(deviation from admcmc.py:59; pass adapt='full' to force the dense recursion)")
        chol = None
        if self.cov_ini is not None:
            chol = self._lower_factor(np.asarray(self.cov_ini, dtype=np.float64))
        self._adapt_used = adapt
        return ops.AmcmcState(st, gamma=self.gamma, t0=self.t0, tadapt=self.tadapt, adapt=adapt, chol_ini=chol)

    def _device_advance(self, st, samp, nsteps, rec, kw):
        ops.amcmc_run(st, samp, nsteps, rec, **kw)

    def _device_export(self, st, samp):
        """Expose the working attributes of admcmc.py:34-36 for the first chain."""
        if samp.xm is None:
            return
        self._Xm = samp.xm[0].double().cpu().numpy()
        c = samp.cov[0].double().cpu().numpy()
        self._cov = c if c.ndim == 2 else np.diag(c)
        if samp.chol is not None and int(samp.prop_kind[0].item()) == 2:
            Lf = samp.chol[0].double().cpu().numpy()
            self._propcov = Lf @ Lf.T

    @staticmethod
    def _lower_factor(cov):
        """Lower-triangular L with L L^T = cov for a positive SEMI-definite cov: Cholesky when it exists, else the
        eigen factor A = V sqrt(S) re-triangularised by a QR of A^T (A A^T = R^T R).  The reference accepts singular
        cov_ini because np.random.multivariate_normal factors through an SVD (admcmc.py:70)."""
        try:
            return np.linalg.cholesky(cov)
        except np.linalg.LinAlgError:
            s, v = np.linalg.eigh(0.5 * (cov + cov.T))
            a = v * np.sqrt(np.clip(s, 0.0, None))[None, :]
            r = np.linalg.qr(a.T, mode='r')
            return r.T

    @staticmethod
    def _factor(cov):
        """A with A A^T = cov for a positive SEMI-definite cov (the reference's numpy draw goes through an SVD,
        admcmc.py:70, so singular matrices such as the initial 0.01 + diag(0.09|0|) are legal)."""
        s, v = torch.linalg.eigh(cov)
        return v * s.clamp_min(0.0).sqrt()[..., None, :]

    # ---- generic-callable adapter: same recursion, batched over chains with torch ops
    def sampler(self, current, imcmc):
        K, P = current.shape
        if imcmc == 0:
            self._gXm = current.clone()
            self._gcov = torch.zeros((K, P, P), dtype=current.dtype, device=current.device)
            if self.cov_ini is not None:
                pc = torch.as_tensor(np.asarray(self.cov_ini), dtype=current.dtype, device=current.device).expand(K, P, P)
            else:
                pc = 0.01 + torch.diag_embed(0.09 * current.abs())
            self._gchol = self._factor(pc)
        else:
            self._gXm = (imcmc * self._gXm + current) / (imcmc + 1.0)
            rt, stt = (imcmc - 1.0) / imcmc, (imcmc + 1.0) / imcmc ** 2
            d = current - self._gXm
            self._gcov = rt * self._gcov + stt * d[:, :, None] * d[:, None, :]
            if imcmc > self.t0 and imcmc % self.tadapt == 0:
                eye = torch.eye(P, dtype=current.dtype, device=current.device)
                self._gchol = self._factor((self.gamma * 2.4 ** 2 / P) * (self._gcov + 1e-8 * eye))
        z = torch.randn((K, P), dtype=current.dtype, device=current.device, generator=self._gen)
        prop = current + torch.einsum('kab,kb->ka', self._gchol, z)
        zero = torch.zeros(K, dtype=current.dtype, device=current.device)
        return prop, zero, zero
