"""CPU baseline port of the reference's evaluation FLOW -- TEST / BENCHMARK INFRASTRUCTURE ONLY.

The numbers in oracle/quinn_oracle.py pin the arithmetic; this file restates how the reference spends
its time on the host so that bench.py's `cpu_baseline` / `--impl reference` legs measure the same work:
per evaluation a fresh flat->tensor unflatten (twice, nn_mcmc.py:56 + nnwrap.py:123), fresh
torch.tensor copies of x and y (nnwrap.py:121-122), a torch.nn.Sequential forward in float64 on all
host threads, the NegLogPost formula (losses.py:197-200) and, for gradients, autograd backward plus the
per-parameter numpy concat (nnwrap.py:144-150).  The AMCMC proposal is the reference's
np.random.multivariate_normal on the dense PxP covariance (admcmc.py:70: an SVD every step).
Never imported by quinn_b200.  parity: pinned through tests/test_oracle_golden.py::test_torch_port_matches_golden.
"""
import numpy as np
import torch


class RefPort:
    def __init__(self, indim, outdim, hls, activ='tanh'):
        torch.set_default_dtype(torch.double)
        act = {'tanh': torch.nn.Tanh, 'relu': torch.nn.ReLU}.get(activ, torch.nn.Identity)
        widths = [indim] + list(hls) + [outdim]
        mods = []
        for l in range(len(widths) - 1):
            if l > 0:
                mods.append(act())
            mods.append(torch.nn.Linear(widths[l], widths[l + 1]))
        self.net = torch.nn.Sequential(*mods).double()
        self.bounds, s = [], 0
        for p in self.net.parameters():
            self.bounds.append((s, s + p.numel()))
            s += p.numel()
        self.pdim = s

    def _unflatten(self, theta):
        for (s, e), p in zip(self.bounds, self.net.parameters()):
            p.data = torch.tensor(theta[s:e]).view(*p.shape)

    def _loss(self, theta, x, ylist, sigma):
        self._unflatten(theta)
        inputs = torch.tensor(x)
        targets = torch.tensor(np.array(ylist))
        self._unflatten(theta)
        pred = self.net(inputs)
        sig = torch.tensor(float(sigma))
        n = len(pred)
        return 0.5 * torch.sum((targets - pred) ** 2) / sig ** 2 + (n / 2) * torch.log(2 * torch.tensor(np.pi)) + n * torch.log(sig)

    def logpost(self, theta, x, ylist, sigma):
        with torch.no_grad():
            return -self._loss(theta, x, ylist, sigma).item()

    def logpostgrad(self, theta, x, ylist, sigma):
        loss = self._loss(theta, x, ylist, sigma)
        loss.backward()
        gs = []
        for p in self.net.parameters():
            gs.append(p.grad.numpy().flatten())
            p.grad = None
        return -np.concatenate(gs)

    def forward(self, theta, x):
        self._unflatten(theta)
        with torch.no_grad():
            return self.net(torch.tensor(x)).numpy()


def amcmc_proposal_draw(theta):
    """One proposal increment the way the reference draws it while its covariance is the initial one
    (admcmc.py:65,70): dense PxP matrix + numpy's SVD-based multivariate_normal."""
    cdim = len(theta)
    propcov = 0.01 + np.diag(0.09 * np.abs(theta))
    return np.random.multivariate_normal(np.zeros(cdim), propcov)
