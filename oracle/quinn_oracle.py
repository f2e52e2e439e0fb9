"""CPU oracle for the QUiNN posterior-sampling hot path  --  TEST INFRASTRUCTURE ONLY.

This file is a numpy (float64) restatement of the reference's algorithm for the
path SURVEY.md section 8 names.  It exists so the CUDA path can be checked; it is
never imported by the product package ``quinn_b200``.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` may import it.

Parity status: PINNED.  The reference ships no golden vectors for this path
(SURVEY.md section 4), so the oracle is pinned against outputs of the reference
itself: ``tests/golden/make_golden.py`` imports ``/root/reference`` (torch CPU,
float64), records log-posteriors, gradients, 1000-step AMCMC/HMC/MALA traces with
their consumed random draws, VI losses and predictive arrays into
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks every function
below against those files.

Where the arithmetic lives in the reference (third party, not under
/root/reference): torch (Linear/addmm, tanh, autograd, distributions.Normal) and
numpy legacy RandomState, both unpinned in the reference's pyproject.toml:29-34;
here torch 2.11.0 / numpy 2.3.5.

Network description used everywhere below (mirrors the flat layout contract of
quinn/nns/nnwrap.py:64-106): ``layers`` is a list of dicts
``{n_in, n_out, w_off, b_off (or -1), act ('tanh'|'relu'|'identity'),
res_step (0.0 = plain layer, else h = h + res_step*act(W h + b))}`` and
``theta`` is the flat parameter vector in ``nnmodel.parameters()`` order with
every tensor flattened C-order; W is (n_out, n_in) row-major at ``w_off``.
"""

import math

import numpy as np

LOG_2PI = math.log(2.0 * math.pi)


# ----------------------------------------------------------------------------
# network description helpers
# ----------------------------------------------------------------------------

def mlp_layers(indim, outdim, hls, biasorno=True, activ='tanh'):
    """Layer list + parameter count of quinn.nns.mlp.MLP (mlp.py:59-86).

    Linear(indim,h0) act Linear(h0,h1) act ... act Linear(h_last,outdim); the
    parameters appear as weight, bias per Linear in order.
    """
    if activ not in ('tanh', 'relu'):
        activ = 'identity'   # mlp.py:56-57: anything else is Identity
    widths = [indim] + list(hls) + [outdim]
    layers, off = [], 0
    for l in range(len(widths) - 1):
        n_in, n_out = widths[l], widths[l + 1]
        w_off = off
        off += n_in * n_out
        b_off = -1
        if biasorno:
            b_off = off
            off += n_out
        act = activ if l < len(widths) - 2 else 'identity'
        layers.append(dict(n_in=n_in, n_out=n_out, w_off=w_off, b_off=b_off,
                           act=act, res_step=0.0))
    return layers, off


def rnet_layers(rdim, nlayers, indim, outdim, biasorno=True, nonlin=True, mlp=False,
                layer_pre=True, layer_post=True, shared=True, poly_order=None):
    """Layer list of quinn.nns.rnet.RNet (rnet.py:90-111 order, :124-164 forward).

    ``shared=True`` is wp_function=Poly(0) (one ww_0/bb_0 used by all nlayers+1
    residual steps, rnet.py:344-347); ``shared=False`` is NonPar(nlayers+1).
    ``poly_order=n`` is wp_function=Poly(n) (Lin = 1, Quad = 2, Cubic = 3; rnet.py:244-347): residual step i uses
    W = sum_m ww_m * t_i^m with t_i = i/(nlayers+1), same for the bias; the layer dict then carries
    ``terms = [(t_i^m, ww_m offset, bb_m offset), ...]``.
    Parameter order: weight_pre, bias_pre, weight_post, bias_post, ww_*, bb_*.
    """
    act = 'tanh' if nonlin else 'identity'
    off = 0
    pre = post = None
    if layer_pre:
        pre = (off, off + rdim * indim)
        off += rdim * indim + rdim
    if layer_post:
        post = (off, off + outdim * rdim)
        off += outdim * rdim + outdim
    npar = 1 if shared else nlayers + 1
    if poly_order is not None:
        npar = poly_order + 1
    ww = []
    for _ in range(npar):
        ww.append(off)
        off += rdim * rdim
    bb = []
    if biasorno:
        for _ in range(npar):
            bb.append(off)
            off += rdim
    step = 1.0 / (nlayers + 1.0)
    layers = []
    if layer_pre:
        layers.append(dict(n_in=indim, n_out=rdim, w_off=pre[0], b_off=pre[1], act=act, res_step=0.0))
    for i in range(nlayers + 1):
        if poly_order is not None:
            t = step * i
            layers.append(dict(n_in=rdim, n_out=rdim, w_off=ww[0], b_off=(bb[0] if biasorno else -1), act=act,
                               res_step=(0.0 if mlp else step),
                               terms=[(t ** m, ww[m], bb[m] if biasorno else -1) for m in range(npar)]))
            continue
        ip = 0 if shared else int((step * i) * npar)      # rnet.py:377 NonPar index rule
        layers.append(dict(n_in=rdim, n_out=rdim, w_off=ww[ip], b_off=(bb[ip] if biasorno else -1),
                           act=act, res_step=(0.0 if mlp else step)))
    if layer_post:
        layers.append(dict(n_in=rdim, n_out=outdim, w_off=post[0], b_off=post[1], act='identity', res_step=0.0))
    return layers, off


def _layer_wb(L, theta):
    """Weight matrix and bias of a layer; polynomial-in-depth layers (rnet.py:344-347) sum their terms in the reference's
    order: val = 0.0; val += pars[m] * t**m."""
    nw = L['n_in'] * L['n_out']
    if 'terms' not in L:
        W = theta[L['w_off']:L['w_off'] + nw].reshape(L['n_out'], L['n_in'])
        b = theta[L['b_off']:L['b_off'] + L['n_out']] if L['b_off'] >= 0 else None
        return W, b
    W, b = 0.0, (0.0 if L['b_off'] >= 0 else None)
    for c, wo, bo in L['terms']:
        W = W + theta[wo:wo + nw].reshape(L['n_out'], L['n_in']) * c
        if b is not None:
            b = b + theta[bo:bo + L['n_out']] * c
    return W, b


def _act(name, z):
    if name == 'tanh':
        return np.tanh(z)
    if name == 'relu':
        return np.maximum(z, 0.0)
    return z


def _dact(name, z, a):
    """derivative of the activation given pre-activation z and value a"""
    if name == 'tanh':
        return 1.0 - a * a
    if name == 'relu':
        return (z > 0).astype(z.dtype)
    return np.ones_like(z)


# ----------------------------------------------------------------------------
# forward / log-posterior / gradient      (nnwrap.py:109-150, losses.py:186-256)
# ----------------------------------------------------------------------------

def forward(layers, theta, x, final=None, keep=False):
    """MLP.forward (mlp.py:92-101) / RNet.forward (rnet.py:124-164) at flat theta.

    x: (N, d) -> (N, o).  With keep=True also returns the per-layer cache for backprop.
    """
    theta = np.asarray(theta, dtype=np.float64)
    h = np.asarray(x, dtype=np.float64)
    cache = []
    for L in layers:
        W, b = _layer_wb(L, theta)
        z = h @ W.T
        if b is not None:
            z = z + b
        a = _act(L['act'], z)
        hin = h
        h = hin + L['res_step'] * a if L['res_step'] != 0.0 else a
        if keep:
            cache.append((hin, z, a))
    pre_final = h
    if final == 'exp':
        h = np.exp(h)
    if keep:
        return h, (cache, pre_final)
    return h


def neg_log_prior(theta, sigma_prior, anchor):
    """NegLogPrior.forward (losses.py:238-256)."""
    theta = np.asarray(theta, dtype=np.float64)
    return np.sum((theta - anchor) ** 2) / 2.0 / sigma_prior ** 2 + (theta.size / 2.0) * math.log(2 * math.pi * sigma_prior ** 2)


def neg_log_post(layers, theta, x, y, sigma, fulldatasize=None, prior=None, final=None):
    """NegLogPost.forward (losses.py:186-206).  Constants use len(predictions)=N
    even when o>1 (losses.py:199-200).  prior = dict(sigma=, anchor=) or None."""
    pred = forward(layers, theta, x, final=final)
    n = pred.shape[0]
    val = 0.5 * np.sum((np.asarray(y, dtype=np.float64) - pred) ** 2) / sigma ** 2
    val += (n / 2.0) * LOG_2PI
    val += n * math.log(sigma)
    if prior is not None:
        val += n * neg_log_prior(theta, prior['sigma'], np.asarray(prior['anchor'], dtype=np.float64)) / fulldatasize
    return val


def logpost(layers, theta, x, y, sigma, fulldatasize=None, prior=None, final=None):
    """NN_MCMC.logpost (nn_mcmc.py:45-71): minus NegLogPost, flat prior there."""
    return -neg_log_post(layers, theta, x, y, sigma, fulldatasize, prior, final)


def logpost_grad(layers, theta, x, y, sigma, fulldatasize=None, prior=None, final=None):
    """NN_MCMC.logpostgrad (nn_mcmc.py:73-98): gradient of logpost wrt flat theta
    (manual reverse mode replacing nnwrap.py:140-150's autograd).  Returns (lp, grad)."""
    theta = np.asarray(theta, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    out, (cache, pre_final) = forward(layers, theta, x, final=final, keep=True)
    n = out.shape[0]
    r = y - out
    lp = -(0.5 * np.sum(r ** 2) / sigma ** 2 + (n / 2.0) * LOG_2PI + n * math.log(sigma))
    g = np.zeros_like(theta)
    da = r / sigma ** 2                       # d lp / d out
    if final == 'exp':
        da = da * out
    for L, (hin, z, a) in zip(reversed(layers), reversed(cache)):
        W, _ = _layer_wb(L, theta)
        scale = L['res_step'] if L['res_step'] != 0.0 else 1.0
        dz = scale * da * _dact(L['act'], z, a)
        dW, db = (dz.T @ hin).ravel(), dz.sum(axis=0)
        for c, wo, bo in L.get('terms', [(1.0, L['w_off'], L['b_off'])]):       # chain rule through W = sum_m c_m ww_m
            g[wo:wo + L['n_in'] * L['n_out']] += c * dW
            if bo >= 0:
                g[bo:bo + L['n_out']] += c * db
        dprev = dz @ W
        da = dprev + da if L['res_step'] != 0.0 else dprev
    if prior is not None:
        anchor = np.asarray(prior['anchor'], dtype=np.float64)
        lp -= n * neg_log_prior(theta, prior['sigma'], anchor) / fulldatasize
        g -= (n / fulldatasize) * (theta - anchor) / prior['sigma'] ** 2
    return lp, g


def diag_fisher(layers, theta, x, y, sigma, fulldatasize=None, prior=None, final=None):
    """Diagonal of NNWrap.calc_hess_diag (nnwrap.py:204-229): the mean over data points of the squared gradient of the
    loss evaluated on ONE point at a time.  For a single point the reference hands NegLogPost a 1-D prediction of
    length o, so ``len(predictions)`` is o there (losses.py:199-204): the prior enters each point's loss with weight
    o / fulldatasize."""
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    acc = np.zeros(np.asarray(theta).size)
    o = y.shape[1]
    for i in range(x.shape[0]):
        _, g = logpost_grad(layers, theta, x[i:i + 1], y[i:i + 1], sigma, final=final)
        g = -g                                   # gradient of the loss (minus log-posterior)
        if prior is not None:
            g = g + (o / fulldatasize) * (np.asarray(theta, dtype=np.float64) - np.asarray(prior['anchor'], dtype=np.float64)) / prior['sigma'] ** 2
        acc += g * g
    return acc / x.shape[0]


# ----------------------------------------------------------------------------
# samplers with externally supplied random draws   (mcmc/*.py)
# ----------------------------------------------------------------------------

def amcmc_moments_update(Xm, cov, current, imcmc):
    """Running mean / covariance recursion of AMCMC.sampler (admcmc.py:52-59)."""
    if imcmc == 0:
        return current.copy(), np.zeros((current.size, current.size))
    Xm = (imcmc * Xm + current) / (imcmc + 1.0)
    rt = (imcmc - 1.0) / imcmc
    st = (imcmc + 1.0) / imcmc ** 2
    d = current - Xm
    cov = rt * cov + st * np.outer(d, d)
    return Xm, cov


def amcmc_initial_propcov(theta0):
    """admcmc.py:65: 0.01 on EVERY entry plus diag(0.09|theta0|)."""
    return 0.01 + np.diag(0.09 * np.abs(theta0))


def run_chain(logpost_fn, param_ini, nmcmc, sampler, draws, grad_fn=None,
              epsilon=0.05, L=3, gamma=0.1, t0=100, tadapt=1000, cov_ini=None, adapt_diag=True):
    """MCMCBase.run (mcmc.py:39-101) driven by recorded draws.

    sampler 'amcmc': draws['xi'][t] is the proposal increment the reference drew with
      np.random.multivariate_normal at step t (admcmc.py:70).  Alternatively draws['z0'][t], draws['z'][t] are
      STANDARD normals and the increment is built here the way the many-chain kernels build it:
      0.1*z0 + sqrt(0.09|theta0|)*z while the proposal covariance is the initial 0.01 + diag(0.09|theta0|)
      (exactly that matrix: rank one plus diagonal), and sqrt(gamma*2.4^2/P*(diag(C)+1e-8))*z after an adaptation
      when adapt_diag (the repo's DIAGONAL deviation from admcmc.py:59-67; '_pscale' is returned for it);
    sampler 'hmc' / 'mala': draws['p'][t] is the momentum np.random.randn(cdim)
      (hmc.py:43, mala.py:42);
    draws['u'][t] is np.random.random_sample() of mcmc.py:75.

    Returns the reference's result dict plus 'accepted' flags and, for AMCMC,
    the final '_Xm', '_cov', '_propcov'.
    """
    current = np.array(param_ini, dtype=np.float64)
    current_U = -logpost_fn(current)
    cmode, pmode = current, -current_U
    samples, alphas, logposts, accepted = [current], [0.0], [-current_U], []
    na = 0
    Xm = cov = propcov = None
    for imcmc in range(nmcmc):
        if sampler == 'amcmc':
            Xm, cov = amcmc_moments_update(Xm, cov, current, imcmc)
            cdim = current.size
            if imcmc == 0:
                propcov = cov_ini if cov_ini is not None else amcmc_initial_propcov(current)
            elif imcmc > t0 and imcmc % tadapt == 0:
                propcov = (gamma * 2.4 ** 2 / cdim) * (cov + 10 ** (-8) * np.eye(cdim))
            if 'xi' in draws:
                proposal = current + draws['xi'][imcmc]
            else:
                if imcmc == 0:
                    pscale, kind = np.sqrt(0.09 * np.abs(current)), 0
                elif adapt_diag and imcmc > t0 and imcmc % tadapt == 0:
                    pscale, kind = np.sqrt((gamma * 2.4 ** 2 / cdim) * (np.diag(cov) + 1e-8)), 1
                common = 0.1 * draws['z0'][imcmc] if kind == 0 else 0.0
                proposal = current + (common + pscale * draws['z'][imcmc])
            K_cur = K_prop = 0.0
        elif sampler == 'hmc':                                    # hmc.py:43-68
            p = np.array(draws['p'][imcmc], dtype=np.float64)
            proposal = current.copy()
            K_cur = np.sum(np.square(p)) / 2
            p = p + epsilon * grad_fn(proposal) / 2
            for jj in range(L):
                proposal = proposal + epsilon * p
                if jj != L - 1:
                    p = p + epsilon * grad_fn(proposal)
            p = p + epsilon * grad_fn(proposal) / 2
            p = -p
            K_prop = np.sum(np.square(p)) / 2
        elif sampler == 'mala':                                   # mala.py:42-51
            p = np.array(draws['p'][imcmc], dtype=np.float64)
            g0 = grad_fn(current)
            proposal = current + 0.5 * epsilon ** 2 * g0 + epsilon * p
            g1 = grad_fn(proposal)
            K_cur = np.sum(np.square(p)) / 2
            p = p + epsilon * (g0 + g1) / 2
            K_prop = np.sum(np.square(p)) / 2
        else:
            raise ValueError(sampler)
        proposed_U = -logpost_fn(proposal)
        with np.errstate(over='ignore', invalid='ignore'):
            mh_prob = np.exp((current_U + K_cur) - (proposed_U + K_prop))    # unclipped, mcmc.py:72
        acc = bool(draws['u'][imcmc] < mh_prob)                               # strict <, NaN rejects
        if acc:
            na += 1
            current = proposal + 0.0
            current_U = proposed_U + 0.0
            if -current_U >= pmode:
                pmode = -current_U
                cmode = current + 0.0
        accepted.append(acc)
        samples.append(current)
        alphas.append(mh_prob)
        logposts.append(-current_U)
    res = dict(chain=np.array(samples), mapparams=cmode, maxpost=pmode,
               accrate=float(na) / max(nmcmc, 1), logpost=np.array(logposts),
               alphas=np.array(alphas), accepted=np.array(accepted, dtype=bool))
    if sampler == 'amcmc':
        res.update(_Xm=Xm, _cov=cov, _propcov=propcov)
        if 'xi' not in draws:
            res.update(_pscale=pscale, _kind=kind)
    return res


# ----------------------------------------------------------------------------
# variational inference (vi/bnet.py, rvar/rvs.py)
# ----------------------------------------------------------------------------

def _normal_logpdf(x, sigma):
    return -x * x / (2.0 * sigma * sigma) - math.log(sigma) - 0.5 * LOG_2PI


def gmm2_logprob(w, pi, sigma1, sigma2):
    """GMM2_1d.log_prob (rvs.py:159-173): log(pi*exp(lp1)+(1-pi)*exp(lp2)) summed, NOT log-sum-exp."""
    p1 = np.exp(_normal_logpdf(w, sigma1))
    p2 = np.exp(_normal_logpdf(w, sigma2))
    return np.sum(np.log(pi * p1 + (1 - pi) * p2))


def gaussian1d_logprob(w, mu, logsigma):
    """Gaussian_1d.log_prob with the logsigma parameterisation BNet uses (rvs.py:120-127, bnet.py:80)."""
    sigma = np.exp(logsigma)
    return np.sum(-math.log(math.sqrt(2 * math.pi)) - np.log(sigma) - ((w - mu) ** 2) / (2 * sigma ** 2))


def vi_loss(layers, mu, rho, eps, x, y, datanoise, num_batches, pi=0.5, sigma1=1.0, sigma2=1.0,
            final=None, want_grad=False):
    """BNet.viloss (bnet.py:181-232) for given standard-normal draws eps[nsam, P].

    w_s = mu + exp(rho)*eps_s (rvs.py:102-108 with logsigma=rho);
    loss = (mean_s log q - mean_s log p)/num_batches
           + B log sd + B/2 log 2pi + B/2 * mean_{s,i,j}(out-y)^2 / sd^2.
    With want_grad returns (loss, dloss/dmu, dloss/drho) by the chain rule
    (what autograd produces in nnfit.py:162).
    """
    mu = np.asarray(mu, dtype=np.float64)
    rho = np.asarray(rho, dtype=np.float64)
    eps = np.asarray(eps, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    nsam = eps.shape[0]
    B, o = y.shape
    sig = np.exp(rho)
    logq = np.zeros(nsam)
    logp = np.zeros(nsam)
    ssq = np.zeros(nsam)
    gmu = np.zeros_like(mu)
    grho = np.zeros_like(rho)
    cnll = 0.5 * B / (nsam * B * o) / datanoise ** 2       # factor on sum of squared residuals
    for s in range(nsam):
        w = mu + sig * eps[s]
        logq[s] = gaussian1d_logprob(w, mu, rho)
        logp[s] = gmm2_logprob(w, pi, sigma1, sigma2)
        out = forward(layers, w, x, final=final)
        ssq[s] = np.sum((out - y) ** 2)
        if want_grad:
            # d ssq / d w via the log-posterior gradient at sigma=1:  lp = -ssq/2 - const
            _, glp = logpost_grad(layers, w, x, y, 1.0, final=final)
            dssq = -2.0 * glp
            p1 = pi * np.exp(_normal_logpdf(w, sigma1))
            p2 = (1 - pi) * np.exp(_normal_logpdf(w, sigma2))
            dlogp = (p1 * (-w / sigma1 ** 2) + p2 * (-w / sigma2 ** 2)) / (p1 + p2)
            dw = cnll * dssq - dlogp / (nsam * num_batches)       # dloss/dw_s (logq's w-dependence handled below)
            # logq: total derivative wrt mu is 0; wrt rho is -1 (explicit -1+eps^2, via w: -eps^2)
            gmu += dw
            grho += dw * sig * eps[s] + (-1.0) / (nsam * num_batches)
    nll = B * math.log(datanoise) + 0.5 * B * LOG_2PI + cnll * np.sum(ssq)
    loss = (logq.mean() - logp.mean()) / num_batches + nll
    if want_grad:
        return loss, gmu, grho
    return loss


# ----------------------------------------------------------------------------
# posterior predictive (solvers/quinn.py:51-104, nn_mcmc.py:180-200)
# ----------------------------------------------------------------------------

def mcmc_thinning_rows(nsamples_total, nens, nburn):
    """Row indices NN_MCMC.predict_ens picks (nn_mcmc.py:194-196)."""
    nevery = int((nsamples_total - nburn) / nens)
    return [nburn + j * nevery for j in range(nens)]


def predict_ens(layers, thetas, x, final=None):
    """Stack of forwards: (M, N, o)  (quinn.py:61-68)."""
    return np.array([forward(layers, th, x, final=final) for th in thetas])


def predict_moments(yens, msc=0):
    """QUiNNBase.predict_mom_sample moments (quinn.py:85-99): mean, var(ddof=1) | None, cov | None."""
    ymean = np.mean(yens, axis=0)
    yvar = ycov = None
    if msc == 1:
        yvar = np.var(yens, axis=0, ddof=1)
    elif msc == 2:
        _, nx, nout = yens.shape
        ycov = np.empty((nx, nx, nout))
        yvar = np.empty((nx, nout))
        for io in range(nout):
            ycov[:, :, io] = np.cov(yens[:, :, io], rowvar=False, ddof=1)
            yvar[:, io] = np.diag(ycov[:, :, io])
    return ymean, yvar, ycov
