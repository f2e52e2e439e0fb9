"""MCMCBase: chain driver with the reference's interface (quinn/mcmc/mcmc.py:9-114), running K chains
at once on the GPU.

``setLogPost`` accepts
  * a ``DeviceLogPost`` handle, or ``NN_MCMC.logpost`` / ``logpostgrad`` bound methods with ``lpinfo=``
    (what the reference's NN_MCMC.fit passes, nn_mcmc.py:130-135): the whole run is ONE fused kernel per
    segment -- propose, evaluate, accept on the device, no host round trip per step (kernel 3);
  * any other Python callable (the reference's own sampler tests pass numpy closures,
    tests/test_mcmc.py:10-22): propose/accept still run on the GPU (torch tensors), the callable is
    invoked once per sweep -- batched over chains if it has attribute ``batched = True`` and takes/returns
    CUDA tensors, else chain by chain on numpy rows.  This adapter exists for API conformance only.

``run(nmcmc, param_ini)``: ``param_ini`` of shape (P,) reproduces the reference's result dict exactly
('chain' (M+1,P), 'mapparams', 'maxpost', 'accrate', 'logpost', 'alphas', mcmc.py:92-99); shape (K,P)
adds a leading K axis to every entry.
"""
import numpy as np
import torch

from .. import ops


_SIDE_STREAMS = {}
_PINNED = {}          # result buffers of the host-pipelined path, reused from one run() to the next


def _pinned_results(K, nstored, P, nmcmc, dtype):
    """Pinned host buffers for the result dict of a (K, nmcmc) run.  Allocating GBs of page-locked memory costs more than
    the chain kernel of a short run, so the buffers of the latest shape are kept (results returned to the caller are
    numpy views of them: a following run() with the same shape overwrites them, pass copy_results=True to run() to keep
    independent arrays)."""
    key = (K, nstored, P, nmcmc, dtype)
    if key not in _PINNED:
        _PINNED.clear()
        pin = lambda *shape, dtype: torch.empty(shape, dtype=dtype, pin_memory=True)      # noqa: E731
        _PINNED[key] = dict(chain=pin(K, nstored, P, dtype=dtype), mapparams=pin(K, P, dtype=dtype),
                            maxpost=pin(K, dtype=torch.float64), accrate=pin(K, dtype=torch.float64),
                            logpost=pin(K, nmcmc + 1, dtype=torch.float64), alphas=pin(K, nmcmc + 1, dtype=torch.float64),
                            accepted=pin(K, nmcmc, dtype=torch.uint8))
    return _PINNED[key]


def _side_streams(device):
    """Two persistent side streams per device (persistent so that the caching allocator's per-stream pools are
    reused from one run() to the next)."""
    key = str(device)
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = [torch.cuda.Stream(device=device) for _ in range(2)]
    return _SIDE_STREAMS[key]


class DeviceLogPost:
    """A log-posterior that lives on the GPU: network + data + likelihood (an ops.Problem)."""

    def __init__(self, problem: ops.Problem):
        self.problem = problem

    def __call__(self, theta, **_):
        lp = ops.logpost(self.problem, theta)
        return float(lp[0].item()) if np.ndim(theta) == 1 else lp

    def grad(self, theta, **_):
        _, g = ops.logpost_grad(self.problem, theta)
        return g[0].double().cpu().numpy() if np.ndim(theta) == 1 else g


class ShardedDataLogPost:
    """Log-posterior of data that is SHARDED over the ranks (north_star: "allreduce of log-likelihood partial sums when N
    exceeds one GPU"): every rank evaluates kernels 1 / 2 on its slice for all chains; the Gaussian log-likelihood and
    its constants are sums over points (losses.py:197-200), so the all-reduced SUM of the per-rank values is the
    full-data log-posterior, bit-identical on every rank.  A batched callable for the generic sampler path (propose /
    accept run redundantly on every rank from the same seed)."""
    batched = True

    def __init__(self, problem_local, n_total):
        self.problem, self.n_total = problem_local, int(n_total)

    def __call__(self, theta, **_):
        from .. import dist
        th = theta if torch.is_tensor(theta) else torch.as_tensor(np.atleast_2d(np.asarray(theta, dtype=np.float64)))
        lp = ops.logpost(self.problem, th.to(self.problem.device)).clone()
        return dist.allreduce_sum_(lp)

    def grad(self, theta, **_):
        from .. import dist
        th = theta if torch.is_tensor(theta) else torch.as_tensor(np.atleast_2d(np.asarray(theta, dtype=np.float64)))
        _, g = ops.logpost_grad(self.problem, th.to(self.problem.device))
        return dist.allreduce_sum_(g.double().clone())


ShardedDataLogPost.grad.batched = True          # the generic driver looks for this flag on the callable it was given


class MCMCBase(object):
    def __init__(self):
        self.logPost = None
        self.logPostGrad = None
        self.postInfo = {}
        self._device_lp = None

    def setLogPost(self, logPost, logPostGrad, **postInfo):
        self.logPost = logPost
        self.logPostGrad = logPostGrad
        self.postInfo = postInfo
        self._device_lp = None
        if isinstance(logPost, DeviceLogPost):
            self._device_lp = logPost
        else:
            owner = getattr(logPost, '__self__', None)
            if owner is not None and hasattr(owner, 'device_logpost') and 'lpinfo' in postInfo:
                self._device_lp = owner.device_logpost(postInfo['lpinfo'])

    # ------------------------------------------------------------------ public driver
    def run(self, nmcmc, param_ini, *, seed=None, store_every=1, replay=None, chain_offset=0, verbose=True,
            keep_on_device=False, copy_results=False, segment_every=None, on_segment=None):
        """nmcmc steps for every row of param_ini.

        Extensions (keyword-only): ``seed`` Philox seed (default: drawn from np.random so np.random.seed
        controls it), ``store_every`` thinning of the returned chain, ``replay=dict(incr=[M,(K,)P],
        unif=[M,(K)])`` to consume recorded draws instead of Philox, ``chain_offset`` global index of the
        first chain (multi-GPU sharding), ``keep_on_device`` return CUDA tensors, ``copy_results`` detach the returned
        arrays of the host-pipelined path from its reusable pinned buffers, ``segment_every`` / ``on_segment(state, recorder,
        t_end)`` one kernel launch per that many steps with a hook after each (running diagnostics on a side stream).
        """
        assert self.logPost is not None
        if not torch.is_tensor(param_ini):          # tensors (e.g. pinned host memory) are passed through untouched
            param_ini = np.asarray(param_ini, dtype=np.float64)
        single = param_ini.ndim == 1
        theta0 = param_ini[None, :] if single else param_ini
        if seed is None:
            seed = int(np.random.randint(1, 2 ** 31 - 1))
        if self._device_lp is not None:
            nsh = 1 if on_segment is not None else self._auto_shards(theta0, replay, keep_on_device)
            if nsh != 1:
                return self._run_fused_sharded(int(nmcmc), theta0, int(seed), int(store_every), int(chain_offset), verbose,
                                               max(nsh, 1), copy_results)
            res = self._run_fused(int(nmcmc), theta0, int(seed), int(store_every), replay, int(chain_offset), verbose,
                                  segment_every, on_segment)
        else:
            res = self._run_generic(int(nmcmc), theta0, int(seed), int(store_every), verbose)
        return self._finish(res, single, keep_on_device)

    # ------------------------------------------------------------------ fused path (kernel 3)
    def _run_fused(self, nmcmc, theta0, seed, store_every, replay, chain_offset, verbose, segment_every=None, on_segment=None):
        prob = self._device_lp.problem
        st = ops.ChainState(prob, theta0)
        K, P = st.K, st.P
        samp = self._device_sampler_state(st)
        incr = unif = None
        if replay is not None:
            incr = ops.as_device(np.asarray(replay['incr']).reshape(nmcmc, K, P), prob.dtype, prob.device)
            unif = ops.as_device(np.asarray(replay['unif']).reshape(nmcmc, K), torch.float64, prob.device)
        nseg = 10 if (verbose and nmcmc >= 10) else 1
        bounds = [int(round(i * nmcmc / nseg)) for i in range(nseg + 1)]
        if segment_every:           # one launch per `segment_every` steps (running diagnostics hook, multi-GPU driver)
            bounds = list(range(0, nmcmc, int(segment_every))) + [nmcmc]
        if store_every > 1:          # segment boundaries must fall on stored steps
            bounds = sorted(set([0, nmcmc] + [b - b % store_every for b in bounds[1:-1]]))
        recs = []
        th0 = st.theta.clone()
        for a, b in zip(bounds[:-1], bounds[1:]):
            if b <= a:
                continue
            rec = ops.Recorder(st, b - a, store_every=store_every)
            kw = dict(seed=seed, chain_offset=chain_offset)
            if incr is not None:
                kw.update(incr=incr[a:b].contiguous(), unif=unif[a:b].contiguous())
            self._device_advance(st, samp, b - a, rec, kw)
            recs.append(rec)
            if on_segment is not None:
                on_segment(st, rec, b)          # must only ENQUEUE work (side stream); rec stays alive until the run returns
            if verbose:
                acc = st.naccept.double().mean().item() / b
                print('%d / %d completed, acceptance rate %lg' % (b, nmcmc, acc))
        self._device_export(st, samp)
        chain = torch.cat([th0[:, None, :]] + [r.samples for r in recs if r.samples is not None], dim=1)
        lp0 = recs[0].logpost0              # log-posterior of the initial state (mcmc.py:55-61), recorded by the kernel
        logpost = torch.cat([lp0[:, None]] + [r.logpost for r in recs], dim=1)
        alphas = torch.cat([torch.zeros((K, 1), dtype=torch.float64, device=prob.device)] + [r.alpha for r in recs], dim=1)
        accepted = torch.cat([r.accepted for r in recs], dim=1)
        return dict(chain=chain, mapparams=st.map_theta, maxpost=st.map_lp, accrate=st.naccept.double() / nmcmc,
                    logpost=logpost, alphas=alphas, accepted=accepted, state=st)

    # ------------------------------------------------------------------ host <-> device pipelining for huge batches
    def _auto_shards(self, theta0, replay, keep_on_device):
        """Chains given in HOST memory and big enough that PCIe time matters (>= 256 MB of states) are processed as 4
        shards on two CUDA streams: the upload of shard i+1 and the download of shard i-1 overlap the kernel of
        shard i.  Philox is keyed by the global chain index, so sharding does not change any result."""
        if replay is not None or keep_on_device:
            return 1
        on_host = (not torch.is_tensor(theta0)) or theta0.device.type == 'cpu'
        prob = self._device_lp.problem
        nbytes = theta0.shape[0] * theta0.shape[1] * prob.x.element_size()
        if not on_host or nbytes < 8 * 2 ** 20:
            return 1
        # >= 8 MB of states: results go straight into pinned host buffers (path below); >= 256 MB: 8 pipelined shards (only
        # the download of the last shard, 1/8 of the results, is not hidden behind a kernel)
        return 8 if (nbytes >= 256 * 2 ** 20 and theta0.shape[0] >= 16) else -1

    def _run_fused_sharded(self, nmcmc, theta0, seed, store_every, chain_offset, verbose, nsh, copy_results=False):
        prob = self._device_lp.problem
        th = theta0 if torch.is_tensor(theta0) else torch.from_numpy(np.ascontiguousarray(theta0))
        K, P = th.shape
        nstored = 1 + (nmcmc // store_every if store_every > 0 else 0)
        host = _pinned_results(K, nstored, P, nmcmc, prob.dtype)
        main = torch.cuda.current_stream(prob.device)
        streams = _side_streams(prob.device)
        from ..dist import shard_range
        for i in range(nsh):
            lo, hi = shard_range(K, i, nsh)
            if hi <= lo:
                continue
            s = streams[i % 2]
            s.wait_stream(main)
            with torch.cuda.stream(s):
                res = self._run_fused(nmcmc, th[lo:hi], seed, store_every, None, chain_offset + lo, False)
                for k, dst in host.items():
                    dst[lo:hi].copy_(res[k], non_blocking=True)
                if i == 0:
                    self.last_state = res.get('state')
                del res
        for s in streams:
            s.synchronize()
        if verbose:
            print('%d / %d completed, acceptance rate %lg' % (nmcmc, nmcmc, host['accrate'].mean().item()))
        out = {k: (v.numpy().copy() if copy_results else v.numpy()) for k, v in host.items()}   # chain / mapparams stay in the compute dtype
        out['accepted'] = out['accepted'].astype(bool)
        return out

    # hooks the samplers implement for the fused path
    def _device_sampler_state(self, st):
        raise NotImplementedError

    def _device_advance(self, st, samp, nsteps, rec, kw):
        raise NotImplementedError

    def _device_export(self, st, samp):
        pass

    # ------------------------------------------------------------------ generic-callable adapter
    def _eval_generic(self, fn, theta):
        """theta: CUDA [K,P] float64 -> CUDA [K] (logPost) or [K,P] (logPostGrad)."""
        if getattr(fn, 'batched', False):
            return fn(theta, **self.postInfo)
        rows = theta.cpu().numpy()
        vals = [np.asarray(fn(r, **self.postInfo), dtype=np.float64) for r in rows]
        return torch.as_tensor(np.stack(vals), dtype=torch.float64, device=theta.device)

    def _run_generic(self, nmcmc, theta0, seed, store_every, verbose):
        if not torch.cuda.is_available():
            raise RuntimeError('quinn_b200 samplers need a CUDA device (there is no CPU fallback)')
        dev = torch.device('cuda')
        gen = torch.Generator(device=dev)
        gen.manual_seed(seed)
        cur = torch.as_tensor(theta0).to(device=dev, dtype=torch.float64).clone()
        K, P = cur.shape
        cur_U = -self._eval_generic(self.logPost, cur)
        cmode, pmode = cur.clone(), -cur_U.clone()
        samples, alphas, logposts, accs = [cur.clone()], [torch.zeros(K, dtype=torch.float64, device=dev)], [-cur_U], []
        na = torch.zeros(K, dtype=torch.float64, device=dev)
        self._gen = gen
        for imcmc in range(nmcmc):
            prop, K_cur, K_prop = self.sampler(cur, imcmc)
            prop_U = -self._eval_generic(self.logPost, prop)
            mh = torch.exp((cur_U + K_cur) - (prop_U + K_prop))
            u = torch.rand(K, dtype=torch.float64, device=dev, generator=gen)
            acc = u < mh
            na += acc
            cur = torch.where(acc[:, None], prop, cur)
            cur_U = torch.where(acc, prop_U, cur_U)
            better = acc & (-cur_U >= pmode)
            pmode = torch.where(better, -cur_U, pmode)
            cmode = torch.where(better[:, None], cur, cmode)
            if (imcmc + 1) % store_every == 0:
                samples.append(cur.clone())
            alphas.append(mh)
            logposts.append(-cur_U)
            accs.append(acc)
            if verbose and nmcmc >= 10 and ((imcmc + 2) % (nmcmc // 10) == 0):
                print('%d / %d completed, acceptance rate %lg' % (imcmc + 2, nmcmc, na.mean().item() / (imcmc + 1)))
        return dict(chain=torch.stack(samples, 1), mapparams=cmode, maxpost=pmode, accrate=na / max(nmcmc, 1),
                    logpost=torch.stack(logposts, 1), alphas=torch.stack(alphas, 1),
                    accepted=torch.stack(accs, 1) if accs else torch.zeros((K, 0), dtype=torch.bool, device=dev))

    def sampler(self, current, imcmc):
        """Generic-path proposal for all chains: (proposal[K,P], current_K[K], proposed_K[K])."""
        raise NotImplementedError("sampler method not implemented in the base class and should be implemented in children.")

    # ------------------------------------------------------------------ result dict
    def _finish(self, res, single, keep_on_device):
        self.last_state = res.pop('state', None)
        if keep_on_device:
            return res
        out = {}
        for k, v in res.items():
            a = v.detach().cpu().numpy() if v.dtype != torch.uint8 and v.dtype != torch.bool else v.cpu().numpy().astype(bool)
            out[k] = a[0] if single else a
        if single:
            out['maxpost'] = float(out['maxpost'])
            out['accrate'] = float(out['accrate'])
        return out
