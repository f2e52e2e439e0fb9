"""Tensor-core gradient paths against the fp64 oracle (manual reverse mode pinned to the reference's autograd by
tests/test_oracle_golden.py) and against the CUDA-core kernel they replace; HMC / MALA chains on them.  Every test runs twice:
   'tcg'  quinn_b200/csrc/qb_tcg.cuh  (3xTF32; widths 32 / 64, tanh / relu)       -- plan code 3
   'tg8'  quinn_b200/csrc/qb_tg8.cuh  (3xFP16 with power-of-two scaling; tanh nets of width 32 (zero-padded to 64) / 64 / 128, the default
          there) -- plan code 4
QB_TG8_64=0 sends the 32- and 64-wide tanh nets back to qb_tcg.cuh."""
import os

import numpy as np
import pytest
import torch

from golden_util import netdesc_from_layers
from oracle import quinn_oracle as qo
from test_gpu_tensorcore import make_net

pytestmark = pytest.mark.gpu

TOL_LP = 1e-5        # relative (north_star: 1e-4 in fp32)
TOL_G = 2e-5         # of the largest gradient entry (north_star: 1e-4)


@pytest.fixture(autouse=True, params=['tg8', 'tcg'])
def grad_path(request):
    old = os.environ.get('QB_TG8_64')
    os.environ['QB_TG8_64'] = '1' if request.param == 'tg8' else '0'
    yield request.param
    if old is None:
        del os.environ['QB_TG8_64']
    else:
        os.environ['QB_TG8_64'] = old


def plan_code(widths, act):
    """Which tensor-core gradient kernel an eligible net takes under the current QB_TG8_64."""
    if widths[1] in (32, 64, 128) and act == 'tanh' and (widths[1] == 128 or os.environ.get('QB_TG8_64', '1') != '0'):
        return 4
    return 3


class no_tcg:
    """Run the enclosed calls on the CUDA-core gradient kernel (the library reads QB_NO_TCG at every launch)."""
    def __enter__(self):
        self.old = os.environ.get('QB_NO_TCG')
        os.environ['QB_NO_TCG'] = '1'

    def __exit__(self, *a):
        if self.old is None:
            del os.environ['QB_NO_TCG']
        else:
            os.environ['QB_NO_TCG'] = self.old


def _data(rs, N, d):
    x = rs.rand(N, d) * 2 - 1
    y = np.sin(x.sum(1, keepdims=True)) + 0.1 * rs.randn(N, 1)
    return x, y


def _check(layers, P, x, y, th, sigma, lp, g, prior=None, nfull=None):
    for k in range(th.shape[0]):
        pk = None
        if prior is not None:
            a = prior['anchor']
            pk = dict(sigma=prior['sigma'], anchor=a[k] if a.ndim == 2 else a)
        rl, rg = qo.logpost_grad(layers, th[k], x, y, sigma, fulldatasize=nfull, prior=pk)
        assert abs(lp[k] - rl) <= TOL_LP * abs(rl), (k, lp[k], rl)
        err = np.abs(g[k] - rg).max() / np.abs(rg).max()
        assert err <= TOL_G, (k, err)


GRAD_SHAPES = [
    # widths, act, N, K
    ([3, 64, 64, 1], 'tanh', 1000, 3),        # config 5 net, few chains -> the data axis is split over blocks
    ([3, 64, 64, 1], 'tanh', 128, 2),         # exactly one tile
    ([3, 64, 64, 1], 'tanh', 129, 2),         # one point in the second tile
    ([3, 64, 64, 1], 'tanh', 1, 2),           # a single data point
    ([3, 64, 64, 1], 'tanh', 300, 1300),      # many chains: one block per chain
    ([2, 32, 32, 1], 'tanh', 1000, 5),        # config 2 net
    ([2, 32, 32, 1], 'relu', 257, 3),
    ([7, 64, 64, 1], 'tanh', 200, 3),         # padded input width 8
    ([5, 32, 32, 1], 'relu', 77, 2),
    ([1, 64, 64, 1], 'relu', 640, 2),
]


@pytest.mark.parametrize('case', range(len(GRAD_SHAPES)))
def test_tc_gradient_matches_oracle(case):
    from quinn_b200 import ops
    widths, act, N, K = GRAD_SHAPES[case]
    rs = np.random.RandomState(300 + case)
    layers, P = make_net(widths, [act, act, 'identity'])
    x, y = _data(rs, N, widths[0])
    th = (0.5 * rs.randn(K, P)).astype(np.float32).astype(np.float64)
    prob = ops.Problem(netdesc_from_layers(layers, P), x, y, 0.2, dtype=torch.float32)
    assert prob.plan_info(K, True)['tensor_core'] == plan_code(widths, act)
    lp, g = ops.logpost_grad(prob, th)
    lp, g = lp.cpu().numpy(), g.double().cpu().numpy()
    idx = np.arange(K) if K <= 5 else np.array([0, K // 2, K - 1])
    _check(layers, P, x, y, th[idx], 0.2, lp[idx], g[idx])
    # and the CUDA-core kernel agrees to fp32 accuracy
    with no_tcg():
        assert prob.plan_info(K, True)['tensor_core'] == 0
        lp2, g2 = ops.logpost_grad(prob, th)
    np.testing.assert_allclose(lp2.cpu().numpy(), lp, rtol=2e-5)
    assert (np.abs(g2.double().cpu().numpy() - g).max(1) <= 1e-4 * np.abs(g).max(1)).all()


def test_tc_gradient_without_biases_and_with_prior():
    from quinn_b200 import ops
    rs = np.random.RandomState(41)
    for bias in (False, True):
        layers, P = make_net([3, 64, 64, 1], ['tanh', 'tanh', 'identity'], bias=bias)
        x, y = _data(rs, 500, 3)
        K = 4
        th = (0.4 * rs.randn(K, P)).astype(np.float32).astype(np.float64)
        for anchor in (0.1 * rs.randn(P), 0.1 * rs.randn(K, P)):
            anchor = anchor.astype(np.float32).astype(np.float64)
            prob = ops.Problem(netdesc_from_layers(layers, P), x, y, 0.3, dtype=torch.float32, prior_sigma=0.7,
                               prior_anchor=anchor, fulldatasize=1200)
            assert prob.plan_info(K, True)['tensor_core'] == plan_code([3, 64, 64, 1], 'tanh')
            lp, g = ops.logpost_grad(prob, th)
            _check(layers, P, x, y, th, 0.3, lp.cpu().numpy(), g.double().cpu().numpy(), prior=dict(sigma=0.7, anchor=anchor),
                   nfull=1200)


def test_tc_gradient_plan_eligibility():
    from quinn_b200 import ops
    rs = np.random.RandomState(1)
    t64 = plan_code([3, 64, 64, 1], 'tanh')
    cases = [([3, 64, 64, 1], ['tanh', 'tanh', 'identity'], t64), ([2, 32, 32, 1], ['relu', 'relu', 'identity'], 3),
             ([3, 64, 64, 1], ['relu', 'relu', 'identity'], 3),          # relu: unbounded activations stay on the tf32 kernel
             ([3, 48, 48, 1], ['tanh', 'tanh', 'identity'], 0),          # width not 32 / 64
             ([3, 64, 32, 1], ['tanh', 'tanh', 'identity'], 0),          # unequal widths
             ([3, 64, 64, 2], ['tanh', 'tanh', 'identity'], 0),          # two outputs
             ([8, 64, 64, 1], ['tanh', 'tanh', 'identity'], 4 if t64 == 4 else 0),   # more than 7 inputs: only the fp16-split kernel
             ([3, 64, 64, 1], ['tanh', 'relu', 'identity'], 0),          # mixed activations
             ([3, 64, 64, 64, 1], ['tanh'] * 3 + ['identity'], 0),       # deeper
             ([10, 128, 128, 1], ['tanh', 'tanh', 'identity'], 4)]       # config 3 / 4 net: the 128-wide kernel (qb_tg8.cuh)
    for widths, acts, want in cases:
        layers, P = make_net(widths, acts)
        x = rs.rand(64, widths[0])
        y = rs.randn(64, widths[-1])
        prob = ops.Problem(netdesc_from_layers(layers, P), x, y, 0.3, dtype=torch.float32)
        assert prob.plan_info(4, True)['tensor_core'] == want, (widths, acts)
        # whatever the path, the gradient is right
        th = 0.3 * rs.randn(2, P)
        lp, g = ops.logpost_grad(prob, th)
        rl, rg = qo.logpost_grad(layers, th[0].astype(np.float32).astype(np.float64), x, y, 0.3)
        assert abs(lp[0].item() - rl) <= 1e-4 * abs(rl)
        assert np.abs(g[0].double().cpu().numpy() - rg).max() <= 2e-3 * np.abs(rg).max()
    prob64 = ops.Problem(netdesc_from_layers(*make_net([3, 64, 64, 1], ['tanh', 'tanh', 'identity'])), rs.rand(64, 3), rs.randn(64, 1), 0.3,
                         dtype=torch.float64)
    assert prob64.plan_info(4, True)['tensor_core'] == 0               # fp64 stays on the CUDA cores


def test_tc_gradient_full_size_config5_and_config2():
    """One direct oracle comparison at the bench sizes: N = 10^4 points = 79 tiles (odd count, ragged last tile)."""
    from quinn_b200 import ops
    for widths, N, sigma in (([3, 64, 64, 1], 10_000, 0.05), ([2, 32, 32, 1], 1_000, 0.02)):
        rs = np.random.RandomState(N)
        layers, P = make_net(widths, ['tanh', 'tanh', 'identity'])
        x = (rs.rand(N, widths[0]) * 2 - 1) * np.pi
        y = np.sum(np.sin(x), axis=1, keepdims=True) + sigma * rs.randn(N, 1)
        K = 200
        th = (rs.rand(K, P) if widths[1] == 64 else 0.1 * rs.randn(K, P)).astype(np.float32).astype(np.float64)
        prob = ops.Problem(netdesc_from_layers(layers, P), x.astype(np.float32), y.astype(np.float32), sigma, dtype=torch.float32)
        assert prob.plan_info(K, True)['tensor_core'] == plan_code(widths, 'tanh')
        lp, g = ops.logpost_grad(prob, th)
        lp, g = lp.cpu().numpy(), g.double().cpu().numpy()
        xs, ys = x.astype(np.float32).astype(np.float64), y.astype(np.float32).astype(np.float64)
        _check(layers, P, xs, ys, th[[0, K - 1]], sigma, lp[[0, K - 1]], g[[0, K - 1]])


@pytest.mark.parametrize('method', ['hmc', 'mala'])
def test_tc_hmc_replay_matches_oracle_chain(method):
    """HMC / MALA on the tensor-core gradient kernel, fed recorded momenta and uniforms: same accept / reject decisions as
    the oracle's restatement of hmc.py:27-70 / mala.py:24-53 up to a tie at fp32 noise, fp32-close log-posteriors."""
    from quinn_b200 import ops
    rs = np.random.RandomState(55)
    layers, P = make_net([3, 64, 64, 1], ['tanh', 'tanh', 'identity'])
    N, K, steps, sigma, eps = 400, 3, 40, 0.3, 2e-3
    x, y = _data(rs, N, 3)
    th0 = (0.3 * rs.randn(K, P)).astype(np.float32).astype(np.float64)
    mom = rs.randn(steps, K, P).astype(np.float32).astype(np.float64)
    u = rs.rand(steps, K)
    prob = ops.Problem(netdesc_from_layers(layers, P), x, y, sigma, dtype=torch.float32)
    assert prob.plan_info(K, True)['tensor_core'] == plan_code([3, 64, 64, 1], 'tanh')
    st = ops.ChainState(prob, th0)
    rec = ops.Recorder(st, steps)
    hm = ops.HmcState(st, epsilon=eps, L=3, method=method)
    ops.hmc_run(st, hm, steps, rec, incr=torch.as_tensor(mom, dtype=torch.float32, device='cuda'), unif=torch.as_tensor(u, device='cuda'))
    acc = rec.accepted.cpu().numpy().astype(bool)
    lps = rec.logpost.cpu().numpy()
    lpf = lambda th: qo.logpost(layers, th, x, y, sigma)                            # noqa: E731
    gf = lambda th: qo.logpost_grad(layers, th, x, y, sigma)[1]                     # noqa: E731
    for k in range(K):
        ref = qo.run_chain(lpf, th0[k], steps, method, dict(p=mom[:, k], u=u[:, k]), grad_fn=gf, epsilon=eps, L=3)
        same = acc[k] == ref['accepted']
        upto = steps
        if not same.all():
            upto = int(np.where(~same)[0][0])
            assert abs(u[upto, k] - ref['alphas'][1 + upto]) <= 2e-3, (k, upto, u[upto, k], ref['alphas'][1 + upto])
        assert upto >= 10
        # an fp32 chain drifts away from the fp64 one (positions and momenta are rounded at every leapfrog update)
        np.testing.assert_allclose(lps[k][:upto], ref['logpost'][1:1 + upto], rtol=2e-4, atol=5e-3)
        assert 0.05 < acc[k].mean() <= 1.0


def test_tc_hmc_philox_equals_cuda_core_chain():
    """Same seed, same Philox momenta: the tensor-core and the CUDA-core HMC kernels walk the same chains (decisions can
    only differ at fp32-noise ties) and every recorded log-posterior equals the oracle's at the stored state."""
    from quinn_b200 import ops
    rs = np.random.RandomState(56)
    layers, P = make_net([2, 32, 32, 1], ['tanh', 'tanh', 'identity'])
    N, K, steps, sigma = 1000, 96, 25, 0.2
    x, y = _data(rs, N, 2)
    th0 = 0.1 * rs.randn(K, P)
    prob = ops.Problem(netdesc_from_layers(layers, P), x, y, sigma, dtype=torch.float32)

    def run():
        st = ops.ChainState(prob, th0)
        rec = ops.Recorder(st, steps, store_every=1)
        ops.hmc_run(st, ops.HmcState(st, epsilon=1e-3, L=3, method='hmc'), steps, rec, seed=17)
        return st, rec

    st, rec = run()
    with no_tcg():
        st2, rec2 = run()
    a1, a2 = rec.accepted.cpu().numpy().astype(bool), rec2.accepted.cpu().numpy().astype(bool)
    assert (a1 != a2).mean() < 0.01
    assert 0.05 < a1.mean() <= 1.0
    same = (a1 == a2).all(axis=1)
    np.testing.assert_allclose(rec.logpost.cpu().numpy()[same], rec2.logpost.cpu().numpy()[same], rtol=5e-4)
    lps = rec.logpost.cpu().numpy()
    samples = rec.samples.double().cpu().numpy()
    for k in (0, K // 2, K - 1):
        for s in (0, steps - 1):
            ref = qo.logpost(layers, samples[k, s], x, y, sigma)
            assert abs(lps[k, s] - ref) <= TOL_LP * abs(ref)
