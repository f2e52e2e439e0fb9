"""Tensor-core value path (quinn_b200/csrc/qb_tc.cuh: tcgen05.mma kind::tf32, operands split hi/lo) against the oracle
and against the CUDA-core kernel it replaces: eligible architectures of every shape class (pipelined one-hidden-GEMM
nets, deeper nets, 128-wide nets, several outputs, missing biases, relu / identity, exp output), ragged and tiny N,
N-splits for few chains, saturation, NaN propagation, fused AMCMC chains (replay and Philox) and the predictive kernel."""
import os

import numpy as np
import pytest
import torch

from golden_util import netdesc_from_layers
from oracle import quinn_oracle as qo

pytestmark = pytest.mark.gpu

TOL32 = 1e-5          # relative, fp32 tensor-core log-posterior vs the fp64 oracle (north_star asks 1e-4)


def make_net(widths, acts, rs=None, bias=True):
    layers, off = [], 0
    for l in range(len(widths) - 1):
        n_in, n_out = widths[l], widths[l + 1]
        w = off
        off += n_in * n_out
        b = -1
        if bias if rs is None else (rs.rand() < 0.8):
            b = off
            off += n_out
        layers.append(dict(n_in=n_in, n_out=n_out, w_off=w, b_off=b, act=acts[l], res_step=0.0))
    return layers, off


class no_tc:
    """Run the enclosed calls on the CUDA-core kernels (the library reads QB_NO_TC at every launch)."""
    def __enter__(self):
        self.old = os.environ.get('QB_NO_TC')
        os.environ['QB_NO_TC'] = '1'

    def __exit__(self, *a):
        if self.old is None:
            del os.environ['QB_NO_TC']
        else:
            os.environ['QB_NO_TC'] = self.old


SHAPES = [
    # widths, activations, N, K
    ([3, 64, 64, 1], ['tanh', 'tanh', 'identity'], 1000, 6),          # config 5 (pipelined path)
    ([3, 64, 64, 1], ['tanh', 'tanh', 'identity'], 128, 3),           # exactly one tile
    ([3, 64, 64, 1], ['tanh', 'tanh', 'identity'], 129, 3),           # one point in the second tile
    ([3, 64, 64, 1], ['tanh', 'tanh', 'identity'], 5, 2),             # far less than a tile
    ([2, 32, 32, 1], ['tanh', 'tanh', 'identity'], 1000, 9),          # config 2 net
    ([1, 16, 48, 1], ['relu', 'relu', 'identity'], 300, 4),           # relu, pipelined, K != N
    ([3, 48, 16, 2], ['tanh', 'tanh', 'tanh'], 257, 4),               # two outputs, tanh on the last layer
    ([7, 64, 64, 4], ['tanh', 'tanh', 'identity'], 400, 3),           # padded input width 16, four outputs
    ([15, 32, 64, 1], ['tanh', 'tanh', 'identity'], 200, 3),          # widest input the path takes
    ([3, 64, 64, 1], ['tanh', 'relu', 'identity'], 300, 3),           # mixed activations -> generic path
    ([5, 16, 48, 32, 2], ['tanh', 'relu', 'identity', 'identity'], 333, 5),     # two tensor-core layers
    ([10, 64, 64, 64, 64, 4], ['tanh'] * 4 + ['identity'], 500, 4),   # three tensor-core layers
    ([10, 128, 128, 1], ['tanh', 'tanh', 'identity'], 700, 3),        # config 3 net: 384 tensor-memory columns
]


@pytest.mark.parametrize('case', range(len(SHAPES)))
def test_tc_logpost_matches_oracle_and_simt(case):
    from quinn_b200 import ops
    widths, acts, N, K = SHAPES[case]
    rs = np.random.RandomState(300 + case)
    layers, P = make_net(widths, acts, rs if case % 3 == 2 else None)
    desc = netdesc_from_layers(layers, P)
    x = rs.rand(N, widths[0]) * 2 - 1
    y = np.sin(x.sum(1, keepdims=True)) + 0.1 * rs.randn(N, widths[-1])
    th = 0.5 * rs.randn(K, P)
    prob = ops.Problem(desc, x, y, 0.1, dtype=torch.float32)
    info = prob.plan_info(K)
    assert info['tensor_core'] in (1, 2), info
    lp_tc = ops.logpost(prob, th).cpu().numpy()
    with no_tc():
        assert prob.plan_info(K)['tensor_core'] == 0
        lp_simt = ops.logpost(prob, th).cpu().numpy()
    for k in range(K):
        ref = qo.logpost(layers, th[k], x, y, 0.1)
        assert abs(lp_tc[k] - ref) <= TOL32 * abs(ref), (case, k, lp_tc[k], ref)
        assert abs(lp_simt[k] - ref) <= TOL32 * abs(ref), (case, k, lp_simt[k], ref)


def test_tc_plan_eligibility():
    """fp64, residual nets, widths that are not multiples of 16, wide outputs and 2-layer nets stay on the CUDA cores."""
    from quinn_b200 import ops
    rs = np.random.RandomState(0)
    x, y = rs.rand(50, 3), rs.rand(50, 1)

    def tc(widths, acts, dtype=torch.float32, res=False, x_=x, y_=y):
        layers, P = make_net(widths, acts)
        if res:
            layers[1]['res_step'] = 0.5
        prob = ops.Problem(netdesc_from_layers(layers, P), x_, y_, 0.1, dtype=dtype)
        return prob.plan_info(100)['tensor_core']

    t3 = ['tanh', 'tanh', 'identity']
    assert tc([3, 64, 64, 1], t3) == 2
    assert tc([3, 64, 64, 1], ['tanh', 'relu', 'identity']) == 1
    assert tc([3, 64, 64, 64, 1], ['tanh'] * 3 + ['identity']) == 1
    assert tc([3, 64, 64, 1], t3, dtype=torch.float64) == 0
    assert tc([3, 64, 64, 1], t3, res=True) == 0
    assert tc([3, 11, 11, 1], t3) == 0
    assert tc([3, 64, 1], ['tanh', 'identity']) == 0
    assert tc([3, 64, 64, 5], t3, y_=rs.rand(50, 5)) == 0
    assert tc([16, 64, 64, 1], t3, x_=rs.rand(50, 16)) == 0


def test_tc_few_chains_split_over_blocks():
    """K = 2 chains, N = 5000: the data axis is split over many blocks and k_finalize adds the partial sums."""
    from quinn_b200 import ops
    rs = np.random.RandomState(5)
    layers, P = make_net([3, 64, 64, 1], ['tanh', 'tanh', 'identity'])
    x = rs.rand(5000, 3) * 2 - 1
    y = np.sin(x.sum(1, keepdims=True))
    th = 0.4 * rs.randn(2, P)
    prob = ops.Problem(netdesc_from_layers(layers, P), x, y, 0.2, dtype=torch.float32)
    info = prob.plan_info(2)
    assert info['tensor_core'] == 2 and info['splits'] > 1
    lp = ops.logpost(prob, th).cpu().numpy()
    for k in range(2):
        ref = qo.logpost(layers, th[k], x, y, 0.2)
        assert abs(lp[k] - ref) <= TOL32 * abs(ref)


def test_tc_saturation_and_nan():
    """Huge pre-activations saturate tanh to +-1 exactly as in the oracle (the shared reciprocal clamps its exponent),
    and a NaN parameter makes the log-posterior NaN (so the chain rejects it, mcmc.py:77 strict <)."""
    from quinn_b200 import ops
    rs = np.random.RandomState(6)
    layers, P = make_net([3, 64, 64, 1], ['tanh', 'tanh', 'identity'])
    x = rs.rand(300, 3) * 2 - 1
    y = rs.randn(300, 1)
    th = rs.randn(3, P)
    th[0] *= 40.0                                  # |z| up to several hundred
    th[1, : 3 * 64] *= 1e4                         # layer 0 saturates completely
    prob = ops.Problem(netdesc_from_layers(layers, P), x, y, 0.5, dtype=torch.float32)
    lp = ops.logpost(prob, th).cpu().numpy()
    for k in range(3):
        ref = qo.logpost(layers, th[k], x, y, 0.5)
        assert np.isfinite(lp[k]) and abs(lp[k] - ref) <= 1e-4 * abs(ref), (k, lp[k], ref)
    th[2, 3 * 64 + 64 + 17] = np.nan               # one hidden weight
    lp = ops.logpost(prob, th).cpu().numpy()
    assert np.isnan(lp[2]) and np.isfinite(lp[0])


@pytest.mark.parametrize('seed', range(12))
def test_tc_random_eligible_architecture(seed):
    from quinn_b200 import ops
    rs = np.random.RandomState(4000 + seed)
    d, o = int(rs.randint(1, 16)), int(rs.randint(1, 5))
    nh = int(rs.randint(2, 5))
    pool = [16, 32, 48, 64, 80, 128] if nh == 2 else [16, 32, 48, 64]      # split weights of every layer stay in smem
    widths = [d] + [int(rs.choice(pool)) for _ in range(nh)] + [o]
    acts = [['tanh', 'relu', 'identity'][rs.randint(3)] for _ in range(nh)] + [['identity', 'tanh'][rs.randint(2)]]
    layers, P = make_net(widths, acts, rs)
    final = 'exp' if rs.rand() < 0.25 else None
    desc = netdesc_from_layers(layers, P, final_exp=final == 'exp')
    N, K = int(rs.choice([1, 31, 128, 200, 515])), int(rs.choice([1, 3, 7]))
    x = rs.rand(N, d) * 2 - 1
    y = rs.randn(N, o) * 0.5 + (1.0 if final else 0.0)
    th = rs.randn(K, P) / np.sqrt(np.maximum(2, max(widths))) * 2
    prob = ops.Problem(desc, x, y, 0.3, dtype=torch.float32)
    assert prob.plan_info(K)['tensor_core'] in (1, 2), widths
    lp = ops.logpost(prob, th).cpu().numpy()
    for k in range(K):
        ref = qo.logpost(layers, th[k], x, y, 0.3, final=final)
        assert abs(lp[k] - ref) <= 2e-5 * max(abs(ref), 1.0), (seed, widths, acts, k, lp[k], ref)


def test_tc_amcmc_replay_matches_oracle_chain():
    """The fused AMCMC kernel on the tensor-core path, fed recorded increments and uniforms, reproduces the oracle's
    restatement of MCMCBase.run (mcmc.py:55-85): same accept / reject decisions, fp32-close log-posteriors."""
    from quinn_b200 import ops
    rs = np.random.RandomState(77)
    layers, P = make_net([3, 64, 64, 1], ['tanh', 'tanh', 'identity'])
    N, K, steps = 300, 4, 60
    x = rs.rand(N, 3) * 2 - 1
    y = np.sin(x.sum(1, keepdims=True)) + 0.05 * rs.randn(N, 1)
    th0 = 0.3 * rs.randn(K, P)
    incr = 0.01 * rs.randn(steps, K, P)
    u = rs.rand(steps, K)
    prob = ops.Problem(netdesc_from_layers(layers, P), x, y, 0.3, dtype=torch.float32)
    assert prob.plan_info(K)['tensor_core'] == 2
    st = ops.ChainState(prob, th0)
    rec = ops.Recorder(st, steps)
    ops.amcmc_run(st, ops.AmcmcState(st, gamma=0.1), steps, rec, incr=torch.as_tensor(incr, dtype=torch.float32, device='cuda'),
                  unif=torch.as_tensor(u, device='cuda'))
    acc = rec.accepted.cpu().numpy().astype(bool)
    lps = rec.logpost.cpu().numpy()
    alphas = rec.alpha.cpu().numpy()
    lpf = lambda th: qo.logpost(layers, th, x, y, 0.3)                   # noqa: E731
    for k in range(K):
        ref = qo.run_chain(lpf, th0[k].astype(np.float32).astype(np.float64), steps, 'amcmc',
                           dict(xi=incr[:, k].astype(np.float32).astype(np.float64), u=u[:, k]))
        # decisions can only differ where u is within fp32 noise of the MH ratio; compare up to such a tie
        same = acc[k] == ref['accepted']
        upto = steps
        if not same.all():
            upto = int(np.where(~same)[0][0])
            assert abs(u[upto, k] - ref['alphas'][1 + upto]) <= 1e-3, (k, upto, u[upto, k], ref['alphas'][1 + upto])
        assert upto >= 10
        np.testing.assert_allclose(lps[k][:upto], ref['logpost'][1:1 + upto], rtol=2e-5)
        a_ref = ref['alphas'][1:1 + upto]
        fin = np.isfinite(a_ref) & (a_ref < 1e3)
        np.testing.assert_allclose(alphas[k][:upto][fin], a_ref[fin], rtol=2e-2, atol=1e-6)


def test_tc_amcmc_philox_self_consistent_and_matches_simt_statistics():
    """Philox-driven AMCMC on the tensor-core path: every recorded log-posterior equals the oracle's value at the stored
    state, and the acceptance rate agrees with the CUDA-core kernel run from the same seed (same proposals)."""
    from quinn_b200 import ops
    rs = np.random.RandomState(78)
    layers, P = make_net([3, 64, 64, 1], ['tanh', 'tanh', 'identity'])
    N, K, steps = 500, 64, 40
    x = rs.rand(N, 3) * 2 - 1
    y = np.sin(x.sum(1, keepdims=True)) + 0.05 * rs.randn(N, 1)
    th0 = rs.rand(K, P)
    prob = ops.Problem(netdesc_from_layers(layers, P), x, y, 0.5, dtype=torch.float32)

    def run():
        st = ops.ChainState(prob, th0)
        rec = ops.Recorder(st, steps, store_every=1)
        ops.amcmc_run(st, ops.AmcmcState(st, gamma=0.01), steps, rec, seed=9)
        return st, rec

    st, rec = run()
    with no_tc():
        st2, rec2 = run()
    lps = rec.logpost.cpu().numpy()
    samples = rec.samples.cpu().numpy().astype(np.float64)
    for k in range(0, K, 16):
        for s in (0, steps // 2, steps - 1):
            ref = qo.logpost(layers, samples[k, s], x, y, 0.5)
            assert abs(lps[k, s] - ref) <= TOL32 * abs(ref)
    a1 = rec.accepted.cpu().numpy().astype(bool)
    a2 = rec2.accepted.cpu().numpy().astype(bool)
    assert (a1 != a2).mean() < 0.01                 # decisions differ only at fp32-noise ties
    assert abs(a1.mean() - a2.mean()) < 0.01
    assert 0.02 < a1.mean() < 0.98


@pytest.mark.parametrize('case', [0, 4, 6, 7, 9, 10, 12])
def test_tc_predict_matches_oracle(case):
    """Kernel 4 on the tensor cores (member-parallel forward + k_moments): outputs, mean and variance (ddof=1,
    quinn.py:95-100) against the oracle, for pipelined (2 and 4 column groups) and deeper nets, ragged N."""
    from quinn_b200 import ops
    widths, acts, N, K = SHAPES[case]
    rs = np.random.RandomState(900 + case)
    layers, P = make_net(widths, acts)
    final = 'exp' if case == 6 else None
    desc = netdesc_from_layers(layers, P, final_exp=final == 'exp')
    M = K + 2
    N = N + 37
    x = rs.rand(N, widths[0]) * 2 - 1
    th = 0.5 * rs.randn(M, P)
    out, mean, var = ops.predict(desc, th, x, dtype=torch.float32, want_out=True, want_moments=True)
    ref = qo.predict_ens(layers, th, x, final=final)
    scale = max(1.0, np.abs(ref).max())
    np.testing.assert_allclose(out.double().cpu().numpy(), ref, rtol=2e-5, atol=2e-5 * scale)
    np.testing.assert_allclose(mean.double().cpu().numpy(), ref.mean(0), rtol=2e-5, atol=2e-5 * scale)
    np.testing.assert_allclose(var.double().cpu().numpy(), ref.var(0, ddof=1), rtol=1e-3, atol=1e-5 * scale * scale)
    with no_tc():
        out2, _, _ = ops.predict(desc, th, x, dtype=torch.float32, want_out=True, want_moments=False)
    np.testing.assert_allclose(out.cpu().numpy(), out2.cpu().numpy(), rtol=2e-5, atol=2e-5 * scale)


def test_tc_predict_moments_only_large():
    """Moments without an output array (scratch path of ops.predict): 32 members x 20000 points of the config-3 net."""
    from quinn_b200 import ops
    rs = np.random.RandomState(33)
    layers, P = make_net([10, 128, 128, 1], ['tanh', 'tanh', 'identity'])
    desc = netdesc_from_layers(layers, P)
    x = rs.rand(20000, 10)
    th = rs.randn(32, P) / np.sqrt(128.0)
    _, mean, var = ops.predict(desc, th, x, dtype=torch.float32, want_out=False, want_moments=True)
    idx = rs.choice(20000, 200, replace=False)
    ref = qo.predict_ens(layers, th, x[idx])
    np.testing.assert_allclose(mean.double().cpu().numpy()[idx], ref.mean(0), rtol=2e-5, atol=2e-5)
    np.testing.assert_allclose(var.double().cpu().numpy()[idx], ref.var(0, ddof=1), rtol=1e-3, atol=1e-8)
