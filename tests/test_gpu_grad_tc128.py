"""128-wide tensor-core gradient path (quinn_b200/csrc/qb_tg8.cuh: forward, back-propagation and both weight-gradient GEMMs
as kind::f16 MMAs on fp16 hi/lo splits with power-of-two scaling) against the fp64 oracle (manual reverse mode pinned to the
reference's autograd by tests/test_oracle_golden.py) and against the CUDA-core kernel it replaces.  BASELINE configs 3 / 4
(MLP 10-128-128-1: ensemble training, VI)."""
import ctypes as C

import numpy as np
import pytest
import torch

from golden_util import netdesc_from_layers
from oracle import quinn_oracle as qo
from test_gpu_tensorcore import make_net
from test_gpu_grad_tc import no_tcg, _check, _data

pytestmark = pytest.mark.gpu

ACTS = ['tanh', 'tanh', 'identity']

SHAPES = [
    # in, N, K
    (10, 1000, 3),       # config 3 / 4 net, few chains -> the data axis is split over blocks
    (10, 128, 2),        # exactly one tile
    (10, 129, 2),        # one point in the second tile
    (10, 1, 2),          # a single data point
    (10, 300, 400),      # many members: one block per member
    (15, 200, 3),        # widest input: all 16 columns of the X image in use
    (11, 150, 2),
    (7, 333, 2),
    (3, 77, 2),
    (1, 640, 2),
]


@pytest.mark.parametrize('case', range(len(SHAPES)))
def test_tc128_gradient_matches_oracle(case):
    from quinn_b200 import ops
    d, N, K = SHAPES[case]
    rs = np.random.RandomState(700 + case)
    layers, P = make_net([d, 128, 128, 1], ACTS)
    x, y = _data(rs, N, d)
    th = (0.5 * rs.randn(K, P)).astype(np.float32).astype(np.float64)
    prob = ops.Problem(netdesc_from_layers(layers, P), x, y, 0.2, dtype=torch.float32)
    info = prob.plan_info(K, True)
    assert info['tensor_core'] == 4 and info['threads'] == 544 and info['tmem_cols'] == 512 and info['smem_bytes'] <= 227 * 1024
    lp, g = ops.logpost_grad(prob, th)
    lp, g = lp.cpu().numpy(), g.double().cpu().numpy()
    idx = np.arange(K) if K <= 5 else np.array([0, K // 2, K - 1])
    if N == 1:
        # one data point, a residual of 0.04 sigma in chain 1 and saturated units: every entry of the gradient is tiny against the
        # bounds that scale the fp16 images, so the lo parts run out of bits (qb_tg8_stage) -- held to north_star's fp32 bar
        for k in idx:
            rl, rg = qo.logpost_grad(layers, th[k], x, y, 0.2)
            assert abs(lp[k] - rl) <= 1e-5 * abs(rl)
            assert np.abs(g[k] - rg).max() <= 1e-4 * np.abs(rg).max()
    else:
        _check(layers, P, x, y, th[idx], 0.2, lp[idx], g[idx])
    # and the CUDA-core kernel agrees to fp32 accuracy
    with no_tcg():
        assert prob.plan_info(K, True)['tensor_core'] == 0
        lp2, g2 = ops.logpost_grad(prob, th)
    np.testing.assert_allclose(lp2.cpu().numpy(), lp, rtol=2e-5)
    if N > 1:       # (the degenerate single-point case is held to the oracle above; both fp32 kernels carry their own error there)
        assert (np.abs(g2.double().cpu().numpy() - g).max(1) <= 1e-4 * np.abs(g).max(1)).all()


def test_tc128_gradient_without_biases_and_with_prior():
    from quinn_b200 import ops
    rs = np.random.RandomState(43)
    for bias in (False, True):
        layers, P = make_net([10, 128, 128, 1], ACTS, bias=bias)
        x, y = _data(rs, 500, 10)
        K = 3
        th = (0.4 * rs.randn(K, P)).astype(np.float32).astype(np.float64)
        for anchor in (0.1 * rs.randn(P), 0.1 * rs.randn(K, P)):
            anchor = anchor.astype(np.float32).astype(np.float64)
            prob = ops.Problem(netdesc_from_layers(layers, P), x, y, 0.3, dtype=torch.float32, prior_sigma=0.7,
                               prior_anchor=anchor, fulldatasize=1200)
            assert prob.plan_info(K, True)['tensor_core'] == 4
            lp, g = ops.logpost_grad(prob, th)
            _check(layers, P, x, y, th, 0.3, lp.cpu().numpy(), g.double().cpu().numpy(), prior=dict(sigma=0.7, anchor=anchor),
                   nfull=1200)


@pytest.mark.parametrize('wscale,xscale,yscale,sigma', [(0.05, 100.0, 30.0, 2.0), (3.0, 1e-3, 1e-2, 0.01), (0.3, 1.0, 1e4, 50.0),
                                                      (1e-3, 1.0, 1.0, 0.1), (0.3, 3e4, 1.0, 1.0)])
def test_tc128_operand_scaling_covers_the_range(wscale, xscale, yscale, sigma):
    """The fp16 operand images are scaled by powers of two derived from max |x|, max |y|, max |W1|, sum |wl| and sigma
    (qb_tg8_stage): inputs, targets, weights and noise levels far from 1 keep the fp32-level accuracy."""
    from quinn_b200 import ops
    rs = np.random.RandomState(5)
    layers, P = make_net([10, 128, 128, 1], ACTS)
    N, K = 400, 2
    x = ((rs.rand(N, 10) * 2 - 1) * xscale).astype(np.float32).astype(np.float64)
    y = ((np.sin(x.sum(1, keepdims=True) / xscale) + 0.1 * rs.randn(N, 1)) * yscale).astype(np.float32).astype(np.float64)
    th = (wscale * rs.randn(K, P)).astype(np.float32).astype(np.float64)
    th[:, :1280] /= xscale                      # keep layer 0 out of saturation
    th = th.astype(np.float32).astype(np.float64)
    prob = ops.Problem(netdesc_from_layers(layers, P), x, y, sigma, dtype=torch.float32)
    assert prob.plan_info(K, True)['tensor_core'] == 4
    lp, g = ops.logpost_grad(prob, th)
    _check(layers, P, x, y, th, sigma, lp.cpu().numpy(), g.double().cpu().numpy())


def test_tc128_plan_eligibility():
    from quinn_b200 import ops
    rs = np.random.RandomState(1)
    cases = [([10, 128, 128, 1], ACTS, 4), ([15, 128, 128, 1], ACTS, 4),
             ([16, 128, 128, 1], ACTS, 0),                                  # more than 15 inputs
             ([10, 128, 128, 2], ACTS, 0),                                  # two outputs
             ([10, 128, 64, 1], ACTS, 0),                                   # unequal widths
             ([10, 128, 128, 1], ['relu', 'relu', 'identity'], 0),          # tanh only (fp16 images need bounded activations)
             ([10, 128, 128, 128, 1], ['tanh'] * 3 + ['identity'], 0)]      # deeper
    for widths, acts, want in cases:
        layers, P = make_net(widths, acts)
        x = rs.rand(64, widths[0])
        y = rs.randn(64, widths[-1])
        prob = ops.Problem(netdesc_from_layers(layers, P), x, y, 0.3, dtype=torch.float32)
        assert prob.plan_info(4, True)['tensor_core'] == want, (widths, acts)
        th = 0.3 * rs.randn(2, P)
        lp, g = ops.logpost_grad(prob, th)
        rl, rg = qo.logpost_grad(layers, th[0].astype(np.float32).astype(np.float64), x, y, 0.3)
        assert abs(lp[0].item() - rl) <= 1e-4 * abs(rl)
        assert np.abs(g[0].double().cpu().numpy() - rg).max() <= 2e-3 * np.abs(rg).max()
    prob64 = ops.Problem(netdesc_from_layers(*make_net([10, 128, 128, 1], ACTS)), rs.rand(64, 10), rs.randn(64, 1), 0.3, dtype=torch.float64)
    assert prob64.plan_info(4, True)['tensor_core'] == 0               # fp64 stays on the CUDA cores


def test_tc128_per_member_data():
    """qb_logpost_members (batched ensemble training: member k reads x + k x_stride, y + k y_stride) on the 128-wide kernel."""
    from quinn_b200 import _lib
    from quinn_b200.ops import _ptr, _stream, qb_dtype
    rs = np.random.RandomState(9)
    layers, P = make_net([10, 128, 128, 1], ACTS)
    desc = netdesc_from_layers(layers, P)
    K, n = 5, 200
    x = (rs.rand(K, n, 10) * 2 - 1).astype(np.float32)
    y = (np.sin(x.sum(2, keepdims=True)) + 0.1 * rs.randn(K, n, 1)).astype(np.float32)
    th = (0.4 * rs.randn(K, P)).astype(np.float32)
    lib = _lib.load()
    cnet = desc.to_c()
    xd, yd, thd = (torch.as_tensor(a, device='cuda') for a in (x, y, th))
    lp = torch.empty(K, dtype=torch.float64, device='cuda')
    g = torch.empty((K, P), dtype=torch.float32, device='cuda')
    need = lib.qb_eval_workspace_bytes(C.byref(cnet), qb_dtype(torch.float32), K, n, 1)
    ws = torch.empty(int(need), dtype=torch.uint8, device='cuda')
    lik = _lib.qb_lik_t(0.25, 0.0, 1.0, None, 0, 0)
    data = _lib.qb_data_t(_ptr(xd), _ptr(yd), n)
    _lib.check(lib.qb_logpost_members(C.byref(cnet), qb_dtype(torch.float32), _ptr(thd), K, C.byref(data), n * 10, n, C.byref(lik),
                                      _ptr(lp), _ptr(g), _ptr(ws), ws.numel(), _stream()), 'qb_logpost_members')
    lp, g = lp.cpu().numpy(), g.double().cpu().numpy()
    for k in range(K):
        _check(layers, P, x[k].astype(np.float64), y[k].astype(np.float64), th[k:k + 1].astype(np.float64), 0.25, lp[k:k + 1], g[k:k + 1])


def test_tc128_full_size_config4():
    """One direct oracle comparison at the bench size of config 4: N = 10^5 points, data axis split over blocks."""
    from quinn_b200 import ops
    N, sigma, K = 100_000, 0.1, 3
    rs = np.random.RandomState(N)
    layers, P = make_net([10, 128, 128, 1], ACTS)
    x = (rs.rand(N, 10) * 2 - 1).astype(np.float32)
    y = (np.sin(x.sum(1, keepdims=True)) + sigma * rs.randn(N, 1)).astype(np.float32)
    th = (0.2 * rs.randn(K, P)).astype(np.float32).astype(np.float64)
    prob = ops.Problem(netdesc_from_layers(layers, P), x, y, sigma, dtype=torch.float32)
    info = prob.plan_info(K, True)
    assert info['tensor_core'] == 4 and info['splits'] > 1
    lp, g = ops.logpost_grad(prob, th)
    lp, g = lp.cpu().numpy(), g.double().cpu().numpy()
    _check(layers, P, x.astype(np.float64), y.astype(np.float64), th[[K - 1]], sigma, lp[[K - 1]], g[[K - 1]])


@pytest.mark.parametrize('method', ['hmc', 'mala'])
def test_tc128_hmc_chain_at_width_128(method):
    """HMC / MALA with the fused leapfrog on the 128-wide kernel (k_hmc_tc128), Philox momenta: the chains equal the CUDA-core
    kernel's (decisions can only differ at fp32-noise ties) and recorded log-posteriors equal the oracle's at the stored states."""
    from quinn_b200 import ops
    rs = np.random.RandomState(77)
    layers, P = make_net([10, 128, 128, 1], ACTS)
    N, K, steps, sigma = 600, 6, 12, 0.3
    x, y = _data(rs, N, 10)
    th0 = 0.1 * rs.randn(K, P)
    prob = ops.Problem(netdesc_from_layers(layers, P), x, y, sigma, dtype=torch.float32)
    assert prob.plan_info(K, True)['tensor_core'] == 4

    def run():
        st = ops.ChainState(prob, th0)
        rec = ops.Recorder(st, steps, store_every=1)
        ops.hmc_run(st, ops.HmcState(st, epsilon=2e-4, L=3, method=method), steps, rec, seed=23)
        return st, rec

    st, rec = run()
    with no_tcg():
        st2, rec2 = run()
    a1, a2 = rec.accepted.cpu().numpy().astype(bool), rec2.accepted.cpu().numpy().astype(bool)
    assert (a1 != a2).mean() < 0.05
    assert 0.05 < a1.mean() <= 1.0
    same = (a1 == a2).all(axis=1)
    assert same.any()
    np.testing.assert_allclose(rec.logpost.cpu().numpy()[same], rec2.logpost.cpu().numpy()[same], rtol=5e-4)
    lps = rec.logpost.cpu().numpy()
    samples = rec.samples.double().cpu().numpy()
    for k in (0, K - 1):
        for s in (0, steps - 1):
            ref = qo.logpost(layers, samples[k, s], x, y, sigma)
            assert abs(lps[k, s] - ref) <= 1e-5 * abs(ref)


def test_tc128_bnet_viloss_fp32_against_the_oracle():
    """BNet.viloss (bnet.py:181-232) in the fp32 throughput mode on the config-4 net: the nsam weight samples go through the
    128-wide tensor-core kernel 2; loss and d/d(mu, rho) against the oracle's restatement with the same draws."""
    from quinn_b200.nns import MLP
    from quinn_b200.vi import BNet
    rs = np.random.RandomState(12)
    torch.manual_seed(0)
    layers, P = qo.mlp_layers(10, 1, (128, 128), True, 'tanh')
    m = MLP(10, 1, (128, 128), activ='tanh')
    b = BNet(m, pi=0.5, sigma1=1.0, sigma2=0.1).float()
    mu = (0.2 * rs.randn(P)).astype(np.float32)
    rho = rs.uniform(-5.0, -4.0, size=P).astype(np.float32)
    off = 0
    with torch.no_grad():
        for i in range(b.nparams):
            n = b.params[2 * i].numel()
            b.params[2 * i].copy_(torch.as_tensor(mu[off:off + n]).view_as(b.params[2 * i]))
            b.params[2 * i + 1].copy_(torch.as_tensor(rho[off:off + n]).view_as(b.params[2 * i + 1]))
            off += n
    N, nsam, datanoise, num_batches = 700, 5, 0.2, 3
    x = (rs.rand(N, 10) * 2 - 1).astype(np.float32)
    y = (np.sin(x.sum(1, keepdims=True)) + 0.1 * rs.randn(N, 1)).astype(np.float32)
    eps = rs.randn(nsam, P).astype(np.float32)
    b.loss_params = [datanoise, nsam, num_batches]
    assert b.params[0].dtype == torch.float32
    loss = b.viloss(torch.as_tensor(x, device='cuda'), torch.as_tensor(y, device='cuda'), eps=eps)
    loss.backward()
    gmu = np.concatenate([b.params[2 * i].grad.double().cpu().numpy().ravel() for i in range(b.nparams)])
    grho = np.concatenate([b.params[2 * i + 1].grad.double().cpu().numpy().ravel() for i in range(b.nparams)])
    rl, rgmu, rgrho = qo.vi_loss(layers, mu.astype(np.float64), rho.astype(np.float64), eps.astype(np.float64), x.astype(np.float64),
                                 y.astype(np.float64), datanoise, num_batches, pi=0.5, sigma1=1.0, sigma2=0.1, want_grad=True)
    assert abs(loss.item() - rl) <= 1e-4 * abs(rl), (loss.item(), rl)
    assert np.abs(gmu - rgmu).max() <= 1e-4 * np.abs(rgmu).max()
    assert np.abs(grho - rgrho).max() <= 1e-4 * np.abs(rgrho).max()


def test_tc128_ensemble_training_fp32_follows_fp64():
    """Batched ensemble training of the config-3 net: a few Adam epochs in fp32 (kernel 2 = the 128-wide
    tensor-core kernel, per-member data) follow the fp64 run (CUDA-core kernels) to fp32 accuracy."""
    from quinn_b200.ens.batched import fit_members
    rs = np.random.RandomState(3)
    layers, P = qo.mlp_layers(10, 1, (128, 128), True, 'tanh')
    desc = netdesc_from_layers(layers, P)
    N, K = 400, 5
    x = rs.rand(N, 10)
    y = np.sin(x.sum(1, keepdims=True))
    th0 = rs.uniform(-0.1, 0.1, size=(K, P))
    subsets = np.stack([rs.permutation(N)[:320] for _ in range(K)])
    res = {}
    for dt in (torch.float64, torch.float32):
        res[dt] = fit_members(desc, th0, x, y, subsets, val=(x, y), nepochs=6, lrate=0.002, dtype=dt, verbose=False)
    h64, h32 = res[torch.float64]['history'].cpu().numpy(), res[torch.float32]['history'].cpu().numpy()
    assert h64.shape == (6, K) and np.isfinite(h32).all()
    np.testing.assert_allclose(h32, h64, rtol=2e-3)
    assert (h64[-1] < h64[0]).all()
