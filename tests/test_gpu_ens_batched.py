"""Batched ensemble training (quinn_b200/ens/batched.py) against the reference's procedure
restated with torch on the CPU: per member an Adam loop over minibatches of its own subset with best-model tracking by
the validation loss evaluated before each update (quinn/nns/nnfit.py:125-166, quinn/solvers/nn_ens.py:51-69)."""
import copy
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def torch_member_fit(net, x, y, val, nepochs, lrate, wd, batch_size, perms):
    """nnfit's loop for ONE member (float64, CPU), minibatch orders given."""
    net = copy.deepcopy(net).double()
    x_, y_ = torch.as_tensor(x), torch.as_tensor(y)
    xv_, yv_ = (x_, y_) if val is None else (torch.as_tensor(val[0]), torch.as_tensor(val[1]))
    mse = torch.nn.MSELoss(reduction='mean')
    opt = torch.optim.Adam(net.parameters(), lr=lrate, weight_decay=wd)
    n = x.shape[0]
    bs = n if batch_size is None or batch_size > n else batch_size
    best_loss, best_model, best_epoch, hist = 1e100, copy.deepcopy(net), 0, []
    for t in range(nepochs):
        perm = torch.arange(n) if perms is None else torch.as_tensor(perms[t])
        for i in range(0, n, bs):
            idx = perm[i:i + bs]
            loss = mse(net(x_[idx]), y_[idx])
            with torch.no_grad():
                crit = mse(net(xv_), yv_).item()
            hist.append(crit)
            if crit < best_loss:
                best_loss, best_model, best_epoch = crit, copy.deepcopy(net), t
            opt.zero_grad()
            loss.backward()
            opt.step()
    flat = lambda m: np.concatenate([p.detach().numpy().ravel() for p in m.parameters()])     # noqa: E731
    return flat(best_model), best_loss, best_epoch, flat(net), np.array(hist)


@pytest.mark.parametrize('mode', ['full_own_val', 'full_shared_val', 'minibatch_own_val', 'minibatch_shared_val_wd'])
def test_fit_members_matches_sequential_adam(mode):
    from quinn_b200.ens.batched import fit_members
    from quinn_b200.netdesc import netdesc_from_module, flatten_module
    from quinn_b200.nns import MLP
    rs = np.random.RandomState(5)
    torch.manual_seed(5)
    net = MLP(2, 2, (16, 16), activ='tanh').double()
    N, K, nsub = 50, 3, 37
    x = rs.rand(N, 2) * 2 - 1
    y = np.stack([np.sin(2 * x.sum(1)), x[:, 0] * x[:, 1]], axis=1) + 0.01 * rs.randn(N, 2)
    subsets = np.stack([rs.permutation(N)[:nsub] for _ in range(K)])
    val = None if 'own_val' in mode else (rs.rand(20, 2) * 2 - 1, rs.randn(20, 2))
    mini = mode.startswith('minibatch')
    nepochs = 25 if mini else 120
    bs = 10 if mini else None
    wd = 1e-3 if mode.endswith('wd') else 0.0
    perms = np.stack([[rs.permutation(nsub) for _ in range(nepochs)] for _ in range(K)]) if mini else None
    # members start from different weights here (the solver starts them equal; this is the stronger test)
    th0 = np.stack([flatten_module(net) + 0.05 * rs.randn(sum(p.numel() for p in net.parameters())) for _ in range(K)])
    desc = netdesc_from_module(net)
    res = fit_members(desc, th0, x, y, subsets, val=val, nepochs=nepochs, lrate=0.02, wd=wd, batch_size=bs, perms=perms,
                      dtype=torch.float64, verbose=False)
    from quinn_b200.netdesc import unflatten_module
    for k in range(K):
        nk = copy.deepcopy(net)
        unflatten_module(nk, th0[k])
        bt, bl, be, ft, hist = torch_member_fit(nk, x[subsets[k]], y[subsets[k]], val, nepochs, 0.02, wd, bs,
                                                None if perms is None else perms[k])
        np.testing.assert_allclose(res['history'][:, k].cpu().numpy(), hist, rtol=1e-7, atol=1e-12)
        np.testing.assert_allclose(res['theta'][k].cpu().numpy(), ft, rtol=1e-6, atol=1e-8)
        assert abs(res['best_loss'][k].item() - bl) <= 1e-8 * bl
        assert int(res['best_epoch'][k].item()) == be
        np.testing.assert_allclose(res['best_theta'][k].cpu().numpy(), bt, rtol=1e-6, atol=1e-8)


def test_nn_ens_fit_batched_equals_member_by_member():
    """NN_Ens.fit with the default MSE / Adam options takes the batched device path; with QB_ENS_SEQUENTIAL=1 it trains
    member by member with torch (nnfit).  Same np.random subsets -> same trained members (full batch)."""
    from quinn_b200.nns import MLP
    from quinn_b200.solvers import NN_Ens
    rs = np.random.RandomState(3)
    x = rs.rand(40, 2) * 2 - 1
    y = np.sin(3 * x[:, :1]) * x[:, 1:]
    xt = rs.rand(11, 2) * 2 - 1
    preds = {}
    for seq in ('', '1'):
        if seq:
            os.environ['QB_ENS_SEQUENTIAL'] = '1'
        try:
            np.random.seed(21)
            torch.manual_seed(21)
            net = MLP(2, 1, (16, 16), activ='tanh')
            ens = NN_Ens(net, nens=3, dfrac=0.75)
            ens.fit(x, y, nepochs=60, lrate=0.01, freq_out=1000)
            assert ens.batched_fit == (seq == '')
            preds[seq] = np.stack([l.predict(xt) for l in ens.learners])
        finally:
            os.environ.pop('QB_ENS_SEQUENTIAL', None)
    assert preds[''].shape == (3, 11, 1)
    assert np.abs(preds[''][0] - preds[''][1]).max() > 1e-6           # members differ (different subsets)
    np.testing.assert_allclose(preds[''], preds['1'], rtol=1e-6, atol=1e-8)


def test_fit_members_fp32_large_ensemble_trains():
    """256 members of the config-3 net shape (smaller data): the loss of every member goes down; fp32 evaluates the
    validation loss on the tensor-core kernel."""
    from quinn_b200.ens.batched import fit_members
    from golden_util import netdesc_from_layers
    from oracle import quinn_oracle as qo
    rs = np.random.RandomState(8)
    layers, P = qo.mlp_layers(10, 1, (32, 32), True, 'tanh')
    desc = netdesc_from_layers(layers, P)
    N, K = 512, 256
    x = rs.rand(N, 10)
    y = np.sin(x.sum(1, keepdims=True))
    th0 = rs.uniform(-0.3, 0.3, size=(K, P))
    subsets = np.stack([rs.permutation(N)[:400] for _ in range(K)])
    res = fit_members(desc, th0, x, y, subsets, val=(x, y), nepochs=40, lrate=0.01, dtype=torch.float32, verbose=False)
    h = res['history'].cpu().numpy()
    assert h.shape == (40, K) and np.isfinite(h).all()
    assert (h[-1] <= h[0]).all() and h[-1].mean() < 0.3 * h[0].mean()
    k = 17
    ref = np.mean((qo.forward(layers, res['best_theta'][k].double().cpu().numpy(), x) - y) ** 2)
    assert abs(res['best_loss'][k].item() - ref) <= 1e-3 * ref


@pytest.mark.parametrize('dtype', [torch.float64, torch.float32])
def test_adam_step_and_row_copy_kernels(dtype):
    """qb_adam_step against torch.optim.Adam on a flat array (ragged length, weight decay, gradient scale) and
    qb_copy_rows_where with all / some / no rows selected."""
    import ctypes as C
    from quinn_b200 import _lib
    from quinn_b200.ops import _ptr, _stream, qb_dtype
    lib = _lib.load()
    torch.manual_seed(0)
    n = 1000 + 37
    th = torch.randn(n, dtype=dtype, device='cuda')
    ref = th.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref], lr=0.03, weight_decay=0.01)
    m, v = torch.zeros_like(th), torch.zeros_like(th)
    for step in range(1, 6):
        g = torch.randn(n, dtype=dtype, device='cuda')
        _lib.check(lib.qb_adam_step(qb_dtype(dtype), _ptr(th), _ptr(g), _ptr(m), _ptr(v), n, 0.03, 0.9, 0.999, 1e-8, 0.01,
                                    step, -0.5, _stream()), 'qb_adam_step')
        ref.grad = -0.5 * g
        opt.step()
    tol = 1e-12 if dtype == torch.float64 else 2e-5
    torch.testing.assert_close(th, ref.detach(), rtol=tol, atol=tol)
    K, P = 5, 300 + 3
    src = torch.randn(K, P, dtype=dtype, device='cuda')
    for sel in ([1, 1, 1, 1, 1], [0, 1, 0, 0, 1], [0, 0, 0, 0, 0]):
        dst = torch.zeros(K, P, dtype=dtype, device='cuda')
        mask = torch.tensor(sel, dtype=torch.uint8, device='cuda')
        _lib.check(lib.qb_copy_rows_where(qb_dtype(dtype), _ptr(dst), _ptr(src), _ptr(mask), K, P, _stream()), 'copy')
        want = src * mask[:, None].to(dtype)
        assert torch.equal(dst, want)
    assert lib.qb_adam_step(qb_dtype(dtype), None, _ptr(th), _ptr(m), _ptr(v), n, 0.1, 0.9, 0.999, 1e-8, 0.0, 1, 1.0, None) != 0
    assert b'qb_adam_step' in lib.qb_last_error()


def test_fit_members_rnet_shared_weights():
    """RNet with Poly(0): one weight matrix is shared by all residual layers, so its gradient is a sum over layers
    (kernel 2) and Adam sees it once in the flat vector, exactly like torch does."""
    from quinn_b200.ens.batched import fit_members
    from quinn_b200.netdesc import netdesc_from_module, flatten_module
    from quinn_b200.nns import RNet, Poly
    rs = np.random.RandomState(12)
    torch.manual_seed(12)
    net = RNet(3, 3, wp_function=Poly(0), indim=1, outdim=1, layer_pre=True, layer_post=True, biasorno=True,
               nonlin=True, mlp=False, final_layer=None).double()
    N, K, nsub, nepochs = 30, 2, 24, 80
    x = rs.rand(N, 1) * 2 * np.pi - np.pi
    y = np.sin(x) + 0.02 * rs.randn(N, 1)
    subsets = np.stack([rs.permutation(N)[:nsub] for _ in range(K)])
    th0 = flatten_module(net)
    res = fit_members(netdesc_from_module(net), th0, x, y, subsets, nepochs=nepochs, lrate=0.02, dtype=torch.float64,
                      verbose=False)
    for k in range(K):
        bt, bl, be, ft, hist = torch_member_fit(net, x[subsets[k]], y[subsets[k]], None, nepochs, 0.02, 0.0, None, None)
        np.testing.assert_allclose(res['history'][:, k].cpu().numpy(), hist, rtol=1e-7, atol=1e-12)
        np.testing.assert_allclose(res['theta'][k].cpu().numpy(), ft, rtol=1e-6, atol=1e-8)
        assert int(res['best_epoch'][k].item()) == be


# ------------------------------------------------------------------------------------------------------------------
# golden fixtures produced by RUNNING THE REFERENCE (tests/golden/make_golden_r2.py): NN_Ens.fit / NN_RMS.fit
# ------------------------------------------------------------------------------------------------------------------
def _golden(name):
    from golden_util import load
    return load(name)


@pytest.mark.parametrize('name', ['ensfit_full', 'ensfit_mini'])
def test_fit_members_reproduces_the_reference_nn_ens_fit(name):
    """The batched device trainer against what the reference's NN_Ens.fit produced (nn_ens.py:51-69 -> nnfit.py:125-166):
    same member subsets and minibatch orders => same validation-loss history, best epoch, best and final weights."""
    from quinn_b200.ens.batched import fit_members
    from quinn_b200.netdesc import netdesc_from_module
    from quinn_b200.nns import MLP
    g = _golden(name + '.npz')
    net = MLP(2, 2, (16, 16), activ='tanh').double()
    desc = netdesc_from_module(net)
    bs = int(g['batch_size'])
    res = fit_members(desc, g['theta0'], g['x'], g['y'], g['subsets'], nepochs=int(g['nepochs']), lrate=float(g['lrate']),
                      batch_size=None if bs < 0 else bs, perms=g['perms'], dtype=torch.float64, verbose=False)
    hist = res['history'].cpu().numpy()                 # [niter, K] validation loss before every update
    np.testing.assert_allclose(hist.T, g['history'][:, :, 3], rtol=1e-7, atol=1e-12)
    np.testing.assert_allclose(res['best_loss'].cpu().numpy(), g['best_loss'], rtol=1e-8)
    assert np.array_equal(res['best_epoch'].cpu().numpy(), g['best_epoch'])
    np.testing.assert_allclose(res['best_theta'].cpu().numpy(), g['best_theta'], rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(res['theta'].cpu().numpy(), g['final_theta'], rtol=1e-6, atol=1e-8)


def test_nn_ens_fit_reproduces_the_reference_under_the_same_seeds():
    """The solver call itself (full batch): np.random.seed / torch.manual_seed as in the generating script."""
    from quinn_b200.nns import MLP
    from quinn_b200.solvers import NN_Ens
    from quinn_b200.netdesc import flatten_module, unflatten_module
    g = _golden('ensfit_full.npz')
    np.random.seed(21)
    torch.manual_seed(21)
    net = MLP(2, 2, (16, 16), activ='tanh')
    unflatten_module(net, g['theta0'])                  # the reference's initial weights (its own torch init stream)
    x, y = g['x'], g['y']
    np.random.rand(50, 2), np.random.randn(50, 2)       # the data draws of the generating script come first
    ens = NN_Ens(net, nens=int(g['nens']), dfrac=float(g['dfrac']))
    ens.fit(x, y, nepochs=int(g['nepochs']), lrate=float(g['lrate']), batch_size=None, freq_out=10 ** 9, freq_plot=10 ** 9)
    assert ens.batched_fit
    best = np.stack([flatten_module(l.best_model) for l in ens.learners])
    np.testing.assert_allclose(best, g['best_theta'], rtol=1e-6, atol=1e-8)


def test_nn_rms_reproduces_the_reference_fit():
    """NN_RMS (nn_rms.py:33-56) on the batched trainer: anchored NegLogPost loss, per-member anchors."""
    from quinn_b200.ens.batched import fit_members
    from quinn_b200.netdesc import netdesc_from_module, flatten_module, unflatten_module
    from quinn_b200.nns import MLP
    from quinn_b200.solvers import NN_RMS
    g = _golden('rms_fit.npz')
    net = MLP(2, 1, (12,), activ='tanh').double()
    desc = netdesc_from_module(net)
    anchors = g['anchors_raw'] * float(g['priorsigma'])
    res = fit_members(desc, g['theta0'], g['x'], g['y'], g['subsets'], nepochs=int(g['nepochs']), lrate=float(g['lrate']),
                      dtype=torch.float64, verbose=False,
                      logpost=dict(sigma=float(g['datanoise']), prior_sigma=float(g['priorsigma']), anchor=anchors))
    np.testing.assert_allclose(res['history'].cpu().numpy().T, g['history'][:, :, 3], rtol=1e-8)
    assert np.array_equal(res['best_epoch'].cpu().numpy(), g['best_epoch'])
    np.testing.assert_allclose(res['best_theta'].cpu().numpy(), g['best_theta'], rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(res['theta'].cpu().numpy(), g['final_theta'], rtol=1e-6, atol=1e-8)
    # the solver class under the generating script's seeds: same subsets, anchors and trained members
    np.random.seed(22)
    torch.manual_seed(22)
    net2 = MLP(2, 1, (12,), activ='tanh')
    unflatten_module(net2, g['theta0'])
    np.random.rand(40, 2), np.random.randn(40, 1)
    rms = NN_RMS(net2, nens=int(g['nens']), dfrac=float(g['dfrac']), datanoise=float(g['datanoise']), priorsigma=float(g['priorsigma']))
    rms.fit(g['x'], g['y'], nepochs=int(g['nepochs']), lrate=float(g['lrate']), freq_out=10 ** 9, freq_plot=10 ** 9)
    assert rms.batched_fit
    np.testing.assert_allclose(rms.anchors, anchors, rtol=0, atol=0)
    best = np.stack([flatten_module(l.best_model) for l in rms.learners])
    np.testing.assert_allclose(best, g['best_theta'], rtol=1e-6, atol=1e-8)
    assert rms.predict_ens(g['x'][:5]).shape == (3, 5, 1)
