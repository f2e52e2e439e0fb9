"""Randomised architectures (depth, widths that are not multiples of the tile, activations, missing biases,
residual blocks, shared weights, exp output, several outputs) and shapes against the oracle: exercises the
non-warp-synchronous, ping-pong and division-based index paths of the kernels that the benchmark nets never take."""
import numpy as np
import pytest
import torch

from golden_util import netdesc_from_layers
from oracle import quinn_oracle as qo

pytestmark = pytest.mark.gpu


def random_net(rs):
    d = int(rs.randint(1, 7))
    o = int(rs.randint(1, 6))
    nh = int(rs.randint(0, 4))
    widths = [d] + [int(rs.choice([1, 2, 3, 5, 8, 12, 17, 24, 33, 40])) for _ in range(nh)] + [o]
    layers, off = [], 0
    acts = ['tanh', 'relu', 'identity']
    for l in range(len(widths) - 1):
        n_in, n_out = widths[l], widths[l + 1]
        w = off
        off += n_in * n_out
        b = -1
        if rs.rand() < 0.8:
            b = off
            off += n_out
        act = acts[rs.randint(3)] if l < len(widths) - 2 else 'identity'
        layers.append(dict(n_in=n_in, n_out=n_out, w_off=w, b_off=b, act=act, res_step=0.0))
        # sometimes follow a hidden layer with 1-2 residual steps, possibly sharing one weight matrix
        if l < len(widths) - 2 and rs.rand() < 0.4:
            share = rs.rand() < 0.5
            nres = int(rs.randint(1, 3))
            ws, bs = off, off + n_out * n_out
            if share:
                off += n_out * n_out + n_out
            for r in range(nres):
                if not share:
                    ws, bs = off, off + n_out * n_out
                    off += n_out * n_out + n_out
                layers.append(dict(n_in=n_out, n_out=n_out, w_off=ws, b_off=bs, act=acts[rs.randint(2)],
                                   res_step=float(rs.choice([0.25, 0.5, 1.0]))))
    return layers, off, d, o


@pytest.mark.parametrize('seed', range(24))
def test_random_architecture_matches_oracle(seed):
    from quinn_b200 import ops
    rs = np.random.RandomState(1000 + seed)
    layers, P, d, o = random_net(rs)
    final = 'exp' if rs.rand() < 0.2 else None
    desc = netdesc_from_layers(layers, P, final_exp=(final == 'exp'))
    N = int(rs.choice([1, 2, 7, 31, 32, 33, 100, 257, 700]))
    K = int(rs.choice([1, 2, 3, 5]))
    x = rs.rand(N, d) * 2 - 1
    y = rs.randn(N, o) * 0.5 + (1.0 if final else 0.0)
    th = 0.6 * rs.randn(K, P) / np.sqrt(max(2, max(L['n_in'] for L in layers))) * 2
    sigma = float(rs.choice([0.1, 0.5, 1.0]))
    use_prior = rs.rand() < 0.3
    kw = dict(prior_sigma=0.8, prior_anchor=0.1 * rs.randn(P), fulldatasize=2 * N + 3) if use_prior else {}
    okw = dict(fulldatasize=2 * N + 3, prior=dict(sigma=0.8, anchor=kw['prior_anchor'])) if use_prior else {}
    for dtype, tl, tg in ((torch.float64, 1e-10, 1e-9), (torch.float32, 2e-4, 5e-3)):
        prob = ops.Problem(desc, x, y, sigma, dtype=dtype, **kw)
        lp = ops.logpost(prob, th).cpu().numpy()
        lp2, g = ops.logpost_grad(prob, th)
        lp2, g = lp2.cpu().numpy(), g.double().cpu().numpy()
        out, mean, var = ops.predict(desc, th, x, dtype=dtype, want_out=True, want_moments=K > 1)
        for k in range(K):
            rl, rg = qo.logpost_grad(layers, th[k], x, y, sigma, final=final, **okw)
            assert abs(lp[k] - rl) <= tl * max(abs(rl), 1.0), (seed, dtype, k, lp[k], rl)
            assert abs(lp2[k] - rl) <= tl * max(abs(rl), 1.0), (seed, dtype, k, lp2[k], rl)
            scale = max(np.abs(rg).max(), 1e-3 * abs(rl), 1e-12)
            assert np.abs(g[k] - rg).max() <= tg * scale, (seed, dtype, k, np.abs(g[k] - rg).max(), scale)
            ref = qo.forward(layers, th[k], x, final=final)
            np.testing.assert_allclose(out[k].double().cpu().numpy(), ref, rtol=50 * tl, atol=50 * tl * max(1.0, np.abs(ref).max()))
        if K > 1:
            refs = qo.predict_ens(layers, th, x, final=final)
            np.testing.assert_allclose(mean.double().cpu().numpy(), refs.mean(0), rtol=50 * tl, atol=50 * tl * max(1.0, np.abs(refs).max()))


@pytest.mark.parametrize('seed', range(10))
@pytest.mark.parametrize('sampler', ['amcmc', 'hmc', 'mala'])
def test_random_architecture_chain_replay(seed, sampler):
    """Fused chain kernels on random nets, fed recorded draws, against the oracle's restatement of MCMCBase.run."""
    from quinn_b200 import ops
    rs = np.random.RandomState(2000 + seed)
    layers, P, d, o = random_net(rs)
    desc = netdesc_from_layers(layers, P)
    N, K, steps = int(rs.choice([5, 40, 130])), 3, 25
    x = rs.rand(N, d) * 2 - 1
    y = rs.randn(N, o) * 0.5
    th0 = 0.4 * rs.randn(K, P)
    sigma = 0.7
    u = rs.rand(steps, K)
    prob = ops.Problem(desc, x, y, sigma, dtype=torch.float64)
    st = ops.ChainState(prob, th0)
    rec = ops.Recorder(st, steps)
    lpf = lambda th: qo.logpost(layers, th, x, y, sigma)                 # noqa: E731
    gf = lambda th: qo.logpost_grad(layers, th, x, y, sigma)[1]          # noqa: E731
    if sampler == 'amcmc':
        incr = 0.03 * rs.randn(steps, K, P)
        ops.amcmc_run(st, ops.AmcmcState(st, gamma=0.1), steps, rec, incr=torch.as_tensor(incr, device='cuda'),
                      unif=torch.as_tensor(u, device='cuda'))
        refs = [qo.run_chain(lpf, th0[k], steps, 'amcmc', dict(xi=incr[:, k], u=u[:, k])) for k in range(K)]
    else:
        incr = rs.randn(steps, K, P)
        eps = 0.02
        ops.hmc_run(st, ops.HmcState(st, epsilon=eps, L=2, method=sampler), steps, rec,
                    incr=torch.as_tensor(incr, device='cuda'), unif=torch.as_tensor(u, device='cuda'))
        refs = [qo.run_chain(lpf, th0[k], steps, sampler, dict(p=incr[:, k], u=u[:, k]), grad_fn=gf, epsilon=eps, L=2)
                for k in range(K)]
    torch.cuda.synchronize()
    for k in range(K):
        # a decision can only differ if u sits within rounding of the MH ratio; ignore such razor-edge cases
        al = refs[k]['alphas'][1:]
        edge = np.abs(u[:, k] - al) < 1e-9 * np.maximum(al, 1e-300)
        if edge.any():
            continue
        assert np.array_equal(rec.accepted[k].cpu().numpy().astype(bool), refs[k]['accepted']), (seed, sampler, k)
        np.testing.assert_allclose(rec.samples[k].cpu().numpy(), refs[k]['chain'][1:], rtol=1e-8, atol=1e-10)
        np.testing.assert_allclose(rec.logpost[k].cpu().numpy(), refs[k]['logpost'][1:], rtol=1e-9)
        np.testing.assert_allclose(rec.logpost0[k].item(), refs[k]['logpost'][0], rtol=1e-10)
