"""Shared, reference-free description of the golden cases.

Both ``tests/golden/make_golden.py`` (which runs the reference) and the parity
tests import this, so the large random inputs (x, y, theta) never have to be
stored: they are regenerated from ``np.random.RandomState(seed)``.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, 'tests', 'golden')
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# name -> network + data spec
NET_CASES = {
    'mlp_tanh_o2':  dict(kind='mlp', indim=3, outdim=2, hls=(8, 6), bias=True, activ='tanh', N=17, sigma=0.3, seed=101, ntheta=6, tscale=0.8),
    'mlp_relu':     dict(kind='mlp', indim=2, outdim=1, hls=(5,), bias=True, activ='relu', N=20, sigma=0.1, seed=102, ntheta=6, tscale=1.0),
    'mlp_lin_nob':  dict(kind='mlp', indim=4, outdim=1, hls=(7, 7, 7), bias=False, activ='lin', N=9, sigma=0.5, seed=103, ntheta=4, tscale=0.5),
    'mlp_c2':       dict(kind='mlp', indim=2, outdim=1, hls=(32, 32), bias=True, activ='tanh', N=50, sigma=0.02, seed=104, ntheta=4, tscale=0.1),
    'mlp_c5':       dict(kind='mlp', indim=3, outdim=1, hls=(64, 64), bias=True, activ='tanh', N=40, sigma=0.05, seed=105, ntheta=3, tscale=0.2),
    'mlp_c3':       dict(kind='mlp', indim=10, outdim=1, hls=(128, 128), bias=True, activ='tanh', N=24, sigma=0.05, seed=106, ntheta=2, tscale=0.1),
    'rnet_c1':      dict(kind='rnet', rdim=3, nlayers=3, indim=1, outdim=1, bias=True, nonlin=True, mlp=False, shared=True, N=13, sigma=0.02, seed=107, ntheta=6, tscale=0.7),
    'rnet_nonpar':  dict(kind='rnet', rdim=4, nlayers=2, indim=2, outdim=2, bias=True, nonlin=True, mlp=False, shared=False, N=11, sigma=0.2, seed=108, ntheta=4, tscale=0.7),
    'rnet_mlpmode': dict(kind='rnet', rdim=5, nlayers=1, indim=2, outdim=1, bias=False, nonlin=True, mlp=True, shared=False, N=12, sigma=0.2, seed=109, ntheta=4, tscale=0.7),
}
# round 2: polynomial-in-depth RNet weights (rnet.py:244-347); fixtures from tests/golden/make_golden_r2.py
NET_CASES_R2 = {
    'rnet_lin':   dict(kind='rnet', rdim=4, nlayers=2, indim=2, outdim=1, bias=True, nonlin=True, mlp=False, poly=1, N=15, sigma=0.2, seed=201, ntheta=4, tscale=0.6),
    'rnet_quad':  dict(kind='rnet', rdim=3, nlayers=3, indim=1, outdim=2, bias=True, nonlin=True, mlp=False, poly=2, N=12, sigma=0.1, seed=202, ntheta=4, tscale=0.6),
    'rnet_cubic': dict(kind='rnet', rdim=5, nlayers=4, indim=2, outdim=1, bias=False, nonlin=True, mlp=True, poly=3, N=14, sigma=0.3, seed=203, ntheta=3, tscale=0.5),
    'rnet_poly4': dict(kind='rnet', rdim=3, nlayers=5, indim=1, outdim=1, bias=True, nonlin=False, mlp=False, poly=4, N=10, sigma=0.2, seed=204, ntheta=3, tscale=0.5),
}


def make_inputs(spec):
    rs = np.random.RandomState(spec['seed'])
    x = rs.rand(spec['N'], spec['indim']) * 2.0 - 1.0
    y = np.sin(x.sum(axis=1, keepdims=True) * np.arange(1, spec['outdim'] + 1)[None, :]) + 0.1 * rs.randn(spec['N'], spec['outdim'])
    return x, y


def make_thetas(spec, pdim):
    rs = np.random.RandomState(spec['seed'] + 1000)
    return spec['tscale'] * rs.randn(spec['ntheta'], pdim)


def oracle_layers(spec):
    """Layer list for oracle/quinn_oracle.py from a NET_CASES spec."""
    from oracle import quinn_oracle as qo
    if spec['kind'] == 'mlp':
        return qo.mlp_layers(spec['indim'], spec['outdim'], spec['hls'], spec['bias'], spec['activ'])
    return qo.rnet_layers(spec['rdim'], spec['nlayers'], spec['indim'], spec['outdim'], biasorno=spec['bias'],
                          nonlin=spec['nonlin'], mlp=spec['mlp'], shared=spec.get('shared', True), poly_order=spec.get('poly'))


def load(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def netdesc_from_spec(spec):
    """quinn_b200 NetDesc for a NET_CASES spec (built from the oracle's layer list, no torch module)."""
    from quinn_b200.netdesc import NetDesc, Layer
    layers, P = oracle_layers(spec)
    return netdesc_from_layers(layers, P)


def netdesc_from_layers(layers, P, final_exp=False):
    from quinn_b200.netdesc import NetDesc, Layer
    ls = [Layer(L['n_in'], L['n_out'], L['w_off'], L['b_off'], L['act'], L['res_step'], terms=L.get('terms')) for L in layers]
    return NetDesc(ls[0].n_in, ls[-1].n_out, P, ls, final_exp)
