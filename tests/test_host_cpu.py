"""CPU-only checks: the C-ABI library loads and exports every symbol the header declares, the flat-layout
descriptor matches the reference's parameter order, and the host-side helpers behave."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from golden_util import ROOT, NET_CASES, oracle_layers


def test_library_exports_every_header_symbol():
    from quinn_b200 import _lib
    hdr = open(os.path.join(ROOT, 'include', 'quinn_b200.h')).read()
    hdr = re.sub(r'/\*.*?\*/', '', hdr, flags=re.S)
    names = set(re.findall(r'\b(qb_[a-z_0-9]+)\s*\(', hdr))
    assert names, 'no prototypes found'
    lib = ctypes.CDLL(_lib.lib_path()) if os.path.exists(_lib.lib_path()) else _lib.load()
    for n in sorted(names):
        assert hasattr(lib, n), f'{n} declared in include/quinn_b200.h but not exported'
    assert names == set(_lib.SYMBOLS), (names ^ set(_lib.SYMBOLS))
    lib = _lib.load()
    assert lib.qb_version() == _lib.QB_ABI_VERSION
    assert lib.qb_launch_count() >= 0


def test_struct_sizes_match_the_header():
    from quinn_b200 import _lib
    assert ctypes.sizeof(_lib.qb_layer_t) == 88
    assert ctypes.sizeof(_lib.qb_net_t) == 32 + 16 * 88
    assert ctypes.sizeof(_lib.qb_data_t) == 24
    assert ctypes.sizeof(_lib.qb_lik_t) == 40


def test_plan_info_runs_without_a_gpu():
    """Launch planning is host code: tile size, threads, shared memory, N-splits."""
    from quinn_b200 import _lib
    from golden_util import netdesc_from_spec
    desc = netdesc_from_spec(NET_CASES['mlp_c5'])
    out = (ctypes.c_int64 * 8)()
    net = desc.to_c()
    assert _lib.load().qb_plan_info(ctypes.byref(net), _lib.QB_F32, 100000, 10000, 0, out) == 0
    assert out[0] in (32, 64, 128, 256) and out[1] % 32 == 0 and 0 < out[2] <= 227 * 1024 and out[3] == 1
    assert _lib.load().qb_plan_info(ctypes.byref(net), _lib.QB_F32, 4, 10000, 1, out) == 0
    assert out[3] > 1                      # few chains -> the data axis is split over blocks
    bad = desc.to_c()
    bad.layers[1].n_in = 7
    assert _lib.load().qb_plan_info(ctypes.byref(bad), _lib.QB_F32, 4, 100, 0, out) != 0
    assert b'n_in' in _lib.load().qb_last_error()


@pytest.mark.parametrize('name', list(NET_CASES))
def test_netdesc_from_module_matches_reference_layout(name):
    """quinn_b200's MLP / RNet register parameters in the reference's order, and netdesc_from_module
    describes them exactly like the oracle's layer list (which is pinned to the reference)."""
    from quinn_b200.nns import MLP, RNet, Poly, NonPar
    from quinn_b200.netdesc import netdesc_from_module
    spec = NET_CASES[name]
    if spec['kind'] == 'mlp':
        m = MLP(spec['indim'], spec['outdim'], spec['hls'], biasorno=spec['bias'], activ=spec['activ'])
    else:
        wp = Poly(0) if spec['shared'] else NonPar(spec['nlayers'] + 1)
        m = RNet(spec['rdim'], spec['nlayers'], wp_function=wp, indim=spec['indim'], outdim=spec['outdim'],
                 layer_pre=True, layer_post=True, biasorno=spec['bias'], nonlin=spec['nonlin'], mlp=spec['mlp'])
    layers, P = oracle_layers(spec)
    d = netdesc_from_module(m)
    assert d.n_params == P == m.numpar()
    assert d.as_oracle_layers() == layers
    # the torch forward of the mirror module equals the oracle forward at the module's own flat parameters
    from oracle import quinn_oracle as qo
    from quinn_b200.netdesc import flatten_module
    x = np.random.RandomState(1).rand(5, spec['indim'])
    ref = qo.forward(layers, flatten_module(m), x)
    np.testing.assert_allclose(m(torch.as_tensor(x)).detach().numpy(), ref, rtol=1e-12, atol=1e-14)


def test_mlp_numpar_and_unsupported_modules():
    from quinn_b200.nns import MLP
    from quinn_b200.netdesc import netdesc_from_module
    assert MLP(2, 1, (5,)).numpar() == 21                      # reference tests/test_mlp.py:45-54
    with pytest.raises(NotImplementedError):
        netdesc_from_module(MLP(2, 1, (5,), bnorm=True))
    with pytest.raises(NotImplementedError):
        netdesc_from_module(MLP(2, 1, (5,), dropout=0.1))
    assert netdesc_from_module(MLP(2, 1, (5,), final_transform='exp')).final_exp


def test_ops_refuse_cpu():
    from quinn_b200 import ops
    from golden_util import netdesc_from_spec
    desc = netdesc_from_spec(NET_CASES['mlp_relu'])
    with pytest.raises(RuntimeError):
        ops.Problem(desc, np.zeros((3, 2)), np.zeros((3, 1)), 0.1, device='cpu')


def test_thinning_rule_and_shard_ranges():
    from quinn_b200 import dist
    from oracle import quinn_oracle as qo
    assert qo.mcmc_thinning_rows(1001, 5, 100) == [100, 280, 460, 640, 820]
    for K, W in [(10, 3), (100000, 8), (7, 8), (1024, 4)]:
        parts = [dist.shard_range(K, r, W) for r in range(W)]
        assert parts[0][0] == 0 and parts[-1][1] == K
        assert all(a[1] == b[0] for a, b in zip(parts[:-1], parts[1:]))
        sizes = [b - a for a, b in parts]
        assert max(sizes) - min(sizes) <= 1


def test_nn_vi_num_batches_formula():
    # nn_vi.py:97-100
    f = lambda ntrn, bs: ntrn if bs == 1 else (ntrn + 1) // bs      # noqa: E731
    assert f(100, 1) == 100 and f(100, 100) == 1 and f(100, 32) == 3 and f(13, 13) == 1


def test_tensor_core_plan_selection_is_host_logic():
    """qb_plan_info reports which value-path kernel a call will take (include/quinn_b200.h): tcgen05 for eligible fp32
    MLPs (config 5: 256 compute threads + the issue warp / 256 tensor-memory columns; config 3: 512 + 32 / 512), CUDA cores for fp64 and for
    ineligible shapes, and for everything when QB_NO_TC=1."""
    import os
    from quinn_b200 import _lib
    from golden_util import netdesc_from_spec
    lib = _lib.load()
    out = (ctypes.c_int64 * 8)()

    def info(name, dtype, K=1000, N=10000, grad=0):
        net = netdesc_from_spec(NET_CASES[name]).to_c()
        assert lib.qb_plan_info(ctypes.byref(net), dtype, K, N, grad, out) == 0
        return list(out)

    c5 = info('mlp_c5', _lib.QB_F32)
    assert c5[0] == 128 and c5[1] == 288 and c5[6] == 2 and c5[7] == 256 and 76 * 1024 < c5[2] <= 227 * 1024
    c3 = info('mlp_c3', _lib.QB_F32)
    assert c3[1] == 544 and c3[6] == 2 and c3[7] == 512
    c2 = info('mlp_c2', _lib.QB_F32)
    assert c2[6] == 2 and c2[7] == 128
    assert info('mlp_c5', _lib.QB_F64)[6] == 0                  # fp64 stays on the CUDA cores
    g5 = info('mlp_c5', _lib.QB_F32, grad=1)                    # gradient path: tcgen05 kernel 2, fp16-split operands (qb_tg8.cuh)
    assert g5[6] == 4 and g5[1] == 288 and g5[7] == 256 and g5[2] <= 113 * 1024
    os.environ['QB_TG8_64'] = '0'
    try:
        g5 = info('mlp_c5', _lib.QB_F32, grad=1)                # ... or the 3xTF32 kernel (qb_tcg.cuh)
        assert g5[6] == 3 and g5[1] == 512 and g5[7] == 512 and g5[2] <= 227 * 1024
        g2 = info('mlp_c2', _lib.QB_F32, grad=1)
        assert g2[6] == 3 and g2[1] == 256 and g2[7] == 256
    finally:
        del os.environ['QB_TG8_64']
    g2 = info('mlp_c2', _lib.QB_F32, grad=1)                    # 32-wide tanh net: zero-padded on the 64-wide fp16-split kernel
    assert g2[6] == 4 and g2[1] == 288 and g2[7] == 256
    g3 = info('mlp_c3', _lib.QB_F32, grad=1)                    # 128-wide: fp16-split kernel 2 (qb_tg8.cuh)
    assert g3[6] == 4 and g3[1] == 544 and g3[7] == 512 and g3[2] <= 227 * 1024
    assert info('mlp_c5', _lib.QB_F64, grad=1)[6] == 0
    os.environ['QB_NO_TCG'] = '1'
    try:
        assert info('mlp_c5', _lib.QB_F32, grad=1)[6] == 0 and info('mlp_c5', _lib.QB_F32)[6] == 2
        assert info('mlp_c3', _lib.QB_F32, grad=1)[6] == 0
    finally:
        del os.environ['QB_NO_TCG']
    assert info('rnet_c1', _lib.QB_F32)[6] == 0                 # residual nets are not eligible
    assert info('mlp_tanh_o2', _lib.QB_F32)[6] == 0             # hidden widths 8, 6: not multiples of 16
    os.environ['QB_NO_TC'] = '1'
    try:
        assert info('mlp_c5', _lib.QB_F32)[6] == 0
    finally:
        del os.environ['QB_NO_TC']
