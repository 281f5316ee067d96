"""CPU tier: the C-ABI library loads, exports every symbol include/bdof.h declares, its host-only
entry points agree with the oracle, and compute entry points fail loudly without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT, rel_l2
from oracle import multislice_oracle as mo


@pytest.fixture(scope='module')
def capi():
    from beyond_dof_b200.build import build_lib
    build_lib(verbose=False)
    from beyond_dof_b200 import capi
    return capi


def test_header_symbols_all_exported(capi):
    hdr = open(os.path.join(ROOT, 'include', 'bdof.h')).read()
    hdr = re.sub(r'/\*.*?\*/', '', hdr, flags=re.S)
    declared = sorted(set(re.findall(r'\b(bdof_[a-z0-9_]+)\s*\(', hdr)))
    assert declared, 'no declarations parsed'
    assert sorted(capi.SYMBOLS) == declared
    for name in declared:
        assert getattr(capi.lib, name) is not None


def test_version_and_sizes(capi):
    assert capi.lib.bdof_version() >= 100
    for n in (64, 128, 256, 512, 1024, 2048, 4096, 8192):
        assert capi.lib.bdof_size_supported(n) == 1
    for n in (18, 48, 72, 100, 45, 1000, 1536):              # mixed-radix passes (genericfft.cu)
        assert capi.lib.bdof_size_supported(n) == 2
    for n in (0, 1, 37, 74, 2049, 3072, 16384):
        assert capi.lib.bdof_size_supported(n) == 0


@pytest.mark.parametrize('args', [(1.0, 0.248, [1., 1., 1.], (64, 64)), (1.0, 1240. / 800, [0.67, 0.67, 0.67], (48, 80)),
                                  (2.5, 0.248, [1.0, 2.0, 2.5], (32, 40)), (1e3, 0.248, [1., 1., 1.], (256, 128))])
def test_kernel_factors_match_oracle(capi, args):
    dist, lam, vox, (ny, nx) = args
    hy = np.empty(ny, np.complex128); hx = np.empty(nx, np.complex128); p0 = np.empty(1, np.complex128)
    v = np.array(vox, np.float64)
    rc = capi.lib.bdof_kernel_factors(dist, lam, v.ctypes.data_as(ctypes.c_void_p), ny, nx, mo.PI_TF,
                                      hy.ctypes.data_as(ctypes.c_void_p), hx.ctypes.data_as(ctypes.c_void_p),
                                      p0.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0
    h = mo.get_kernel(dist, lam, vox, [ny, nx, 4])
    assert rel_l2(p0[0] * np.outer(hy, hx), h) < 1e-12
    p0o, hyo, hxo = mo.kernel_factors(dist, lam, vox, [ny, nx, 4])
    assert rel_l2(hy, hyo) < 1e-13 and rel_l2(hx, hxo) < 1e-13 and abs(p0[0] - p0o) < 1e-13


def test_python_util_matches_oracle_and_factorisation():
    from beyond_dof_b200 import util
    h = util.get_kernel(1.0, 0.248, [1., 1., 1.], [64, 48, 8])
    assert np.array_equal(h, mo.get_kernel(1.0, 0.248, [1., 1., 1.], [64, 48, 8]))
    p0, hy, hx = util.factor_kernel(h)
    assert abs(abs(p0) - 1) < 1e-12 and rel_l2(p0 * np.outer(hy, hx), h) < 1e-12
    u, v = util.gen_mesh([0.5, 0.5], (32, 32))
    nonsep = np.exp(1j * 40 * np.sqrt(1 - 0.06 * (u ** 2 + v ** 2)))
    assert util.factor_kernel(nonsep) is None


def test_bad_arguments_are_rejected(capi):
    h = ctypes.c_void_p()
    assert capi.lib.bdof_plan_create(ctypes.byref(h), 0, 64, 1, 1, 0, None) == -1          # BDOF_E_BADARG
    assert capi.lib.bdof_plan_create(ctypes.byref(h), 74, 64, 1, 1, 0, None) == -2         # BDOF_E_UNSUPPORTED
    assert b'power of two' in capi.lib.bdof_last_error()
    assert capi.lib.bdof_forward(None, None, None, None) == -1


def test_no_cpu_fallback_without_gpu(capi):
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    h = ctypes.c_void_p()
    rc = capi.lib.bdof_plan_create(ctypes.byref(h), 64, 64, 1, 1, 0, None)
    assert rc > 0                                     # a cudaError_t: no device, and no silent CPU path
    import beyond_dof_b200 as bd
    z = np.zeros((1, 64, 64, 2), np.float32)
    with pytest.raises(RuntimeError):
        bd.multislice_propagate_batch_numpy(z, z, np.ones((64, 64)), np.zeros((64, 64)), 5000, 1e-7, obj_batch_shape=z.shape)
