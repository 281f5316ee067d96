"""Mixed-radix line pass (csrc/genericfft.cuh): the kernel's __host__ __device__ index arithmetic stepped thread by
thread on the CPU (tests/genfft_host.cu, built here with nvcc as host code) against numpy.fft, for every pass variant
of the LineParams contract and the reference's probe sizes (72, 18: tensorflow_recon/reconstruct_ptycho.py)."""
import ctypes
import os
import shutil
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
V = dict(ROW_CONV_T=0, ROW_CONV=1, ROW_CONV_ADJ=2, ROW_FWD=3, ROW_INV=4, COL_CONV=5, COL_FWD=6, COL_INV=7, COL_CONV2D=8)


@pytest.fixture(scope='module')
def host():
    nvcc = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(nvcc):
        pytest.skip('nvcc not available')
    out = os.path.join(HERE, '_build')
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, 'libgenfft_host.so')
    src = os.path.join(HERE, 'genfft_host.cu')
    hdr = os.path.join(HERE, '..', 'beyond_dof_b200', 'csrc', 'genericfft.cuh')
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.run([nvcc, '-O2', '-std=c++17', '--expt-relaxed-constexpr', '-diag-suppress', '20011,20014', '-shared',
                        '-Xcompiler', '-fPIC', '-o', so, src], check=True, capture_output=True)
    lib = ctypes.CDLL(so)
    vp, i32, i64, f32 = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_float
    lib.genfft_host_run.argtypes = [i32, i32, i64, i32, i64, i32, i32, i32, i32, f32, vp, vp, vp, vp, vp, vp]
    lib.genfft_host_factorize.argtypes = [i32, vp]
    return lib


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def run(lib, variant, field, h=None, db=None, psi=None, in_shift=0, out_shift=0, k=0.0):
    """field [B, ny, nx] complex64 -> (out, grad)"""
    B, ny, nx = field.shape
    col = variant >= 5
    n = ny if col else nx
    out = np.zeros_like(field)
    grad = np.zeros((B, ny, nx, 2), np.float32) if db is not None and variant == V['ROW_CONV_ADJ'] else None
    rc = lib.genfft_host_run(n, variant, B * (nx if col else ny), nx if col else ny, ny * nx, nx if col else 1, 1 if col else nx,
                             in_shift, out_shift, k, _p(field), _p(out), _p(h), _p(db), _p(grad), _p(psi))
    assert rc == 0
    return out, grad


def rel(a, b):
    return float(np.linalg.norm((a - b).ravel()) / np.linalg.norm(b.ravel()))


def test_factorisation(host):
    r = (ctypes.c_int * 12)()
    for n, ok in ((72, True), (18, True), (2, True), (2048, True), (1000, True), (63, True), (37, False), (74, False), (4096, False), (1, False)):
        ns = host.genfft_host_factorize(n, r)
        assert (ns > 0) == ok
        if ok:
            assert int(np.prod(r[:ns])) == n


@pytest.mark.parametrize('shape', [(2, 18, 72), (1, 72, 18), (3, 10, 45), (1, 7, 96), (1, 6, 1000), (2, 63, 8)])
def test_every_variant_against_numpy(host, shape):
    rng = np.random.default_rng(5)
    B, ny, nx = shape
    f = (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(np.complex64)
    f64 = f.astype(np.complex128)
    hx = np.exp(1j * rng.standard_normal(nx)).astype(np.complex64) / np.float32(nx)
    hy = np.exp(1j * rng.standard_normal(ny)).astype(np.complex64) / np.float32(ny)
    db = np.stack([rng.random(shape) * 0.3, rng.random(shape) * 0.05], -1).astype(np.float32)
    k = np.float32(1.7)
    t = np.exp(1j * k * db[..., 0].astype(np.float64) - k * db[..., 1].astype(np.float64))
    tol = 3e-6
    # rows
    out, _ = run(host, V['ROW_CONV'], f, h=hx)
    assert rel(out, np.fft.ifft(np.fft.fft(f64, axis=2) * hx, axis=2) * nx) < tol
    out, _ = run(host, V['ROW_CONV_T'], f, h=hx, db=db, k=k)
    assert rel(out, np.fft.ifft(np.fft.fft(f64 * t, axis=2) * hx, axis=2) * nx) < tol
    out, _ = run(host, V['ROW_FWD'], f, out_shift=nx // 2)
    assert rel(out, np.fft.fftshift(np.fft.fft(f64, axis=2), axes=2)) < tol
    out, _ = run(host, V['ROW_INV'], f, in_shift=nx // 2)
    assert rel(out, np.fft.ifft(np.fft.ifftshift(f64, axes=2), axis=2) * nx) < tol
    psi = (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(np.complex64)
    out, grad = run(host, V['ROW_CONV_ADJ'], f, h=hx, db=db, psi=psi, k=k)
    gu = np.fft.ifft(np.fft.fft(f64, axis=2) * hx, axis=2) * nx
    w = np.conj(gu) * (psi.astype(np.complex128) * t)
    assert rel(out, gu * np.conj(t)) < tol
    assert rel(grad[..., 0], -k * w.imag) < tol and rel(grad[..., 1], -k * w.real) < tol
    # columns
    out, _ = run(host, V['COL_CONV'], f, h=hy)
    assert rel(out, np.fft.ifft(np.fft.fft(f64, axis=1) * hy[None, :, None], axis=1) * ny) < tol
    out, _ = run(host, V['COL_FWD'], f, out_shift=ny // 2)
    assert rel(out, np.fft.fftshift(np.fft.fft(f64, axis=1), axes=1)) < tol
    out, _ = run(host, V['COL_INV'], f, in_shift=ny // 2)
    assert rel(out, np.fft.ifft(np.fft.ifftshift(f64, axes=1), axis=1) * ny) < tol
    H2 = (np.exp(1j * rng.standard_normal((ny, nx))) / ny).astype(np.complex64)
    out, _ = run(host, V['COL_CONV2D'], f, h=H2)
    assert rel(out, np.fft.ifft(np.fft.fft(f64, axis=1) * H2[None], axis=1) * ny) < tol
