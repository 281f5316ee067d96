"""CPU tier: the host-side model of the realised fp32 transform pair (bdof_debug_fft_gain, csrc/bdof.cu: conv_diag_gain).

The multiplier tables are divided by the per-bin gain g_k of one realised convolution IFFT(h FFT(x)); g_k comes from a
closed form over the small in-register DFTs and the twiddle table.  Here the same quantity is derived independently: the
FULL realised transform matrix is assembled in Python from the butterflies' recursive structure (regfft.cuh) with
fp32-rounded constants, and g_k is read off by projecting the realised forward row / inverse column of bin k on the exact
ones.  (That the structure matches what the kernels execute is shown on the GPU: test_gpu_depth.py, config-3 shard.)"""
import ctypes

import numpy as np
import pytest


def _tw32(M, R):
    M %= R
    if (4 * M) % R == 0:
        return [1, -1j, -1, 1j][(4 * M) // R]
    a = 2 * np.pi * M / R
    return complex(np.float32(np.cos(a)), -np.float32(np.sin(a)))


def _regfft(v):
    R = len(v)
    if R == 1:
        return v
    if R == 2:
        return [v[0] + v[1], v[0] - v[1]]
    if R == 4:
        t0, t1, t2, t3 = v[0] + v[2], v[0] - v[2], v[1] + v[3], (v[1] - v[3]) * (-1j)
        return [t0 + t2, t1 + t3, t0 - t2, t1 - t3]
    A = 2 if R == 8 else 4
    B = R // A
    v = list(v)
    for n2 in range(B):
        t = _regfft([v[B * n1 + n2] for n1 in range(A)])
        for k1 in range(A):
            v[B * k1 + n2] = t[k1] * _tw32((n2 * k1) % R, R)
    out = [0] * R
    for k1 in range(A):
        u = _regfft([v[B * k1 + n2] for n2 in range(B)])
        for k2 in range(B):
            out[k1 + A * k2] = u[k2]
    return out


def _small(R):
    M = np.zeros((R, R), complex)
    for r in range(R):
        e = [0j] * R
        e[r] = 1
        M[:, r] = _regfft(e)
    return M


def _realised_matrix(N, R1, R2):
    A1, A2 = _small(R1), _small(R2)
    F = np.zeros((N, N), complex)
    for k1 in range(R1):
        for j in range(R2):
            a = -2 * np.pi * ((j * k1) % N) / N
            W = 1.0 if j == 0 else complex(np.float32(np.cos(a)), np.float32(np.sin(a)))
            for k2 in range(R2):
                F[k1 + R1 * k2, j + R2 * np.arange(R1)] = A2[k2, j] * W * A1[k1, :]
    return F


@pytest.mark.parametrize('cfg', [(64, 8, 8), (128, 16, 8), (256, 16, 16)])
def test_gain_closed_form_matches_full_matrix(cfg):
    from beyond_dof_b200 import capi
    N, R1, R2 = cfg
    F = _realised_matrix(N, R1, R2)
    Fex = np.exp(-2j * np.pi * np.outer(np.arange(N), np.arange(N)) / N)
    assert 1e-9 < np.abs(F - Fex).max() < 1e-5                       # a DFT up to fp32 rounding of its constants
    Finv = np.conj(F) / N                                            # the kernels' conj trick
    gain = np.array([(Fex[k] @ Finv[:, k]) * (F[k] @ np.conj(Fex[k]) / N) for k in range(N)])
    out = np.zeros(2 * N)
    assert capi.lib.bdof_debug_fft_gain(N, out.ctypes.data_as(ctypes.c_void_p)) == 0
    g = out[0::2] + 1j * out[1::2]
    assert 1e-9 < np.abs(g - 1).max() < 2e-7
    assert np.abs(g - gain).max() < 1e-13


def test_gain_is_one_where_the_model_does_not_apply():
    from beyond_dof_b200 import capi
    for n in (4096, 8192, 74):
        out = np.zeros(2 * n)
        assert capi.lib.bdof_debug_fft_gain(n, out.ctypes.data_as(ctypes.c_void_p)) == 0
        assert np.all(out[0::2] == 1.0) and np.all(out[1::2] == 0.0)


def _stockham_matrix(n):
    """The mixed-radix passes' transform (genericfft.cuh: radices 4.., 2, 3, 5, 7, autosort, full fp32 table W_n^q) applied
    to the identity in complex128 -- written from the algorithm's definition, vectorised over the work items."""
    radix, m = [], n
    while m % 4 == 0:
        radix.append(4); m //= 4
    for p in (2, 3, 5, 7):
        while m % p == 0:
            radix.append(p); m //= p
    assert m == 1
    a = -2 * np.pi * np.arange(n) / n
    tw = np.cos(a).astype(np.float32).astype(np.float64) + 1j * np.sin(a).astype(np.float32).astype(np.float64)
    x = np.eye(n, dtype=np.complex128)                  # x[element, basis vector]
    ns = 1
    for R in radix:
        mm = n // R
        j = np.arange(mm)
        k = j % ns
        o = (j // ns) * ns * R + k
        v = [x[j + r * mm] * np.where(k > 0, tw[(r * k * (n // (ns * R))) % n], 1.0)[:, None] if r else x[j] for r in range(R)]
        y = np.zeros_like(x)
        if R == 2:
            y[o], y[o + ns] = v[0] + v[1], v[0] - v[1]
        elif R == 4:
            s0, d0, s1, d1 = v[0] + v[2], v[0] - v[2], v[1] + v[3], v[1] - v[3]
            y[o], y[o + ns], y[o + 2 * ns], y[o + 3 * ns] = s0 + s1, d0 - 1j * d1, s0 - s1, d0 + 1j * d1
        else:
            for q in range(R):
                acc = v[0].copy()
                for r in range(1, R):
                    acc += v[r] * tw[((r * q) % R) * (n // R)]
                y[o + q * ns] = acc
        x = y
        ns *= R
    return x


@pytest.mark.parametrize('n', [72, 18, 100, 63, 1000])
def test_mixed_radix_gain_matches_full_matrix(n):
    from beyond_dof_b200 import capi
    F = _stockham_matrix(n)
    Fex = np.exp(-2j * np.pi * np.outer(np.arange(n), np.arange(n)) / n)
    assert np.abs(F - Fex).max() < 1e-5
    Finv = np.conj(F) / n
    gain = np.array([(Fex[k] @ Finv[:, k]) * (F[k] @ np.conj(Fex[k]) / n) for k in range(n)])
    out = np.zeros(2 * n)
    assert capi.lib.bdof_debug_fft_gain(n, out.ctypes.data_as(ctypes.c_void_p)) == 0
    g = out[0::2] + 1j * out[1::2]
    assert 1e-9 < np.abs(g - 1).max() < 3e-7
    assert np.abs(g - gain).max() < 1e-12
