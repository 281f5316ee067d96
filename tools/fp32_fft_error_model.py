"""Where the depth-dependent error of a complex64 multislice chain comes from ("fp32 error budget").

CPU-only experiment on the strong-object case of tests/test_gpu_depth.py (128 x 256 field, 512 slices, delta <= 1e-3):
  (1) the chain in complex64 with pocketfft's single-precision FFT (any fp32 FFT library is in this class),
  (2) fp32 STORAGE between slices but exact (float64) FFTs,
  (3) float64 ARITHMETIC throughout, but a two-stage Stockham FFT whose constants (stage twiddles and the small-DFT
      matrices of the in-register butterflies) are rounded to fp32: the deterministic operator error alone,
  (4) the same with the per-frequency diagonal gain of the realised FFT pair divided out of the multiplier table.
Printed: relative L2 error of the exit intensity / field against the complex128 oracle.

    python tools/fp32_fft_error_model.py            # about 2 minutes on one core
"""
import os
import sys

import numpy as np
import scipy.fft as sf

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import multislice_oracle as mo                     # noqa: E402  (analysis tool, not product code)
from beyond_dof_b200.util import kernel_factors                # noqa: E402


def rel(a, b):
    return np.linalg.norm((a - b).ravel()) / np.linalg.norm(b.ravel())


def r32(z):
    return z.astype(np.complex64).astype(np.complex128)


def fft_matrix(N, R1, R2, round_tw=True, round_small=True):
    """two-stage transform n = R2 n1 + n2, k = k1 + R1 k2 as an explicit matrix"""
    def small(R):
        n = np.arange(R)
        W = np.exp(-2j * np.pi * np.outer(n, n) / R)
        return r32(W) if round_small else W
    A, Bm = small(R1), small(R2)
    n2, k1 = np.arange(R2), np.arange(R1)
    T = np.exp(-2j * np.pi * np.outer(n2, k1) / N)
    if round_tw:
        T = r32(T)
    F = np.zeros((N, N), complex)
    for a in range(R1):
        for b in range(R2):
            for n1 in range(R1):
                F[a + R1 * b, R2 * n1 + n2] = Bm[b, n2] * T[n2, a] * A[a, n1]
    return F


def conv_operator(N, R1, R2, h, correct_diagonal=False, **kw):
    F = fft_matrix(N, R1, R2, **kw)
    Finv = np.conj(F) / N                                       # the kernels' conj trick
    if correct_diagonal:
        Fex = np.exp(-2j * np.pi * np.outer(np.arange(N), np.arange(N)) / N)
        gain = np.array([(Fex[k] @ Finv[:, k]) * (F[k] @ np.conj(Fex[k]) / N) for k in range(N)])
        h = h / gain
    return Finv @ np.diag(h) @ F


def main():
    shape = (1, 128, 256, 512)
    B, Y, X, Z = shape
    gd, gb = mo.random_phantom(shape, seed=99, delta_scale=1e-3, beta_scale=2e-5)
    pr, pi = mo.gaussian_probe(shape[1:3], 40., 30., 0.5)
    ref = mo.multislice_forward(gd, gb, pr, pi, 5000, 1e-7, propagate_last=True)[0]
    p0, hy, hx = kernel_factors(1.0, 0.248, [1, 1, 1], [Y, X, Z])
    hxs, hys = np.fft.ifftshift(hx), np.fft.ifftshift(hy)
    H = np.outer(hys, hxs)
    k = 2 * mo.PI_TF * 1.0 / 0.248
    probe = (np.zeros((Y, X), np.complex64) + (pr + 1j * pi))

    def t_of(i):
        return np.exp(1j * k * gd[0, ..., i].astype(np.float64)) * np.exp(-k * gb[0, ..., i].astype(np.float64))

    def report(name, psi):
        psi = psi.astype(np.complex128) * p0 ** Z
        print('%-58s intensity %.2e  field %.2e' % (name, rel(np.abs(psi) ** 2, np.abs(ref) ** 2), rel(psi, ref)), flush=True)

    psi = probe.astype(np.complex64)
    H32 = H.astype(np.complex64)
    for i in range(Z):
        psi = sf.ifft2(sf.fft2(psi * t_of(i).astype(np.complex64)) * H32)
    report('(1) complex64 chain, pocketfft single precision', psi)
    psi = probe.astype(np.complex64)
    for i in range(Z):
        psi = np.fft.ifft2(np.fft.fft2((psi * t_of(i).astype(np.complex64)).astype(np.complex128)) * H32).astype(np.complex64)
    report('(2) fp32 storage, exact FFT', psi)
    for name, kw in (('(3) float64 arithmetic, fp32-rounded FFT constants', {}),
                     ('(3b) ... stage twiddles only', {'round_small': False}),
                     ('(4) (3) + diagonal gain divided out of h', {'correct_diagonal': True})):
        Cx, Cy = conv_operator(X, 16, 16, hxs, **kw), conv_operator(Y, 16, 8, hys, **kw)
        psi = probe.astype(np.complex128)
        for i in range(Z):
            psi = Cy @ (psi * t_of(i)) @ Cx.T
        report(name, psi)


if __name__ == '__main__':
    main()
