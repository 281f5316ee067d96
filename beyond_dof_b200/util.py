"""Host-side math helpers with the reference's names and conventions (float64, NumPy).

  gen_mesh, get_kernel  <- tensorflow_recon/util.py:156-185, cnn_propagator/util.py:73-102
"""
import numpy as np

# tensorflow_recon/constants.py:90 ; cnn_propagator/util.py:20
PI = 3.14159265359
PI_CNN = 3.1415927


def gen_mesh(max, shape):
    """Endpoint-inclusive frequency mesh (util.py:156-162)."""
    yy = np.linspace(-max[0], max[0], shape[0])
    xx = np.linspace(-max[1], max[1], shape[1])
    return np.meshgrid(xx, yy)


def get_kernel(dist_nm, lmbda_nm, voxel_nm, grid_shape, pi=PI):
    """Centred Fresnel transfer function for the TF algorithm (util.py:165-185), complex128 [ny,nx]."""
    k = 2 * pi / lmbda_nm
    u_max = 1. / (2. * voxel_nm[0])
    v_max = 1. / (2. * voxel_nm[1])
    u, v = gen_mesh([v_max, u_max], grid_shape[0:2])
    return np.exp(1j * k * dist_nm) * np.exp(-1j * pi * lmbda_nm * dist_nm * (u ** 2 + v ** 2))


def kernel_factors(dist_nm, lmbda_nm, voxel_nm, grid_shape, pi=PI):
    """get_kernel == phase0 * outer(hy, hx): the separable factors the pass kernels consume."""
    k = 2 * pi / lmbda_nm
    u_max = 1. / (2. * voxel_nm[0])
    v_max = 1. / (2. * voxel_nm[1])
    yy = np.linspace(-v_max, v_max, grid_shape[0])
    xx = np.linspace(-u_max, u_max, grid_shape[1])
    hy = np.exp(-1j * pi * lmbda_nm * dist_nm * yy ** 2)
    hx = np.exp(-1j * pi * lmbda_nm * dist_nm * xx ** 2)
    return complex(np.exp(1j * k * dist_nm)), hy, hx


def factor_kernel(h, tol=1e-6):
    """Rank-1 factorisation of a caller-supplied centred H (util.py:459-461 accepts any h).
    Returns (phase0, hy, hx) with |phase0| = 1 when ||H - phase0 hy hx^T|| <= tol ||H||, else None."""
    h = np.asarray(h, dtype=np.complex128)
    r, c = h.shape[0] // 2, h.shape[1] // 2
    pivot = h[r, c]
    if abs(pivot) == 0:
        return None
    hx = h[r, :] / pivot
    hy = h[:, c] / pivot
    if np.linalg.norm(h - pivot * np.outer(hy, hx)) > tol * np.linalg.norm(h):
        return None
    # move the modulus of the pivot into hx so that phase0 is a pure phase
    return complex(pivot / abs(pivot)), hy, hx * abs(pivot)
