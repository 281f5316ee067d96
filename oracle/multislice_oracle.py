"""CPU oracle for the Fresnel multislice hot path -- TEST INFRASTRUCTURE ONLY.

This file is a complex128 NumPy restatement of the reference algorithm
(mdw771/beyond_dof).  It exists to check the CUDA path; it is never the thing
shipped or measured.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.

Parity pinning: the reference repository has no tests and no golden vectors
(SURVEY.md section 4).  The restatement is pinned instead against outputs of
the reference's own functions executed unmodified in the build container
(``oracle/gen_golden.py`` -> ``tests/golden/*.npz``) for the forward paths
(a-1 NumPy multislice, a-2 get_kernel, a-5 real-space "cnn" multislice).
The TF-graph variant (a-3) and every gradient (a-8) cannot be executed here
(TensorFlow 1.x / HIPS autograd are absent): for those the restatement is
pinned against torch.autograd (complex128, CPU) and finite differences, i.e.
"parity unpinned by the reference itself" for a-3 / a-8.

Reference citations (paths relative to the reference checkout):
  tensorflow_recon/npfuncs.py:16-63      multislice_propagate_batch_numpy
  tensorflow_recon/util.py:156-185       gen_mesh, get_kernel
  tensorflow_recon/util.py:432-508       multislice_propagate_batch (TF graph)
  cnn_propagator/propagation.py:18-133   multislice_propagate_cnn
  cnn_propagator/fullfield.py:93-121     calculate_loss (full field)
  tensorflow_recon/ptychography.py:37-97 rotate_and_project (ptychography)
"""
import numpy as np
from numpy.fft import fft2, ifft2, fftshift, ifftshift

# tensorflow_recon/constants.py:90 and cnn_propagator/util.py:20
PI_TF = 3.14159265359
PI_CNN = 3.1415927


def gen_mesh(max_, shape):
    """tensorflow_recon/util.py:156-162 -- endpoint-inclusive frequency grid."""
    yy = np.linspace(-max_[0], max_[0], shape[0])
    xx = np.linspace(-max_[1], max_[1], shape[1])
    return np.meshgrid(xx, yy)


def get_kernel(dist_nm, lmbda_nm, voxel_nm, grid_shape, pi=PI_TF):
    """Centred Fresnel transfer function H (tensorflow_recon/util.py:165-185).

    H[v, u] = exp(i k d) exp(-i pi lambda d (u^2 + v^2)); u runs along axis 1
    over linspace(+-1/(2 voxel[0])), v along axis 0 over +-1/(2 voxel[1]).
    """
    k = 2 * pi / lmbda_nm
    u_max = 1. / (2. * voxel_nm[0])
    v_max = 1. / (2. * voxel_nm[1])
    u, v = gen_mesh([v_max, u_max], grid_shape[0:2])
    return np.exp(1j * k * dist_nm) * np.exp(-1j * pi * lmbda_nm * dist_nm * (u ** 2 + v ** 2))


def kernel_factors(dist_nm, lmbda_nm, voxel_nm, grid_shape, pi=PI_TF):
    """Separable form of get_kernel: H = phase0 * outer(hy, hx).

    Returns (phase0, hy[ny], hx[nx]) in complex128, all centred (DC in the middle).
    """
    k = 2 * pi / lmbda_nm
    u_max = 1. / (2. * voxel_nm[0])
    v_max = 1. / (2. * voxel_nm[1])
    yy = np.linspace(-v_max, v_max, grid_shape[0])
    xx = np.linspace(-u_max, u_max, grid_shape[1])
    hy = np.exp(-1j * pi * lmbda_nm * dist_nm * yy ** 2)
    hx = np.exp(-1j * pi * lmbda_nm * dist_nm * xx ** 2)
    return np.exp(1j * k * dist_nm), hy, hx


def _propagate(wavefront, h):
    """One transfer-function step exactly as written in the reference:
    ifft2(ifftshift(fftshift(fft2(psi)) * H))   (npfuncs.py:41)."""
    return ifft2(ifftshift(fftshift(fft2(wavefront), axes=[1, 2]) * h, axes=[1, 2]))


def _free_prop(wavefront, free_prop_cm, lmbda_nm, voxel_nm, grid_shape, pi):
    """Free-space step after the object (npfuncs.py:43-61, util.py:490-507).
    'inf' -> far field; float -> one more TF-kernel step (the TF/IR switch is
    forced to 'TF' in the batch functions)."""
    if free_prop_cm is None:
        return wavefront
    if isinstance(free_prop_cm, str):
        if free_prop_cm != 'inf':
            raise ValueError('free_prop_cm must be None, "inf" or a float')
        return fftshift(fft2(wavefront), axes=[1, 2])
    dist_nm = free_prop_cm * 1e7
    h = get_kernel(dist_nm, lmbda_nm, voxel_nm, grid_shape, pi=pi)
    return _propagate(wavefront, h)


def multislice_forward(grid_delta_batch, grid_beta_batch, probe_real, probe_imag, energy_ev,
                       psize_cm, free_prop_cm=None, obj_batch_shape=None, h=None,
                       propagate_last=False, pi=PI_TF, return_slices=False):
    """Forward multislice, [B,Y,X,Z] inputs, complex128 arithmetic.

    propagate_last=False : NumPy semantics (npfuncs.py:35-41, last slice only modulates)
    propagate_last=True  : TF semantics (util.py:464-488, every slice propagates,
                           except the n_slice == 1 special case which only modulates)
    return_slices: also return psi entering every slice (for the adjoint).
    """
    # float64 always: a float32 array times the Python complex 1j*k would stay complex64 in NumPy 2
    grid_delta_batch = np.asarray(grid_delta_batch, dtype=np.float64)
    grid_beta_batch = np.asarray(grid_beta_batch, dtype=np.float64)
    if obj_batch_shape is None:
        obj_batch_shape = grid_delta_batch.shape
    batch = obj_batch_shape[0]
    grid_shape = list(obj_batch_shape[1:])
    voxel_nm = np.array([psize_cm] * 3) * 1.e7
    # npfuncs.py:21-22: the probe is accumulated into a complex64 array (so it is rounded to
    # fp32 once); the first multiply with the float64-derived c promotes everything to complex128.
    wavefront = np.zeros([batch, obj_batch_shape[1], obj_batch_shape[2]], dtype=np.complex64)
    wavefront += (np.asarray(probe_real) + 1j * np.asarray(probe_imag))
    wavefront = wavefront.astype(np.complex128)
    lmbda_nm = 1240. / energy_ev
    n_slice = obj_batch_shape[-1]
    delta_nm = voxel_nm[-1]
    if h is None:
        h = get_kernel(delta_nm, lmbda_nm, voxel_nm, grid_shape, pi=pi)
    k = 2. * pi * delta_nm / lmbda_nm
    slices = []
    for i in range(n_slice):
        if return_slices:
            slices.append(wavefront)
        c = np.exp(1j * k * grid_delta_batch[:, :, :, i]) * np.exp(-k * grid_beta_batch[:, :, :, i])
        wavefront = wavefront * c
        if propagate_last:
            do_prop = n_slice > 1
        else:
            do_prop = i < n_slice - 1
        if do_prop:
            wavefront = _propagate(wavefront, h)
    wavefront = _free_prop(wavefront, free_prop_cm, lmbda_nm, voxel_nm, grid_shape, pi)
    if return_slices:
        return wavefront, slices
    return wavefront


def multislice_propagate_batch_numpy(grid_delta_batch, grid_beta_batch, probe_real, probe_imag,
                                     energy_ev, psize_cm, free_prop_cm=None, obj_batch_shape=None):
    """Restatement of tensorflow_recon/npfuncs.py:16-63 (NumPy semantics)."""
    return multislice_forward(grid_delta_batch, grid_beta_batch, probe_real, probe_imag, energy_ev,
                              psize_cm, free_prop_cm, obj_batch_shape, propagate_last=False, pi=PI_TF)


def multislice_propagate_batch_numpy_cnn(grid_delta_batch, grid_beta_batch, probe_real, probe_imag, energy_ev, psize_cm,
                                         free_prop_cm=None, obj_batch_shape=None):
    """Restatement of cnn_propagator/np_funcs.py:15-65: the same loop with PI = 3.1415927 (np_funcs.py:12), returning
    (wavefront, probe_array) where probe_array[i] is the field after slice i (np_funcs.py:41), before the free-space step."""
    gd = np.asarray(grid_delta_batch, dtype=np.float64)
    gb = np.asarray(grid_beta_batch, dtype=np.float64)
    if obj_batch_shape is None:
        obj_batch_shape = gd.shape
    grid_shape = list(obj_batch_shape[1:])
    voxel_nm = np.array([psize_cm] * 3) * 1.e7
    wavefront = np.zeros([obj_batch_shape[0], obj_batch_shape[1], obj_batch_shape[2]], dtype=np.complex64)
    wavefront += (np.asarray(probe_real) + 1j * np.asarray(probe_imag))
    wavefront = wavefront.astype(np.complex128)
    lmbda_nm = 1240. / energy_ev
    n_slice = obj_batch_shape[-1]
    h = get_kernel(voxel_nm[-1], lmbda_nm, voxel_nm, grid_shape, pi=PI_CNN)
    k = 2. * PI_CNN * voxel_nm[-1] / lmbda_nm
    probe_array = []
    for i in range(n_slice):
        wavefront = wavefront * (np.exp(1j * k * gd[:, :, :, i]) * np.exp(-k * gb[:, :, :, i]))
        if i < n_slice - 1:
            wavefront = _propagate(wavefront, h)
        probe_array.append(wavefront)
    wavefront = _free_prop(wavefront, free_prop_cm, lmbda_nm, voxel_nm, grid_shape, PI_CNN)
    return wavefront, np.array(probe_array)


def multislice_propagate_batch(grid_delta_batch, grid_beta_batch, probe_real, probe_imag, energy_ev,
                               psize_cm, h=None, free_prop_cm=None, obj_batch_shape=None):
    """Restatement of tensorflow_recon/util.py:432-508, type='plane' (TF semantics),
    evaluated in complex128 instead of TF's complex64."""
    return multislice_forward(grid_delta_batch, grid_beta_batch, probe_real, probe_imag, energy_ev,
                              psize_cm, free_prop_cm, obj_batch_shape, h=h, propagate_last=True, pi=PI_TF)


def get_kernel_ir(dist_nm, lmbda_nm, voxel_nm, grid_shape, pi=PI_TF):
    """tensorflow_recon/util.py:188-216, NumPy branch: H = fftshift(fft2(h)) dx dy of the impulse response h."""
    size_nm = np.array(voxel_nm) * np.array(grid_shape)
    k = 2 * pi / lmbda_nm
    ymin, xmin = np.array(size_nm)[:2] / -2.
    dy, dx = voxel_nm[0:2]
    x = np.arange(xmin, xmin + size_nm[1], dx)
    y = np.arange(ymin, ymin + size_nm[0], dy)
    x, y = np.meshgrid(x, y)
    h = np.exp(1j * k * dist_nm) / (1j * lmbda_nm * dist_nm) * np.exp(1j * k / (2 * dist_nm) * (x ** 2 + y ** 2))
    return fftshift(fft2(h)) * voxel_nm[0] * voxel_nm[1]


def multislice_propagate_unbatched(grid_delta, grid_beta, probe_real, probe_imag, energy_ev, psize_cm, h=None,
                                   free_prop_cm=None, pad=None, pi=PI_TF):
    """Restatement of tensorflow_recon/util.py:360-429 (un-batched [Y,X,Z] TF function) in complex128, line by line:
    every slice propagates; the TF / IR orderings are chosen by the sampling criterion per slice (:396-404) and for the
    free-space step (:416-427), where the 'IR' branch applies fft2 twice as written in the reference."""
    grid_delta = np.asarray(grid_delta, dtype=np.float64)
    grid_beta = np.asarray(grid_beta, dtype=np.float64)
    if pad is not None:
        grid_delta = np.pad(grid_delta, pad, 'constant')
        grid_beta = np.pad(grid_beta, pad, 'constant')
    voxel_nm = np.array([psize_cm] * 3) * 1.e7
    wavefront = np.zeros([grid_delta.shape[0], grid_delta.shape[1]], dtype=np.complex64)
    wavefront = (wavefront + (np.asarray(probe_real) + 1j * np.asarray(probe_imag)).astype(np.complex64)).astype(np.complex128)
    lmbda_nm = 1240. / energy_ev
    mean_voxel_nm = np.prod(voxel_nm) ** (1. / 3)
    size_nm = np.array(grid_delta.shape) * voxel_nm
    n_slice = grid_delta.shape[-1]
    delta_nm = voxel_nm[-1]
    if h is None:
        h = get_kernel(delta_nm, lmbda_nm, voxel_nm, grid_delta.shape, pi=pi)
    k = 2. * pi * delta_nm / lmbda_nm
    sh = lambda a: fftshift(a, axes=[-2, -1])
    ish = lambda a: ifftshift(a, axes=[-2, -1])
    for i in range(n_slice):
        c = np.exp(1j * k * grid_delta[:, :, i]) * np.exp(-k * grid_beta[:, :, i])
        wavefront = wavefront * c
        l = np.prod(size_nm) ** (1. / 3)
        crit_samp = lmbda_nm * delta_nm / l
        if mean_voxel_nm > crit_samp:
            wavefront = ifft2(ish(sh(fft2(wavefront)) * h))
        else:
            wavefront = fft2(sh(wavefront))
            wavefront = ish(ifft2(wavefront * h))
    if free_prop_cm is not None:
        if free_prop_cm == 'inf':
            wavefront = sh(fft2(wavefront))
        else:
            dist_nm = free_prop_cm * 1e7
            l = np.prod(size_nm) ** (1. / 3)
            crit_samp = lmbda_nm * dist_nm / l
            if mean_voxel_nm > crit_samp:
                hf = get_kernel(dist_nm, lmbda_nm, voxel_nm, list(grid_delta.shape), pi=pi)
                wavefront = ifft2(ish(sh(fft2(wavefront)) * hf))
            else:
                hf = get_kernel_ir(dist_nm, lmbda_nm, voxel_nm, list(grid_delta.shape), pi=pi)
                wavefront = sh(fft2(wavefront)) * hf
                wavefront = ish(fft2(wavefront))
    return wavefront


# ---------------------------------------------------------------------------------------------
# loss head and hand adjoint.  No adjoint exists in the reference: it
# differentiates with TF / autograd.  Pinned against torch.autograd in tests/test_oracle.py.
# ---------------------------------------------------------------------------------------------

def loss_mag(wavefront, target_mag):
    """loss = mean((|psi| - |y|)^2)  (fullfield.py:115, cnn fullfield.py:106, ptychography.py:79).
    Returns (loss, G) with G = dL/dRe + i dL/dIm."""
    mag = np.abs(wavefront)
    diff = mag - target_mag
    loss = np.mean(diff ** 2)
    with np.errstate(invalid='ignore', divide='ignore'):
        g = (2.0 / diff.size) * diff * np.where(mag > 0, wavefront / mag, 0)
    return loss, g


def multislice_adjoint(grid_delta_batch, grid_beta_batch, slices, grad_exit, energy_ev, psize_cm,
                       free_prop_cm=None, h=None, propagate_last=False, pi=PI_TF):
    """Back-propagate G = dL/dRe(psi_out) + i dL/dIm(psi_out) through the chain.

    Returns (grad_delta, grad_beta, grad_probe) with grad_* in [B,Y,X,Z] and
    grad_probe = G at the entrance plane summed over the batch (the probe is shared).
    """
    grid_delta_batch = np.asarray(grid_delta_batch, dtype=np.float64)
    grid_beta_batch = np.asarray(grid_beta_batch, dtype=np.float64)
    batch, ny, nx, n_slice = grid_delta_batch.shape
    voxel_nm = np.array([psize_cm] * 3) * 1.e7
    lmbda_nm = 1240. / energy_ev
    delta_nm = voxel_nm[-1]
    grid_shape = [ny, nx, n_slice]
    if h is None:
        h = get_kernel(delta_nm, lmbda_nm, voxel_nm, grid_shape, pi=pi)
    k = 2. * pi * delta_nm / lmbda_nm
    g = np.asarray(grad_exit, dtype=np.complex128)

    def adj_propagate(g, hh):
        # adjoint of ifft2(ifftshift(fftshift(fft2(.)) * H)) = same chain with conj(H)
        return ifft2(ifftshift(fftshift(fft2(g), axes=[1, 2]) * np.conj(hh), axes=[1, 2]))

    if free_prop_cm is not None:
        if isinstance(free_prop_cm, str):
            # psi_out = fftshift(fft2 psi): adjoint = N * ifft2(ifftshift(G))
            g = ifft2(ifftshift(g, axes=[1, 2])) * (ny * nx)
        else:
            hf = get_kernel(free_prop_cm * 1e7, lmbda_nm, voxel_nm, grid_shape, pi=pi)
            g = adj_propagate(g, hf)
    gd = np.zeros(grid_delta_batch.shape, dtype=np.float64)
    gb = np.zeros(grid_delta_batch.shape, dtype=np.float64)
    for i in range(n_slice - 1, -1, -1):
        if propagate_last:
            do_prop = n_slice > 1
        else:
            do_prop = i < n_slice - 1
        if do_prop:
            g = adj_propagate(g, h)
        t = np.exp(1j * k * grid_delta_batch[:, :, :, i]) * np.exp(-k * grid_beta_batch[:, :, :, i])
        u = slices[i] * t
        w = np.conj(g) * u
        gd[:, :, :, i] = -k * w.imag
        gb[:, :, :, i] = -k * w.real
        g = np.conj(t) * g
    return gd, gb, g.sum(axis=0)


def loss_and_grad(grid_delta_batch, grid_beta_batch, probe_real, probe_imag, energy_ev, psize_cm,
                  target_mag, free_prop_cm=None, h=None, propagate_last=False, pi=PI_TF):
    """loss = mean((|psi_out| - target)^2) and its gradient w.r.t. delta, beta."""
    psi, slices = multislice_forward(grid_delta_batch, grid_beta_batch, probe_real, probe_imag,
                                     energy_ev, psize_cm, free_prop_cm, None, h=h,
                                     propagate_last=propagate_last, pi=pi, return_slices=True)
    loss, g = loss_mag(psi, target_mag)
    gd, gb, gp = multislice_adjoint(grid_delta_batch, grid_beta_batch, slices, g, energy_ev, psize_cm,
                                    free_prop_cm, h=h, propagate_last=propagate_last, pi=pi)
    return loss, gd, gb, psi


# ---------------------------------------------------------------------------------------------
# real-space ("cnn") multislice, cnn_propagator/propagation.py:18-133
# ---------------------------------------------------------------------------------------------

def cnn_kernel(energy_ev, psize_cm, grid_shape_yxz, kernel_size):
    """Truncated real-space propagator (propagation.py:35-44): IFFT of H built on the
    (grid_shape - 1) grid, centred, cropped to kernel_size^2."""
    lmbda_nm = 1240. / energy_ev
    voxel_nm = np.array(psize_cm) * 1.e7
    delta_nm = voxel_nm[-1]
    gs = np.array(grid_shape_yxz) - 1
    kern = get_kernel(delta_nm, lmbda_nm, voxel_nm, gs, pi=PI_CNN)
    kern = fftshift(ifft2(ifftshift(kern)))
    mid = ((np.array(kern.shape) - 1) / 2).astype('int')
    half = int((kernel_size - 1) / 2)
    return kern[mid[0] - half:mid[0] + half + 1, mid[1] - half:mid[1] + half + 1]


def multislice_propagate_cnn(grid_delta, grid_beta, probe_real, probe_imag, energy_ev, psize_cm,
                             kernel_size=17, free_prop_cm=None):
    """Restatement of cnn_propagator/propagation.py:18-133 (FFT-free direct convolution is
    replaced by an equivalent explicit shift-and-add; arithmetic order differs only in the
    summation order of the kernel taps)."""
    assert kernel_size % 2 == 1, 'kernel_size must be an odd number.'
    grid_delta = np.asarray(grid_delta, dtype=np.float64)
    grid_beta = np.asarray(grid_beta, dtype=np.float64)
    n_batch, shape_y, shape_x, n_slice = grid_delta.shape
    lmbda_nm = 1240. / energy_ev
    voxel_nm = np.array(psize_cm) * 1.e7
    delta_nm = voxel_nm[-1]
    k = 2. * np.pi * delta_nm / lmbda_nm        # propagation.py:25 uses full-precision pi here
    grid_shape = np.array(grid_delta.shape[1:])
    kern = cnn_kernel(energy_ev, psize_cm, grid_shape, kernel_size)
    pad = (kernel_size - 1) // 2
    probe = np.tile(np.asarray(probe_real) + 1j * np.asarray(probe_imag), [n_batch, 1, 1]).astype(np.complex128)
    edge_val = 1.0
    initial = probe[0, 0, 0]
    ksum = kern.sum()
    for i in range(n_slice):
        c = np.exp(1j * k * grid_delta[:, :, :, i] - k * grid_beta[:, :, :, i])
        probe = probe * c
        padded = np.pad(probe, [[0, 0], [pad, pad], [pad, pad]], mode='constant', constant_values=edge_val)
        out = np.zeros_like(probe)
        # true convolution, 'valid': out[y,x] = sum_{a,b} kern[a,b] * padded[y + 2p - a, x + 2p - b]
        for a in range(kernel_size):
            for b in range(kernel_size):
                out += kern[a, b] * padded[:, 2 * pad - a:2 * pad - a + shape_y, 2 * pad - b:2 * pad - b + shape_x]
        probe = out
        edge_val = ksum * edge_val
    final = probe[0, 0, 0]
    probe = probe * (initial / final)
    if free_prop_cm is not None:
        if isinstance(free_prop_cm, str):
            probe = fftshift(fft2(probe), axes=[1, 2])
        else:
            hf = get_kernel(free_prop_cm * 1e7, lmbda_nm, voxel_nm, grid_shape, pi=PI_CNN)
            probe = _propagate(probe, hf)
    return probe


def cnn_loss_and_grad(grid_delta, grid_beta, probe_real, probe_imag, energy_ev, psize_cm, target_mag, kernel_size=17,
                      free_prop_cm=None):
    """loss = mean((|psi| - target)^2) through multislice_propagate_cnn and its gradient w.r.t. delta, beta [B,Y,X,Z] -- what
    autograd.grad(calculate_loss) differentiates in cnn_propagator/fullfield.py:93-121,329 and ptychography.py:30-81,248.
    Hand adjoint: per slice psi' = conv_valid(pad(psi t, edge), K) is linear in u = psi t with a constant edge, so
    G_u[y,x] = sum_ab conj(K[a,b]) G'[y - p + a, x - p + b]; the final rescaling by psi_0[0,0,0] / psi_Z[0,0,0] (batch element
    0's corner pixel, propagation.py:79,109-110) couples every batch element to that one pixel.  Pinned against
    torch.autograd in tests/test_oracle.py (no reference gradient exists to run: HIPS autograd is absent)."""
    gd = np.asarray(grid_delta, dtype=np.float64)
    gb = np.asarray(grid_beta, dtype=np.float64)
    B, Y, X, Z = gd.shape
    lmbda_nm = 1240. / energy_ev
    voxel_nm = np.array(psize_cm) * 1.e7
    k = 2. * np.pi * voxel_nm[-1] / lmbda_nm
    kern = cnn_kernel(energy_ev, psize_cm, np.array([Y, X, Z]), kernel_size)
    pad = (kernel_size - 1) // 2
    probe = np.tile(np.asarray(probe_real) + 1j * np.asarray(probe_imag), [B, 1, 1]).astype(np.complex128)
    initial = probe[0, 0, 0]
    edge_val = 1.0 + 0j
    ksum = kern.sum()
    slices = []
    for i in range(Z):
        slices.append(probe)
        u = probe * np.exp(1j * k * gd[..., i] - k * gb[..., i])
        padded = np.pad(u, [[0, 0], [pad, pad], [pad, pad]], mode='constant', constant_values=edge_val)
        out = np.zeros_like(u)
        for a in range(kernel_size):
            for b in range(kernel_size):
                out += kern[a, b] * padded[:, 2 * pad - a:2 * pad - a + Y, 2 * pad - b:2 * pad - b + X]
        probe = out
        edge_val = ksum * edge_val
    f = probe[0, 0, 0]
    psi = probe * (initial / f)
    hf = None
    if free_prop_cm is not None:
        if isinstance(free_prop_cm, str):
            exit_wave = fftshift(fft2(psi), axes=[1, 2])
        else:
            hf = get_kernel(free_prop_cm * 1e7, lmbda_nm, voxel_nm, np.array([Y, X, Z]), pi=PI_CNN)
            exit_wave = _propagate(psi, hf)
    else:
        exit_wave = psi
    loss, g = loss_mag(exit_wave, target_mag)
    if free_prop_cm is not None:
        if isinstance(free_prop_cm, str):
            g = ifft2(ifftshift(g, axes=[1, 2])) * (Y * X)
        else:
            g = ifft2(ifftshift(fftshift(fft2(g), axes=[1, 2]) * np.conj(hf), axes=[1, 2]))
    # rescaling psi = s psi_Z, s = initial / f, f = psi_Z[0,0,0]
    s_ = initial / f
    g_f = np.sum(np.conj(-initial * probe / f ** 2) * g)
    g = np.conj(s_) * g
    g[0, 0, 0] += g_f
    g_d = np.zeros_like(gd)
    g_b = np.zeros_like(gd)
    for i in range(Z - 1, -1, -1):
        gp = np.pad(g, [[0, 0], [pad, pad], [pad, pad]], mode='constant', constant_values=0)
        gu = np.zeros_like(g)
        for a in range(kernel_size):
            for b in range(kernel_size):
                gu += np.conj(kern[a, b]) * gp[:, a:a + Y, b:b + X]          # G'[y - p + a, x - p + b]
        t = np.exp(1j * k * gd[..., i] - k * gb[..., i])
        w = np.conj(gu) * (slices[i] * t)
        g_d[..., i] = -k * w.imag
        g_b[..., i] = -k * w.real
        g = np.conj(t) * gu
    return loss, g_d, g_b, exit_wave


# ---------------------------------------------------------------------------------------------
# model heads (rotation = identity, theta = 0: rotation is a "next" row)
# ---------------------------------------------------------------------------------------------

def fullfield_loss(obj_delta, obj_beta, prj_batch, probe_real, probe_imag, energy_ev, psize_cm,
                   free_prop_cm=None, minibatch_size=1, propagate_last=True):
    """tensorflow_recon/fullfield.py:92-116 with theta = 0 for every batch element."""
    d = np.broadcast_to(obj_delta, (minibatch_size,) + obj_delta.shape)
    b = np.broadcast_to(obj_beta, (minibatch_size,) + obj_beta.shape)
    psi = multislice_forward(d, b, probe_real, probe_imag, energy_ev, psize_cm, free_prop_cm,
                             propagate_last=propagate_last)
    return np.mean((np.abs(psi) - np.abs(prj_batch)) ** 2), psi


def gaussian_probe(probe_size, mag_sigma, phase_sigma, phase_max):
    """tensorflow_recon/ptychography.py:271-281 -- r measured from (size-1)/2."""
    py = np.arange(probe_size[0]) - (probe_size[0] - 1.) / 2
    px = np.arange(probe_size[1]) - (probe_size[1] - 1.) / 2
    pxx, pyy = np.meshgrid(px, py)
    mag = np.exp(-(pxx ** 2 + pyy ** 2) / (2 * mag_sigma ** 2))
    phase = phase_max * np.exp(-(pxx ** 2 + pyy ** 2) / (2 * phase_sigma ** 2))
    return mag * np.cos(phase), mag * np.sin(phase)


def ptycho_windows(obj_yxz, probe_pos, probe_size):
    """Zero-pad and cut one probe_size window per scan position
    (tensorflow_recon/ptychography.py:45-76; note :59 uses probe_size_half[0] for the x pad)."""
    probe_pos = np.asarray(probe_pos).astype(int)
    half = (np.array(probe_size) / 2).astype('int')
    obj_size = obj_yxz.shape
    pad_arr = np.array([[0, 0], [0, 0]])
    o = obj_yxz
    if probe_pos[:, 0].min() - half[0] < 0:
        pl = half[0] - probe_pos[:, 0].min()
        o = np.pad(o, ((pl, 0), (0, 0), (0, 0)), mode='constant'); pad_arr[0, 0] = pl
    if probe_pos[:, 0].max() + half[0] > obj_size[0]:
        pl = probe_pos[:, 0].max() + half[0] - obj_size[0]
        o = np.pad(o, ((0, pl), (0, 0), (0, 0)), mode='constant'); pad_arr[0, 1] = pl
    if probe_pos[:, 1].min() - half[1] < 0:
        pl = half[1] - probe_pos[:, 1].min()
        o = np.pad(o, ((0, 0), (pl, 0), (0, 0)), mode='constant'); pad_arr[1, 0] = pl
    if probe_pos[:, 1].max() + half[1] > obj_size[1]:
        pl = probe_pos[:, 1].max() + half[0] - obj_size[1]
        o = np.pad(o, ((0, 0), (0, pl), (0, 0)), mode='constant'); pad_arr[1, 1] = pl
    wins = []
    for pos in probe_pos:
        y0 = int(pos[0]) + pad_arr[0, 0] - half[0]
        x0 = int(pos[1]) + pad_arr[1, 0] - half[1]
        wins.append(o[y0:y0 + probe_size[0], x0:x0 + probe_size[1], :])
    return np.stack(wins), pad_arr


def ptycho_loss(obj_delta, obj_beta, probe_pos, prj, probe_real, probe_imag, probe_size, energy_ev,
                psize_cm, n_dp_batch=20, scale_by_npos=True):
    """tensorflow_recon/ptychography.py:37-97 at theta = 0 (TF semantics, far field)."""
    wd, _ = ptycho_windows(obj_delta, probe_pos, probe_size)
    wb, _ = ptycho_windows(obj_beta, probe_pos, probe_size)
    outs = []
    for s in range(0, len(probe_pos), n_dp_batch):
        outs.append(multislice_forward(wd[s:s + n_dp_batch], wb[s:s + n_dp_batch], probe_real, probe_imag,
                                       energy_ev, psize_cm, 'inf', propagate_last=True))
    ex = np.concatenate(outs, 0)
    loss = np.mean((np.abs(ex) - np.abs(prj)) ** 2)
    if scale_by_npos:
        loss = loss * len(probe_pos)
    return loss, ex


# ---------------------------------------------------------------------------------------------
# synthetic phantoms
# ---------------------------------------------------------------------------------------------

def zone_plate_phantom(n=512, n_slice=100, lmbda_nm=0.248, focal_nm=7742.0, n_zones=30,
                       delta=1.0e-4, beta=1.0e-5, voxel_nm=1.0):
    """Binary Fresnel zone plate, axially invariant, centred at (n-1)/2 (config 1)."""
    c = (n - 1) / 2.
    yy, xx = np.meshgrid(np.arange(n) - c, np.arange(n) - c, indexing='ij')
    r = np.sqrt(xx ** 2 + yy ** 2) * voxel_nm
    nn = np.arange(n_zones + 1)
    rn = np.sqrt(nn * lmbda_nm * focal_nm + (nn * lmbda_nm / 2.) ** 2)
    zone = np.searchsorted(rn, r, side='right')           # zone index 1..n_zones inside, >n_zones outside
    mask = (zone % 2 == 1) & (zone <= n_zones)
    d2 = np.where(mask, delta, 0.0)
    b2 = np.where(mask, beta, 0.0)
    gd = np.repeat(d2[None, :, :, None], n_slice, axis=3)
    gb = np.repeat(b2[None, :, :, None], n_slice, axis=3)
    return gd, gb


def random_phantom(shape_byxz, seed=1234, delta_scale=1e-5, beta_scale=1e-6):
    """Random delta/beta phantom (config 2): float32 uniform draws, fixed seed."""
    rng = np.random.default_rng(seed)
    gd = rng.random(shape_byxz, dtype=np.float32) * np.float32(delta_scale)
    gb = rng.random(shape_byxz, dtype=np.float32) * np.float32(beta_scale)
    return gd, gb


# ---------------------------------------------------------------------------------------------
# model-level gradients (hand adjoint + window scatter-add); pinned against torch.autograd
# ---------------------------------------------------------------------------------------------

def fullfield_loss_and_grad(obj_delta, obj_beta, prj_batch, probe_real, probe_imag, energy_ev, psize_cm,
                            free_prop_cm=None, propagate_last=True):
    """Data term of tensorflow_recon/fullfield.py:92-116 at theta = 0 and its gradient w.r.t. the
    [Y,X,Z] object (sum of the per-batch-element gradients)."""
    B = len(prj_batch)
    d = np.broadcast_to(obj_delta, (B,) + obj_delta.shape).astype(np.float64)
    b = np.broadcast_to(obj_beta, (B,) + obj_beta.shape).astype(np.float64)
    loss, gd, gb, psi = loss_and_grad(d, b, probe_real, probe_imag, energy_ev, psize_cm, np.abs(prj_batch),
                                      free_prop_cm=free_prop_cm, propagate_last=propagate_last)
    return loss, gd.sum(0), gb.sum(0), psi


def ptycho_loss_and_grad(obj_delta, obj_beta, probe_pos, prj, probe_real, probe_imag, probe_size, energy_ev,
                         psize_cm, scale_by_npos=True):
    """tensorflow_recon/ptychography.py:37-97 at theta = 0 with the gradient w.r.t. the object."""
    probe_pos = np.asarray(probe_pos).astype(int)
    wd, pad_arr = ptycho_windows(obj_delta, probe_pos, probe_size)
    wb, _ = ptycho_windows(obj_beta, probe_pos, probe_size)
    loss, gd, gb, psi = loss_and_grad(wd.astype(np.float64), wb.astype(np.float64), probe_real, probe_imag,
                                      energy_ev, psize_cm, np.abs(prj), free_prop_cm='inf', propagate_last=True)
    scale = len(probe_pos) if scale_by_npos else 1
    half = (np.array(probe_size) / 2).astype('int')
    Y, X, Z = obj_delta.shape
    g_d = np.zeros((Y, X, Z)); g_b = np.zeros((Y, X, Z))
    for n, pos in enumerate(probe_pos):
        y0, x0 = int(pos[0]) - half[0], int(pos[1]) - half[1]
        ys = slice(max(y0, 0), min(y0 + probe_size[0], Y)); xs = slice(max(x0, 0), min(x0 + probe_size[1], X))
        g_d[ys, xs] += gd[n, ys.start - y0:ys.stop - y0, xs.start - x0:xs.stop - x0]
        g_b[ys, xs] += gb[n, ys.start - y0:ys.stop - y0, xs.start - x0:xs.stop - x0]
    return loss * scale, g_d * scale, g_b * scale, psi


# ---------------------------------------------------------------------------------------------
# the steps either side of the hot path in the cnn_propagator drivers
# ---------------------------------------------------------------------------------------------
def rotation_lookup(array_size, theta):
    """Nearest-neighbour rotation table of ONE angle about axis 0, in the (axis 1, axis 2) plane
    (cnn_propagator/util.py:295-336, save_rotation_lookup): for every rotated pixel (x, z), flattened with z
    fastest, the clipped source pixel (x_old, z_old).  array_size = [Y, X, Z]."""
    image_center = [np.floor(x / 2) for x in array_size]
    coord1 = np.arange(array_size[1])
    coord2 = np.arange(array_size[2])
    coord2_vec = np.tile(coord2, array_size[1])                                           # util.py:303
    coord1_vec = np.tile(coord1, array_size[2])
    coord1_vec = np.reshape(coord1_vec, [array_size[1], array_size[2]])
    coord1_vec = np.reshape(np.transpose(coord1_vec), [-1])                               # util.py:305-307
    coord1_vec = coord1_vec - image_center[1]
    coord2_vec = coord2_vec - image_center[2]
    coord_new = np.stack([coord1_vec, coord2_vec]).astype(np.float32)                     # util.py:318
    m_rot = np.array([[np.cos(theta), -np.sin(theta)],
                      [np.sin(theta), np.cos(theta)]])
    coord_old = np.matmul(m_rot, coord_new)
    coord1_old = np.round(coord_old[0, :] + image_center[1]).astype(int)
    coord2_old = np.round(coord_old[1, :] + image_center[2]).astype(int)
    coord1_old = np.clip(coord1_old, 0, array_size[1] - 1)
    coord2_old = np.clip(coord2_old, 0, array_size[2] - 1)
    return np.stack([coord1_old, coord2_old], axis=1)                                     # util.py:333


def apply_rotation(obj, coord_old):
    """obj_rot[y, x, z, c] = obj[y, x_old(x, z), z_old(x, z), c]   (cnn_propagator/util.py:374-402)."""
    s = obj.shape
    c1 = coord_old[:, 0].reshape(s[1], s[2])
    c2 = coord_old[:, 1].reshape(s[1], s[2])
    return obj[:, c1, c2, ...]


def apply_rotation_adjoint(grad_rot, coord_old):
    """Transpose of apply_rotation (what autograd does to the fancy index): scatter-add into the source pixels."""
    s = grad_rot.shape
    c1 = coord_old[:, 0].reshape(s[1], s[2])
    c2 = coord_old[:, 1].reshape(s[1], s[2])
    out = np.zeros_like(grad_rot)
    np.add.at(out, (slice(None), c1, c2), grad_rot)
    return out


def _tf_bilinear_taps(shape_yxz, theta):
    """Source coordinates and weights of tf.contrib.image.rotate(..., 'BILINEAR') on [Y, X, Z, C]: images of height X and
    width Z (TF 1.x contrib/image: angles_to_projective_transforms, then ProjectiveGenerator::bilinear_interpolation with
    fill value 0).  TensorFlow cannot run here: restated from the published algorithm, checked against
    scipy.ndimage.affine_transform(order=1, mode='grid-constant') -- parity unpinned by the reference."""
    Y, X, Z = shape_yxz
    H, W = X, Z
    c, s = np.cos(theta), np.sin(theta)
    x_off = ((W - 1) - (c * (W - 1) - s * (H - 1))) / 2.0
    y_off = ((H - 1) - (s * (W - 1) + c * (H - 1))) / 2.0
    ho, wo = np.meshgrid(np.arange(H, dtype=np.float64), np.arange(W, dtype=np.float64), indexing='ij')
    wi = c * wo - s * ho + x_off
    hi = s * wo + c * ho + y_off
    wf, hf = np.floor(wi), np.floor(hi)
    taps = []
    for dh, wh in ((0, hf + 1 - hi), (1, hi - hf)):
        for dw, ww in ((0, wf + 1 - wi), (1, wi - wf)):
            h, w = (hf + dh).astype(int), (wf + dw).astype(int)
            ok = (h >= 0) & (h < H) & (w >= 0) & (w < W)
            taps.append((np.clip(h, 0, H - 1), np.clip(w, 0, W - 1), wh * ww * ok))
    return taps


def tf_rotate_bilinear(obj, theta):
    """tf_rotate(obj [Y, X, Z, C], theta, interpolation='BILINEAR')   (tensorflow_recon/fullfield.py:96)."""
    out = np.zeros(obj.shape, dtype=np.float64)
    for h, w, wt in _tf_bilinear_taps(obj.shape[:3], theta):
        out += obj[:, h, w, ...] * wt[None, ..., None]
    return out


def tf_rotate_bilinear_adjoint(grad_rot, theta):
    """Transpose of tf_rotate_bilinear (what TF's autodiff of the rotation computes)."""
    out = np.zeros(grad_rot.shape, dtype=np.float64)
    for h, w, wt in _tf_bilinear_taps(grad_rot.shape[:3], theta):
        np.add.at(out, (slice(None), h, w), grad_rot * wt[None, ..., None])
    return out


def apply_gradient_adam(x, g, i_batch, m=None, v=None, step_size=0.001, b1=0.9, b2=0.999, eps=1e-8):
    """cnn_propagator/util.py:280-291 (first call: zero moments)."""
    g = np.array(g)
    if m is None or v is None:
        m = np.zeros_like(x)
        v = np.zeros_like(x)
    m = (1 - b1) * g + b1 * m
    v = (1 - b2) * (g ** 2) + b2 * v
    mhat = m / (1 - b1 ** (i_batch + 1))
    vhat = v / (1 - b2 ** (i_batch + 1))
    x = x - step_size * mhat / (np.sqrt(vhat) + eps)
    return x, m, v


def tomo_loss_and_grad(obj_delta, obj_beta, theta_batch, prj_batch, probe_real, probe_imag, energy_ev, psize_cm,
                       free_prop_cm=None, propagate_last=False, rotation='nearest'):
    """Rotate (nearest-neighbour table) -> multislice -> mean((|psi| - |prj|)^2) over the minibatch, and its
    gradient w.r.t. the UNROTATED object (calculate_loss of cnn_propagator/fullfield.py:93-107 with the FFT
    propagator in place of the real-space one)."""
    Y, X, Z = obj_delta.shape
    obj = np.stack([obj_delta, obj_beta], axis=3).astype(np.float64)
    if rotation == 'bilinear':              # the TF driver (tensorflow_recon/fullfield.py:92-116)
        rot = np.stack([tf_rotate_bilinear(obj, th) for th in theta_batch])
        loss, gd, gb, psi = loss_and_grad(rot[..., 0], rot[..., 1], probe_real, probe_imag, energy_ev, psize_cm,
                                          prj_batch, free_prop_cm=free_prop_cm, propagate_last=propagate_last)
        g = np.zeros_like(obj)
        for b, th in enumerate(theta_batch):
            g += tf_rotate_bilinear_adjoint(np.stack([gd[b], gb[b]], axis=3), th)
        return loss, g[..., 0], g[..., 1], psi
    tabs = [rotation_lookup([Y, X, Z], th) for th in theta_batch]
    rot = np.stack([apply_rotation(obj, t) for t in tabs])                       # [B, Y, X, Z, 2]
    loss, gd, gb, psi = loss_and_grad(rot[..., 0], rot[..., 1], probe_real, probe_imag, energy_ev, psize_cm,
                                      prj_batch, free_prop_cm=free_prop_cm, propagate_last=propagate_last)
    g = np.zeros_like(obj)
    for b, t in enumerate(tabs):
        g += apply_rotation_adjoint(np.stack([gd[b], gb[b]], axis=3), t)
    return loss, g[..., 0], g[..., 1], psi
