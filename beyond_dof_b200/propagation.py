"""Drop-in multislice entry points with the reference's names, arguments and semantics.

  multislice_propagate_batch_numpy  <- tensorflow_recon/npfuncs.py:16-63 (and cnn_propagator/np_funcs.py:15-65)
  multislice_propagate_batch        <- tensorflow_recon/util.py:432-508  (type='plane')
  multislice_propagate              <- tensorflow_recon/util.py:360-429
  multislice_propagate_cnn          <- cnn_propagator/propagation.py:18-133

Arrays in: NumPy or torch (CPU or CUDA), object layout [B,Y,X,Z] with z fastest, exactly as the
reference takes them.  Arrays out: the same container kind, [B,Y,X] complex64.  All arithmetic
runs on the GPU through libbdof (complex64); there is no CPU path.
"""
import ctypes

import numpy as np
import torch

from . import capi
from .capi import lib, check
from .plan import MultislicePlan, _ptr, _hptr
from .util import PI, PI_CNN, get_kernel

_PLAN_CACHE = {}
_PLAN_CACHE_MAX = 8                    # plans
_PLAN_CACHE_MAX_BYTES = 32 << 30       # device memory the cached plans may own together (work fields + slice stores)


def _cached_plan(key, make):
    """Memoised plans, evicted oldest first when there are more than _PLAN_CACHE_MAX of them or when their workspaces add up
    to more than _PLAN_CACHE_MAX_BYTES (a 2048^2 x 256 plan with a slice store owns 8.7 GB).  Plans re-bind to the
    caller's current stream on every call (MultislicePlan.use_current_stream), so reuse across streams is safe."""
    p = _PLAN_CACHE.get(key)
    if p is None:
        p = make()
        _PLAN_CACHE[key] = p
        total = sum(q.workspace_bytes() for q in _PLAN_CACHE.values())
        while len(_PLAN_CACHE) > 1 and (len(_PLAN_CACHE) > _PLAN_CACHE_MAX or total > _PLAN_CACHE_MAX_BYTES):
            old_key = next(iter(_PLAN_CACHE))
            if old_key == key:
                break
            total -= _PLAN_CACHE.pop(old_key).workspace_bytes()
    return p


def clear_plan_cache():
    _PLAN_CACHE.clear()


def _is_torch(x):
    return isinstance(x, torch.Tensor)


def _device():
    if not torch.cuda.is_available():
        raise RuntimeError('beyond_dof_b200 needs a CUDA device: the multislice path has no CPU fallback')
    return torch.device('cuda', torch.cuda.current_device())


def _to_dev(x, dtype):
    if _is_torch(x):
        return x.to(_device(), dtype)
    return torch.as_tensor(np.ascontiguousarray(x)).to(_device(), dtype)


def _probe_c64(probe_real, probe_imag, shape):
    pr = _to_dev(probe_real, torch.float32)
    pi = _to_dev(probe_imag, torch.float32)
    return torch.complex(pr.expand(shape).contiguous(), pi.expand(shape).contiguous())


def _h_key(h):
    if h is None:
        return None
    hh = h.detach().cpu().numpy() if _is_torch(h) else np.asarray(h)
    return (hh.shape, hash(hh.tobytes()))


def _run_forward(grid_delta_batch, grid_beta_batch, probe_real, probe_imag, energy_ev, psize_cm, free_prop_cm,
                 obj_batch_shape, propagate_last, h=None):
    shape = tuple(int(s) for s in (obj_batch_shape if obj_batch_shape is not None else grid_delta_batch.shape))
    B, Y, X, Z = shape
    key = ('fwd', shape, float(energy_ev), float(psize_cm), free_prop_cm, propagate_last, _h_key(h),
           torch.cuda.current_device())
    hh = None if h is None else (h.detach().cpu().numpy() if _is_torch(h) else np.asarray(h))
    plan = _cached_plan(key, lambda: MultislicePlan(Y, X, B, Z, energy_ev, psize_cm, free_prop_cm=free_prop_cm,
                                                     propagate_last=propagate_last, h=hh))
    as_torch = _is_torch(grid_delta_batch)
    if not as_torch and not _is_torch(probe_real):
        # host arrays in, host array out: the whole trip runs behind one C-ABI call
        probe = (np.zeros((Y, X), dtype=np.complex64) + (np.asarray(probe_real) + 1j * np.asarray(probe_imag))).astype(np.complex64)
        return plan.forward_host(np.asarray(grid_delta_batch), np.asarray(grid_beta_batch), probe)
    db = plan.pack(_to_dev(grid_delta_batch, torch.float32), _to_dev(grid_beta_batch, torch.float32))
    out = plan.forward(db, _probe_c64(probe_real, probe_imag, (Y, X)))
    if as_torch:
        return out if grid_delta_batch.is_cuda else out.cpu()
    return out.cpu().numpy()


def multislice_propagate_batch_numpy(grid_delta_batch, grid_beta_batch, probe_real, probe_imag, energy_ev, psize_cm,
                                     free_prop_cm=None, obj_batch_shape=None):
    """NumPy-semantics forward multislice (npfuncs.py:16-63): the last slice modulates but does not
    propagate; free_prop_cm in {None, 'inf', float cm}.  Returns [B,Y,X] complex64."""
    return _run_forward(grid_delta_batch, grid_beta_batch, probe_real, probe_imag, energy_ev, psize_cm, free_prop_cm,
                        obj_batch_shape, propagate_last=False)


def multislice_propagate_batch(grid_delta_batch, grid_beta_batch, probe_real, probe_imag, energy_ev, psize_cm, h=None,
                               free_prop_cm=None, obj_batch_shape=None, type='plane', **kwargs):
    """TF-semantics forward multislice (util.py:432-508): every slice propagates (n_slice == 1 only
    modulates); optional caller-supplied centred kernel h."""
    if type != 'plane':
        raise NotImplementedError("type='projection' (paraxial magnification, util.py:473-475) is outside the "
                                  "FFT multislice hot path")
    return _run_forward(grid_delta_batch, grid_beta_batch, probe_real, probe_imag, energy_ev, psize_cm, free_prop_cm,
                        obj_batch_shape, propagate_last=True, h=h)


def get_kernel_ir(dist_nm, lmbda_nm, voxel_nm, grid_shape):
    """Impulse-response Fresnel kernel, H = fftshift(fft2(h)) dx dy (tensorflow_recon/util.py:188-216), float64 on the host;
    the reference's np.arange grid is kept literally (its length is what np.arange makes of the float end point)."""
    size_nm = np.array(voxel_nm) * np.array(grid_shape)
    k = 2 * PI / lmbda_nm
    ymin, xmin = np.array(size_nm)[:2] / -2.
    dy, dx = voxel_nm[0:2]
    x = np.arange(xmin, xmin + size_nm[1], dx)
    y = np.arange(ymin, ymin + size_nm[0], dy)
    x, y = np.meshgrid(x, y)
    h = np.exp(1j * k * dist_nm) / (1j * lmbda_nm * dist_nm) * np.exp(1j * k / (2 * dist_nm) * (x ** 2 + y ** 2))
    return np.fft.fftshift(np.fft.fft2(h), axes=(-2, -1)) * voxel_nm[0] * voxel_nm[1]


def propagation_algorithm(dist_nm, lmbda_nm, voxel_nm, grid_shape):
    """'TF' or 'IR': the sampling criterion of multislice_propagate (util.py:396-404, 416-419):
    TF when the mean voxel exceeds lambda d / L, L = (prod of the grid's physical sides)^(1/3)."""
    voxel_nm = np.asarray(voxel_nm, dtype=np.float64)
    mean_voxel_nm = np.prod(voxel_nm) ** (1. / 3)
    size_nm = np.array(grid_shape) * voxel_nm
    crit_samp = lmbda_nm * dist_nm / (np.prod(size_nm) ** (1. / 3))
    return 'TF' if mean_voxel_nm > crit_samp else 'IR'


def multislice_propagate(grid_delta, grid_beta, probe_real, probe_imag, energy_ev, psize_cm, h=None, free_prop_cm=None,
                         pad=None):
    """Un-batched [Y,X,Z] variant (util.py:360-429).  Differences from the batched function that are reproduced here:
      * EVERY slice propagates, also when n_slice == 1 (no modulate-only special case, util.py:406-408);
      * per slice and for a finite free_prop_cm the algorithm is chosen by the sampling criterion (propagation_algorithm):
        - per slice, 'IR' = ifftshift(ifft2(fft2(fftshift(psi)) * h)) with the SAME transfer-function kernel h (util.py:401-404):
          for even sides that is the convolution whose multiplier is h in natural FFT order (no ifftshift of h);
        - free space, 'IR' = ifftshift(fft2(fftshift(fft2(psi)) * get_kernel_ir(d))) -- the reference applies a FORWARD
          transform twice there (util.py:424-427); reproduced literally.
    Sides must be even for the IR orderings (fftshift == ifftshift); odd sides raise."""
    as_torch = _is_torch(grid_delta)
    if pad is not None:
        if as_torch:
            flat = [int(v) for pr in reversed(list(pad)) for v in pr]
            grid_delta = torch.nn.functional.pad(grid_delta, flat)
            grid_beta = torch.nn.functional.pad(grid_beta, flat)
        else:
            grid_delta = np.pad(grid_delta, pad, 'constant')
            grid_beta = np.pad(grid_beta, pad, 'constant')
    Y, X, Z = (int(v) for v in grid_delta.shape)
    voxel_nm = np.array([psize_cm] * 3) * 1.e7
    lmbda_nm = 1240. / energy_ev
    shape3 = [Y, X, Z]
    slice_alg = propagation_algorithm(voxel_nm[-1], lmbda_nm, voxel_nm, shape3)
    free_alg = None
    if free_prop_cm is not None and not isinstance(free_prop_cm, str):
        free_alg = propagation_algorithm(free_prop_cm * 1e7, lmbda_nm, voxel_nm, shape3)
    if (slice_alg == 'IR' or free_alg == 'IR') and (Y % 2 or X % 2):
        raise NotImplementedError('the IR orderings of multislice_propagate are implemented for even field sides only')
    hh = None if h is None else (h.detach().cpu().numpy() if _is_torch(h) else np.asarray(h))
    if slice_alg == 'IR':
        # multiplier h applied in natural FFT order: hand the plan fftshift(h), whose ifftshift is h again
        hk = get_kernel(voxel_nm[-1], lmbda_nm, voxel_nm, shape3) if hh is None else hh
        hh = np.fft.fftshift(hk, axes=(-2, -1))
    plan_free = free_prop_cm if free_alg != 'IR' else None
    key = ('fwd1', (Y, X, Z), float(energy_ev), float(psize_cm), plan_free, slice_alg, _h_key(hh), torch.cuda.current_device())
    dev = _device()
    d = _to_dev(grid_delta, torch.float32)[None]
    b = _to_dev(grid_beta, torch.float32)[None]
    if Z == 1:
        # one slice that DOES propagate: append a vacuum slice and drop its propagation with the NumPy semantics
        plan1 = _cached_plan(key + ('z1',), lambda: MultislicePlan(Y, X, 1, 2, energy_ev, psize_cm, free_prop_cm=plan_free,
                                                                    propagate_last=False, h=hh))
        zero = torch.zeros_like(d)
        db = plan1.pack(torch.cat([d, zero], dim=3), torch.cat([b, zero], dim=3))
        out = plan1.forward(db, _probe_c64(probe_real, probe_imag, (Y, X)))
    else:
        plan = _cached_plan(key, lambda: MultislicePlan(Y, X, 1, Z, energy_ev, psize_cm, free_prop_cm=plan_free, propagate_last=True,
                                                         h=hh))
        db = plan.pack(d, b)
        out = plan.forward(db, _probe_c64(probe_real, probe_imag, (Y, X)))
    if free_alg == 'IR':
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        far = _cached_plan(('far', (Y, X), torch.cuda.current_device()),
                           lambda: MultislicePlan(Y, X, 1, 1, energy_ev, psize_cm, free_prop_cm='inf'))
        H = torch.as_tensor(get_kernel_ir(free_prop_cm * 1e7, lmbda_nm, voxel_nm, shape3)).to(dev, torch.complex64).contiguous()
        spec = far.free_prop(out)                                  # fftshift(fft2(psi))
        check(lib.bdof_field_multiply(_ptr(spec), _ptr(H), _ptr(spec), 1, Y * X, st))
        out = far.free_prop(spec)                                  # ifftshift(fft2(.)) == fftshift(fft2(.)) for even sides
    out = out[0]
    if as_torch:
        return out if grid_delta.is_cuda else out.cpu()
    return out.cpu().numpy()


def cnn_kernel(energy_ev, psize_cm, grid_shape_yxz, kernel_size):
    """The cropped real-space propagator of propagation.py:35-44 (float64 on the host)."""
    lmbda_nm = 1240. / energy_ev
    voxel_nm = np.array(psize_cm) * 1.e7
    kern = get_kernel(voxel_nm[-1], lmbda_nm, voxel_nm, np.array(grid_shape_yxz) - 1, pi=PI_CNN)
    kern = np.fft.fftshift(np.fft.ifft2(np.fft.ifftshift(kern)))
    mid = ((np.array(kern.shape) - 1) / 2).astype('int')
    half = int((kernel_size - 1) / 2)
    return kern[mid[0] - half:mid[0] + half + 1, mid[1] - half:mid[1] + half + 1]


def _cnn_free_plan(Y, X, B, energy_ev, psize_cm, free_prop_cm):
    """free-space step of the cnn propagator (propagation.py:112-128: 'inf' or the TF kernel with PI = 3.1415927) as a plan"""
    key = ('cnnfree', (B, Y, X), float(energy_ev), tuple(np.atleast_1d(psize_cm).astype(float)), free_prop_cm, torch.cuda.current_device())
    return _cached_plan(key, lambda: MultislicePlan(Y, X, B, 1, energy_ev, psize_cm, free_prop_cm=free_prop_cm, pi=PI_CNN))


def _cnn_setup(grid_delta, grid_beta, probe_real, probe_imag, energy_ev, psize_cm, kernel_size):
    assert kernel_size % 2 == 1, 'kernel_size must be an odd number.'
    B, Y, X, Z = (int(s) for s in grid_delta.shape)
    dev = _device()
    voxel_nm = np.array(psize_cm) * 1.e7
    lmbda_nm = 1240. / energy_ev
    k_dz = 2. * np.pi * voxel_nm[-1] / lmbda_nm
    kern = np.ascontiguousarray(cnn_kernel(energy_ev, psize_cm, (Y, X, Z), kernel_size), dtype=np.complex128)
    d = _to_dev(grid_delta, torch.float32).contiguous()
    b = _to_dev(grid_beta, torch.float32).contiguous()
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    db = torch.empty((Z, B, Y, X, 2), dtype=torch.float32, device=dev)
    check(lib.bdof_pack_db(_ptr(d), _ptr(b), _ptr(db), B, Y, X, Z, st))
    probe = _probe_c64(probe_real, probe_imag, (Y, X))
    return (B, Y, X, Z), dev, k_dz, kern, db, probe, st


def multislice_propagate_cnn(grid_delta, grid_beta, probe_real, probe_imag, energy_ev, psize_cm, kernel_size=17,
                             free_prop_cm=None, debug=False):
    """Real-space ("cnn") multislice (propagation.py:18-133): per slice modulate, pad with the tracked
    plane-wave edge value, convolve with the kernel_size^2 crop of IFFT(H); rescale by the corner pixel.
    debug=True also returns (probe_array, seconds): the magnitude of the field after every slice (propagation.py:107,130)."""
    import time
    t0 = time.time()
    as_torch = _is_torch(grid_delta)
    (B, Y, X, Z), dev, k_dz, kern, db, probe, st = _cnn_setup(grid_delta, grid_beta, probe_real, probe_imag, energy_ev, psize_cm,
                                                              kernel_size)
    probe_array = None
    if debug:
        slices = torch.empty((Z + 1, B, Y, X), dtype=torch.complex64, device=dev)
        check(lib.bdof_cnn_forward_store(_ptr(db), _ptr(probe), _ptr(slices), B, Y, X, Z, _hptr(kern), kernel_size, float(k_dz), st))
        out = slices[Z]
        probe_array = slices[1:].abs()
    else:
        out = torch.empty((B, Y, X), dtype=torch.complex64, device=dev)
        work = torch.empty((2, B, Y, X), dtype=torch.complex64, device=dev)
        check(lib.bdof_cnn_forward(_ptr(db), _ptr(probe), _ptr(out), _ptr(work), B, Y, X, Z, _hptr(kern), kernel_size,
                                   float(k_dz), st))
    # propagation.py:79,109-110: rescale by the corner pixel of batch element 0
    out = out * (probe[0, 0] / out[0, 0, 0])
    if free_prop_cm is not None:
        out = _cnn_free_plan(Y, X, B, energy_ev, psize_cm, free_prop_cm).free_prop(out)
    conv = (lambda t: t if grid_delta.is_cuda else t.cpu()) if as_torch else (lambda t: t.cpu().numpy())
    res = conv(out)
    if debug:
        pa = conv(probe_array)
        return res, (pa if as_torch else list(pa)), time.time() - t0
    return res


def cnn_loss_and_grad(grid_delta, grid_beta, probe_real, probe_imag, energy_ev, psize_cm, target_mag, kernel_size=17,
                      free_prop_cm=None):
    """loss = mean((|psi| - target)^2) through multislice_propagate_cnn and its gradient (g_delta, g_beta) [B,Y,X,Z] -- the pair
    autograd.grad(calculate_loss, [0, 1]) returns in the cnn_propagator drivers (fullfield.py:93-121,329; ptychography.py:30-81,
    248) for already rotated / cut batches.  The chain runs on bdof_cnn_forward_store / bdof_cnn_adjoint; the corner-pixel
    rescaling (propagation.py:79,109-110) and its gradient, which couples every batch element to pixel [0,0,0], are applied here."""
    as_torch = _is_torch(grid_delta)
    (B, Y, X, Z), dev, k_dz, kern, db, probe, st = _cnn_setup(grid_delta, grid_beta, probe_real, probe_imag, energy_ev, psize_cm,
                                                              kernel_size)
    slices = torch.empty((Z + 1, B, Y, X), dtype=torch.complex64, device=dev)
    check(lib.bdof_cnn_forward_store(_ptr(db), _ptr(probe), _ptr(slices), B, Y, X, Z, _hptr(kern), kernel_size, float(k_dz), st))
    psi_z = slices[Z]
    initial, f = probe[0, 0].to(torch.complex128), psi_z[0, 0, 0].to(torch.complex128)
    s_ = initial / f
    psi = (psi_z * s_.to(torch.complex64)).contiguous()
    fplan = _cnn_free_plan(Y, X, B, energy_ev, psize_cm, free_prop_cm)
    exit_wave = fplan.free_prop(psi) if free_prop_cm is not None else psi
    loss, g = fplan.loss_mag(exit_wave, _to_dev(target_mag, torch.float32))
    if free_prop_cm is not None:
        g = fplan.free_prop_adjoint(g)
    g_f = torch.sum(torch.conj(-initial * psi_z.to(torch.complex128) / f ** 2) * g.to(torch.complex128))
    G = torch.empty((2, B, Y, X), dtype=torch.complex64, device=dev)
    G[0] = torch.conj(s_).to(torch.complex64) * g
    G[0, 0, 0, 0] += g_f.to(torch.complex64)
    check(lib.bdof_cnn_adjoint(_ptr(db), _ptr(slices), _ptr(G), _ptr(db), B, Y, X, Z, _hptr(kern), kernel_size, float(k_dz), st))
    g_d = torch.empty((B, Y, X, Z), dtype=torch.float32, device=dev)
    g_b = torch.empty((B, Y, X, Z), dtype=torch.float32, device=dev)
    check(lib.bdof_unpack_db(_ptr(db), _ptr(g_d), _ptr(g_b), B, Y, X, Z, st))
    if as_torch and not grid_delta.is_cuda:
        g_d, g_b = g_d.cpu(), g_b.cpu()
    elif not as_torch:
        g_d, g_b = g_d.cpu().numpy(), g_b.cpu().numpy()
    return loss, (g_d, g_b), exit_wave
