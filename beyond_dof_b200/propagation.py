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


def multislice_propagate(grid_delta, grid_beta, probe_real, probe_imag, energy_ev, psize_cm, h=None, free_prop_cm=None,
                         pad=None):
    """Un-batched [Y,X,Z] variant (util.py:360-429); all shipped configurations take its TF branch."""
    if pad is not None:
        if _is_torch(grid_delta):
            flat = [int(v) for pr in reversed(list(pad)) for v in pr]
            grid_delta = torch.nn.functional.pad(grid_delta, flat)
            grid_beta = torch.nn.functional.pad(grid_beta, flat)
        else:
            grid_delta = np.pad(grid_delta, pad, 'constant')
            grid_beta = np.pad(grid_beta, pad, 'constant')
    out = _run_forward(grid_delta[None], grid_beta[None], probe_real, probe_imag, energy_ev, psize_cm, free_prop_cm,
                       None, propagate_last=True, h=h)
    return out[0]


def cnn_kernel(energy_ev, psize_cm, grid_shape_yxz, kernel_size):
    """The cropped real-space propagator of propagation.py:35-44 (float64 on the host)."""
    lmbda_nm = 1240. / energy_ev
    voxel_nm = np.array(psize_cm) * 1.e7
    kern = get_kernel(voxel_nm[-1], lmbda_nm, voxel_nm, np.array(grid_shape_yxz) - 1, pi=PI_CNN)
    kern = np.fft.fftshift(np.fft.ifft2(np.fft.ifftshift(kern)))
    mid = ((np.array(kern.shape) - 1) / 2).astype('int')
    half = int((kernel_size - 1) / 2)
    return kern[mid[0] - half:mid[0] + half + 1, mid[1] - half:mid[1] + half + 1]


def multislice_propagate_cnn(grid_delta, grid_beta, probe_real, probe_imag, energy_ev, psize_cm, kernel_size=17,
                             free_prop_cm=None, debug=False):
    """Real-space ("cnn") multislice (propagation.py:18-133): per slice modulate, pad with the tracked
    plane-wave edge value, convolve with the kernel_size^2 crop of IFFT(H); rescale by the corner pixel."""
    assert kernel_size % 2 == 1, 'kernel_size must be an odd number.'
    as_torch = _is_torch(grid_delta)
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
    out = torch.empty((B, Y, X), dtype=torch.complex64, device=dev)
    work = torch.empty((2, B, Y, X), dtype=torch.complex64, device=dev)
    check(lib.bdof_cnn_forward(_ptr(db), _ptr(probe), _ptr(out), _ptr(work), B, Y, X, Z, _hptr(kern), kernel_size,
                               float(k_dz), st))
    # propagation.py:79,109-110: rescale by the corner pixel
    out = out * (probe[0, 0] / out[0, 0, 0])
    if free_prop_cm is not None:
        if isinstance(free_prop_cm, str):
            out = torch.fft.fftshift(torch.fft.fft2(out), dim=(1, 2))
        else:
            hf = torch.as_tensor(np.fft.ifftshift(get_kernel(free_prop_cm * 1e7, lmbda_nm, voxel_nm, (Y, X, Z), pi=PI_CNN))).to(dev, torch.complex64)
            out = torch.fft.ifft2(torch.fft.fft2(out) * hf)
    if as_torch:
        res = out if grid_delta.is_cuda else out.cpu()
    else:
        res = out.cpu().numpy()
    if debug:
        return res, None, 0.0
    return res
