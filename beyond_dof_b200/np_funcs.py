"""Drop-in for cnn_propagator/np_funcs.py:15-65 -- the variant of multislice_propagate_batch_numpy that also returns the
field after every slice:

    wavefront, probe_array = multislice_propagate_batch_numpy(grid_delta_batch, grid_beta_batch, probe_real, probe_imag,
                                                              energy_ev, psize_cm, free_prop_cm=None, obj_batch_shape=None)

probe_array[i] ([n_slice, B, Y, X]) is the wavefield after slice i (modulated, and propagated except for the last slice;
np_funcs.py:36-41), before the free-space step.  PI = 3.1415927 as in that file (np_funcs.py:12).  The slices are stepped one
by one on the GPU (bdof_slice_step: per-pass kernels); the global phase exp(i k dz) per propagation, which the engine keeps
out of its fp32 chain, is restored in float64 on every returned field.
"""
import numpy as np
import torch

from .plan import MultislicePlan
from .propagation import _cached_plan, _is_torch, _probe_c64, _to_dev
from .util import kernel_factors

PI = 3.1415927                                  # cnn_propagator/np_funcs.py:12


def multislice_propagate_batch_numpy(grid_delta_batch, grid_beta_batch, probe_real, probe_imag, energy_ev, psize_cm,
                                     free_prop_cm=None, obj_batch_shape=None):
    shape = tuple(int(v) for v in (obj_batch_shape if obj_batch_shape is not None else grid_delta_batch.shape))
    B, Y, X, Z = shape
    key = ('npf', shape, float(energy_ev), float(psize_cm), free_prop_cm, torch.cuda.current_device() if torch.cuda.is_available() else -1)
    plan = _cached_plan(key, lambda: MultislicePlan(Y, X, B, Z, energy_ev, psize_cm, free_prop_cm=free_prop_cm,
                                                     propagate_last=False, pi=PI, stepwise=True))
    db = plan.pack(_to_dev(grid_delta_batch, torch.float32), _to_dev(grid_beta_batch, torch.float32))
    field = _probe_c64(probe_real, probe_imag, (Y, X)).unsqueeze(0).expand(B, Y, X).contiguous()
    p0 = complex(kernel_factors(plan.voxel_nm[-1], plan.lmbda_nm, plan.voxel_nm, [Y, X, Z], pi=PI)[0])
    phase = 1.0 + 0.0j
    probe_array = torch.empty((Z, B, Y, X), dtype=torch.complex64, device=field.device)
    for i in range(Z):
        prop = i < Z - 1
        field = plan.slice_step(field, db[i], propagate=prop, index=i)
        if prop:
            phase *= p0
        torch.mul(field, phase, out=probe_array[i])
    wavefront = plan.free_prop(field) * phase     # the free-space step applies its own global phase (bdof_free_prop)
    if _is_torch(grid_delta_batch):
        cpu = not grid_delta_batch.is_cuda
        return (wavefront.cpu(), probe_array.cpu()) if cpu else (wavefront, probe_array)
    return wavefront.cpu().numpy(), probe_array.cpu().numpy()
