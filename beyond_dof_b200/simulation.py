"""Dataset simulators of the reference with the multislice on the GPU (the callers of the NumPy hot path).

  create_fullfield_data_numpy           <- tensorflow_recon/simulation.py:80-161   (identical file in cnn_propagator/)
  create_ptychography_data_batch_numpy  <- tensorflow_recon/simulation.py:283-386

Same arguments and the same steps as the reference: read grid_delta.npy / grid_beta.npy from `phantom_path`, rotate the object
for every angle with scipy.ndimage.rotate (the reference's own call: degrees, cubic spline, reshape=False, axes=(1, 2); host
side, as there), project every rotated object (batch of angles, or batches of probe windows) with
multislice_propagate_batch_numpy -- here the libbdof engine -- and store the complex64 exit waves as `exchange/data` of an
HDF5 file.  h5py is not a dependency of this package: when it cannot be imported the array is written next to the requested
name as `<fname>.npy`; either way it is also returned.  TIFF monitor outputs (dxchange) are not written.
probe_type 'point' (spherical-wave propagator) is outside the FFT multislice path and raises.
"""
import os

import numpy as np
from scipy.ndimage import rotate as sp_rotate

from .propagation import multislice_propagate_batch_numpy
from .util import PI


def mag_phase_to_real_imag(mag, phase):
    """tensorflow_recon/util.py: real = mag cos(phase), imag = mag sin(phase)."""
    return mag * np.cos(phase), mag * np.sin(phase)


def _gaussian_probe(shape, mag_sigma, phase_sigma, phase_max):
    py = np.arange(shape[0]) - (shape[0] - 1.) / 2
    px = np.arange(shape[1]) - (shape[1] - 1.) / 2
    pxx, pyy = np.meshgrid(px, py)
    probe_mag = np.exp(-(pxx ** 2 + pyy ** 2) / (2 * mag_sigma ** 2))
    probe_phase = phase_max * np.exp(-(pxx ** 2 + pyy ** 2) / (2 * phase_sigma ** 2))
    return mag_phase_to_real_imag(probe_mag, probe_phase)


def _save(save_folder, fname, arr):
    """`exchange/data` of an HDF5 file (simulation.py:131-133), or <fname>.npy when h5py is not installed"""
    os.makedirs(save_folder, exist_ok=True)
    path = os.path.join(save_folder, fname)
    try:
        import h5py
    except ImportError:
        np.save(path + '.npy', arr)
        return path + '.npy'
    with h5py.File(path, 'w') as f:
        f.create_group('exchange').create_dataset('data', data=arr)
    return path


def _load_obj(phantom_path):
    grid_delta = np.load(os.path.join(phantom_path, 'grid_delta.npy'))
    grid_beta = np.load(os.path.join(phantom_path, 'grid_beta.npy'))
    obj = np.zeros(np.append(grid_delta.shape, 2))
    obj[:, :, :, 0] = grid_delta
    obj[:, :, :, 1] = grid_beta
    return obj, grid_delta.shape


def create_fullfield_data_numpy(energy_ev, psize_cm, free_prop_cm, n_theta, phantom_path, save_folder, fname, batch_size=1,
                                probe_type='plane', wavefront_initial=None, theta_st=0, theta_end=2 * PI, monitor_output=False,
                                **kwargs):
    """simulation.py:80-161.  Returns the [n_theta, Y, X] complex64 projections it wrote."""
    if probe_type == 'point':
        raise NotImplementedError("probe_type='point' (multislice_propagate_spherical_numpy) is outside the FFT multislice path")
    obj, img_dim = _load_obj(phantom_path)
    theta_ls = -np.linspace(theta_st, theta_end, n_theta) / np.pi * 180
    n_batch = np.ceil(float(n_theta) / batch_size)
    theta_batch = np.array_split(theta_ls, n_batch)
    if probe_type == 'plane':
        probe_real = np.ones([img_dim[0], img_dim[1]], dtype='float32')
        probe_imag = np.zeros([img_dim[0], img_dim[1]], dtype='float32')
    elif probe_type == 'fixed':
        probe_mag, probe_phase = wavefront_initial
        probe_real, probe_imag = mag_phase_to_real_imag(probe_mag, probe_phase)
    elif probe_type == 'gaussian':
        probe_real, probe_imag = _gaussian_probe(obj.shape, kwargs['probe_mag_sigma'], kwargs['probe_phase_sigma'], kwargs['probe_phase_max'])
    else:
        raise ValueError('Invalid wavefront type. Choose from \'plane\', \'point\', or \'fixed\'.')
    dat = np.zeros((n_theta, img_dim[0], img_dim[1]), dtype=np.complex64)
    for i_batch, this_theta_batch in enumerate(theta_batch):
        obj_rot_batch = np.array([sp_rotate(obj, theta, reshape=False, axes=(1, 2)) for theta in this_theta_batch])
        wave_out = multislice_propagate_batch_numpy(obj_rot_batch[:, :, :, :, 0], obj_rot_batch[:, :, :, :, 1], probe_real, probe_imag,
                                                    energy_ev, psize_cm, free_prop_cm=free_prop_cm,
                                                    obj_batch_shape=obj_rot_batch.shape[:-1])
        # (the reference indexes with i_batch * batch_size although array_split makes uneven batches; kept)
        dat[i_batch * batch_size:i_batch * batch_size + batch_size, :, :] = wave_out
    _save(save_folder, fname, dat)
    return dat


def create_ptychography_data_batch_numpy(energy_ev, psize_cm, n_theta, phantom_path, save_folder, fname, probe_pos,
                                         probe_type='gaussian', probe_size=(72, 72), wavefront_initial=None,
                                         theta_st=0, theta_end=2 * PI, probe_circ_mask=0.9, minibatch_size=20, **kwargs):
    """simulation.py:283-386.  If probe_type is 'gaussian', supply 'probe_mag_sigma', 'probe_phase_sigma', 'probe_phase_max'.
    probe_circ_mask must be None: the reference's masking calls tomopy.circ_mask, which simulation.py never imports (NameError
    there) and which is not a dependency here.  Returns the [n_theta, n_pos, py, px] complex64 far-field waves it wrote."""
    if probe_circ_mask is not None:
        raise NotImplementedError('probe_circ_mask needs tomopy.circ_mask (not imported by the reference either): pass probe_circ_mask=None')
    if probe_type != 'gaussian':
        raise ValueError("the reference only builds probe_type='gaussian' here (simulation.py:366-377)")
    probe_pos = np.array(probe_pos)
    n_pos = len(probe_pos)
    minibatch_size = min([minibatch_size, n_pos])
    n_batch = np.ceil(float(n_pos) / minibatch_size)
    probe_pos_batches = np.array_split(probe_pos, n_batch)
    obj, img_dim = _load_obj(phantom_path)
    probe_size_half = (np.array(probe_size) / 2).astype('int')
    theta_ls = np.rad2deg(-np.linspace(theta_st, theta_end, n_theta))
    probe_real, probe_imag = _gaussian_probe(probe_size, kwargs['probe_mag_sigma'], kwargs['probe_phase_sigma'], kwargs['probe_phase_max'])
    dat = np.zeros((n_theta, n_pos, probe_size[0], probe_size[1]), dtype=np.complex64)
    for ii, theta in enumerate(theta_ls):
        obj_rot = sp_rotate(obj, theta, reshape=False, axes=(1, 2))
        pad_arr = np.array([[0, 0], [0, 0]])
        if probe_pos[:, 0].min() - probe_size_half[0] < 0:
            pad_len = probe_size_half[0] - probe_pos[:, 0].min()
            obj_rot = np.pad(obj_rot, ((pad_len, 0), (0, 0), (0, 0), (0, 0)), mode='constant')
            pad_arr[0, 0] = pad_len
        if probe_pos[:, 0].max() + probe_size_half[0] > img_dim[0]:
            pad_len = probe_pos[:, 0].max() + probe_size_half[0] - img_dim[0]
            obj_rot = np.pad(obj_rot, ((0, pad_len), (0, 0), (0, 0), (0, 0)), mode='constant')
            pad_arr[0, 1] = pad_len
        if probe_pos[:, 1].min() - probe_size_half[1] < 0:
            pad_len = probe_size_half[1] - probe_pos[:, 1].min()
            obj_rot = np.pad(obj_rot, ((0, 0), (pad_len, 0), (0, 0), (0, 0)), mode='constant')
            pad_arr[1, 0] = pad_len
        if probe_pos[:, 1].max() + probe_size_half[1] > img_dim[1]:
            pad_len = probe_pos[:, 1].max() + probe_size_half[0] - img_dim[1]          # [0], as written in the reference (:305)
            obj_rot = np.pad(obj_rot, ((0, 0), (0, pad_len), (0, 0), (0, 0)), mode='constant')
            pad_arr[1, 1] = pad_len
        outs = []
        for pos_batch in probe_pos_batches:
            subs = []
            for pos in pos_batch:
                pos = np.array(pos, dtype=int)
                y0 = pos[0] + pad_arr[0, 0] - probe_size_half[0]
                x0 = pos[1] + pad_arr[1, 0] - probe_size_half[1]
                subs.append(obj_rot[y0:y0 + probe_size[0], x0:x0 + probe_size[1], :, :])
            subs = np.array(subs)
            outs.append(multislice_propagate_batch_numpy(subs[..., 0], subs[..., 1], probe_real, probe_imag, energy_ev, psize_cm,
                                                          free_prop_cm='inf',
                                                          obj_batch_shape=[len(pos_batch), probe_size[0], probe_size[1], img_dim[-1]]))
        dat[ii] = np.vstack(outs)
    _save(save_folder, fname, dat)
    return dat
