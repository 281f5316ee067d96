"""Generate golden vectors by running the UNMODIFIED reference functions.  TEST INFRASTRUCTURE.

Run in the build container only (needs /root/reference; the GPU box does not have it):

    python oracle/gen_golden.py            # writes tests/golden/*.npz

The reference modules are imported from where they lie; the packages they import but do not
use on this path (tensorflow, dxchange, h5py, matplotlib, autograd, pyfftw) are replaced by
stubs that never touch the arithmetic.  Each case stores the seed /
recipe of its inputs plus the reference output, so the fixtures stay small.
"""
import os
import sys
import subprocess
import json

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get('BDOF_REFERENCE', '/root/reference')
OUT = os.path.join(ROOT, 'tests', 'golden')

# -------------------------------------------------------------------------------------------
# child scripts: the two reference directories both define `util`, so each runs in its own
# interpreter.
# -------------------------------------------------------------------------------------------

CHILD_FFT = r'''
import sys, json
import numpy as np
from unittest.mock import MagicMock
for m in ("tensorflow", "dxchange", "h5py", "matplotlib", "matplotlib.pyplot"):
    sys.modules[m] = MagicMock()
sys.path.insert(0, sys.argv[1] + "/tensorflow_recon")
import npfuncs, util
out = sys.argv[2]
sys.path.insert(0, sys.argv[3]); sys.path.insert(0, sys.argv[3] + "/oracle")
import multislice_oracle as mo

res = {}
# a-2: get_kernel on several grids (square, non-square, anisotropic voxels)
for name, args in {
    "k64": (1.0, 0.248, [1., 1., 1.], [64, 64, 64]),
    "k48x80": (1.0, 1240. / 800, [0.67, 0.67, 0.67], [48, 80, 8]),
    "kaniso": (2.5, 0.248, [1.0, 2.0, 2.5], [32, 40, 4]),
    "kfree": (1e-4 * 1e7, 0.248, [1., 1., 1.], [64, 64, 64]),
}.items():
    res["kernel_" + name] = util.get_kernel(*args)
    res["kernel_" + name + "_args"] = np.array(json.dumps(args))

def run(tag, gd, gb, pr, pi, energy, psize, free):
    psi = npfuncs.multislice_propagate_batch_numpy(gd, gb, pr, pi, energy, psize,
                                                   free_prop_cm=free, obj_batch_shape=gd.shape)
    res["psi_" + tag] = psi

# a-1 case A: the reference's own 64^3 delta fixture, beta = 0.1 delta, plane wave
gdf = np.load(sys.argv[1] + "/tensorflow_recon/grid_delta.npy")
vals, lab = np.unique(gdf, return_inverse=True)      # 3 distinct values: store as labels (12 KB)
res["fixture64_delta_values"] = vals
res["fixture64_delta_labels"] = lab.reshape(gdf.shape).astype(np.uint8)
run("fixture64", gdf[None], 0.1 * gdf[None], np.ones([64, 64]), np.zeros([64, 64]), 5000, 1e-7, None)
# case B: random phantom, batch 2, non-square 48x80x12, gaussian probe, 800 eV, 0.67 nm
gd, gb = mo.random_phantom((2, 48, 80, 12), seed=11, delta_scale=3e-4, beta_scale=3e-5)
pr, pi = mo.gaussian_probe((48, 80), 9., 9., 0.5)
run("rand48x80", gd.astype(np.float64), gb.astype(np.float64), pr, pi, 800, 0.67e-7, None)
# case C: far field
run("rand48x80_inf", gd.astype(np.float64), gb.astype(np.float64), pr, pi, 800, 0.67e-7, 'inf')
# case D: finite free-space distance
gd, gb = mo.random_phantom((1, 64, 64, 8), seed=12, delta_scale=1e-5, beta_scale=1e-6)
run("rand64_free", gd.astype(np.float64), gb.astype(np.float64), np.ones([64, 64]), np.zeros([64, 64]), 5000, 1e-7, 1e-4)
# case E: single slice
gd, gb = mo.random_phantom((3, 32, 32, 1), seed=13, delta_scale=1e-3, beta_scale=1e-4)
run("rand32_1slice", gd.astype(np.float64), gb.astype(np.float64), np.ones([32, 32]), np.zeros([32, 32]), 5000, 1e-7, None)
# case F: zone plate 128^2 x 20 (scaled config 1)
gd, gb = mo.zone_plate_phantom(n=128, n_slice=20, n_zones=8)
run("zp128", gd, gb, np.ones([128, 128]), np.zeros([128, 128]), 5000, 1e-7, None)
np.savez_compressed(out, **res)
'''

CHILD_CNN = r'''
import sys, types
import numpy as np
import scipy.signal
from unittest.mock import MagicMock
for m in ("tensorflow", "dxchange", "h5py", "matplotlib", "matplotlib.pyplot", "tqdm"):
    sys.modules[m] = MagicMock()
import tqdm
tqdm.trange = range
ag = types.ModuleType("autograd"); ag.numpy = np; ag.grad = lambda *a, **k: None
agn = np
ags = types.ModuleType("autograd.scipy"); agss = types.ModuleType("autograd.scipy.signal")
def convolve(A, B, axes=None, mode='valid'):
    assert axes == ([1, 2], [0, 1]) and mode == 'valid'
    return np.stack([scipy.signal.convolve(a, B, mode='valid', method='direct') for a in A])
agss.convolve = convolve; ags.signal = agss; ag.scipy = ags
sys.modules.update({"autograd": ag, "autograd.numpy": np, "autograd.numpy.random": np.random,
                    "autograd.scipy": ags, "autograd.scipy.signal": agss})
sys.path.insert(0, sys.argv[1] + "/cnn_propagator")
import propagation
out = sys.argv[2]
sys.path.insert(0, sys.argv[3] + "/oracle")
import multislice_oracle as mo
res = {}
gd, gb = mo.random_phantom((2, 32, 40, 6), seed=21, delta_scale=1e-4, beta_scale=1e-5)
for ks in (5, 17):
    for free in (None, 'inf'):
        psi = propagation.multislice_propagate_cnn(gd.astype(np.float64), gb.astype(np.float64),
                                                   np.ones([32, 40]), np.zeros([32, 40]), 5000,
                                                   [1e-7] * 3, kernel_size=ks, free_prop_cm=free)
        res["cnn_ks%d_%s" % (ks, free)] = psi
np.savez_compressed(out, **res)
'''


CHILD_ROT = r'''
import sys, types, os, tempfile
import numpy as np
from unittest.mock import MagicMock
for m in ("tensorflow", "dxchange", "h5py", "matplotlib", "matplotlib.pyplot", "tqdm"):
    sys.modules[m] = MagicMock()
ag = types.ModuleType("autograd"); ag.numpy = np; ag.grad = lambda *a, **k: None
sys.modules.update({"autograd": ag, "autograd.numpy": np, "autograd.numpy.random": np.random})
np.int = int                                   # util.py:327 uses the alias NumPy removed in 1.24
sys.path.insert(0, sys.argv[1] + "/cnn_propagator")
import util
out = sys.argv[2]
res = {}
os.chdir(tempfile.mkdtemp())
# 8f-1: rotation tables and their application (save_rotation_lookup / apply_rotation, util.py:295-402)
for tag, size, n_theta in (("a", [6, 16, 16], 7), ("b", [3, 12, 12], 5)):
    coords = util.save_rotation_lookup(size, n_theta, dest_folder="lk_" + tag)
    res["rot_%s_size" % tag] = np.array(size)
    res["rot_%s_coords" % tag] = np.stack(coords).astype(np.int32)
    rng = np.random.default_rng(61)
    obj = rng.standard_normal(size + [2])
    res["rot_%s_obj" % tag] = obj
    res["rot_%s_out" % tag] = np.stack([util.apply_rotation(obj, util.read_origin_coords("lk_" + tag, i), "lk_" + tag)
                                        for i in range(n_theta)])
# 8f-2: Adam (apply_gradient_adam, util.py:280-291), three consecutive updates
rng = np.random.default_rng(62)
x = rng.standard_normal((2, 5, 4, 3)) * 1e-6
m = v = None
xs = []
for i in range(3):
    g = rng.standard_normal(x.shape) * 1e-3
    res["adam_g%d" % i] = g
    if i == 0:
        res["adam_x0"] = x
        # the reference's first call trips over np.zeros_like(None) (util.py:285); it is only ever called with
        # m, v = None on the first minibatch, where zero moments are what was meant
        m = np.zeros_like(x); v = np.zeros_like(x)
    x, m, v = util.apply_gradient_adam(x, g, i, m, v, step_size=1e-7)
    xs.append(x)
res["adam_x"] = np.stack(xs); res["adam_m"] = m; res["adam_v"] = v
np.savez_compressed(out, **res)
'''


CHILD_NPF = r'''
import sys, types
import numpy as np
from unittest.mock import MagicMock
for m in ("tensorflow", "dxchange", "h5py", "matplotlib", "matplotlib.pyplot", "tqdm"):
    sys.modules[m] = MagicMock()
ag = types.ModuleType("autograd"); ag.numpy = np; ag.grad = lambda *a, **k: None
pf = types.ModuleType("pyfftw"); pfi = types.ModuleType("pyfftw.interfaces"); pfi.numpy_fft = np.fft; pf.interfaces = pfi
sys.modules.update({"autograd": ag, "autograd.numpy": np, "autograd.numpy.random": np.random,
                    "pyfftw": pf, "pyfftw.interfaces": pfi, "pyfftw.interfaces.numpy_fft": np.fft})
np.int = int
sys.path.insert(0, sys.argv[1] + "/cnn_propagator")
import np_funcs                              # cnn_propagator/np_funcs.py:15-65: returns (wavefront, probe_array)
out = sys.argv[2]
sys.path.insert(0, sys.argv[3] + "/oracle")
import multislice_oracle as mo
res = {}
gd, gb = mo.random_phantom((2, 32, 40, 5), seed=23, delta_scale=3e-4, beta_scale=3e-5)
pr, pi = mo.gaussian_probe((32, 40), 7., 7., 0.5)
for tag, free in (("none", None), ("inf", "inf"), ("free", 2e-6)):
    wf, pa = np_funcs.multislice_propagate_batch_numpy(gd.astype(np.float64), gb.astype(np.float64), pr, pi, 5000, 1e-7,
                                                       free_prop_cm=free, obj_batch_shape=gd.shape)
    res["npf_wavefront_" + tag] = wf
    if free is None:
        res["npf_probe_array"] = pa
np.savez_compressed(out, **res)
'''

CHILD_IR = r'''
import sys, json
import numpy as np
from unittest.mock import MagicMock
for m in ("tensorflow", "dxchange", "h5py", "matplotlib", "matplotlib.pyplot"):
    sys.modules[m] = MagicMock()
sys.path.insert(0, sys.argv[1] + "/tensorflow_recon")
import util                                   # get_kernel_ir, tensorflow_recon/util.py:188-216 (NumPy branch; the TIFF write is mocked)
out = sys.argv[2]
res = {}
for name, args in {
    "ir64": (1e-4 * 1e7, 0.248, [1., 1., 1.], [64, 64, 64]),
    "ir48x80": (3e-4 * 1e7, 1240. / 800, [0.67, 0.67, 0.67], [48, 80, 8]),
    "ir128_aniso": (2e-4 * 1e7, 0.248, [1.0, 2.0, 2.5], [128, 64, 4]),
}.items():
    res["kernel_" + name] = util.get_kernel_ir(*args)
    res["kernel_" + name + "_args"] = np.array(json.dumps(args))
np.savez_compressed(out, **res)
'''

CHILD_SIM = r'''
import sys, os, types, tempfile
import numpy as np
from unittest.mock import MagicMock
for m in ("tensorflow", "tensorflow.contrib", "tensorflow.contrib.image", "dxchange", "matplotlib", "matplotlib.pyplot"):
    sys.modules[m] = MagicMock()
# in-memory stand-in for h5py that keeps what the simulators write to exchange/data
STORE = {}
class _Dat:
    def __init__(self, shape, dtype): self.a = np.zeros(shape, dtype=dtype)
    def __setitem__(self, k, v): self.a[k] = v
class _Grp:
    def create_dataset(self, name, shape=None, dtype=None, **kw):
        d = _Dat(shape, dtype); STORE[name] = d; return d
class _File:
    def __init__(self, *a, **k): pass
    def create_group(self, name): return _Grp()
    def close(self): pass
h5 = types.ModuleType("h5py"); h5.File = _File
sys.modules["h5py"] = h5
sys.path.insert(0, sys.argv[1] + "/tensorflow_recon")
import simulation                      # tensorflow_recon/simulation.py, unmodified
out = sys.argv[2]
sys.path.insert(0, sys.argv[3] + "/oracle")
import multislice_oracle as mo
res = {}
tmp = tempfile.mkdtemp()
rng = np.random.default_rng(41)
Y, X, Z = 24, 24, 8
yy, xx, zz = np.meshgrid(np.arange(Y), np.arange(X), np.arange(Z), indexing="ij")
blob = ((yy - 11) ** 2 + (xx - 13) ** 2 + 4 * (zz - 4) ** 2 < 60) * 1.0 + ((yy - 6) ** 2 + (xx - 7) ** 2 < 9) * 0.5
gd = blob * 3e-4 * (1 + 0.2 * rng.random((Y, X, Z)))
gb = blob * 3e-5
np.save(os.path.join(tmp, "grid_delta.npy"), gd); np.save(os.path.join(tmp, "grid_beta.npy"), gb)
res["phantom_delta"] = gd; res["phantom_beta"] = gb
simulation.create_fullfield_data_numpy(5000, 1e-7, None, 4, tmp, tmp, "ff_plane.h5", batch_size=2, probe_type="plane",
                                       theta_st=0, theta_end=np.pi)
res["ff_plane"] = STORE["data"].a.copy()
simulation.create_fullfield_data_numpy(5000, 1e-7, 1e-4, 3, tmp, tmp, "ff_gauss.h5", batch_size=1, probe_type="gaussian",
                                       theta_st=0, theta_end=2 * np.pi, probe_mag_sigma=8., probe_phase_sigma=8., probe_phase_max=0.5)
res["ff_gauss_free"] = STORE["data"].a.copy()
pos = [(3, 3), (12, 12), (20, 22), (9, 15), (0, 23)]
simulation.create_ptychography_data_batch_numpy(5000, 1e-7, 2, tmp, tmp, "pty.h5", pos, probe_type="gaussian", probe_size=(18, 18),
                                                theta_st=0, theta_end=np.pi / 3, probe_circ_mask=None, minibatch_size=3,
                                                probe_mag_sigma=3., probe_phase_sigma=3., probe_phase_max=0.5)
res["pty_pos"] = np.array(pos)
res["pty"] = STORE["data"].a.copy()
np.savez_compressed(out, **res)
'''


def main():
    os.makedirs(OUT, exist_ok=True)
    only = sys.argv[1:]
    for name, code in (('ref_fft.npz', CHILD_FFT), ('ref_cnn.npz', CHILD_CNN), ('ref_rot.npz', CHILD_ROT), ('ref_npfuncs_cnn.npz', CHILD_NPF), ('ref_ir.npz', CHILD_IR), ('ref_sim.npz', CHILD_SIM)):
        if only and name not in only:
            continue
        path = os.path.join(OUT, name)
        subprocess.run([sys.executable, '-c', code, REF, path, ROOT], check=True)
        print('wrote', path, os.path.getsize(path), 'bytes')


if __name__ == '__main__':
    main()
