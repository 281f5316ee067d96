"""CPU tier: pin the oracle (oracle/multislice_oracle.py) against the golden vectors produced by
running the reference's own functions (oracle/gen_golden.py), the known-answer anchors recorded
in SURVEY.md 8c, and torch.autograd for the hand adjoint."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import multislice_oracle as mo
from conftest import rel_l2


@pytest.fixture(scope='module')
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, 'ref_fft.npz'))


@pytest.fixture(scope='module')
def gold_cnn(golden_dir):
    return np.load(os.path.join(golden_dir, 'ref_cnn.npz'))


def fixture64(gold):
    return gold['fixture64_delta_values'][gold['fixture64_delta_labels']]


@pytest.mark.parametrize('name', ['k64', 'k48x80', 'kaniso', 'kfree'])
def test_get_kernel_matches_reference(gold, name):
    args = json.loads(str(gold['kernel_%s_args' % name]))
    h = mo.get_kernel(*args)
    assert h.shape == gold['kernel_' + name].shape
    assert np.array_equal(h, gold['kernel_' + name])          # same float64 ops -> bit exact
    p0, hy, hx = mo.kernel_factors(*args)
    assert rel_l2(p0 * np.outer(hy, hx), h) < 1e-13            # separable to rounding; phases reach ~1e3 rad


def test_kernel_known_answers():
    # SURVEY.md 8c anchors
    h = mo.get_kernel(1.0, 0.248, [1, 1, 1], [64, 64, 64])
    assert abs(h[0, 0] - (0.9825897919540318 - 0.185788322420254j)) < 1e-14
    assert abs(h[32, 32] - (0.9795496939908097 + 0.20120237822280107j)) < 1e-14


def test_forward_fixture64(gold):
    gd = fixture64(gold)
    psi = mo.multislice_propagate_batch_numpy(gd[None], 0.1 * gd[None], np.ones([64, 64]), np.zeros([64, 64]),
                                              5000, 1e-7, None, (1, 64, 64, 64))
    assert rel_l2(psi, gold['psi_fixture64']) < 1e-14
    # SURVEY.md 8c anchors
    assert abs(np.sum(np.abs(psi) ** 2) - 4095.9518576500595) < 1e-8
    assert abs(psi[0, 0, 0] - (0.9807559047911851 + 0.19523787326342254j)) < 1e-12
    assert abs(psi[0, 32, 32] - (0.9807332762611662 + 0.1953300000993423j)) < 1e-12


def test_forward_random_cases(gold):
    gd, gb = mo.random_phantom((2, 48, 80, 12), seed=11, delta_scale=3e-4, beta_scale=3e-5)
    pr, pi = mo.gaussian_probe((48, 80), 9., 9., 0.5)
    for free, key in ((None, 'psi_rand48x80'), ('inf', 'psi_rand48x80_inf')):
        psi = mo.multislice_propagate_batch_numpy(gd.astype(np.float64), gb.astype(np.float64), pr, pi, 800,
                                                  0.67e-7, free, gd.shape)
        assert rel_l2(psi, gold[key]) < 1e-14
    gd, gb = mo.random_phantom((1, 64, 64, 8), seed=12, delta_scale=1e-5, beta_scale=1e-6)
    psi = mo.multislice_propagate_batch_numpy(gd.astype(np.float64), gb.astype(np.float64), np.ones([64, 64]),
                                              np.zeros([64, 64]), 5000, 1e-7, 1e-4, gd.shape)
    assert rel_l2(psi, gold['psi_rand64_free']) < 1e-14
    gd, gb = mo.random_phantom((3, 32, 32, 1), seed=13, delta_scale=1e-3, beta_scale=1e-4)
    psi = mo.multislice_propagate_batch_numpy(gd.astype(np.float64), gb.astype(np.float64), np.ones([32, 32]),
                                              np.zeros([32, 32]), 5000, 1e-7, None, gd.shape)
    assert rel_l2(psi, gold['psi_rand32_1slice']) < 1e-14
    gd, gb = mo.zone_plate_phantom(n=128, n_slice=20, n_zones=8)
    psi = mo.multislice_propagate_batch_numpy(gd, gb, np.ones([128, 128]), np.zeros([128, 128]), 5000, 1e-7,
                                              None, gd.shape)
    assert rel_l2(psi, gold['psi_zp128']) < 1e-14


def test_cnn_matches_reference(gold_cnn):
    gd, gb = mo.random_phantom((2, 32, 40, 6), seed=21, delta_scale=1e-4, beta_scale=1e-5)
    for ks in (5, 17):
        for free in (None, 'inf'):
            psi = mo.multislice_propagate_cnn(gd.astype(np.float64), gb.astype(np.float64), np.ones([32, 40]),
                                              np.zeros([32, 40]), 5000, [1e-7] * 3, kernel_size=ks, free_prop_cm=free)
            assert rel_l2(psi, gold_cnn['cnn_ks%d_%s' % (ks, free)]) < 1e-12


def test_shift_folding_identity():
    # ifft2(ifftshift(fftshift(fft2 u) H)) == ifft2(fft2(u) * ifftshift(H)), also for odd sizes
    rng = np.random.default_rng(0)
    for shape in ((16, 16), (9, 12), (15, 7)):
        u = rng.standard_normal((2,) + shape) + 1j * rng.standard_normal((2,) + shape)
        h = mo.get_kernel(1.0, 0.248, [1, 1, 1], list(shape) + [4])
        a = mo._propagate(u, h)
        b = np.fft.ifft2(np.fft.fft2(u) * np.fft.ifftshift(h))
        assert rel_l2(b, a) < 1e-14


def _torch_forward(gd, gb, probe, k, h, propagate_last, free, hfree):
    psi = probe[None].expand(gd.shape[0], -1, -1)
    n_slice = gd.shape[-1]
    hs = torch.fft.ifftshift(h)
    for i in range(n_slice):
        psi = psi * torch.exp(1j * k * gd[..., i]) * torch.exp(-k * gb[..., i])
        do_prop = (n_slice > 1) if propagate_last else (i < n_slice - 1)
        if do_prop:
            psi = torch.fft.ifft2(torch.fft.fft2(psi) * hs)
    if free == 'inf':
        psi = torch.fft.fftshift(torch.fft.fft2(psi), dim=(1, 2))
    elif free is not None:
        psi = torch.fft.ifft2(torch.fft.fft2(psi) * torch.fft.ifftshift(hfree))
    return psi


@pytest.mark.parametrize('propagate_last', [False, True])
@pytest.mark.parametrize('free', [None, 'inf', 1e-4])
def test_hand_adjoint_vs_torch_autograd(propagate_last, free):
    shape = (2, 16, 24, 5)
    gd, gb = mo.random_phantom(shape, seed=3, delta_scale=2e-3, beta_scale=2e-4)
    gd = gd.astype(np.float64); gb = gb.astype(np.float64)
    pr, pi = mo.gaussian_probe(shape[1:3], 5., 4., 0.7)
    energy, psize = 5000, 1e-7
    rng = np.random.default_rng(5)
    target = rng.random(shape[:3]) + 0.5
    loss, g_d, g_b, psi = mo.loss_and_grad(gd, gb, pr, pi, energy, psize, target, free_prop_cm=free,
                                           propagate_last=propagate_last)
    lmbda = 1240. / energy
    h = mo.get_kernel(1.0, lmbda, [1., 1., 1.], list(shape[1:]))
    hfree = mo.get_kernel(free * 1e7, lmbda, [1., 1., 1.], list(shape[1:])) if isinstance(free, float) else None
    k = 2 * mo.PI_TF * 1.0 / lmbda
    tgd = torch.tensor(gd, requires_grad=True); tgb = torch.tensor(gb, requires_grad=True)
    tpsi = _torch_forward(tgd, tgb, torch.tensor((pr + 1j * pi).astype(np.complex64).astype(np.complex128)), k, torch.tensor(h), propagate_last, free,
                          None if hfree is None else torch.tensor(hfree))
    tloss = torch.mean((tpsi.abs() - torch.tensor(target)) ** 2)
    tloss.backward()
    assert abs(loss - tloss.item()) < 1e-13 * max(1, abs(loss))
    assert rel_l2(psi, tpsi.detach().numpy()) < 1e-13
    assert rel_l2(g_d, tgd.grad.numpy()) < 1e-11
    assert rel_l2(g_b, tgb.grad.numpy()) < 1e-11


def test_adjoint_finite_difference():
    shape = (1, 8, 8, 3)
    gd, gb = mo.random_phantom(shape, seed=8, delta_scale=1e-2, beta_scale=1e-3)
    gd = gd.astype(np.float64); gb = gb.astype(np.float64)
    one, zero = np.ones(shape[1:3]), np.zeros(shape[1:3])
    target = np.full(shape[:3], 0.9)
    loss, g_d, g_b, _ = mo.loss_and_grad(gd, gb, one, zero, 5000, 1e-7, target)
    eps = 1e-7
    for idx in [(0, 1, 2, 0), (0, 5, 5, 1), (0, 7, 0, 2)]:
        for arr, g in ((gd, g_d), (gb, g_b)):
            a = arr.copy(); a[idx] += eps
            lp = mo.loss_and_grad(a if arr is gd else gd, a if arr is gb else gb, one, zero, 5000, 1e-7, target)[0]
            a[idx] -= 2 * eps
            lm = mo.loss_and_grad(a if arr is gd else gd, a if arr is gb else gb, one, zero, 5000, 1e-7, target)[0]
            fd = (lp - lm) / (2 * eps)
            assert abs(fd - g[idx]) < 1e-5 * max(abs(g[idx]), 1e-3)


def test_energy_conservation_and_vacuum():
    # beta = 0 and |H| = 1: energy is conserved; vacuum plane wave picks up exp(i k dz) per propagation
    gd, _ = mo.random_phantom((1, 32, 32, 6), seed=4, delta_scale=1e-3)
    psi = mo.multislice_propagate_batch_numpy(gd.astype(np.float64), np.zeros(gd.shape), np.ones([32, 32]),
                                              np.zeros([32, 32]), 5000, 1e-7, None, gd.shape)
    assert abs(np.sum(np.abs(psi) ** 2) - 32 * 32) < 1e-9
    z = np.zeros((1, 32, 32, 6))
    psi = mo.multislice_propagate_batch_numpy(z, z, np.ones([32, 32]), np.zeros([32, 32]), 5000, 1e-7, None, z.shape)
    assert np.allclose(np.abs(psi), 1.0, atol=1e-12)


def test_ptycho_windows_and_loss_shapes():
    rng = np.random.default_rng(2)
    obj = rng.random((24, 28, 3))
    pos = [(0, 0), (5, 7), (23, 27), (12, 14)]
    wins, pad = mo.ptycho_windows(obj, pos, (8, 8))
    assert wins.shape == (4, 8, 8, 3)
    assert np.array_equal(wins[1], obj[1:9, 3:11])
    assert np.all(wins[0][:4, :, :] == 0) and np.all(wins[0][:, :4, :] == 0)
    assert np.array_equal(wins[0][4:, 4:], obj[:4, :4])


# ---------------------------------------------------------------------------------------------
# rotation tables, their application and Adam, pinned by the reference's own functions
# (oracle/gen_golden.py CHILD_ROT runs cnn_propagator/util.py unmodified)
# ---------------------------------------------------------------------------------------------
@pytest.fixture(scope='module')
def gold_rot(golden_dir):
    return np.load(os.path.join(golden_dir, 'ref_rot.npz'))


@pytest.mark.parametrize('tag', ['a', 'b'])
def test_rotation_lookup_and_apply_match_reference(gold_rot, tag):
    size = [int(v) for v in gold_rot['rot_%s_size' % tag]]
    coords = gold_rot['rot_%s_coords' % tag]
    n_theta = coords.shape[0]
    obj = gold_rot['rot_%s_obj' % tag]
    for i, theta in enumerate(np.linspace(0, 2 * np.pi, n_theta)):        # util.py:321
        tab = mo.rotation_lookup(size, theta)
        assert np.array_equal(tab, coords[i])
        assert np.array_equal(mo.apply_rotation(obj, tab), gold_rot['rot_%s_out' % tag][i])


def test_rotation_adjoint_is_the_transpose():
    rng = np.random.default_rng(3)
    size = [4, 10, 10]
    tab = mo.rotation_lookup(size, 0.7)
    a = rng.standard_normal(size + [2])
    b = rng.standard_normal(size + [2])
    lhs = np.sum(mo.apply_rotation(a, tab) * b)
    rhs = np.sum(a * mo.apply_rotation_adjoint(b, tab))
    assert abs(lhs - rhs) < 1e-12 * max(1.0, abs(lhs))


def test_adam_matches_reference(gold_rot):
    x = gold_rot['adam_x0']
    m = v = None
    for i in range(3):
        x, m, v = mo.apply_gradient_adam(x, gold_rot['adam_g%d' % i], i, m, v, step_size=1e-7)
        assert np.array_equal(x, gold_rot['adam_x'][i])
    assert np.array_equal(m, gold_rot['adam_m']) and np.array_equal(v, gold_rot['adam_v'])


def test_bilinear_rotation_restatement_against_scipy_and_its_transpose():
    # tf.contrib.image.rotate cannot run here (TF 1.x): the restatement is checked against an independent implementation of
    # the same affine resampling (scipy, linear, zero fill with interpolation across the border) and <A x, y> == <x, A^T y>
    from scipy import ndimage
    rng = np.random.default_rng(91)
    for (Y, X, Z), theta in (((3, 16, 16), 0.4), ((2, 12, 20), 2.1), ((1, 9, 7), -1.0), ((2, 8, 8), 0.0)):
        obj = rng.random((Y, X, Z, 2))
        rot = mo.tf_rotate_bilinear(obj, theta)
        c, s = np.cos(theta), np.sin(theta)
        H, W = X, Z
        x_off = ((W - 1) - (c * (W - 1) - s * (H - 1))) / 2.0
        y_off = ((H - 1) - (s * (W - 1) + c * (H - 1))) / 2.0
        for y in range(Y):
            for ch in range(2):
                ref = ndimage.affine_transform(obj[y, :, :, ch], np.array([[c, s], [-s, c]]), offset=[y_off, x_off], order=1,
                                               mode='grid-constant', cval=0.0)
                assert np.abs(rot[y, :, :, ch] - ref).max() < 1e-12
        g = rng.standard_normal(obj.shape)
        lhs = np.sum(rot * g)
        rhs = np.sum(obj * mo.tf_rotate_bilinear_adjoint(g, theta))
        assert abs(lhs - rhs) < 1e-10 * max(1.0, abs(lhs))
    assert np.allclose(mo.tf_rotate_bilinear(obj, 0.0), obj)


def test_np_funcs_variant_matches_reference(golden_dir):
    # cnn_propagator/np_funcs.py:15-65 executed unmodified (oracle/gen_golden.py, CHILD_NPF): (wavefront, probe_array)
    g = np.load(os.path.join(golden_dir, 'ref_npfuncs_cnn.npz'))
    gd, gb = mo.random_phantom((2, 32, 40, 5), seed=23, delta_scale=3e-4, beta_scale=3e-5)
    pr, pi = mo.gaussian_probe((32, 40), 7., 7., 0.5)
    for tag, free in (('none', None), ('inf', 'inf'), ('free', 2e-6)):
        wf, pa = mo.multislice_propagate_batch_numpy_cnn(gd, gb, pr, pi, 5000, 1e-7, free_prop_cm=free, obj_batch_shape=gd.shape)
        assert rel_l2(wf, g['npf_wavefront_' + tag]) < 1e-13
        if free is None:
            assert pa.shape == (5, 2, 32, 40) and rel_l2(pa, g['npf_probe_array']) < 1e-13
            assert np.array_equal(pa[-1], wf)


def test_get_kernel_ir_matches_reference_golden(golden_dir):
    # tensorflow_recon/util.py:188-216 executed unmodified (oracle/gen_golden.py, CHILD_IR) vs the oracle restatement and the
    # host function the product uses for the IR free-space step of multislice_propagate
    import json
    from beyond_dof_b200.propagation import get_kernel_ir
    g = np.load(os.path.join(golden_dir, 'ref_ir.npz'))
    names = [k for k in g.files if not k.endswith('_args')]
    assert len(names) == 3
    for k in names:
        args = json.loads(str(g[k + '_args']))
        assert np.array_equal(mo.get_kernel_ir(*args), g[k])
        assert np.array_equal(get_kernel_ir(*args), g[k])


def test_unbatched_oracle_reduces_to_batched_tf_semantics():
    # where the sampling criterion picks 'TF' everywhere the un-batched function is the batched TF one with B = 1
    gd, gb = mo.random_phantom((1, 32, 48, 5), seed=3, delta_scale=3e-4, beta_scale=3e-5)
    pr, pi = mo.gaussian_probe((32, 48), 8., 8., 0.5)
    a = mo.multislice_propagate_unbatched(gd[0], gb[0], pr, pi, 5000, 1e-7, free_prop_cm='inf')
    b = mo.multislice_propagate_batch(gd, gb, pr, pi, 5000, 1e-7, free_prop_cm='inf')[0]
    assert rel_l2(a, b) < 1e-14
    # ... except for a single slice, which the un-batched function propagates (util.py:406-408 vs :484-488)
    a1 = mo.multislice_propagate_unbatched(gd[0, ..., :1], gb[0, ..., :1], pr, pi, 5000, 1e-7)
    b1 = mo.multislice_propagate_batch(gd[..., :1], gb[..., :1], pr, pi, 5000, 1e-7)[0]
    assert rel_l2(a1, b1) > 1e-3


@pytest.mark.parametrize('free', [None, 'inf', 2e-6])
def test_cnn_hand_adjoint_matches_torch_autograd(free):
    # a-5 gradient: the reference differentiates multislice_propagate_cnn with HIPS autograd (cnn_propagator/fullfield.py:329);
    # the oracle's hand adjoint (incl. the corner-pixel rescaling that couples the batch) against torch.autograd in complex128
    B, Y, X, Z, ks = 2, 12, 14, 3, 5
    psize = [1e-7] * 3
    gd, gb = mo.random_phantom((B, Y, X, Z), seed=5, delta_scale=2e-3, beta_scale=2e-4)
    gd, gb = gd.astype(np.float64), gb.astype(np.float64)
    pr, pi = mo.gaussian_probe((Y, X), 6., 6., 0.5)
    pr = pr + 0.5
    rng = np.random.default_rng(6)
    target = rng.random((B, Y, X)) * (10 if free == 'inf' else 1) + 0.2
    loss, g_d, g_b, psi = mo.cnn_loss_and_grad(gd, gb, pr, pi, 5000, psize, target, kernel_size=ks, free_prop_cm=free)
    assert rel_l2(psi, mo.multislice_propagate_cnn(gd, gb, pr, pi, 5000, psize, kernel_size=ks, free_prop_cm=free)) < 1e-13
    # torch restatement of propagation.py:77-128
    k = 2. * np.pi * 1.0 / (1240. / 5000)
    kern = torch.as_tensor(mo.cnn_kernel(5000, psize, np.array([Y, X, Z]), ks))
    pad = (ks - 1) // 2
    td = torch.tensor(gd, requires_grad=True)
    tb = torch.tensor(gb, requires_grad=True)
    probe = torch.as_tensor(np.tile(pr + 1j * pi, [B, 1, 1]))
    initial = probe[0, 0, 0]
    edge = torch.tensor(1.0 + 0j, dtype=torch.complex128)
    for i in range(Z):
        u = probe * torch.exp(torch.complex(-k * tb[..., i], k * td[..., i]))
        padded = edge * torch.ones((B, Y + 2 * pad, X + 2 * pad), dtype=torch.complex128)
        padded = padded.clone()
        padded[:, pad:pad + Y, pad:pad + X] = u
        out = torch.zeros_like(u)
        for a in range(ks):
            for b in range(ks):
                out = out + kern[a, b] * padded[:, 2 * pad - a:2 * pad - a + Y, 2 * pad - b:2 * pad - b + X]
        probe = out
        edge = kern.sum() * edge
    probe = probe * (initial / probe[0, 0, 0])
    if free == 'inf':
        probe = torch.fft.fftshift(torch.fft.fft2(probe), dim=(1, 2))
    elif free is not None:
        hf = torch.as_tensor(mo.get_kernel(free * 1e7, 1240. / 5000, np.array(psize) * 1e7, [Y, X, Z], pi=mo.PI_CNN))
        probe = torch.fft.ifft2(torch.fft.ifftshift(torch.fft.fftshift(torch.fft.fft2(probe), dim=(1, 2)) * hf, dim=(1, 2)))
    tl = ((probe.abs() - torch.as_tensor(target)) ** 2).mean()
    tl.backward()
    assert abs(tl.item() - loss) < 1e-12 * abs(loss)
    assert rel_l2(g_d, td.grad.numpy()) < 1e-10 and rel_l2(g_b, tb.grad.numpy()) < 1e-10
