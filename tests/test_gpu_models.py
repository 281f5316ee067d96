"""GPU tier: full-field and ptychography forward model + loss + gradient against the oracle, the real-space 'cnn'
propagator against the reference, and the steps either side of the hot path (rotation, Adam)."""
import os

import numpy as np
import pytest
import torch

from conftest import rel_l2
from oracle import multislice_oracle as mo

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def bd():
    import beyond_dof_b200 as pkg
    assert torch.cuda.is_available()
    return pkg


def test_fullfield_loss_and_grad(bd):
    Y, X, Z, B = 64, 64, 8, 3
    rng = np.random.default_rng(50)
    od = np.clip(rng.normal(8.7e-7, 1e-7, (Y, X, Z)), 0, None).astype(np.float32) * 100
    ob = np.clip(rng.normal(5.1e-8, 1e-8, (Y, X, Z)), 0, None).astype(np.float32) * 100
    gt_d, gt_b = mo.random_phantom((1, Y, X, Z), seed=51, delta_scale=2e-3, beta_scale=4e-3)   # strong contrast: O(1) misfit
    one, zero = np.ones((Y, X)), np.zeros((Y, X))
    prj = mo.multislice_propagate_batch(gt_d.astype(np.float64), gt_b.astype(np.float64), one, zero, 5000, 1e-7, free_prop_cm=1e-4)
    prj_b = np.repeat(prj, B, axis=0)
    lo, gdo, gbo, _ = mo.fullfield_loss_and_grad(od, ob, prj_b, one, zero, 5000, 1e-7, free_prop_cm=1e-4)
    loss, (g_d, g_b), ex = bd.fullfield_loss_and_grad(od, ob, np.zeros(B), prj_b, one, zero, 5000, 1e-7, free_prop_cm=1e-4)
    assert abs(loss.item() - lo) < 1e-5 * abs(lo)
    assert rel_l2(g_d.cpu().numpy(), gdo) < 1e-4 and rel_l2(g_b.cpu().numpy(), gbo) < 1e-4
    # regularisers (fullfield.py:389-396)
    loss2, (g_d2, _), _ = bd.fullfield_loss_and_grad(od, ob, np.zeros(B), prj_b, one, zero, 5000, 1e-7, free_prop_cm=1e-4,
                                                     alpha_d=1e-3, alpha_b=1e-4, gamma=1e-5)
    tod = torch.tensor(od, dtype=torch.float64, requires_grad=True)
    reg = 1e-3 * tod.abs().sum() + 1e-5 * sum((torch.roll(tod, 1, a) - tod).abs().sum() for a in range(3))
    reg.backward()
    assert abs((loss2 - loss).item() - (reg.item() + 1e-4 * np.abs(ob).sum())) < 1e-6 * reg.item()
    assert rel_l2((g_d2 - g_d).cpu().numpy(), tod.grad.numpy()) < 1e-5


def test_ptycho_loss_and_grad_with_padding(bd):
    Y, X, Z = 96, 112, 6
    probe_size = (64, 64)
    od, ob = mo.random_phantom((Y, X, Z), seed=60, delta_scale=3e-4, beta_scale=3e-5)
    gt_d, gt_b = mo.random_phantom((Y, X, Z), seed=61, delta_scale=5e-3, beta_scale=5e-3)          # strong contrast: O(1) misfit
    pr, pi = mo.gaussian_probe(probe_size, 6., 6., 0.5)
    # positions include windows that overhang every edge (zero padding, ptychography.py:45-61)
    pos = [(0, 0), (10, 100), (95, 111), (48, 56), (40, 40), (90, 5)]
    _, prj = mo.ptycho_loss(gt_d, gt_b, pos, np.zeros((len(pos),) + probe_size), pr, pi, probe_size, 5000, 1e-7)
    lo, gdo, gbo, _ = mo.ptycho_loss_and_grad(od, ob, pos, prj, pr, pi, probe_size, 5000, 1e-7)
    loss, (g_d, g_b) = bd.ptycho_loss_and_grad(od, ob, 0.0, pos, prj, pr, pi, probe_size, 5000, 1e-7)
    assert abs(loss.item() - lo) < 1e-5 * abs(lo)
    assert rel_l2(g_d.cpu().numpy(), gdo) < 1e-4 and rel_l2(g_b.cpu().numpy(), gbo) < 1e-4
    # oracle model head agrees with the literal restatement of the TF loss
    l2, _ = mo.ptycho_loss(od, ob, pos, prj, pr, pi, probe_size, 5000, 1e-7)
    assert abs(l2 - lo) < 1e-12 * abs(lo)


@pytest.mark.parametrize('probe_size', [(72, 72), (18, 18), (64, 64)])
def test_ptycho_rotated_object_and_reference_probe_sizes(bd, probe_size):
    # rotated object (apply_rotation before the window cut, cnn_propagator/ptychography.py:32-34) and the reference's own
    # probe sizes (72 x 72: reconstruct_ptycho.py; 18 x 18: the 4x down-sampled pass) -> mixed-radix line passes
    Y, X, Z = 100, 40, 40                         # the reference's rotation tables need a square (x, z) cross-section
    od, ob = mo.random_phantom((Y, X, Z), seed=64, delta_scale=3e-4, beta_scale=3e-5)
    gt_d, gt_b = mo.random_phantom((Y, X, Z), seed=65, delta_scale=5e-3, beta_scale=5e-3)
    pr, pi = mo.gaussian_probe(probe_size, 6., 6., 0.5)
    pos = [(2, 3), (50, 20), (99, 39), (30, 35), (71, 12)]
    theta = 0.7
    tab = mo.rotation_lookup([Y, X, Z], theta)

    def rot(a, b):
        r = mo.apply_rotation(np.stack([a, b], axis=3).astype(np.float64), tab)
        return r[..., 0], r[..., 1]
    _, prj = mo.ptycho_loss(*rot(gt_d, gt_b), pos, np.zeros((len(pos),) + probe_size), pr, pi, probe_size, 5000, 1e-7)
    lo, gdr, gbr, _ = mo.ptycho_loss_and_grad(*rot(od, ob), pos, prj, pr, pi, probe_size, 5000, 1e-7)
    go = mo.apply_rotation_adjoint(np.stack([gdr, gbr], axis=3), tab)
    loss, (g_d, g_b) = bd.ptycho_loss_and_grad(od, ob, theta, pos, prj, pr, pi, probe_size, 5000, 1e-7)
    assert abs(loss.item() - lo) < 1e-5 * abs(lo)
    assert rel_l2(g_d.cpu().numpy(), go[..., 0]) < 1e-4 and rel_l2(g_b.cpu().numpy(), go[..., 1]) < 1e-4


def test_ptycho_gradient_matches_torch_autograd_small(bd):
    # independent check of oracle + GPU against torch.autograd through a complex128 torch restatement
    Y, X, Z = 80, 80, 3
    probe_size = (64, 64)
    od, ob = mo.random_phantom((Y, X, Z), seed=62, delta_scale=1e-3, beta_scale=1e-4)
    pr, pi = mo.gaussian_probe(probe_size, 8., 8., 0.3)
    pos = [(32, 32), (40, 44), (47, 47)]
    rng = np.random.default_rng(63)
    prj = rng.random((3,) + probe_size) * 30
    tod = torch.tensor(od.astype(np.float64), requires_grad=True); tob = torch.tensor(ob.astype(np.float64), requires_grad=True)
    h = torch.tensor(np.fft.ifftshift(mo.get_kernel(1.0, 0.248, [1., 1., 1.], [64, 64, Z])))
    k = 2 * mo.PI_TF / 0.248
    probe = torch.tensor((pr + 1j * pi).astype(np.complex64).astype(np.complex128))
    loss_t = 0
    outs = []
    for (py, px) in pos:
        wd = tod[py - 32:py + 32, px - 32:px + 32]; wb = tob[py - 32:py + 32, px - 32:px + 32]
        psi = probe
        for i in range(Z):
            psi = psi * torch.exp(1j * k * wd[..., i]) * torch.exp(-k * wb[..., i])
            psi = torch.fft.ifft2(torch.fft.fft2(psi) * h)
        outs.append(torch.fft.fftshift(torch.fft.fft2(psi)))
    ex = torch.stack(outs)
    loss_t = torch.mean((ex.abs() - torch.tensor(prj)) ** 2) * len(pos)
    loss_t.backward()
    loss, (g_d, g_b) = bd.ptycho_loss_and_grad(od, ob, 0.0, pos, prj, pr, pi, probe_size, 5000, 1e-7)
    assert abs(loss.item() - loss_t.item()) < 1e-5 * abs(loss_t.item())
    assert rel_l2(g_d.cpu().numpy(), tod.grad.numpy()) < 1e-4 and rel_l2(g_b.cpu().numpy(), tob.grad.numpy()) < 1e-4


def test_cnn_propagator_matches_reference(bd, golden_dir):
    gold = np.load(os.path.join(golden_dir, 'ref_cnn.npz'))
    gd, gb = mo.random_phantom((2, 32, 40, 6), seed=21, delta_scale=1e-4, beta_scale=1e-5)
    for ks in (5, 17):
        for free in (None, 'inf'):
            psi = bd.multislice_propagate_cnn(gd, gb, np.ones([32, 40]), np.zeros([32, 40]), 5000, [1e-7] * 3,
                                              kernel_size=ks, free_prop_cm=free)
            ref = gold['cnn_ks%d_%s' % (ks, free)]
            assert rel_l2(np.abs(psi) ** 2, np.abs(ref) ** 2) < 1e-5
            assert rel_l2(psi, ref) < 1e-5


# ---------------------------------------------------------------------------------------------
# rotation (nearest-neighbour table of the cnn_propagator drivers), its transpose, Adam
# ---------------------------------------------------------------------------------------------
@pytest.fixture(scope='module')
def gold_rot(golden_dir):
    return np.load(os.path.join(golden_dir, 'ref_rot.npz'))


@pytest.mark.parametrize('tag', ['a', 'b'])
def test_apply_rotation_matches_reference_bit_exact(bd, gold_rot, tag):
    from beyond_dof_b200 import rotation
    size = [int(v) for v in gold_rot['rot_%s_size' % tag]]
    coords = gold_rot['rot_%s_coords' % tag]
    obj = gold_rot['rot_%s_obj' % tag].astype(np.float32)
    for i, theta in enumerate(np.linspace(0, 2 * np.pi, coords.shape[0])):
        assert np.array_equal(rotation.rotation_table(size, theta), coords[i])          # host table == save_rotation_lookup
        out = rotation.apply_rotation(obj, coords[i], 'ignored')
        assert out.dtype == np.float32 and np.array_equal(out, gold_rot['rot_%s_out' % tag][i].astype(np.float32))


def test_rotation_adjoint_matches_oracle_ragged(bd):
    from beyond_dof_b200 import rotation
    rng = np.random.default_rng(71)
    for (Y, X, Z), theta in (((5, 33, 33), 0.3), ((3, 64, 64), 2.1), ((2, 40, 40), 4.0), ((1, 8, 8), 0.0), ((11, 160, 160), 0.7), ((9, 130, 130), 1.55), ((7, 101, 101), 2.4)):
        g = rng.standard_normal((Y, X, Z, 2)).astype(np.float32)
        tab = mo.rotation_lookup([Y, X, Z], theta)
        ref = mo.apply_rotation_adjoint(g.astype(np.float64), tab)
        dev_tab = rotation.device_table([Y, X, Z], theta, torch.device('cuda'))
        g_db = torch.as_tensor(g).cuda().permute(2, 0, 1, 3).contiguous()
        for atomic in (False, True):                                                # inverse-list gather / atomic scatter-add
            acc = torch.zeros_like(g_db)
            rotation.rotate_db_adjoint(g_db, dev_tab, acc, atomic=atomic)
            rotation.rotate_db_adjoint(g_db, dev_tab, acc, atomic=atomic)          # accumulates
            assert rel_l2(acc.permute(1, 2, 0, 3).cpu().numpy(), 2 * ref) < 1e-6
        rot = rotation.rotate_db(g_db, dev_tab)
        assert np.array_equal(rot.permute(1, 2, 0, 3).cpu().numpy(), mo.apply_rotation(g, tab))
    # the reference only rotates [dim_y, dim_x, dim_x] objects: other shapes accept the identity only
    assert np.array_equal(rotation.rotation_table([2, 17, 40], 0.0)[:, 0], np.repeat(np.arange(17), 40))
    with pytest.raises(ValueError):
        rotation.rotation_table([2, 17, 40], 0.5)


def test_rotation_adjoint_batch_sums_the_minibatch_in_one_pass(bd):
    """bdof_rotate_adjoint_csr_batch == the per-angle transposes summed; overwrite mode needs no zero-fill; more angles than
    BDOF_ROT_MAX_ANGLES are chunked; bit-reproducible."""
    from beyond_dof_b200 import rotation
    rng = np.random.default_rng(73)
    dev = torch.device('cuda')
    for (Y, X, Z), n_ang in (((5, 33, 33), 3), ((19, 40, 40), 18), ((8, 64, 64), 10), ((6, 150, 150), 5), ((3, 97, 97), 4)):
        thetas = rng.random(n_ang) * 2 * np.pi
        tabs = [rotation.device_table([Y, X, Z], float(t), dev) for t in thetas]
        g = torch.as_tensor(rng.standard_normal((Z, n_ang, Y, X, 2)).astype(np.float32)).cuda()
        ref = torch.zeros((Z, Y, X, 2), device='cuda')
        for b in range(n_ang):
            rotation.rotate_db_adjoint(g[:, b], tabs[b], ref)
        out = torch.full((Z, Y, X, 2), float('nan'), device='cuda')
        rotation.rotate_db_adjoint_batch(g, tabs, out, accumulate=False)
        assert not torch.isnan(out).any()
        assert rel_l2(out.cpu().numpy(), ref.cpu().numpy()) < 1e-6         # same terms, one running sum instead of per-angle sums
        out2 = torch.empty_like(out)
        rotation.rotate_db_adjoint_batch(g, tabs, out2, accumulate=False)
        assert torch.equal(out, out2)                                      # deterministic
        again = out.clone()
        rotation.rotate_db_adjoint_batch(g, tabs, again, accumulate=True)
        assert rel_l2(again.cpu().numpy(), 2 * ref.cpu().numpy()) < 1e-6
        o64 = sum(mo.apply_rotation_adjoint(g[:, b].permute(1, 2, 0, 3).cpu().numpy().astype(np.float64), mo.rotation_lookup([Y, X, Z], float(thetas[b])))
                  for b in range(n_ang))
        assert rel_l2(out.permute(1, 2, 0, 3).cpu().numpy(), o64) < 1e-6


def test_adam_step_matches_reference(bd, gold_rot):
    from beyond_dof_b200 import rotation
    x = torch.as_tensor(gold_rot['adam_x0'].astype(np.float32)).cuda()
    m = v = None
    for i in range(3):
        g = torch.as_tensor(gold_rot['adam_g%d' % i].astype(np.float32)).cuda()
        x, m, v = rotation.adam_step(x, g, i, m, v, step_size=1e-7)
        assert rel_l2(x.cpu().numpy(), gold_rot['adam_x'][i]) < 2e-6                      # fp32 vs the reference's float64
    assert rel_l2(m.cpu().numpy(), gold_rot['adam_m']) < 1e-6 and rel_l2(v.cpu().numpy(), gold_rot['adam_v']) < 1e-6


def test_fullfield_loss_and_grad_with_rotation(bd):
    Y, X, Z, B = 64, 64, 64, 3
    rng = np.random.default_rng(72)
    od = (rng.random((Y, X, Z)) * 4e-4).astype(np.float32)
    ob = (rng.random((Y, X, Z)) * 4e-5).astype(np.float32)
    theta = np.array([0.0, 0.6, 2.5])
    one, zero = np.ones((Y, X)), np.zeros((Y, X))
    target = rng.random((B, Y, X)) + 0.5
    lo, gdo, gbo, psio = mo.tomo_loss_and_grad(od, ob, theta, target, one, zero, 5000, 1e-7, free_prop_cm=None, propagate_last=True)
    loss, (g_d, g_b), ex = bd.fullfield_loss_and_grad(od, ob, theta, target, one, zero, 5000, 1e-7, propagate_last=True)
    assert rel_l2(np.abs(ex.cpu().numpy()) ** 2, np.abs(psio) ** 2) < 1e-5
    assert abs(loss.item() - lo) < 1e-5 * abs(lo)
    assert rel_l2(g_d.cpu().numpy(), gdo) < 1e-4 and rel_l2(g_b.cpu().numpy(), gbo) < 1e-4


def test_tomography_objective_descends(bd):
    # a few Adam steps of the full loop (rotate -> multislice -> loss -> adjoint -> back-rotate -> Adam) reduce the loss
    from beyond_dof_b200.models import TomographyObjective, pack_object
    Y = X = Z = 64
    rng = np.random.default_rng(73)
    gt_d = (rng.random((Y, X, Z)) * 2e-3).astype(np.float32)
    gt_b = (rng.random((Y, X, Z)) * 2e-4).astype(np.float32)
    theta = np.linspace(0, np.pi, 4)
    one, zero = np.ones((Y, X)), np.zeros((Y, X))
    obj = np.stack([gt_d, gt_b], axis=3).astype(np.float64)
    rot = np.stack([mo.apply_rotation(obj, mo.rotation_lookup([Y, X, Z], t)) for t in theta])
    prj = np.abs(mo.multislice_propagate_batch(rot[..., 0], rot[..., 1], one, zero, 5000, 1e-7)).astype(np.float32)
    start = pack_object(np.zeros_like(gt_d), np.zeros_like(gt_b))
    probe = torch.ones((Y, X), dtype=torch.complex64, device='cuda')
    tomo = TomographyObjective(start, probe, 5000, 1e-7, minibatch_size=4, step_size=2e-5)
    losses = [tomo.step(theta, torch.as_tensor(prj)) for _ in range(12)]
    assert losses[-1] < 0.5 * losses[0]


def test_finite_support_clip_and_shrink_wrap(bd):
    # obj <- clip(obj * mask, 0); mask <- mask * (delta > 1e-15)   (cnn_propagator/fullfield.py:359-368)
    from beyond_dof_b200 import rotation
    rng = np.random.default_rng(81)
    obj = (rng.standard_normal((5, 7, 9, 2)) * 1e-5).astype(np.float32)
    mask = (rng.random((5, 7, 9)) > 0.3).astype(np.float32)
    x = torch.as_tensor(obj).cuda()
    m = torch.as_tensor(mask).cuda()
    rotation.finite_support(x, m, shrink_threshold=1e-15)
    ref = np.clip(obj * mask[..., None], 0, None)
    assert np.array_equal(x.cpu().numpy(), ref)
    assert np.array_equal(m.cpu().numpy(), mask * (ref[..., 0] > 1e-15))
    y = torch.as_tensor(obj).cuda()
    rotation.finite_support(y)                       # clip only
    assert np.array_equal(y.cpu().numpy(), np.clip(obj, 0, None))


def test_ptycho_position_losses_and_dynamic_dropping(bd):
    Y, X, Z = 96, 96, 4
    probe_size = (72, 72)
    od, ob = mo.random_phantom((Y, X, Z), seed=82, delta_scale=3e-4, beta_scale=3e-5)
    pr, pi = mo.gaussian_probe(probe_size, 6., 6., 0.5)
    pos = [(40, 40), (10, 80), (60, 50), (48, 48)]
    _, prj = mo.ptycho_loss(od, ob, pos, np.zeros((len(pos),) + probe_size), pr, pi, probe_size, 5000, 1e-7)
    prj = np.abs(prj)
    prj[1] *= 1.5                                    # positions 1 and 2 do not fit the data
    prj[2] += 0.3
    table = bd.ptycho_position_losses(od, ob, 0.0, pos, prj, pr, pi, probe_size, 5000, 1e-7, n_dp_batch=3)
    ref = [mo.ptycho_loss(od, ob, [p], prj[j:j + 1], pr, pi, probe_size, 5000, 1e-7, scale_by_npos=False)[0] for j, p in enumerate(pos)]
    t = table.cpu().numpy()
    assert np.all(np.abs(t[1:3] - np.array(ref[1:3])) < 1e-4 * np.array(ref[1:3]))
    assert t[0] < 1e-6 * t[1] and t[3] < 1e-6 * t[1]            # exact fits: only fp32 rounding left
    assert list(bd.dynamic_dropping(table, dropping_threshold=1e-3 * t[1])) == [1, 2]


def test_tomography_objective_regularisers_and_support(bd):
    # L1 + TV terms (tensorflow_recon/fullfield.py:389-396) and the finite-support mask in the update loop
    from beyond_dof_b200.models import TomographyObjective, pack_object, unpack_object
    Y = X = Z = 64
    rng = np.random.default_rng(83)
    od = (rng.random((Y, X, Z)) * 4e-4).astype(np.float32)
    ob = (rng.random((Y, X, Z)) * 4e-5).astype(np.float32)
    theta = np.array([0.0, 1.1])
    one, zero = np.ones((Y, X)), np.zeros((Y, X))
    target = (rng.random((2, Y, X)) + 0.5).astype(np.float32)
    lo, gdo, gbo, _ = mo.tomo_loss_and_grad(od, ob, theta, target, one, zero, 5000, 1e-7, free_prop_cm=None, propagate_last=True)
    a_d, a_b, gam = 1e-2, 2e-2, 3e-3
    tv = sum(np.abs(np.roll(od.astype(np.float64), 1, ax) - od).sum() for ax in range(3))
    lo_reg = lo + a_d * np.abs(od).sum() + a_b * np.abs(ob).sum() + gam * tv
    tod = torch.tensor(od.astype(np.float64), requires_grad=True)
    (a_d * tod.abs().sum() + gam * sum((torch.roll(tod, 1, ax) - tod).abs().sum() for ax in range(3))).backward()
    probe = torch.ones((Y, X), dtype=torch.complex64, device='cuda')
    mask = torch.zeros((Z, Y, X), device='cuda'); mask[:, 8:56, 8:56] = 1
    tomo = TomographyObjective(pack_object(od, ob), probe, 5000, 1e-7, minibatch_size=2, step_size=1e-6, mask=mask,
                               alpha_d=a_d, alpha_b=a_b, gamma=gam)
    tgt = torch.as_tensor(target).cuda()
    loss = tomo.loss_and_grad(theta, tgt)
    assert abs(float(loss) - lo_reg) < 1e-5 * abs(lo_reg)
    g_d, g_b = unpack_object(tomo.grad)
    assert rel_l2(g_d.cpu().numpy(), gdo + tod.grad.numpy()) < 1e-4
    assert rel_l2(g_b.cpu().numpy(), gbo + a_b * np.sign(ob)) < 1e-4
    tomo.step(theta, tgt)
    d, b = unpack_object(tomo.obj)
    outside = np.ones((Y, X, Z), bool); outside[8:56, 8:56, :] = False
    assert np.all(d.cpu().numpy()[outside] == 0) and np.all(b.cpu().numpy()[outside] == 0)
    assert d.min().item() >= 0 and b.min().item() >= 0


@pytest.mark.parametrize('shape,theta', [((5, 24, 24), 0.4), ((3, 20, 36), 2.1), ((2, 33, 17), -1.0), ((4, 16, 16), 0.0)])
def test_bilinear_rotation_and_transpose_match_restatement(bd, shape, theta):
    # tf.contrib.image.rotate(..., 'BILINEAR') of the TF drivers (tensorflow_recon/fullfield.py:96): GPU kernels against the
    # oracle restatement (itself checked against scipy in tests/test_oracle.py; TF 1.x cannot run here)
    from beyond_dof_b200 import rotation
    Y, X, Z = shape
    rng = np.random.default_rng(92)
    obj = rng.random((Y, X, Z, 2)).astype(np.float32)
    db = torch.as_tensor(obj).cuda().permute(2, 0, 1, 3).contiguous()                 # [Z, Y, X, 2]
    batch = torch.zeros((Z, 3, Y, X, 2), device='cuda')
    rotation.rotate_db_bilinear(db, theta, out=batch[:, 1])                            # strided batch-element view
    got = batch[:, 1].permute(1, 2, 0, 3).cpu().numpy()
    ref = mo.tf_rotate_bilinear(obj.astype(np.float64), theta)
    assert np.abs(got - ref).max() < 2e-5                                             # fp32 coordinates: weights differ by ~1e-6 * size
    assert float(batch[:, 0].abs().max()) == 0 and float(batch[:, 2].abs().max()) == 0
    g = rng.standard_normal((Y, X, Z, 2)).astype(np.float32)
    batch[:, 1] = torch.as_tensor(g).cuda().permute(2, 0, 1, 3)
    acc = torch.ones((Z, Y, X, 2), device='cuda')
    rotation.rotate_db_bilinear_adjoint(batch[:, 1], theta, acc)                       # accumulates
    gref = mo.tf_rotate_bilinear_adjoint(g.astype(np.float64), theta) + 1.0
    assert np.abs(acc.permute(1, 2, 0, 3).cpu().numpy() - gref).max() < 1e-4


def test_fullfield_loss_and_grad_with_bilinear_rotation(bd):
    Y, X, Z, B = 64, 64, 64, 3
    rng = np.random.default_rng(93)
    od = (rng.random((Y, X, Z)) * 4e-4).astype(np.float32)
    ob = (rng.random((Y, X, Z)) * 4e-5).astype(np.float32)
    theta = np.array([0.0, 0.6, 2.5])
    one, zero = np.ones((Y, X)), np.zeros((Y, X))
    target = rng.random((B, Y, X)) + 0.5
    lo, gdo, gbo, psio = mo.tomo_loss_and_grad(od, ob, theta, target, one, zero, 5000, 1e-7, free_prop_cm=None, propagate_last=True,
                                               rotation='bilinear')
    loss, (g_d, g_b), ex = bd.fullfield_loss_and_grad(od, ob, theta, target, one, zero, 5000, 1e-7, propagate_last=True,
                                                      rotation='bilinear')
    assert rel_l2(np.abs(ex.cpu().numpy()) ** 2, np.abs(psio) ** 2) < 1e-5
    assert abs(loss.item() - lo) < 1e-5 * abs(lo)
    assert rel_l2(g_d.cpu().numpy(), gdo) < 1e-4 and rel_l2(g_b.cpu().numpy(), gbo) < 1e-4


def test_ptycho_with_bilinear_rotation(bd):
    # TF ptychography driver: tf_rotate(..., 'BILINEAR') before the window cut (tensorflow_recon/ptychography.py:39); the
    # bilinear rotation has no square-cross-section restriction
    Y, X, Z = 90, 80, 6
    probe_size = (64, 64)
    od, ob = mo.random_phantom((Y, X, Z), seed=94, delta_scale=3e-4, beta_scale=3e-5)
    gt_d, gt_b = mo.random_phantom((Y, X, Z), seed=95, delta_scale=5e-3, beta_scale=5e-3)
    pr, pi = mo.gaussian_probe(probe_size, 6., 6., 0.5)
    pos = [(3, 5), (45, 40), (89, 79), (30, 60)]
    theta = 0.3

    def rot(a, b):
        r = mo.tf_rotate_bilinear(np.stack([a, b], axis=3).astype(np.float64), theta)
        return r[..., 0], r[..., 1]
    _, prj = mo.ptycho_loss(*rot(gt_d, gt_b), pos, np.zeros((len(pos),) + probe_size), pr, pi, probe_size, 5000, 1e-7)
    lo, gdr, gbr, _ = mo.ptycho_loss_and_grad(*rot(od, ob), pos, prj, pr, pi, probe_size, 5000, 1e-7)
    go = mo.tf_rotate_bilinear_adjoint(np.stack([gdr, gbr], axis=3), theta)
    loss, (g_d, g_b) = bd.ptycho_loss_and_grad(od, ob, theta, pos, prj, pr, pi, probe_size, 5000, 1e-7, rotation='bilinear')
    assert abs(loss.item() - lo) < 1e-5 * abs(lo)
    assert rel_l2(g_d.cpu().numpy(), go[..., 0]) < 1e-4 and rel_l2(g_b.cpu().numpy(), go[..., 1]) < 1e-4


def test_patch_gather_add_is_the_deterministic_scatter(bd):
    # the ordered gather accumulates exactly what the atomic scatter-add does, bit-reproducibly
    import ctypes
    from beyond_dof_b200.capi import lib, check
    from beyond_dof_b200.plan import _ptr
    Z, OY, OX, n, py, px = 5, 70, 150, 37, 24, 40
    g = torch.Generator(device='cuda').manual_seed(3)
    patches = torch.randn((Z, n, py, px, 2), device='cuda', generator=g)
    pos = torch.stack([torch.randint(-10, OY - 5, (n,), generator=torch.Generator().manual_seed(4)),
                       torch.randint(-20, OX - 5, (n,), generator=torch.Generator().manual_seed(5))], 1).to(torch.int32).cuda().contiguous()
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    ref = torch.zeros((Z, OY, OX, 2), dtype=torch.float64, device='cuda')
    for i in range(n):                                             # plain accumulation in float64
        y0, x0 = int(pos[i, 0]), int(pos[i, 1])
        ys, xs = slice(max(y0, 0), min(y0 + py, OY)), slice(max(x0, 0), min(x0 + px, OX))
        ref[:, ys, xs] += patches[:, i, ys.start - y0:ys.stop - y0, xs.start - x0:xs.stop - x0].double()
    outs = []
    for _ in range(3):
        out = torch.zeros((Z, OY, OX, 2), dtype=torch.float32, device='cuda')
        check(lib.bdof_patch_gather_add(_ptr(patches), Z, OY, OX, _ptr(pos), n, py, px, _ptr(out), st))
        outs.append(out)
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
    assert rel_l2(outs[0].cpu().numpy(), ref.cpu().numpy()) < 2e-7
    # accumulates (+=) and agrees with the atomic version
    at = torch.zeros_like(outs[0])
    check(lib.bdof_patch_scatter_add(_ptr(patches), Z, OY, OX, _ptr(pos), n, py, px, _ptr(at), st))
    assert rel_l2(at.cpu().numpy(), outs[0].cpu().numpy()) < 1e-6
    check(lib.bdof_patch_gather_add(_ptr(patches), Z, OY, OX, _ptr(pos), n, py, px, _ptr(outs[0]), st))
    assert rel_l2(outs[0].cpu().numpy(), 2 * ref.cpu().numpy()) < 2e-7


@pytest.mark.parametrize('probe_size', [(64, 64), (72, 72)])
def test_ptychography_objective_step_matches_oracle_update(bd, probe_size):
    # one update of the reconstruction loop (cnn_propagator/ptychography.py:292-310) against the oracle gradient + Adam;
    # 64 x 64: the resident kernels reading the object through the windows, 72 x 72 (the reference's default): cut windows
    from beyond_dof_b200.models import PtychographyObjective, pack_object, unpack_object
    Y, X, Z = 96, 96, 6
    od, ob = mo.random_phantom((Y, X, Z), seed=90, delta_scale=3e-4, beta_scale=3e-5)
    gt_d, gt_b = mo.random_phantom((Y, X, Z), seed=91, delta_scale=5e-3, beta_scale=5e-3)
    pr, pi = mo.gaussian_probe(probe_size, 6., 6., 0.5)
    pos = np.array([(32, 32), (38, 50), (60, 40), (48, 48), (30, 64), (64, 64), (0, 0), (90, 95)])
    _, prj = mo.ptycho_loss(gt_d, gt_b, pos, np.zeros((len(pos),) + probe_size), pr, pi, probe_size, 5000, 1e-7)
    lo, gdo, gbo, _ = mo.ptycho_loss_and_grad(od, ob, pos, prj, pr, pi, probe_size, 5000, 1e-7)
    obj = pack_object(od, ob)
    pty = PtychographyObjective(obj, torch.as_tensor((pr + 1j * pi).astype(np.complex64)), probe_size, 5000, 1e-7,
                                n_pos_per_step=len(pos), step_size=1e-7)
    assert pty.windowed == (probe_size == (64, 64))
    loss = pty.step(pos, torch.as_tensor(np.abs(prj).astype(np.float32)).pin_memory())
    assert abs(loss - lo) < 2e-5 * abs(lo)
    g_d, g_b = unpack_object(pty.grad)
    assert rel_l2(g_d.cpu().numpy(), gdo) < 1e-4 and rel_l2(g_b.cpu().numpy(), gbo) < 1e-4
    # the update is Adam on THAT gradient (first step: x - step g / (|g| + eps), which is ill-conditioned in g where |g| ~ eps, so
    # the expected update is formed from the gradient the GPU produced) followed by the clip
    xd, _, _ = mo.apply_gradient_adam(od.astype(np.float64), g_d.cpu().numpy().astype(np.float64), 0, None, None, step_size=1e-7)
    xb, _, _ = mo.apply_gradient_adam(ob.astype(np.float64), g_b.cpu().numpy().astype(np.float64), 0, None, None, step_size=1e-7)
    nd, nb = unpack_object(pty.obj)
    assert rel_l2(nd.cpu().numpy(), np.clip(xd, 0, None)) < 2e-6 and rel_l2(nb.cpu().numpy(), np.clip(xb, 0, None)) < 2e-6
    # bit-reproducible: the same step from the same state gives the same gradient
    obj2 = pack_object(od, ob)
    pty2 = PtychographyObjective(obj2, torch.as_tensor((pr + 1j * pi).astype(np.complex64)), probe_size, 5000, 1e-7,
                                 n_pos_per_step=len(pos), step_size=1e-7)
    pty2.step(pos, torch.as_tensor(np.abs(prj).astype(np.float32)).pin_memory())
    assert torch.equal(pty.grad, pty2.grad) and torch.equal(pty.obj, pty2.obj)


@pytest.mark.parametrize('ks', [5, 17])
@pytest.mark.parametrize('free', [None, 'inf', 2e-6])
def test_cnn_propagator_gradient_matches_oracle(bd, ks, free):
    # a-5 gradient (autograd.grad(calculate_loss) in cnn_propagator/fullfield.py:329): CUDA adjoint of the real-space chain,
    # incl. the corner-pixel rescaling, against the oracle's hand adjoint (itself pinned to torch.autograd)
    shape = (2, 40, 48, 6)
    gd, gb = mo.random_phantom(shape, seed=31, delta_scale=1e-3, beta_scale=1e-4)
    pr, pi = mo.gaussian_probe(shape[1:3], 16., 14., 0.5)
    pr = pr + 0.5
    rng = np.random.default_rng(32)
    target = rng.random(shape[:3]) * (20 if free == 'inf' else 1) + 0.3
    lo, gdo, gbo, psio = mo.cnn_loss_and_grad(gd, gb, pr, pi, 5000, [1e-7] * 3, target, kernel_size=ks, free_prop_cm=free)
    loss, (g_d, g_b), psi = bd.cnn_loss_and_grad(gd, gb, pr, pi, 5000, [1e-7] * 3, target, kernel_size=ks, free_prop_cm=free)
    assert rel_l2(psi.cpu().numpy(), psio) < 1e-5
    assert abs(loss.item() - lo) < 2e-5 * abs(lo)
    assert rel_l2(g_d, gdo) < 1e-4 and rel_l2(g_b, gbo) < 1e-4


def test_cnn_propagator_debug_returns_slice_magnitudes(bd):
    # propagation.py:107,130: debug=True returns (probe, [|probe| after every slice], seconds)
    shape = (2, 32, 40, 4)
    gd, gb = mo.random_phantom(shape, seed=33, delta_scale=1e-3, beta_scale=1e-4)
    one, zero = np.ones(shape[1:3]), np.zeros(shape[1:3])
    res, pa, secs = bd.multislice_propagate_cnn(gd, gb, one, zero, 5000, [1e-7] * 3, kernel_size=5, debug=True)
    plain = bd.multislice_propagate_cnn(gd, gb, one, zero, 5000, [1e-7] * 3, kernel_size=5)
    assert rel_l2(res, plain) < 1e-6 and len(pa) == 4 and pa[0].shape == shape[:3] and secs > 0
    # the magnitude after the last slice, rescaled like the result, is |result|
    assert rel_l2(pa[-1] * abs(1.0 / (res[0, 0, 0] / plain[0, 0, 0])) * np.abs(plain[0, 0, 0]) / pa[-1][0, 0, 0], np.abs(plain)) < 1e-5


def test_simulators_match_reference_outputs(bd, golden_dir, tmp_path):
    # create_fullfield_data_numpy / create_ptychography_data_batch_numpy (tensorflow_recon/simulation.py:80-161,
    # 283-386) executed unmodified in the build container (oracle/gen_golden.py CHILD_SIM, in-memory stand-in for h5py) vs the
    # drop-ins, which rotate with the same scipy call and project on the GPU
    from beyond_dof_b200 import simulation
    g = np.load(os.path.join(golden_dir, 'ref_sim.npz'))
    ph = str(tmp_path)
    np.save(os.path.join(ph, 'grid_delta.npy'), g['phantom_delta'])
    np.save(os.path.join(ph, 'grid_beta.npy'), g['phantom_beta'])
    out = simulation.create_fullfield_data_numpy(5000, 1e-7, None, 4, ph, ph, 'ff_plane.h5', batch_size=2, probe_type='plane',
                                                 theta_st=0, theta_end=np.pi)
    assert out.shape == g['ff_plane'].shape and out.dtype == np.complex64
    assert rel_l2(out, g['ff_plane']) < 1e-5 and rel_l2(np.abs(out) ** 2, np.abs(g['ff_plane']) ** 2) < 1e-5
    out = simulation.create_fullfield_data_numpy(5000, 1e-7, 1e-4, 3, ph, ph, 'ff_gauss.h5', batch_size=1, probe_type='gaussian',
                                                 theta_st=0, theta_end=2 * np.pi, probe_mag_sigma=8., probe_phase_sigma=8.,
                                                 probe_phase_max=0.5)
    assert rel_l2(out, g['ff_gauss_free']) < 1e-5
    out = simulation.create_ptychography_data_batch_numpy(5000, 1e-7, 2, ph, ph, 'pty.h5', [tuple(p) for p in g['pty_pos']],
                                                          probe_type='gaussian', probe_size=(18, 18), theta_st=0, theta_end=np.pi / 3,
                                                          probe_circ_mask=None, minibatch_size=3, probe_mag_sigma=3.,
                                                          probe_phase_sigma=3., probe_phase_max=0.5)
    assert out.shape == g['pty'].shape
    assert rel_l2(out, g['pty']) < 1e-5 and rel_l2(np.abs(out) ** 2, np.abs(g['pty']) ** 2) < 1e-5
    # the file lands where the reference writes it (HDF5 exchange/data, or <fname>.npy without h5py)
    assert os.path.exists(os.path.join(ph, 'pty.h5')) or os.path.exists(os.path.join(ph, 'pty.h5.npy'))
    with pytest.raises(NotImplementedError):
        simulation.create_fullfield_data_numpy(5000, 1e-7, None, 1, ph, ph, 'x.h5', probe_type='point')
