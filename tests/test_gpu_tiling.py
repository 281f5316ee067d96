"""GPU tier: single-slice stepping, the tiling scheme against the global-FFT oracle (error vs halo),
and the z-bucket events of the adjoint."""
import numpy as np
import pytest
import torch

from conftest import rel_l2
from oracle import multislice_oracle as mo

pytestmark = pytest.mark.gpu


def test_slice_step_equals_forward_chain():
    from beyond_dof_b200.plan import MultislicePlan
    B, Y, X, Z = 2, 128, 64, 5
    gd, gb = mo.random_phantom((B, Y, X, Z), seed=80, delta_scale=3e-4, beta_scale=3e-5)
    plan = MultislicePlan(Y, X, B, Z, 5000, 1e-7)
    db = plan.pack(torch.as_tensor(gd).cuda(), torch.as_tensor(gb).cuda())
    probe = torch.ones((Y, X), dtype=torch.complex64, device='cuda')
    ref = plan.forward(db, probe)
    f = probe[None].expand(B, Y, X).contiguous()
    n_prop = 0
    for i in range(Z):
        prop = i < Z - 1
        f = plan.slice_step(f, db[i].contiguous(), propagate=prop)
        n_prop += prop
    phase = np.exp(1j * plan.k_dz * n_prop)          # the global phase is left out of slice_step
    assert rel_l2((f * complex(phase)).cpu().numpy(), ref.cpu().numpy()) < 2e-6


@pytest.mark.parametrize('halo', [32, 64])
def test_tiled_multislice_vs_global_oracle(halo):
    # scaled-down twin of BASELINE config 5: global 384^2 field, 256^2 local tiles, axially varying random object
    from beyond_dof_b200 import tiling
    from beyond_dof_b200.plan import MultislicePlan
    NY = NX = 384; L = 256; Z = 12
    layout = tiling.TileLayout(NY, NX, L, halo, 1)
    gd, gb = mo.random_phantom((1, NY, NX, Z), seed=81, delta_scale=3e-4, beta_scale=3e-5)
    pr, pi = mo.gaussian_probe((NY, NX), 60., 60., 0.5)
    ref = mo.multislice_forward(gd.astype(np.float64), gb.astype(np.float64), pr, pi, 5000, 1e-7, propagate_last=True)[0]
    dbg = torch.stack([torch.as_tensor(gd[0]), torch.as_tensor(gb[0])], -1).permute(2, 0, 1, 3).cuda()   # [Z,NY,NX,2]
    db_tiles = tiling.scatter_to_tiles(dbg.permute(0, 3, 1, 2), layout, 0)        # [n, Z, 2, L, L]
    db_tiles = db_tiles.permute(1, 0, 3, 4, 2).contiguous()                          # [Z, n, L, L, 2]
    probe = torch.as_tensor((pr + 1j * pi).astype(np.complex64)).cuda()
    f = tiling.scatter_to_tiles(probe, layout, 0)
    plan = MultislicePlan(L, L, layout.n_tiles, Z, 5000, 1e-7)
    out = tiling.tiled_multislice(db_tiles, f, layout, lambda fld, d, p: plan.slice_step(fld, d.contiguous(), propagate=p),
                                  n_slice=Z, propagate_last=True, rank=0)
    got = tiling.gather_from_tiles([out], layout).cpu().numpy()
    err = rel_l2(np.abs(got) ** 2, np.abs(ref) ** 2)
    print('tiling halo %d: rel-L2 intensity error %.3e' % (halo, err))
    # the error is the truncation of the Fresnel kernel's slowly decaying tails at `halo` pixels
    # ("tiling error vs halo"); it is an approximation, not a parity failure
    assert err < (2e-3 if halo == 32 else 1e-3)


def test_adjoint_bucket_events_fire_in_sweep_order():
    from beyond_dof_b200.plan import MultislicePlan
    B, Y, X, Z = 1, 64, 64, 10
    plan = MultislicePlan(Y, X, B, Z, 5000, 1e-7, store_slices=True)
    buckets = plan.set_gradient_buckets(4)
    assert [(a, b) for a, b, _ in buckets] == [(7, 10), (4, 7), (1, 4), (0, 1)]
    db = torch.rand((Z, B, Y, X, 2), device='cuda') * 1e-4
    psi = plan.forward(db, torch.ones((Y, X), dtype=torch.complex64, device='cuda'))
    _, g = plan.loss_mag(psi, torch.full((B, Y, X), 0.9, device='cuda'))
    grad = torch.empty_like(db)
    plan.adjoint(db, g, grad_out=grad)
    torch.cuda.synchronize()
    assert all(ev.query() for _, _, ev in buckets)
