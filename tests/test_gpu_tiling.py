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


# ---------------------------------------------------------------------------------------------------------------------------
# production tiling path (csrc/tilehalo.cu, tiling.TiledMultislice): blocks with aprons in peer-mapped memory
# ---------------------------------------------------------------------------------------------------------------------------
def _zone_plate_db(ny, nx, f_nm=20000.0, n_zones=40, delta=3e-4, beta=3e-5, lmbda=0.248):
    """(delta, beta) of a binary zone plate centred in the [ny, nx] field as a function of periodic global coordinates"""
    def fn(y0, x0, h, w):
        ys = (torch.arange(y0, y0 + h, device='cuda') % ny).double() - (ny - 1) / 2
        xs = (torch.arange(x0, x0 + w, device='cuda') % nx).double() - (nx - 1) / 2
        r2 = ys[:, None] ** 2 + xs[None, :] ** 2
        # zone index n with r_n^2 = n lambda f + (n lambda / 2)^2  ~  n lambda f
        n = torch.floor(r2 / (lmbda * f_nm)).long()
        mask = ((n % 2) == 1) & (n <= n_zones)
        out = torch.zeros((h, w, 2), dtype=torch.float32, device='cuda')
        out[..., 0] = mask * delta
        out[..., 1] = mask * beta
        return out
    return fn


def _pattern(ny, nx):
    def fn(y0, x0, h, w):
        ys = torch.arange(y0, y0 + h, device='cuda') % ny
        xs = torch.arange(x0, x0 + w, device='cuda') % nx
        return torch.complex((ys[:, None] * nx + xs[None, :]).float(), (ys[:, None] - 2 * xs[None, :]).float())
    return fn


def _halo_pattern_check(rank, world, grid, ny, nx, halo, lengths):
    """cut -> paste -> halo exchange reproduce the periodic global pattern in interior AND apron, exactly"""
    import ctypes
    from beyond_dof_b200 import tiling
    from beyond_dof_b200.capi import lib, check
    tm = tiling.TiledMultislice(ny, nx, grid, halo, 5000, 1e-7, 4, _zone_plate_db(ny, nx), lengths=lengths)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    pat = _pattern(ny, nx)
    y0, x0 = (rank // grid[1]) * tm.by, (rank % grid[1]) * tm.bx
    full = pat(y0 - tm.apron, x0 - tm.apron, tm.rows, tm.pitch)
    ok = True
    for rep in range(3):                                   # ping-pong a few times, as the slice loop does
        src, dst = rep % 2, (rep + 1) % 2
        tm.buf[src].copy_(full)
        tm.buf[dst].zero_()
        torch.cuda.synchronize()
        if world > 1:
            import torch.distributed as dist
            dist.barrier()                                 # (test only: nobody pushes into an apron that is still being cleared)
        check(lib.bdof_tiles_cut(tm._h, src, tm._p(tm.origin), tm.n_tiles, tm.ly, tm.lx, tm._p(tm.tiles[0]), st))
        check(lib.bdof_tiles_paste(tm._h, dst, tm._p(tm.tiles[0]), tm._p(tm.origin), tm._p(tm.own), tm.n_tiles, tm.ly, tm.lx, st))
        check(lib.bdof_tiles_halo_exchange(tm._h, dst, st))
        torch.cuda.synchronize()
        ok = ok and bool(torch.equal(tm.buf[dst], full))
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
    n_tiles, red = tm.n_tiles, tm.redundancy
    tm.close()
    return ok, n_tiles, red


@pytest.mark.parametrize('cfg', [((1, 1), 512, 768, 16, (256,)), ((1, 1), 256, 256, 8, (128, 256))])
def test_tile_cut_paste_halo_single_rank(cfg):
    grid, ny, nx, halo, lengths = cfg
    ok, n_tiles, red = _halo_pattern_check(0, 1, grid, ny, nx, halo, lengths)
    assert ok and n_tiles >= 1 and red >= 1.0


def _tiles_worker(rank, world):
    grid = {2: (1, 2), 4: (2, 2)}[world]
    return _halo_pattern_check(rank, world, grid, 512, 1024, 16, (256,))


def _tiles_physics_worker(rank, world):
    from beyond_dof_b200 import tiling
    grid = {2: (1, 2), 4: (2, 2)}[world]
    ny = nx = 1024
    tm = tiling.TiledMultislice(ny, nx, grid, 48, 5000, 1e-7, 10, _zone_plate_db(ny, nx), lengths=(256, 512))
    out = (tm.run() * complex(tm.total_phase)).cpu().numpy()
    tm.close()
    return rank, out


@pytest.mark.parametrize('world', [2, 4])
def test_tile_halo_exchange_between_ranks(world):
    # the ranks are processes (sharing cuda:0 on a one-GPU box, one device each otherwise): peer-mapped aprons over CUDA IPC
    from test_gpu_dist import spawn
    res = spawn(_tiles_worker, world)
    assert all(r[0] for r in res)


def test_tiled_multislice_ranks_vs_global_fft():
    # 2 x 2 ranks, 1024^2 zone plate, 10 slices: the tiled result against the exact global-FFT engine (itself oracle-checked)
    import beyond_dof_b200 as bd
    from test_gpu_dist import spawn
    ny = nx = 1024
    Z = 10
    db = _zone_plate_db(ny, nx)(0, 0, ny, nx)
    gd = db[..., 0][None, :, :, None].expand(1, ny, nx, Z).contiguous()
    gb = db[..., 1][None, :, :, None].expand(1, ny, nx, Z).contiguous()
    ref = bd.multislice_propagate_batch_numpy(gd, gb, np.ones((ny, nx)), np.zeros((ny, nx)), 5000, 1e-7, obj_batch_shape=(1, ny, nx, Z))[0].cpu().numpy()
    res = dict(spawn(_tiles_physics_worker, 4))
    full = np.block([[res[0], res[1]], [res[2], res[3]]])
    err = rel_l2(np.abs(full) ** 2, np.abs(ref) ** 2)
    assert err < 2e-3, err                                  # halo 48: the truncation error of the Fresnel kernel's tails


@pytest.mark.parametrize('halo', [8, 24, 40, 60])
def test_tiled_multislice_error_vs_halo_single_rank(halo):
    # the halo-vs-error study of SURVEY 8d/e on a 2048 x 2048 x 32 twin of config 5 (one rank, many windows, periodic wrap
    # through the rank's own apron), against the exact global-FFT engine; measured errors are recorded for DESIGN.md
    import json
    import os
    import beyond_dof_b200 as bd
    from beyond_dof_b200 import tiling
    from conftest import ROOT
    ny = nx = 2048
    Z = 32
    fn = _zone_plate_db(ny, nx, f_nm=80000.0, n_zones=200)
    db = fn(0, 0, ny, nx)
    gd = db[..., 0][None, :, :, None].expand(1, ny, nx, Z).contiguous()
    gb = db[..., 1][None, :, :, None].expand(1, ny, nx, Z).contiguous()
    ref = bd.multislice_propagate_batch_numpy(gd, gb, np.ones((ny, nx)), np.zeros((ny, nx)), 5000, 1e-7, obj_batch_shape=(1, ny, nx, Z))[0]
    tm = tiling.TiledMultislice(ny, nx, (1, 1), halo, 5000, 1e-7, Z, fn, lengths=(256,))
    out = tm.run() * complex(tm.total_phase)
    err = rel_l2((out.abs() ** 2).cpu().numpy(), (ref.abs() ** 2).cpu().numpy())
    path = os.path.join(ROOT, 'gpurun_out', 'tiling_halo_error.json')
    try:
        rec = json.load(open(path))
    except Exception:
        rec = {}
    eff = int((tm.ly - int(tm.own[:, 2].max())) // 2)             # windows are centred on what they own: the halo they really see
    rec[str(halo)] = {'intensity_rel_l2': err, 'effective_halo': eff, 'n_tiles': tm.n_tiles, 'window': tm.ly, 'redundancy': tm.redundancy,
                      'apron': tm.apron}
    os.makedirs(os.path.dirname(path), exist_ok=True)
    json.dump(rec, open(path, 'w'), indent=1, sort_keys=True)
    tm.close()
    assert err < {8: 5e-3, 24: 3e-3, 40: 2e-3, 60: 1e-3}[halo]
