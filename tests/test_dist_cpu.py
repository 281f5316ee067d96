"""CPU tier, world_size 2 over gloo: the host logic of the multi-GPU paths --
sharding, the gradient all-reduce (plain and z-bucketed), and the tile/halo exchange of the tiling
scheme -- with stand-in local compute (the NumPy oracle / a periodic stencil) so no GPU is needed."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import rel_l2


def _free_port():
    s = socket.socket(); s.bind(('127.0.0.1', 0)); p = s.getsockname()[1]; s.close(); return p


def _run(rank, world, port, fn, ret):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        ret[rank] = fn(rank, world)
    finally:
        dist.destroy_process_group()


def spawn(fn, world=2):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_run, args=(world, _free_port(), fn, ret), nprocs=world, join=True)
    return [ret[r] for r in range(world)]


def test_sharding_covers_everything_once():
    from beyond_dof_b200 import dist as bd
    for n in (0, 1, 7, 16, 1024):
        for w in (1, 2, 3, 8):
            rr = np.concatenate([bd.shard_round_robin(n, r, w) for r in range(w)])
            cc = np.concatenate([bd.shard_contiguous(n, r, w) for r in range(w)])
            assert sorted(rr.tolist()) == list(range(n)) and cc.tolist() == list(range(n))
            sizes = [len(bd.shard_contiguous(n, r, w)) for r in range(w)]
            assert max(sizes) - min(sizes) <= 1
    assert bd.world() == (0, 1)


def _dp_ptycho(rank, world):
    from beyond_dof_b200 import dist as bd
    from oracle import multislice_oracle as mo
    Y, X, Z = 80, 80, 2
    ps = (64, 64)
    od, ob = mo.random_phantom((Y, X, Z), seed=70, delta_scale=1e-3, beta_scale=1e-4)
    pr, pi = mo.gaussian_probe(ps, 8., 8., 0.3)
    pos = [(32, 32), (40, 44), (47, 47), (36, 41), (45, 33)]
    prj = np.random.default_rng(71).random((len(pos),) + ps) * 30

    def local(idx, gbuf):
        # per-item losses sum to the loss over all positions (mean over n_pos x py x px, times n_pos)
        l, gd, gb, _ = mo.ptycho_loss_and_grad(od, ob, [pos[i] for i in idx], prj[idx], pr, pi, ps, 5000, 1e-7, scale_by_npos=False)
        w = len(idx) / len(pos) * len(pos)
        gbuf[..., 0] += torch.as_tensor(gd * w); gbuf[..., 1] += torch.as_tensor(gb * w)
        return l * w
    g = torch.zeros((Y, X, Z, 2), dtype=torch.float64)
    loss, g = bd.data_parallel_step(len(pos), local, g, average=False)
    return loss, g.numpy()


def test_data_parallel_gradient_equals_single_process():
    from oracle import multislice_oracle as mo
    res = spawn(_dp_ptycho, 2)
    Y, X, Z = 80, 80, 2
    ps = (64, 64)
    od, ob = mo.random_phantom((Y, X, Z), seed=70, delta_scale=1e-3, beta_scale=1e-4)
    pr, pi = mo.gaussian_probe(ps, 8., 8., 0.3)
    pos = [(32, 32), (40, 44), (47, 47), (36, 41), (45, 33)]
    prj = np.random.default_rng(71).random((len(pos),) + ps) * 30
    l, gd, gb, _ = mo.ptycho_loss_and_grad(od, ob, pos, prj, pr, pi, ps, 5000, 1e-7, scale_by_npos=True)
    for loss_r, g_r in res:                                   # every rank ends with the same total
        assert abs(loss_r - l) < 1e-10 * abs(l)
        assert rel_l2(g_r[..., 0], gd) < 1e-12 and rel_l2(g_r[..., 1], gb) < 1e-12


def _bucketed(rank, world):
    from beyond_dof_b200 import dist as bd
    g = torch.arange(10 * 6, dtype=torch.float32).reshape(10, 3, 2) * (rank + 1)
    buckets = [(7, 10, None), (4, 7, None), (1, 4, None), (0, 1, None)]      # completion order of the adjoint sweep
    works = bd.allreduce_gradient(g, average=True, buckets=buckets)
    bd.finish_allreduce(g, works)
    t = torch.zeros(3) + rank
    bd.broadcast_object_(t, src=0)
    return g.numpy(), bd.allreduce_scalar(rank + 1.0), t.numpy()


def test_bucketed_allreduce_mean_and_broadcast():
    res = spawn(_bucketed, 2)
    base = np.arange(60, dtype=np.float32).reshape(10, 3, 2)
    for g, s, t in res:
        assert np.allclose(g, base * 1.5) and s == 3.0 and np.all(t == 0)


def _stencil(t):
    # periodic 5-point-ish stencil INSIDE the local array (support 1 pixel): stands in for the local FFT step
    return 0.2 * (t + torch.roll(t, 1, -1) + torch.roll(t, -1, -1) + torch.roll(t, 1, -2) + torch.roll(t, -1, -2))


def _tiled(rank, world):
    from beyond_dof_b200 import tiling
    ny, nx, L, h = 24, 36, 16, 2                       # interior 12 -> 2 x 3 tiles
    layout = tiling.TileLayout(ny, nx, L, h, world)
    g = torch.arange(ny * nx, dtype=torch.float64).reshape(ny, nx).sin()
    tiles = tiling.scatter_to_tiles(g, layout, rank)
    db = torch.zeros((1, tiles.shape[0], L, L, 2))
    out = tiling.tiled_multislice(db, tiles, layout, lambda f, d, p: _stencil(f) if p else f, n_slice=5, propagate_last=True, rank=rank)
    return out.numpy()


def test_tiled_halo_exchange_matches_global_periodic_operator():
    from beyond_dof_b200 import tiling
    res = spawn(_tiled, 2)
    ny, nx, L, h = 24, 36, 16, 2
    layout = tiling.TileLayout(ny, nx, L, h, 2)
    assert layout.n_tiles == 6 and sorted(layout.tiles_of(0).tolist() + layout.tiles_of(1).tolist()) == list(range(6))
    g = torch.arange(ny * nx, dtype=torch.float64).reshape(ny, nx).sin()
    ref = g.clone()
    for _ in range(5):
        ref = _stencil(ref)
    got = tiling.gather_from_tiles([torch.as_tensor(r) for r in res], layout)
    assert torch.equal(got, ref) or rel_l2(got.numpy(), ref.numpy()) < 1e-15
    # single-rank layout (all strips are local copies, incl. a 1-wide tile grid that is its own neighbour)
    lay1 = tiling.TileLayout(12, 36, 16, 2, 1)
    t1 = tiling.scatter_to_tiles(g[:12], lay1, 0)
    out = tiling.tiled_multislice(torch.zeros((1, 3, 16, 16, 2)), t1, lay1, lambda f, d, p: _stencil(f), n_slice=3, propagate_last=True, rank=0)
    ref = g[:12].clone()
    for _ in range(3):
        ref = _stencil(ref)
    assert rel_l2(tiling.gather_from_tiles([out], lay1).numpy(), ref.numpy()) < 1e-15


def test_tile_layout_validation():
    from beyond_dof_b200 import tiling
    with pytest.raises(ValueError):
        tiling.TileLayout(100, 100, 64, 8, 1)          # 100 is not a multiple of 48
    with pytest.raises(ValueError):
        tiling.TileLayout(64, 64, 64, 32, 1)           # no interior left
    lay = tiling.TileLayout(96, 48, 64, 8, 4)
    assert (lay.ty, lay.tx, lay.interior) == (2, 1, 48)
    assert [len(lay.tiles_of(r)) for r in range(4)] == [1, 1, 0, 0]


def _pick(rank, world):
    from beyond_dof_b200 import dist as bd
    return bd.pick_exchange()


def test_exchange_policy_and_dynamic_dropping():
    # two ranks: copy engines; anything else: NCCL (measured on B200); no process group: NCCL path (a no-op)
    from beyond_dof_b200 import dist as bd
    assert bd.pick_exchange() == 'nccl'
    assert spawn(_pick, 2) == ['ce', 'ce']
    assert spawn(_pick, 3) == ['nccl'] * 3
    # the reference drops the scan positions whose loss is below the threshold (cnn_propagator/ptychography.py:340)
    from beyond_dof_b200.models import dynamic_dropping
    table = np.array([1e-3, 2e-5, 8e-5, 7.9e-5, 0.5])
    assert dynamic_dropping(table).tolist() == [0, 2, 4]
    assert dynamic_dropping(torch.as_tensor(table), dropping_threshold=1e-2).tolist() == [4]


def test_axis_tiles_cover_the_block_with_the_requested_halo():
    # host logic of the production tiling path (tiling.axis_tiles): windows of ONE power-of-two length, contiguous owned shares
    # that need not be equal nor divide the block, at least `halo` pixels visible beyond every share, even window starts
    from beyond_dof_b200.tiling import axis_tiles, FAST_LENGTHS
    for block in (1024, 2048, 4096, 8192, 16384, 3000):
        for halo in (4, 8, 16, 32, 64, 100):
            L, tiles, apron = axis_tiles(block, halo)
            assert L in FAST_LENGTHS
            assert tiles[0][1] == 0 and tiles[-1][1] + tiles[-1][2] == block
            for (s0, o0, w0), (s1, o1, w1) in zip(tiles, tiles[1:]):
                assert o0 + w0 == o1                                   # shares are contiguous
            for s, o, w in tiles:
                assert s % 2 == 0 and o - s >= halo and (s + L) - (o + w) >= halo
                assert s >= -apron and s + L <= block + apron
            assert apron % 2 == 0
            # no other length does the same job with less transformed length
            for L2 in FAST_LENGTHS:
                if L2 > 2 * (halo + 1):
                    n2 = -(-block // (L2 - 2 * (halo + 1)))
                    assert n2 * L2 >= len(tiles) * L
    # restricted choice of lengths (bench --tile-lengths)
    L, tiles, _ = axis_tiles(8192, 8, (2048,))
    assert L == 2048 and len(tiles) == 5


def test_back_rotation_list_entries_split_exactly():
    """The back-rotation kernel splits a reader-list entry d = z * nx + x with a multiply and a shift (bdof.cu: RotDiv); the same
    arithmetic evaluated on the host must be the exact division for every d a table can hold."""
    import ctypes
    from beyond_dof_b200 import capi
    z, x = ctypes.c_int(), ctypes.c_int()
    rng = np.random.default_rng(5)
    for nx in (1, 2, 3, 18, 64, 72, 97, 256, 1000, 2048, 4097, 46340):
        top = min(nx * nx, 2 ** 31 - 1)
        ds = np.unique(np.concatenate([np.arange(0, min(top, 3000)), rng.integers(0, top, 3000), [top - 1, max(top - nx, 0), 2 ** 31 - 1]]))
        for d in ds:
            assert capi.lib.bdof_debug_rot_split(nx, int(d), ctypes.byref(z), ctypes.byref(x)) == 0
            assert (z.value, x.value) == (int(d) // nx, int(d) % nx), (nx, int(d))
