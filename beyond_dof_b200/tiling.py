"""Tiling-based multislice for lateral fields too large for one GPU (Ali, Du et al., Opt. Express 2020;
the enabling idea in this snapshot is the finite real-space kernel of cnn_propagator/propagation.py:18-133:
finite support => finite halo).

The global [NY, NX] field is cut into tiles of `interior` x `interior` pixels; every tile is stored with
a halo of `halo` pixels on each side, local size L = interior + 2*halo (a supported FFT length).  Each
slice is (1) modulated and propagated locally on every padded tile with the exact FFT propagator of the
line kernels (the tile's own periodic wrap only contaminates the outer `support` pixels of the halo),
(2) followed by a halo refresh: every halo strip is overwritten with the neighbour tile's interior edge,
with wrap-around neighbours at the global border so that the global periodic boundary of the reference's
FFT propagator is reproduced.  The refresh runs as two phases (x strips, then y strips over the full
width) so corners need no separate messages.

Tiles are dealt to ranks in contiguous blocks; strips between tiles of the same rank are device copies,
strips between ranks go over NCCL point-to-point (NVLink) in one batched isend/irecv per phase.
Everything here is host logic over torch tensors: it runs over gloo/CPU in the tests and NCCL/CUDA in
production.  The approximation is the truncation of the Fresnel kernel's tails at `halo` pixels: see
DESIGN.md ("tiling error vs halo") for the measured error.
"""
import numpy as np
import torch
import torch.distributed as dist

from .dist import world


class TileLayout:
    """Geometry of the decomposition and ownership of tiles."""

    def __init__(self, ny, nx, local_n, halo, world_size=1):
        self.ny, self.nx, self.local_n, self.halo = int(ny), int(nx), int(local_n), int(halo)
        self.interior = self.local_n - 2 * self.halo
        if self.interior <= 0 or self.halo < 0:
            raise ValueError('halo too large for the local tile size')
        if self.halo > self.interior:
            raise ValueError('halo must not exceed the tile interior (a strip comes from ONE neighbour)')
        if self.ny % self.interior or self.nx % self.interior:
            raise ValueError('global field %dx%d is not a multiple of the tile interior %d' % (ny, nx, self.interior))
        self.ty, self.tx = self.ny // self.interior, self.nx // self.interior
        self.n_tiles = self.ty * self.tx
        self.world_size = int(world_size)
        base, extra = divmod(self.n_tiles, self.world_size)
        self.owner = np.empty(self.n_tiles, dtype=np.int64)
        self.local_index = np.empty(self.n_tiles, dtype=np.int64)
        k = 0
        for r in range(self.world_size):
            cnt = base + (1 if r < extra else 0)
            self.owner[k:k + cnt] = r
            self.local_index[k:k + cnt] = np.arange(cnt)
            k += cnt

    def tiles_of(self, rank):
        return np.nonzero(self.owner == rank)[0]

    def tile_id(self, iy, ix):
        return (iy % self.ty) * self.tx + (ix % self.tx)

    def coords(self, tile):
        return divmod(int(tile), self.tx)


def scatter_to_tiles(global_arr, layout, rank=0):
    """Cut this rank's tiles (with halos, periodic wrap at the global border) out of a global
    [..., NY, NX] array -> [n_local, ..., L, L]."""
    h, it, L = layout.halo, layout.interior, layout.local_n
    out = []
    for tile in layout.tiles_of(rank):
        iy, ix = layout.coords(tile)
        ys = (torch.arange(iy * it - h, iy * it - h + L, device=global_arr.device)) % layout.ny
        xs = (torch.arange(ix * it - h, ix * it - h + L, device=global_arr.device)) % layout.nx
        out.append(global_arr.index_select(-2, ys).index_select(-1, xs))
    if not out:
        return global_arr.new_zeros((0,) + tuple(global_arr.shape[:-2]) + (L, L))
    return torch.stack(out).contiguous()


def gather_from_tiles(tiles_by_rank, layout):
    """Assemble the global [..., NY, NX] array from the interiors; tiles_by_rank[r] = rank r's [n_local, ..., L, L]."""
    h, it = layout.halo, layout.interior
    first = next(t for t in tiles_by_rank if t.shape[0] > 0)
    out = first.new_zeros(tuple(first.shape[1:-2]) + (layout.ny, layout.nx))
    for r, tiles in enumerate(tiles_by_rank):
        for j, tile in enumerate(layout.tiles_of(r)):
            iy, ix = layout.coords(tile)
            out[..., iy * it:(iy + 1) * it, ix * it:(ix + 1) * it] = tiles[j][..., h:h + it, h:h + it]
    return out


def _exchange_phase(tiles, layout, rank, axis):
    """Refresh the two halo strips along `axis` (-1: x / columns, -2: y / rows) of every local tile."""
    h, L = layout.halo, layout.local_n
    if h == 0:
        return
    mine = layout.tiles_of(rank)
    # rows/cols taking part: x phase only touches interior rows (corners come from the y phase, which
    # copies full-width strips that already contain the fresh x halos)
    def strip(t, lo, hi):
        if axis == -1:
            return t[..., h:L - h, lo:hi]
        return t[..., lo:hi, :]
    sends, recvs, copies = [], [], []
    for j, tile in enumerate(mine):
        iy, ix = layout.coords(tile)
        for side in (-1, +1):
            nb = layout.tile_id(iy, ix + side) if axis == -1 else layout.tile_id(iy + side, ix)
            # my halo on `side` <- neighbour's interior edge facing me
            dst = strip(tiles[j], 0, h) if side < 0 else strip(tiles[j], L - h, L)
            src_lo, src_hi = (L - 2 * h, L - h) if side < 0 else (h, 2 * h)
            owner = int(layout.owner[nb])
            if owner == rank:
                copies.append((dst, strip(tiles[int(layout.local_index[nb])], src_lo, src_hi)))
            else:
                recvs.append((owner, int(tile), side, dst))
            # and the neighbour on `side` needs MY interior edge facing it (as its halo on the opposite side)
            if owner != rank:
                my_lo, my_hi = (h, 2 * h) if side < 0 else (L - 2 * h, L - h)
                sends.append((owner, int(nb), -side, strip(tiles[j], my_lo, my_hi).contiguous()))
    # local copies first need the sources untouched by this phase: sources are interior columns/rows,
    # destinations are halo columns/rows -> disjoint, any order is fine
    staged = [(d, s.clone()) for d, s in copies]
    for d, s in staged:
        d.copy_(s)
    if sends or recvs:
        # deterministic message order on both sides: sort by (receiving tile, side)
        sends.sort(key=lambda m: (m[0], m[1], m[2]))
        recvs.sort(key=lambda m: (m[0], m[1], m[2]))
        bufs = [torch.empty_like(m[3]) for m in recvs]
        ops = [dist.P2POp(dist.isend, m[3], m[0]) for m in sends] + \
              [dist.P2POp(dist.irecv, b, m[0]) for b, m in zip(bufs, recvs)]
        for req in dist.batch_isend_irecv(ops):
            req.wait()
        for b, m in zip(bufs, recvs):
            m[3].copy_(b)


def halo_exchange(tiles, layout, rank=None):
    """Overwrite every halo of this rank's tiles [n_local, ..., L, L] with the neighbours' interior edges."""
    if rank is None:
        rank, _ = world()
    _exchange_phase(tiles, layout, rank, -1)
    _exchange_phase(tiles, layout, rank, -2)
    return tiles


def tiled_multislice(db_tiles, field_tiles, layout, step_fn, n_slice, propagate_last=False, rank=None):
    """Run the slice loop on this rank's tiles.

    db_tiles:    [Z, n_local, L, L, 2] (delta, beta) tiles with halos (static, scattered once), or
                 [1, n_local, L, L, 2] for an axially repeating object
    field_tiles: [n_local, L, L] complex64 entrance wave tiles (with halos)
    step_fn(field, db_slice, propagate) -> new field: the local slice step (MultislicePlan.slice_step)
    Returns the exit wave tiles; valid data is the interior (use gather_from_tiles).
    """
    if rank is None:
        rank, _ = world()
    z_bcast = db_tiles.shape[0] == 1
    for i in range(n_slice):
        prop = True if propagate_last else (i < n_slice - 1)
        field_tiles = step_fn(field_tiles, db_tiles[0 if z_bcast else i], prop)
        if prop:
            halo_exchange(field_tiles, layout, rank)
    return field_tiles
