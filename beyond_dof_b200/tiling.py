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


# ---------------------------------------------------------------------------------------------------------------------------
# Production path (BASELINE config 5): blocks with aprons in peer-mapped memory, strips stored straight into the neighbours'
# aprons by a kernel (libbdof: csrc/tilehalo.cu).  The functions above remain as the small reference implementation over
# torch.distributed point-to-point (CPU / gloo tests).
# ---------------------------------------------------------------------------------------------------------------------------
FAST_LENGTHS = (256, 512, 1024, 2048, 4096, 8192)


def axis_tiles(block, halo, lengths=FAST_LENGTHS):
    """Cover one axis of a block of `block` pixels with n windows of ONE power-of-two length L (the register-resident FFT
    kernels), each owning a contiguous share of the block and seeing at least `halo` pixels beyond it on both sides.  The shares
    need not be equal and the window need not divide anything: L and n minimise the transformed length n * L.
    Returns (L, [(window_start, own_start, own_len)], apron) with positions relative to the block; windows may start at -apron."""
    best = None
    for L in lengths:
        if L <= 2 * (halo + 1):
            continue
        n = -(-block // (L - 2 * (halo + 1)))             # + 1: starts are rounded to even pixels below
        if best is None or n * L < best[0] * best[1]:
            best = (n, L)
    if best is None:
        raise ValueError('halo %d too large for the available FFT lengths' % halo)
    n, L = best
    edges = [round(k * block / n) for k in range(n + 1)]
    out, apron = [], 0
    for k in range(n):
        own = edges[k + 1] - edges[k]
        start = edges[k] - (L - own) // 2
        start -= start % 2                       # even starts: window rows stay 16-byte aligned in the block buffer (bulk copies)
        out.append((start, edges[k], own))
        apron = max(apron, -start, start + L - block)
    apron += apron % 2
    return L, out, apron


def window_kernel_factors(dist_nm, lmbda_nm, voxel_nm, n_global, n_window, pi=None):
    """Factor of the reference transfer function for a WINDOW of n_window pixels of a field of n_global pixels.
    get_kernel (tensorflow_recon/util.py:165-185) samples u on linspace(-1/(2 dx), 1/(2 dx), N): relative to the true DFT
    frequency f of a bin it evaluates H at u(f) = f N/(N-1) + 1/(2 (N-1) dx), a stretch and an offset that depend on N.  A
    window that transformed with its OWN n_window would therefore propagate over a slightly different distance than the global
    field (0.4 % at 256 vs 0.006 % at 16384: 2.6e-4 on the intensity, measured); the windows use the global N instead, so that
    the tiled result converges to the global-FFT reference as the halo grows."""
    from .util import PI
    pi = PI if pi is None else pi
    f = (np.arange(n_window) - n_window // 2) / (n_window * voxel_nm)
    u = f * n_global / (n_global - 1.0) + 1.0 / (2.0 * (n_global - 1.0) * voxel_nm)
    return np.exp(-1j * pi * lmbda_nm * dist_nm * u ** 2)


class TiledMultislice:
    """Tiling-based forward multislice of ONE large field over the ranks of the default process group (one process per GPU).

    grid = (gy, gx) blocks; this rank owns block (rank // gx, rank % gx) of the [NY, NX] field.  Every block is covered by local
    FFT windows (axis_tiles) that see `halo` pixels beyond what they own; per slice the windows are cut from the block (with its
    apron), stepped with the exact FFT propagator (bdof_slice_step), their owned parts pasted back, and the block's border strips
    stored into the neighbours' aprons over NVLink peer memory (bdof_tiles_halo_exchange).  The approximation is the truncation
    of the Fresnel kernel at `halo` pixels (error vs halo).

    db_block_fn(y0, x0, h, w) -> [h, w, 2] float32 CUDA tensor with (delta, beta) of the global pixels y0..y0+h-1, x0..x0+w-1
    taken PERIODICALLY (the caller wraps indices): an axially repeating object (z_broadcast), evaluated once at set-up.
    """

    def __init__(self, ny, nx, grid, halo, energy_ev, psize_cm, n_slice, db_block_fn, propagate_last=False, group=None,
                 lengths=FAST_LENGTHS):
        import ctypes
        from .capi import lib, check
        from .plan import MultislicePlan
        from .dist import _DeviceBuffer
        self._lib, self._check = lib, check
        rank, w = world()
        gy, gx = int(grid[0]), int(grid[1])
        if gy * gx != w:
            raise ValueError('tile grid %dx%d needs %d ranks, the process group has %d' % (gy, gx, gy * gx, w))
        if ny % gy or nx % gx:
            raise ValueError('the field must split evenly into the block grid')
        self.rank, self.gy, self.gx = rank, gy, gx
        self.ny, self.nx, self.by, self.bx = int(ny), int(nx), ny // gy, nx // gx
        self.n_slice, self.propagate_last = int(n_slice), bool(propagate_last)
        self.device = torch.device('cuda', torch.cuda.current_device())
        self.ly, ty, ay = axis_tiles(self.by, halo, lengths)
        self.lx, tx, ax = axis_tiles(self.bx, halo, lengths)
        self.apron = max(ay, ax)
        self.n_tiles = len(ty) * len(tx)
        origin = [(sy, sx) for (sy, _, _) in ty for (sx, _, _) in tx]
        own = [(oy, ox, hy, hx) for (_, oy, hy) in ty for (_, ox, hx) in tx]
        self.origin = torch.tensor(origin, dtype=torch.int32, device=self.device).contiguous()
        pitch = self.bx + 2 * self.apron
        # element offset of every window's first row in a block buffer (the row pass reads the windows in place: fused cut)
        self.win_offsets = torch.tensor([(self.apron + sy) * pitch + self.apron + sx for (sy, sx) in origin], dtype=torch.int64,
                                        device=self.device).contiguous()
        self.fused_cut = (self.bx % 2 == 0)
        self.own = torch.tensor(own, dtype=torch.int32, device=self.device).contiguous()
        self.redundancy = self.n_tiles * self.ly * self.lx / float(self.by * self.bx)
        self._h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            check(lib.bdof_tiles_create(ctypes.byref(self._h), rank, gy, gx, self.by, self.bx, self.apron))
            hb = lib.bdof_tiles_handle_bytes()
            mine = ctypes.create_string_buffer(hb)
            check(lib.bdof_tiles_export(self._h, mine))
            if w > 1:
                handles = [None] * w
                dist.all_gather_object(handles, bytes(mine.raw), group=group)
                check(lib.bdof_tiles_connect(self._h, ctypes.c_char_p(b''.join(handles))))
            else:
                check(lib.bdof_tiles_connect(self._h, ctypes.c_char_p(bytes(mine.raw))))
            self.pitch = self.bx + 2 * self.apron
            self.rows = self.by + 2 * self.apron
            self.buf = []
            for which in (0, 1):
                ptr = ctypes.c_void_p()
                check(lib.bdof_tiles_block_ptr(self._h, which, ctypes.byref(ptr)))
                t = torch.as_tensor(_DeviceBuffer(ptr.value, self.rows * self.pitch * 2, self), device=self.device)
                self.buf.append(torch.view_as_complex(t.view(self.rows, self.pitch, 2)))
        # one plan for all windows of this rank: single slice steps with the exact FFT propagator
        voxel_nm, lmbda_nm = psize_cm * 1e7, 1240. / energy_ev
        fac = (complex(self.plan_phase0(energy_ev, psize_cm)),
               window_kernel_factors(voxel_nm, lmbda_nm, voxel_nm, self.ny, self.ly),
               window_kernel_factors(voxel_nm, lmbda_nm, voxel_nm, self.nx, self.lx))
        self.plan = MultislicePlan(self.ly, self.lx, self.n_tiles, max(self.n_slice, 2), energy_ev, psize_cm, device=self.device,
                                   propagate_last=self.propagate_last, stepwise=True, factors=fac)
        self.phase0 = complex(self.plan_phase0(energy_ev, psize_cm))
        # (delta, beta) windows, cut once (axially repeating object): block with apron from the caller, then the same cut kernel
        y0 = (rank // gx) * self.by - self.apron
        x0 = (rank % gx) * self.bx - self.apron
        dbb = db_block_fn(y0, x0, self.rows, self.pitch).to(self.device, torch.float32).contiguous()
        assert tuple(dbb.shape) == (self.rows, self.pitch, 2)
        self.db_tiles = torch.empty((self.n_tiles, self.ly, self.lx, 2), dtype=torch.float32, device=self.device)
        self.tiles = [torch.empty((self.n_tiles, self.ly, self.lx), dtype=torch.complex64, device=self.device) for _ in range(2)]
        st = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        self.buf[0].copy_(torch.view_as_complex(dbb))
        check(lib.bdof_tiles_cut(self._h, 0, self._p(self.origin), self.n_tiles, self.ly, self.lx, self._p(self.db_tiles), st))
        torch.cuda.current_stream(self.device).synchronize()
        if w > 1:
            dist.barrier(group=group)             # nobody stores into a peer that has not opened the handles yet
        self._group = group

    @staticmethod
    def _p(t):
        import ctypes
        return ctypes.c_void_p(t.data_ptr())

    @staticmethod
    def plan_phase0(energy_ev, psize_cm):
        from .util import PI
        dz_nm = psize_cm * 1e7
        return np.exp(1j * 2 * PI / (1240. / energy_ev) * dz_nm)

    def interior(self, which):
        a = self.apron
        return self.buf[which][a:a + self.by, a:a + self.bx]

    def run(self, probe_block_fn=None):
        """Propagate through all slices.  probe_block_fn(y0, x0, h, w) -> complex64 [h, w] entrance wave of those global pixels
        (periodic); default: plane wave.  Returns this rank's block of the exit wave [by, bx] complex64 (a view of the buffer)."""
        import ctypes
        lib, check = self._lib, self._check
        st = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        cur = 0
        if probe_block_fn is None:
            self.buf[0].fill_(1.0)
        else:
            y0 = (self.rank // self.gx) * self.by - self.apron
            x0 = (self.rank % self.gx) * self.bx - self.apron
            self.buf[0].copy_(probe_block_fn(y0, x0, self.rows, self.pitch))
        n_prop = 0
        for i in range(self.n_slice):
            prop = True if self.propagate_last else (i < self.n_slice - 1)
            if prop and self.fused_cut:
                # the row pass reads the windows straight out of the block buffer
                self.plan.use_current_stream()
                check(lib.bdof_slice_step_windows(self.plan._h, self._p(self.buf[cur]), self.pitch, self._p(self.win_offsets),
                                                  self._p(self.db_tiles), self._p(self.tiles[1]), i))
            else:
                check(lib.bdof_tiles_cut(self._h, cur, self._p(self.origin), self.n_tiles, self.ly, self.lx, self._p(self.tiles[0]), st))
                self.plan.slice_step(self.tiles[0], self.db_tiles, out=self.tiles[1], propagate=prop, index=i)
            check(lib.bdof_tiles_paste(self._h, cur ^ 1, self._p(self.tiles[1]), self._p(self.origin), self._p(self.own), self.n_tiles,
                                       self.ly, self.lx, st))
            if prop and i < self.n_slice - 1:
                check(lib.bdof_tiles_halo_exchange(self._h, cur ^ 1, st))
            n_prop += 1 if prop else 0
            cur ^= 1
        self.total_phase = self.phase0 ** n_prop          # the engine keeps exp(i k dz) per propagation out of its fp32 chain
        return self.interior(cur)

    def close(self):
        if getattr(self, '_h', None) is not None and self._h.value:
            torch.cuda.synchronize(self.device)
            if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
                dist.barrier(group=self._group)
            self.buf = []
            self._lib.bdof_tiles_destroy(self._h)
            import ctypes
            self._h = ctypes.c_void_p()
