"""CPU tier: the geometric facts the tiled rotation kernels rely on (csrc/bdof.cu: ROT_T = 32, ROT_BOX = 48), checked on the
reference-pinned lookup tables (rotation.rotation_table == save_rotation_lookup, tests/golden/ref_rot.npz).

 - gather: the sources of every 32 x 32 tile of rotated pixels lie in a box of at most 48 x 48 once the box starts on an even x
   (TMA boxes start on 16 bytes), for every angle -- so the plain-gather fall-back is never taken for a rotation table;
 - transpose: with the box centred on the mean position of the first readers of the tile's ordinary cells (the rule of
   k_rot_origins), the readers of interior cells lie inside the box up to a per-cent of stragglers one or two pixels outside near
   45 degrees, i.e. practically only the clipped border cells' tails take the cooperative global-memory path."""
import numpy as np
import pytest

from beyond_dof_b200 import rotation

T, BOX = 32, 48


@pytest.mark.parametrize('n', [64, 96, 256])
def test_sources_of_a_tile_fit_the_box(n):
    for theta in np.linspace(0, 2 * np.pi, 37):
        tab = rotation.rotation_table([4, n, n], float(theta)).reshape(n, n, 2).transpose(1, 0, 2)      # [z][x] -> (x_old, z_old)
        xo, zo = tab[..., 0], tab[..., 1]
        for z0 in range(0, n, T):
            for x0 in range(0, n, T):
                sx, sz = xo[z0:z0 + T, x0:x0 + T], zo[z0:z0 + T, x0:x0 + T]
                xmn = int(sx.min()) & ~1
                assert sz.max() - sz.min() + 1 <= BOX and sx.max() - xmn + 1 <= BOX, (n, theta, z0, x0)


@pytest.mark.parametrize('n', [64, 256])
def test_readers_of_ordinary_cells_fit_the_box_centred_on_their_mean(n):
    worst_tail, n_readers, n_outside, worst_miss = 0, 0, 0, 0
    for theta in np.linspace(0.05, 2 * np.pi, 19):
        tab = rotation.rotation_table([4, n, n], float(theta)).reshape(n, n, 2).transpose(1, 0, 2)
        src = (tab[..., 1].astype(np.int64) * n + tab[..., 0]).reshape(-1)         # source cell of rotated pixel z * n + x
        order = np.argsort(src, kind='stable')
        counts = np.bincount(src, minlength=n * n)
        offsets = np.concatenate([[0], np.cumsum(counts)])
        cz, cx = np.divmod(np.arange(n * n), n)
        for z0 in range(0, n, T):
            for x0 in range(0, n, T):
                cells = np.array([(z * n + x) for z in range(z0, min(z0 + T, n)) for x in range(x0, min(x0 + T, n))])
                ordinary = cells[(counts[cells] > 0) & (counts[cells] <= 2)]
                if len(ordinary) == 0:
                    continue
                first = order[offsets[ordinary]]
                zmn = max(0, min(int(np.sum(first // n) // len(ordinary)) - BOX // 2, n - BOX))
                xmn = max(0, min(int(np.sum(first % n) // len(ordinary)) - BOX // 2, n - BOX)) & ~1
                interior = ordinary[(cz[ordinary] > 0) & (cz[ordinary] < n - 1) & (cx[ordinary] > 0) & (cx[ordinary] < n - 1)]
                for c in interior:
                    for d in order[offsets[c]:offsets[c + 1]]:
                        z, x = divmod(int(d), n)
                        n_readers += 1
                        miss = max(zmn - z, z - (zmn + BOX - 1), xmn - x, x - (xmn + BOX - 1), 0)
                        n_outside += miss > 0
                        worst_miss = max(worst_miss, miss)
                # cells with more than two readers exist only on the border (clipped coordinates)
                long_cells = cells[counts[cells] > 2]
                assert np.all((cz[long_cells] == 0) | (cz[long_cells] == n - 1) | (cx[long_cells] == 0) | (cx[long_cells] == n - 1))
                if len(long_cells):
                    worst_tail = max(worst_tail, int(counts[long_cells].max()))
    # near 45 degrees the readers of a tile span 32 sqrt(2) + 2 = 47.3 pixels: with the integer mean and the even start a few of them
    # miss the 48-wide box by a pixel or two and take the work-list path -- rare, and never wrong
    assert worst_miss <= 2 and n_outside < 0.01 * n_readers, (worst_miss, n_outside, n_readers)
    assert worst_tail < n                                                  # the tails are lines of clipped pixels, not areas
