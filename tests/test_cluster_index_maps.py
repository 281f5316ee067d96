"""CPU tier: the index algebra of the cluster-resident kernels (csrc/clusterfft.cuh), restated in NumPy.

A field of N x N lives in the registers of C CTAs, R = N / C lines each, T threads per line, E = N / T elements per thread.
x steps: l = tid // T, t = tid % T, element q = (row c R + l, column t + T q); y steps: l = tid % R, t = tid // R, element q =
(row t + T q, column c R + l).  cluster_transpose pushes element q of every thread to CTA (T q) // R at a buffer offset that the
receiving thread reads back as ITS element of the other layout.  The test checks, for the shipped configurations, that
 - the target CTA of an element depends on q only (it is a compile-time constant in the kernel),
 - every buffer slot of every CTA is written exactly once (N * R float2 = the byte count the receiver's mbarrier expects),
 - what a thread reads is the field element its new layout says it holds,
 - a warp's pushes are contiguous runs (two of 128 bytes or one of 256 for T = 16; four of 64 bytes for T = 8)."""
import numpy as np
import pytest


def maps(N, C, T, col, cta, tid):
    R, E = N // C, N // T
    l, t = (tid % R, tid // R) if col else (tid // T, tid % T)
    q = np.arange(E)
    y = (t + T * q) if col else np.full(E, cta * R + l)
    x = np.full(E, cta * R + l) if col else (t + T * q)
    return l, t, y, x


@pytest.mark.parametrize('cfg', [(256, 8, 16, 16), (128, 4, 8, 16)])        # N, C, T, R1 (LineCfg<256,16,16,16,1>, LineCfg<128,8,16,8,1>)
@pytest.mark.parametrize('from_col', [False, True])
def test_cluster_transpose_is_a_bijection_onto_the_other_layout(cfg, from_col):
    N, C, T, R1 = cfg
    R, E, XP = N // C, N // T, N + N // R1
    threads = R * T
    assert T <= R
    field = np.arange(N * N).reshape(N, N)                                # element id = y * N + x
    buf_elems = max(R * XP, N * R)
    recv = -np.ones((C, buf_elems), dtype=np.int64)
    writes = np.zeros((C, buf_elems), dtype=np.int64)
    for cta in range(C):
        for tid in range(threads):
            l, t, y, x = maps(N, C, T, from_col, cta, tid)
            for q in range(E):
                peer, off = (T * q) // R, (T * q) % R
                assert t + off < R
                if from_col:        # -> x-step buffer of the row's owner: [row % R][column], pitch XP
                    assert (t + T * q) // R == peer
                    slot = (t + off) * XP + cta * R + l
                else:               # -> y-step buffer of the column's owner: [row][column % R]
                    assert (t + T * q) // R == peer
                    slot = (cta * R + l) * R + t + off
                recv[peer, slot] = field[y[q], x[q]]
                writes[peer, slot] += 1
    assert writes.max() == 1 and (writes.sum(axis=1) == N * R).all()       # 64 KB per CTA and step at N = 256
    for cta in range(C):
        for tid in range(threads):
            l, t, y, x = maps(N, C, T, not from_col, cta, tid)             # the layout after the transpose
            for q in range(E):
                slot = (l * XP + t + T * q) if from_col else ((t + T * q) * R + l)
                assert recv[cta, slot] == field[y[q], x[q]]


@pytest.mark.parametrize('cfg', [(256, 8, 16, 16), (128, 4, 8, 16)])
def test_cluster_pushes_of_a_warp_are_contiguous_runs(cfg):
    N, C, T, R1 = cfg
    R, XP = N // C, N + N // R1
    for from_col in (False, True):
        for q in (0, 1, 5):
            off = (T * q) % R
            slots = []
            for tid in range(32):                                          # warp 0 of CTA 1
                l, t = (tid % R, tid // R) if from_col else (tid // T, tid % T)
                slots.append((t + off) * XP + 1 * R + l if from_col else (1 * R + l) * R + t + off)
            runs = np.split(np.array(slots), np.where(np.diff(slots) != 1)[0] + 1)
            if from_col:
                assert len(runs) == 1 and len(runs[0]) == 32               # 32 adjacent columns of one row: 256 bytes
            else:
                assert all(len(r) == T for r in runs) and len(runs) == 32 // T
