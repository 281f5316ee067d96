"""GPU tier: the copy-engine gradient exchange (csrc/dpexchange.cu, dist.CopyEngineExchange) with world_size 2 and 3.
The ranks are separate processes that share cuda:0 (CUDA IPC works between processes on one device), with gloo carrying
the IPC handles, so the whole protocol -- peer-mapped buffers, 4-byte flag copies, stream memory waits, shard sums -- runs on
a single-GPU box; on a multi-GPU box every rank takes its own device."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(('127.0.0.1', 0)); p = s.getsockname()[1]; s.close(); return p


def _run(rank, world, port, fn, ret):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank % torch.cuda.device_count())
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


def _values(step, rank, shape):
    g = torch.Generator().manual_seed(1000 * step + rank)
    return torch.randn(shape, generator=g)


def _exchange_raw(rank, world):
    from beyond_dof_b200.dist import CopyEngineExchange
    shape = (12, 1, 16, 24, 2)                         # 12 z slices in 4 buckets
    ex = CopyEngineExchange(shape, n_buckets=4)
    errs = []
    side = torch.cuda.Stream()
    for step in range(4):
        mine = _values(step, rank, shape).cuda()
        want = sum(_values(step, r, shape) for r in range(world)) / world
        buckets = []
        # the producer fills the buckets from the top on a side stream, as the adjoint sweep does
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for j in range(4):
                z_hi = 12 - 3 * j
                ex.grad[z_hi - 3:z_hi].copy_(mine[z_hi - 3:z_hi])
                ev = torch.cuda.Event(); ev.record(side)
                buckets.append((z_hi - 3, z_hi, ev))
        ex.exchange(buckets)
        ex.finish()
        got = ex.grad.clone()
        torch.cuda.synchronize()
        errs.append(float((got.cpu() - want).abs().max()))
        dist.barrier()
    # single-bucket form (ready at the call)
    mine = _values(99, rank, shape).cuda()
    ex.grad.copy_(mine)
    ex.exchange(None)
    ex.finish()
    want = sum(_values(99, r, shape) for r in range(world)) / world
    errs.append(float((ex.grad.cpu() - want).abs().max()))
    digest = ex.grad.cpu().numpy().tobytes()
    ex.close()
    return errs, digest


@pytest.mark.parametrize('world', [2, 3])
def test_copy_engine_exchange_is_the_mean_and_identical_on_every_rank(world):
    res = spawn(_exchange_raw, world)
    for errs, _ in res:
        assert max(errs) < 1e-6
    assert all(r[1] == res[0][1] for r in res)          # bit-identical gradients on every rank


def _objective(rank, world):
    # data-parallel FullfieldObjective (forward + loss + adjoint + exchange overlapped with the sweep) against the
    # same fields evaluated one by one in a single process
    from beyond_dof_b200.models import FullfieldObjective
    Z, B, Y, X = 16, 1, 128, 64
    probe = torch.ones((Y, X), dtype=torch.complex64, device='cuda')

    def inputs(r):
        g = torch.Generator().manual_seed(50 + r)
        db = torch.rand((Z, B, Y, X, 2), generator=g) * torch.tensor([4e-4, 4e-5])
        tgt = torch.rand((B, Y, X), generator=g) + 0.5
        return db.cuda(), tgt.cuda()
    db, tgt = inputs(rank)
    obj = FullfieldObjective(db, probe, 5000, 1e-7)
    obj.enable_data_parallel(n_buckets=4, exchange='ce')
    for _ in range(3):
        loss = obj.step_device(tgt)
    torch.cuda.synchronize()
    got = obj.grad.clone()
    ref = torch.zeros_like(got)
    for r in range(world):
        dbr, tr = inputs(r)
        o = FullfieldObjective(dbr, probe, 5000, 1e-7)
        o.step_device(tr)
        ref += o.grad
    ref /= world
    torch.cuda.synchronize()
    err = float((got - ref).norm() / ref.norm())
    dist.barrier()
    obj._ce.close()
    return err


def test_data_parallel_objective_matches_serial_average():
    for err in spawn(_objective, 2):
        assert err < 1e-6


def _tomography(rank, world):
    # data-parallel TomographyObjective: the back-rotation runs in z buckets, each all-reduced on the communication stream under
    # the next one; against the same update with the ranks' minibatches evaluated in one process and averaged
    from beyond_dof_b200.models import TomographyObjective
    n, mb = 64, 2
    g = torch.Generator().manual_seed(7)
    obj0 = torch.rand((n, n, n, 2), generator=g) * torch.tensor([8e-5, 5e-6])
    probe = torch.ones((n, n), dtype=torch.complex64, device='cuda')
    thetas = np.linspace(0.2, 2.9, mb * world)
    prj = [0.9 + 0.1 * torch.rand((mb, n, n), generator=torch.Generator().manual_seed(90 + r)) for r in range(world)]
    tomo = TomographyObjective(obj0.clone().cuda(), probe, 5000, 1e-7, minibatch_size=mb, free_prop_cm=None, propagate_last=True, step_size=1e-7)
    tomo.enable_data_parallel(exchange='nccl', n_buckets=2)        # ('nccl' = torch.distributed all-reduce; the group here is gloo)
    tomo.step(thetas[rank * mb:(rank + 1) * mb], prj[rank].cuda())
    got_grad, got_obj = tomo.grad.clone(), tomo.obj.clone()
    ref = torch.zeros_like(got_grad)
    for r in range(world):
        t = TomographyObjective(obj0.clone().cuda(), probe, 5000, 1e-7, minibatch_size=mb, free_prop_cm=None, propagate_last=True, step_size=1e-7)
        t.loss_and_grad(thetas[r * mb:(r + 1) * mb], prj[r].cuda())
        ref += t.grad
    ref /= world
    torch.cuda.synchronize()
    e_grad = float((got_grad - ref).norm() / ref.norm())
    # every rank holds the same object after the update
    mine = got_obj.cpu()
    gathered = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine)
    same = all(torch.equal(gathered[0], x) for x in gathered)
    return e_grad, same


def test_data_parallel_tomography_buckets_match_serial_average():
    for e_grad, same in spawn(_tomography, 2):
        assert e_grad < 1e-6 and same


def _ptychography(rank, world):
    # data-parallel PtychographyObjective: the window accumulation runs in z buckets, each all-reduced under the next; against
    # the mean of the ranks' gradients evaluated in one process
    from beyond_dof_b200.models import PtychographyObjective
    Z, OY, OX, n = 8, 96, 96, 4
    g = torch.Generator().manual_seed(11)
    obj0 = torch.rand((Z, OY, OX, 2), generator=g) * torch.tensor([3e-4, 3e-5])
    xs = torch.linspace(-1, 1, 64)
    probe = torch.exp(-(xs[:, None] ** 2 + xs[None, :] ** 2) / 0.08).to(torch.complex64)
    pos = [np.array([(32 + 7 * (r * n + i), 40 + 5 * i) for i in range(n)]) for r in range(world)]
    prj = [1.0 + torch.rand((n, 64, 64), generator=torch.Generator().manual_seed(70 + r)) for r in range(world)]
    pty = PtychographyObjective(obj0.clone().cuda(), probe, (64, 64), 5000, 1e-7, n_pos_per_step=n, n_pos_total=n * world, step_size=1e-7)
    pty.enable_data_parallel(n_buckets=4)
    pty.loss_and_grad(pos[rank], prj[rank].cuda())
    torch.cuda.synchronize()
    got = pty.grad.clone()
    ref = torch.zeros_like(got)
    for r in range(world):
        q = PtychographyObjective(obj0.clone().cuda(), probe, (64, 64), 5000, 1e-7, n_pos_per_step=n, n_pos_total=n * world, step_size=1e-7)
        q.loss_and_grad(pos[r], prj[r].cuda())
        ref += q.grad
    ref /= world
    torch.cuda.synchronize()
    return float((got - ref).norm() / ref.norm())


def test_data_parallel_ptychography_buckets_match_serial_average():
    for err in spawn(_ptychography, 2):
        assert err < 1e-6
