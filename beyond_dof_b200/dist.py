"""Data-parallel layer of the reconstruction loops: one process per GPU, torch.distributed for the plumbing.

The reference's only parallelism strategy is data parallel: projection angles
(full field) or (angle, scan position) pairs (ptychography) are sharded over ranks and the object
gradient is all-reduced every optimiser step:

  dataset .shard(hvd.size(), hvd.rank())                 tensorflow_recon/fullfield.py:221-224
  hvd.DistributedOptimizer (allreduce-mean)              tensorflow_recon/fullfield.py:412,444
  each rank takes minibatch_size positions of one angle  cnn_propagator/ptychography.py:292-297
  comm.Allreduce(this_grads, grads); grads / size        cnn_propagator/fullfield.py:348-351, ptychography.py:302-306

Here the exchange is one NCCL all-reduce of the slice-major object gradient [Z, ..., 2] (fp32), optionally
issued bucket by bucket along z on a communication stream while the adjoint sweep is still producing
the remaining slices (a slice's gradient is final once the backward sweep has passed it).
Everything above CopyEngineExchange is host logic: it runs unchanged over gloo on CPU tensors (tests) and over
NCCL on CUDA tensors.  CopyEngineExchange is the production exchange on one NVLink box: the same mean over ranks, moved
by the copy engines between peer-mapped buffers (libbdof, csrc/dpexchange.cu) so that no communication kernel competes
with the persistent sweep kernels for SMs; torch.distributed only carries the 64-byte IPC handles at set-up.
"""
import ctypes

import numpy as np
import torch
import torch.distributed as dist


def world():
    """(rank, world_size) of the default process group, (0, 1) when not initialised
    (the single-process stand-in the reference ships as pseudo.py:3-33)."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_round_robin(n_items, rank=None, world_size=None):
    """tf.data .shard(size, rank) semantics (fullfield.py:221-224): every size-th item starting at rank."""
    r, w = world()
    rank = r if rank is None else rank
    world_size = w if world_size is None else world_size
    return np.arange(rank, n_items, world_size)


def shard_contiguous(n_items, rank=None, world_size=None):
    """Contiguous blocks, as the ptychography driver hands out scan positions of one angle
    (cnn_propagator/ptychography.py:292-297).  The first n_items % world ranks get one extra item."""
    r, w = world()
    rank = r if rank is None else rank
    world_size = w if world_size is None else world_size
    base, extra = divmod(n_items, world_size)
    start = rank * base + min(rank, extra)
    return np.arange(start, start + base + (1 if rank < extra else 0))


def allreduce_gradient(grad, average=True, buckets=None, comm_stream=None):
    """Sum (or mean) of the object gradient over ranks, in place.

    grad: tensor whose FIRST axis is z (slice-major native layout) or any tensor when buckets is None.
    buckets: optional [(z_lo, z_hi, event)] from MultislicePlan.set_gradient_buckets(): each bucket is
    reduced on comm_stream as soon as its event (recorded inside the adjoint sweep) has fired, so the
    exchange overlaps the rest of the sweep.  Returns the list of async work handles (empty when
    world_size == 1); call finish_allreduce() before using the gradient.
    """
    rank, w = world()
    if w == 1:
        return []
    works = []
    # NCCL averages inside the collective (no extra pass over the gradient); gloo only sums
    fused_avg = average and grad.is_cuda and dist.get_backend() == 'nccl'
    op = dist.ReduceOp.AVG if fused_avg else dist.ReduceOp.SUM
    if buckets is None:
        works.append(dist.all_reduce(grad, op=op, async_op=True))
    else:
        for z_lo, z_hi, ev in buckets:
            if comm_stream is not None and grad.is_cuda:
                comm_stream.wait_event(ev)
                with torch.cuda.stream(comm_stream):
                    works.append(dist.all_reduce(grad[z_lo:z_hi], op=op, async_op=True))
            else:
                works.append(dist.all_reduce(grad[z_lo:z_hi], op=op, async_op=True))
    if average and not fused_avg:
        works.append(('scale', 1.0 / w))
    return works


def finish_allreduce(grad, works, comm_stream=None):
    """Wait for the handles returned by allreduce_gradient and apply the 1/world scaling."""
    scale = None
    for wk in works:
        if isinstance(wk, tuple):
            scale = wk[1]
        else:
            wk.wait()
    if comm_stream is not None and grad.is_cuda:
        torch.cuda.current_stream().wait_stream(comm_stream)
    if scale is not None:
        grad.mul_(scale)
    return grad


def allreduce_scalar(value, average=False):
    """Sum (mean) of a Python/0-d scalar over ranks, e.g. the per-rank partial loss
    (cnn_propagator/ptychography.py:338 keeps a per-rank loss table)."""
    rank, w = world()
    if w == 1:
        return float(value)
    backend = dist.get_backend()
    dev = torch.device('cuda', torch.cuda.current_device()) if backend == 'nccl' else torch.device('cpu')
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item()) / (w if average else 1)


def broadcast_object_(tensor, src=0):
    """hvd.broadcast_global_variables(0) (fullfield.py:481): every rank starts from rank 0's object."""
    rank, w = world()
    if w > 1:
        dist.broadcast(tensor, src=src)
    return tensor


def data_parallel_step(n_items, local_loss_and_grad, grad_buffer, shard='contiguous', average=True):
    """One data-parallel loss/gradient evaluation.

    n_items: number of work items of this step (scan positions of one angle, or angles of a minibatch).
    local_loss_and_grad(indices, grad_buffer) -> local loss: evaluates the model on this rank's items,
        ACCUMULATING the object gradient into grad_buffer (pre-zeroed here) and returning the sum of the
        per-item losses (so that sum over ranks = loss over all items).
    Returns (loss_total, grad_buffer) with the gradient summed (average=False) or averaged over ranks as
    the reference does (grads / size, cnn_propagator/ptychography.py:306).
    """
    idx = shard_contiguous(n_items) if shard == 'contiguous' else shard_round_robin(n_items)
    grad_buffer.zero_()
    local = local_loss_and_grad(idx, grad_buffer) if len(idx) else 0.0
    works = allreduce_gradient(grad_buffer, average=average)
    finish_allreduce(grad_buffer, works)
    return allreduce_scalar(float(local)), grad_buffer


def auto_sm_reserve(batch, ny, nx, n_sm=148, max_reserve=24):
    """Largest number of SMs (<= max_reserve) the persistent sweep kernels can leave to NCCL without adding a round of tiles
    (8 lines per tile up to 2048-long lines, 4 for 4096, 16 below 256): e.g. 2048^2, B = 1: the y kernels run 256 column
    tiles in two rounds on anything from 128 to 148 SMs, so 20 SMs are free for the collective at no cost there."""
    def rounds(lines, length, ctas):
        lpc = 4 if length >= 4096 else (8 if length >= 256 else 16)
        tiles = max(1, batch * lines // lpc)
        return -(-tiles // ctas)
    best = 0
    for r in range(0, max_reserve + 1):
        if rounds(ny, nx, n_sm - r) == rounds(ny, nx, n_sm) and rounds(nx, ny, n_sm - r) == rounds(nx, ny, n_sm):
            best = r
    return best


def pick_exchange():
    """'ce' or 'nccl' for the gradient exchange of this process group.  Measured on B200 / NVLink 5 (8.6 GB gradient,
    tools/dp_diag.py): between TWO GPUs one copy-engine stream sustains ~550 GB/s and the exchange costs no SMs (step 29.9 ms
    against 34.7 ms with NCCL); with three or more peers concurrent copy-engine transfers do not add up (4 GPUs: 340 GB/s
    aggregate, exchange alone 36 ms against NCCL's 19 ms), so NCCL (which drives NVLink from SM kernels) is used there."""
    if not (dist.is_available() and dist.is_initialized()):
        return 'nccl'
    return 'ce' if dist.get_world_size() == 2 else 'nccl'        # 'hybrid' (NCCL reduce-scatter + copy-engine all-gather) is opt-in


class _DeviceBuffer:
    """CUDA-array-interface view of memory owned by libbdof (torch.as_tensor wraps it without a copy)."""

    def __init__(self, ptr, n_float32, owner):
        self.__cuda_array_interface__ = {'shape': (int(n_float32),), 'typestr': '<f4', 'data': (int(ptr), False), 'version': 3,
                                         'strides': None}
        self._owner = owner


class CopyEngineExchange:
    """Mean of the object gradient over the ranks of one NVLink box, bucket by bucket, with the copy engines.

    ex = CopyEngineExchange(grad_shape, n_buckets)      # collective: every rank, after init_process_group (any backend)
    ex.grad                                             # float32 CUDA tensor the adjoint must write its gradient into
    ex.exchange(buckets)                                # buckets = MultislicePlan.set_gradient_buckets(n): [(z_lo, z_hi, event)]
                                                        #   (or None: one bucket, ready once the current stream reaches this call)
    ex.finish()                                         # current stream waits for the averaged gradient

    Replaces hvd.DistributedOptimizer's allreduce (tensorflow_recon/fullfield.py:412) / comm.Allreduce + grads / size
    (cnn_propagator/fullfield.py:348-351).  Every rank ends with bit-identical values (a shard is summed once, by its owner,
    in rank order, and broadcast)."""

    def __init__(self, grad_shape, n_buckets=8, group=None, device=None, gather_only=False):
        from .capi import lib, check
        self._lib, self._check = lib, check
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        n = int(np.prod(grad_shape))
        self.n_buckets = int(n_buckets)
        self._h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            check(lib.bdof_dp_create(ctypes.byref(self._h), self.rank, self.world, n * 4, self.n_buckets, 1 if gather_only else 0))
            hb = lib.bdof_dp_handle_bytes()
            mine = ctypes.create_string_buffer(hb)
            check(lib.bdof_dp_export(self._h, mine))
            handles = [None] * self.world
            dist.all_gather_object(handles, bytes(mine.raw), group=group)
            blob = b''.join(handles)
            assert len(blob) == hb * self.world
            check(lib.bdof_dp_connect(self._h, ctypes.c_char_p(blob)))
            ptr = ctypes.c_void_p()
            check(lib.bdof_dp_grad_ptr(self._h, ctypes.byref(ptr)))
            self.grad = torch.as_tensor(_DeviceBuffer(ptr.value, n, self), device=self.device).view(*grad_shape)
        self._ready = torch.cuda.Event()
        dist.barrier(group=group)             # nobody pushes into a peer that has not opened the handles yet... or freed them
        self._group = group

    def exchange(self, buckets=None):
        g = self.grad
        if buckets is None:
            self._ready.record(torch.cuda.current_stream(self.device))
            self._check(self._lib.bdof_dp_bucket(self._h, 0, g.numel() * 4, ctypes.c_void_p(self._ready.cuda_event)))
            return
        row = g[0].numel() * 4                 # bytes per z slice
        for z_lo, z_hi, ev in buckets:
            self._check(self._lib.bdof_dp_bucket(self._h, z_lo * row, (z_hi - z_lo) * row, ctypes.c_void_p(ev.cuda_event)))

    def reduce_scatter_gather(self, buckets, comm_stream):
        """Three or more GPUs: per bucket an in-place NCCL reduce-scatter (mean) on comm_stream leaves rank r with shard r of
        the bucket, which the copy engines then broadcast peer by peer (bdof_dp_gather) -- the SM kernels of NCCL run for
        half the traffic of an all-reduce and the all-gather half costs no SMs."""
        g = self.grad
        row = g[0].numel() * 4
        if not hasattr(self, '_rs_events'):
            self._rs_events = [torch.cuda.Event() for _ in range(self.n_buckets)]
        for j, (z_lo, z_hi, ev) in enumerate(buckets):
            seg = g[z_lo:z_hi].view(-1)
            shard = seg.numel() // self.world
            comm_stream.wait_event(ev)
            with torch.cuda.stream(comm_stream):
                wk = dist.reduce_scatter_tensor(seg[self.rank * shard:(self.rank + 1) * shard], seg, op=dist.ReduceOp.AVG,
                                                group=self._group, async_op=True)
                wk.wait()
                self._rs_events[j].record(comm_stream)
            self._check(self._lib.bdof_dp_gather(self._h, z_lo * row, (z_hi - z_lo) * row, ctypes.c_void_p(self._rs_events[j].cuda_event)))

    def finish(self):
        self._check(self._lib.bdof_dp_finish(self._h, ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)))
        return self.grad

    def close(self):
        if getattr(self, '_h', None) is not None and self._h.value:
            torch.cuda.synchronize(self.device)
            if dist.is_initialized():
                dist.barrier(group=self._group)   # peers may still be copying into / out of my buffers
            self.grad = None
            self._lib.bdof_dp_destroy(self._h)
            self._h = ctypes.c_void_p()
