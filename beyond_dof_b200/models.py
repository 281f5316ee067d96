"""Forward models + loss + gradient for the two reconstruction drivers of the reference.

  fullfield_loss_and_grad <- rotate_and_project_batch  tensorflow_recon/fullfield.py:92-116
                             calculate_loss            cnn_propagator/fullfield.py:93-121
                             + autodiff gradient       tensorflow_recon/fullfield.py:428-435
  ptycho_loss_and_grad    <- rotate_and_project        tensorflow_recon/ptychography.py:37-97
                             calculate_loss            cnn_propagator/ptychography.py:30-81
                             + autodiff gradient       cnn_propagator/ptychography.py:248

The reference differentiates with TF / autograd; here the gradient is the hand-written adjoint
of the multislice chain (libbdof), returned as the 2-tuple (g_delta, g_beta) that
`loss_grad = grad(calculate_loss, [0, 1])` returns in the reference.
Object rotation follows the cnn_propagator drivers: the nearest-neighbour lookup of save_rotation_lookup /
apply_rotation (cnn_propagator/util.py:295-402), applied on the GPU to the native slice-major
object, and its transpose for the gradient.  (tf.contrib.image.rotate of the TF driver is not reproduced.)
The ptychography model rotates the same way (calculate_loss, cnn_propagator/ptychography.py:30-34).
"""
import ctypes

import numpy as np
import torch

from .capi import lib, check
from .plan import MultislicePlan, _ptr
from .propagation import _cached_plan, _device, _to_dev, _probe_c64
from . import rotation as _rot


def total_variation_3d(arr):
    """cnn_propagator/util.py:61-70 (periodic first differences, L1)."""
    res = torch.sum(torch.abs(torch.roll(arr, 1, 0) - arr))
    res = res + torch.sum(torch.abs(torch.roll(arr, 1, 1) - arr))
    res = res + torch.sum(torch.abs(torch.roll(arr, 1, 2) - arr))
    return res


def _tv_grad(arr):
    g = torch.zeros_like(arr)
    for ax in range(3):
        s = torch.sign(torch.roll(arr, 1, ax) - arr)
        g += torch.roll(s, -1, ax) - s
    return g


def _scalar_theta(theta):
    th = np.atleast_1d(np.asarray(theta.cpu() if isinstance(theta, torch.Tensor) else theta, dtype=np.float64))
    if th.size != 1:
        raise ValueError('the ptychography model evaluates one rotation angle per call (ptychography.py:292-297)')
    return float(th[0])


def fullfield_loss_and_grad(obj_delta, obj_beta, theta_batch, prj_batch, probe_real, probe_imag, energy_ev, psize_cm,
                            free_prop_cm=None, alpha_d=None, alpha_b=None, gamma=0.0, propagate_last=True,
                            want_grad=True, rotation='nearest'):
    """loss = mean((|psi_exit| - |prj|)^2) [+ alpha_d |delta|_1 + alpha_b |beta|_1 + gamma TV(delta)]
    over a minibatch of projection angles; returns (loss, (g_delta, g_beta), exit_wave).

    obj_delta, obj_beta: [Y,X,Z] float32; prj_batch: [B,Y,X] complex or magnitude.
    propagate_last=True follows the TF driver (fullfield.py:107-109); False the NumPy simulator.
    rotation: 'nearest' = the lookup tables of the cnn_propagator driver (apply_rotation), 'bilinear' = tf.contrib.image.rotate
    of the TF driver (fullfield.py:96).
    """
    th = np.atleast_1d(np.asarray(theta_batch.cpu() if isinstance(theta_batch, torch.Tensor) else theta_batch, dtype=np.float64))
    B = len(th)
    dev = _device()
    od = _to_dev(obj_delta, torch.float32)
    ob = _to_dev(obj_beta, torch.float32)
    Y, X, Z = od.shape
    key = ('ff', (B, Y, X, Z), float(energy_ev), float(psize_cm), free_prop_cm, propagate_last, dev.index)
    plan = _cached_plan(key, lambda: MultislicePlan(Y, X, B, Z, energy_ev, psize_cm, free_prop_cm=free_prop_cm,
                                                     propagate_last=propagate_last, store_slices=True))
    # every batch element sees the object rotated by its angle (apply_rotation, cnn_propagator/fullfield.py:95-100)
    obj_db = pack_object(od, ob)
    db = torch.empty((Z, B, Y, X, 2), dtype=torch.float32, device=dev)
    if rotation not in ('nearest', 'bilinear'):
        raise ValueError("rotation must be 'nearest' or 'bilinear'")
    tabs = [_rot.device_table([Y, X, Z], t, dev) for t in th] if rotation == 'nearest' else None
    for b in range(B):
        if tabs is not None:
            _rot.rotate_db(obj_db, tabs[b], out=db[:, b])
        else:
            _rot.rotate_db_bilinear(obj_db, th[b], out=db[:, b])
    probe = _probe_c64(probe_real, probe_imag, (Y, X))
    plan.set_t_stash(db if want_grad else None)     # the in-place adjoint finds t_i where it will write the gradient
    exit_wave = plan.forward(db, probe)
    prj = _to_dev(prj_batch, torch.complex64 if (isinstance(prj_batch, torch.Tensor) and prj_batch.is_complex())
                  or np.iscomplexobj(prj_batch) else torch.float32)
    target = prj.abs().to(torch.float32)
    loss, g_exit = plan.loss_mag(exit_wave, target, want_grad=want_grad)
    loss = loss.clone()
    g_d = g_b = None
    g_obj = None
    if want_grad:
        plan.adjoint(db, g_exit)                       # db now holds (dL/ddelta, dL/dbeta) per ROTATED batch element
        g_obj = torch.zeros_like(obj_db)
        for b in range(B):
            if tabs is not None:
                _rot.rotate_db_adjoint(db[:, b], tabs[b], g_obj)
            else:
                _rot.rotate_db_bilinear_adjoint(db[:, b], th[b], g_obj)
    if (alpha_d is not None and alpha_d != 0) or (alpha_b is not None and alpha_b != 0) or gamma:
        # L1 / TV regularisers (fullfield.py:389-396) and their gradients in one fused pass over the native object
        loss = _rot.regularizers(obj_db, g_obj, loss, alpha_d, alpha_b, gamma)
    if want_grad:
        g_d, g_b = unpack_object(g_obj)
    return loss, (g_d, g_b), exit_wave


class _HostPipe:
    """Pinned staging + worker threads for moving pageable NumPy volumes across PCIe at more than the ~6-10 GB/s a single
    cudaMemcpy from pageable memory sustains: worker threads copy row chunks into (out of) pinned buffers (NumPy releases the
    GIL for plain copies), a copy stream moves them with cudaMemcpyAsync, and the layout conversion (bdof_pack_db_rows /
    bdof_unpack_db_rows) runs per chunk on that stream, all overlapped chunk by chunk."""
    _inst = {}

    def __init__(self, dev, chunk_bytes=128 << 20, depth=3, threads=8):
        from concurrent.futures import ThreadPoolExecutor
        self.dev = dev
        self.chunk_bytes = int(chunk_bytes)
        self.depth = depth
        self.pool = ThreadPoolExecutor(max_workers=threads)
        self.threads = threads
        self.stream = torch.cuda.Stream(device=dev)
        n = self.chunk_bytes // 4
        self.pin = [[torch.empty(n, dtype=torch.float32).pin_memory() for _ in range(2)] for _ in range(depth)]
        self.devbuf = [[torch.empty(n, dtype=torch.float32, device=dev) for _ in range(2)] for _ in range(depth)]
        self.events = [torch.cuda.Event() for _ in range(depth)]

    @classmethod
    def get(cls, dev):
        key = (dev.index,)
        if key not in cls._inst:
            cls._inst[key] = cls(dev)
        return cls._inst[key]

    def _par_copy(self, dst, src):
        """dst[:] = src (flat float32 NumPy views of equal length) with all worker threads"""
        n = dst.shape[0]
        step = -(-n // self.threads)
        futs = [self.pool.submit(np.copyto, dst[i:i + step], src[i:i + step]) for i in range(0, n, step)]
        for f in futs:
            f.result()

    def upload_packed(self, od, ob, db):
        """od, ob: contiguous float32 NumPy [Y,X,Z]; db: [Z,1,Y,X,2] CUDA tensor, filled in the native layout."""
        Y, X, Z = od.shape
        rows_per = max(1, self.chunk_bytes // (X * Z * 4))
        fd, fb = od.reshape(-1), ob.reshape(-1)
        main = torch.cuda.current_stream(self.dev)
        self.stream.wait_stream(main)
        k = 0
        for r0 in range(0, Y, rows_per):
            nr = min(rows_per, Y - r0)
            n = nr * X * Z
            slot = k % self.depth
            if k >= self.depth:
                self.events[slot].synchronize()                # the device has consumed this slot's pinned buffers
            pd, pb = self.pin[slot]
            self._par_copy(pd.numpy()[:n], fd[r0 * X * Z:r0 * X * Z + n])
            self._par_copy(pb.numpy()[:n], fb[r0 * X * Z:r0 * X * Z + n])
            with torch.cuda.stream(self.stream):
                dd, dbb = self.devbuf[slot]
                dd[:n].copy_(pd[:n], non_blocking=True)
                dbb[:n].copy_(pb[:n], non_blocking=True)
                check(lib.bdof_pack_db_rows(_ptr(dd), _ptr(dbb), _ptr(db), Y, r0, nr, X, Z, ctypes.c_void_p(self.stream.cuda_stream)))
                self.events[slot].record(self.stream)
            k += 1
        main.wait_stream(self.stream)

    def download_unpacked(self, db, out_d, out_b):
        """db: [Z,1,Y,X,2] CUDA tensor -> out_d, out_b contiguous float32 NumPy [Y,X,Z]."""
        Y, X, Z = out_d.shape
        rows_per = max(1, self.chunk_bytes // (X * Z * 4))
        fd, fb = out_d.reshape(-1), out_b.reshape(-1)
        main = torch.cuda.current_stream(self.dev)
        self.stream.wait_stream(main)
        chunks = [(r0, min(rows_per, Y - r0)) for r0 in range(0, Y, rows_per)]

        def issue(k):
            r0, nr = chunks[k]
            n = nr * X * Z
            slot = k % self.depth
            with torch.cuda.stream(self.stream):
                dd, dbb = self.devbuf[slot]
                check(lib.bdof_unpack_db_rows(_ptr(db), _ptr(dd), _ptr(dbb), Y, r0, nr, X, Z, ctypes.c_void_p(self.stream.cuda_stream)))
                pd, pb = self.pin[slot]
                pd[:n].copy_(dd[:n], non_blocking=True)
                pb[:n].copy_(dbb[:n], non_blocking=True)
                self.events[slot].record(self.stream)
        for k in range(min(self.depth, len(chunks))):
            issue(k)
        for k, (r0, nr) in enumerate(chunks):
            n = nr * X * Z
            slot = k % self.depth
            self.events[slot].synchronize()
            pd, pb = self.pin[slot]
            self._par_copy(fd[r0 * X * Z:r0 * X * Z + n], pd.numpy()[:n])
            self._par_copy(fb[r0 * X * Z:r0 * X * Z + n], pb.numpy()[:n])
            if k + self.depth < len(chunks):
                issue(k + self.depth)
        main.wait_stream(self.stream)


def fullfield_loss_and_grad_host(obj_delta, obj_beta, prj_batch, probe_real, probe_imag, energy_ev, psize_cm, free_prop_cm=None,
                                 propagate_last=True, out=None):
    """The cnn_propagator driver's call loss_grad(obj_delta, obj_beta, ...) (cnn_propagator/fullfield.py:329,346) for ONE
    un-rotated field with everything in HOST memory: NumPy delta/beta [Y,X,Z] in, loss and NumPy gradients out.  The volumes
    cross PCIe through the pinned, multi-threaded chunk pipeline of _HostPipe (H2D + layout conversion overlapped; gradient
    layout conversion + D2H overlapped); the multislice itself runs in place on the packed object.
    out: optional (g_delta, g_beta) float32 NumPy arrays to fill."""
    dev = _device()
    od = np.ascontiguousarray(obj_delta, dtype=np.float32)
    ob = np.ascontiguousarray(obj_beta, dtype=np.float32)
    Y, X, Z = od.shape
    key = ('ffh', (1, Y, X, Z), float(energy_ev), float(psize_cm), free_prop_cm, propagate_last, dev.index)
    plan = _cached_plan(key, lambda: MultislicePlan(Y, X, 1, Z, energy_ev, psize_cm, free_prop_cm=free_prop_cm,
                                                     propagate_last=propagate_last, store_slices=True))
    plan.use_current_stream()
    pipe = _HostPipe.get(dev)
    db = torch.empty((Z, 1, Y, X, 2), dtype=torch.float32, device=dev)
    pipe.upload_packed(od, ob, db)
    probe = _probe_c64(probe_real, probe_imag, (Y, X))
    plan.set_t_stash(db)
    exit_wave = plan.forward(db, probe)
    is_cplx = (isinstance(prj_batch, torch.Tensor) and prj_batch.is_complex()) or np.iscomplexobj(prj_batch)
    target = _to_dev(prj_batch, torch.complex64 if is_cplx else torch.float32).abs().to(torch.float32).reshape(1, Y, X)
    loss, g_exit = plan.loss_mag(exit_wave, target)
    plan.adjoint(db, g_exit)
    g_d, g_b = out if out is not None else (np.empty((Y, X, Z), np.float32), np.empty((Y, X, Z), np.float32))
    pipe.download_unpacked(db, g_d, g_b)
    plan.set_t_stash(None)
    return float(loss.item()), (g_d, g_b)


def pack_object(obj_delta, obj_beta):
    """[Y,X,Z] delta/beta -> slice-major interleaved object [Z,Y,X,2] on the current device."""
    od = _to_dev(obj_delta, torch.float32).contiguous()
    ob = _to_dev(obj_beta, torch.float32).contiguous()
    Y, X, Z = od.shape
    out = torch.empty((Z, Y, X, 2), dtype=torch.float32, device=od.device)
    check(lib.bdof_pack_db(_ptr(od), _ptr(ob), _ptr(out), 1, Y, X, Z, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    return out


def unpack_object(db_obj):
    Z, Y, X, _ = db_obj.shape
    d = torch.empty((Y, X, Z), dtype=torch.float32, device=db_obj.device)
    b = torch.empty((Y, X, Z), dtype=torch.float32, device=db_obj.device)
    check(lib.bdof_unpack_db(_ptr(db_obj), _ptr(d), _ptr(b), 1, Y, X, Z, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    return d, b


def ptycho_loss_and_grad(obj_delta, obj_beta, theta, probe_pos_batch, prj_batch, probe_real, probe_imag, probe_size,
                         energy_ev, psize_cm, n_dp_batch=None, scale_by_npos=True, n_pos_total=None, want_grad=True,
                         db_obj=None, grad_obj_out=None, rotation='nearest', deterministic=True):
    """Ptychography forward model + loss + gradient for one rotation angle theta (radians; the object is rotated with the
    reference's nearest-neighbour table first, cnn_propagator/ptychography.py:32-34, and the gradient rotated back).

    probe_pos_batch: integer (y, x) scan positions; the window of size probe_size starts at
    pos - int(probe_size/2) and is zero-padded outside the object (ptychography.py:45-76).
    prj_batch: [n_pos, py, px] measured far-field amplitudes (complex or magnitude).
    loss = mean((|Psi| - |prj|)^2) (* n_pos_total if scale_by_npos, ptychography.py:94).
    Returns (loss, (g_delta, g_beta)); with db_obj / grad_obj_out (slice-major [Z,Y,X,2]) the
    object and its gradient stay in the native layout and (loss, grad_obj_out) is returned.
    deterministic: accumulate the window gradients as an ordered gather (bdof_patch_gather_add: bit-reproducible) instead of
    fp32 atomics (bdof_patch_scatter_add).
    """
    th = _scalar_theta(theta)
    dev = _device()
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    native = db_obj is not None
    if not native:
        db_obj = pack_object(obj_delta, obj_beta)
    db_obj = db_obj.contiguous()
    Z, OY, OX, _ = db_obj.shape
    tab = None
    obj_unrot = db_obj
    if rotation not in ('nearest', 'bilinear'):
        raise ValueError("rotation must be 'nearest' or 'bilinear'")
    if th != 0.0 and rotation == 'nearest':
        tab = _rot.device_table([OY, OX, Z], th, dev)
        db_obj = _rot.rotate_db(db_obj, tab)
    elif th != 0.0:
        db_obj = _rot.rotate_db_bilinear(db_obj.contiguous(), th)        # tf_rotate(..., 'BILINEAR'), ptychography.py:39
    py, px = int(probe_size[0]), int(probe_size[1])
    pos = np.asarray(probe_pos_batch.cpu() if isinstance(probe_pos_batch, torch.Tensor) else probe_pos_batch).astype(np.int64)
    n = len(pos)
    half = (np.array([py, px]) / 2).astype('int')                 # ptychography.py:184
    origin = torch.as_tensor((pos - half[None, :]).astype(np.int32)).to(dev).contiguous()
    key = ('pty', (n, py, px, Z), float(energy_ev), float(psize_cm), dev.index)
    plan = _cached_plan(key, lambda: MultislicePlan(py, px, n, Z, energy_ev, psize_cm, free_prop_cm='inf',
                                                     propagate_last=True, store_slices=True))
    patches = torch.empty((Z, n, py, px, 2), dtype=torch.float32, device=dev)
    probe = _probe_c64(probe_real, probe_imag, (py, px))
    is_cplx = (isinstance(prj_batch, torch.Tensor) and prj_batch.is_complex()) or np.iscomplexobj(prj_batch)
    target = _to_dev(prj_batch, torch.complex64 if is_cplx else torch.float32).abs().to(torch.float32)
    scale = float(n_pos_total if n_pos_total is not None else n) if scale_by_npos else 1.0
    windowed = plan.is_resident()
    if windowed:
        # 64 x 64 probes: the resident kernels read the object straight through the windows; `patches` only holds the
        # transmission stash and then the per-window gradients
        plan.set_windows((Z, OY, OX), origin)
        plan.set_t_stash(patches if want_grad else None)
        exit_wave = plan.forward(db_obj, probe)
    else:
        check(lib.bdof_patch_gather(_ptr(db_obj), Z, OY, OX, _ptr(origin), n, py, px, _ptr(patches), st))
        plan.set_t_stash(patches if want_grad else None)
        exit_wave = plan.forward(patches, probe)
    loss, g_exit = plan.loss_mag(exit_wave, target, want_grad=want_grad, loss_scale=scale)
    loss = loss.clone()
    if not want_grad:
        return loss, None
    if windowed:
        plan.adjoint(db_obj, g_exit, grad_out=patches)
    else:
        plan.adjoint(patches, g_exit)
    if grad_obj_out is None:
        grad_obj_out = torch.zeros_like(db_obj)
    scatter = lib.bdof_patch_gather_add if deterministic else lib.bdof_patch_scatter_add
    if db_obj is obj_unrot:
        check(scatter(_ptr(patches), Z, OY, OX, _ptr(origin), n, py, px, _ptr(grad_obj_out), st))
    else:
        g_rot = torch.zeros_like(db_obj)
        check(scatter(_ptr(patches), Z, OY, OX, _ptr(origin), n, py, px, _ptr(g_rot), st))
        if tab is not None:
            _rot.rotate_db_adjoint(g_rot, tab, grad_obj_out)
        else:
            _rot.rotate_db_bilinear_adjoint(g_rot, th, grad_obj_out)
    if native:
        return loss, grad_obj_out
    return loss, unpack_object(grad_obj_out)


def ptycho_position_losses(obj_delta, obj_beta, theta, probe_pos, prj, probe_real, probe_imag, probe_size, energy_ev,
                           psize_cm, n_dp_batch=256, db_obj=None):
    """Loss of every scan position on its own, mean((|Psi_j| - |prj_j|)^2) -- the loss table of the reference's "dynamic
    dropping" pass (cnn_propagator/ptychography.py:323-342, one calculate_loss call per position there; batched here)."""
    pos = np.asarray(probe_pos.cpu() if isinstance(probe_pos, torch.Tensor) else probe_pos).astype(np.int64)
    if db_obj is None:
        db_obj = pack_object(obj_delta, obj_beta)
    out = []
    for s0 in range(0, len(pos), n_dp_batch):
        sl = slice(s0, min(s0 + n_dp_batch, len(pos)))
        ex = _ptycho_exit_waves(db_obj, theta, pos[sl], probe_real, probe_imag, probe_size, energy_ev, psize_cm)
        p = prj[sl]
        is_cplx = (isinstance(p, torch.Tensor) and p.is_complex()) or np.iscomplexobj(p)
        target = _to_dev(p, torch.complex64 if is_cplx else torch.float32).abs().to(torch.float32)
        out.append(((ex.abs() - target) ** 2).mean(dim=(1, 2)))
    return torch.cat(out)


def dynamic_dropping(loss_table, dropping_threshold=8e-5):
    """Indices of the scan positions to KEEP: the reference drops those whose loss is below the threshold
    (ptychography.py:340-342; its np.delete result is discarded there, i.e. the shipped code drops nothing)."""
    lt = loss_table.detach().cpu().numpy() if isinstance(loss_table, torch.Tensor) else np.asarray(loss_table)
    return np.where(lt >= dropping_threshold)[0]


def _ptycho_exit_waves(db_obj, theta, pos, probe_real, probe_imag, probe_size, energy_ev, psize_cm):
    dev = _device()
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    th = _scalar_theta(theta)
    Z, OY, OX, _ = db_obj.shape
    if th != 0.0:
        db_obj = _rot.rotate_db(db_obj, _rot.device_table([OY, OX, Z], th, dev))
    py, px = int(probe_size[0]), int(probe_size[1])
    n = len(pos)
    half = (np.array([py, px]) / 2).astype('int')
    origin = torch.as_tensor((pos - half[None, :]).astype(np.int32)).to(dev).contiguous()
    key = ('ptyf', (n, py, px, Z), float(energy_ev), float(psize_cm), dev.index)
    plan = _cached_plan(key, lambda: MultislicePlan(py, px, n, Z, energy_ev, psize_cm, free_prop_cm='inf',
                                                     propagate_last=True, store_slices=False))
    patches = torch.empty((Z, n, py, px, 2), dtype=torch.float32, device=dev)
    check(lib.bdof_patch_gather(_ptr(db_obj), Z, OY, OX, _ptr(origin), n, py, px, _ptr(patches), st))
    return plan.forward(patches, _probe_c64(probe_real, probe_imag, (py, px)))


class FullfieldObjective:
    """One full-field reconstruction step with the object resident on the GPU in the native
    slice-major layout -- what the optimisation loop of reconstruct_fullfield calls every minibatch
    (tensorflow_recon/fullfield.py:497-556: feed the projection batch, get loss + gradient).

    db_obj: [Z,B,Y,X,2] float32 CUDA tensor (delta, beta) of the (rotated) object per batch element.
    step(prj_mag_host) copies this step's measured projection magnitudes host->device, runs
    forward + loss + adjoint into self.grad and returns the loss as a Python float (one D2H read).
    """

    def __init__(self, db_obj, probe, energy_ev, psize_cm, free_prop_cm=None, propagate_last=False, in_place=False):
        Z, B, Y, X, _ = db_obj.shape
        self.plan = MultislicePlan(Y, X, B, Z, energy_ev, psize_cm, free_prop_cm=free_prop_cm,
                                   propagate_last=propagate_last, store_slices=True, device=db_obj.device)
        self.db = db_obj
        self.probe = probe.to(db_obj.device, torch.complex64).contiguous()
        # in_place: the adjoint overwrites db_obj with the gradient (halves the footprint: 4096^2 x 512 fits in 180 GB)
        self.in_place = in_place
        self.grad = db_obj if in_place else torch.empty_like(db_obj)
        self.target = torch.empty((B, Y, X), dtype=torch.float32, device=db_obj.device)
        self.exit = torch.empty((B, Y, X), dtype=torch.complex64, device=db_obj.device)
        self.loss_host = torch.empty((), dtype=torch.float64).pin_memory()
        # the forward leaves the transmission of every slice where the adjoint will write that slice's gradient
        self.plan.set_t_stash(self.grad)

    def enable_data_parallel(self, n_buckets=None, exchange='auto', sm_reserve=0):
        """Average the object gradient over the ranks of the default process group every step, bucket by
        bucket along z while the adjoint sweep is still running.  exchange='ce': copy engines over NVLink peer
        memory (dist.CopyEngineExchange; the gradient then lives in the exchange's exportable buffer);
        'nccl': NCCL all-reduce on a communication stream; 'hybrid': NCCL reduce-scatter + copy-engine all-gather;
        'auto': dist.pick_exchange().
        sm_reserve: SMs the persistent sweep kernels leave to the NCCL kernels (they own every SM otherwise, one 226 KB CTA
        each); -1 = the largest number that does not add a round of tiles to the sweep kernels for this field shape
        (dist.auto_sm_reserve) whenever the exchange runs NCCL kernels.  Default 0: measured on 8 x B200 (2048^2 x 256, one field
        per exchange) the step takes 41.0 ms with 0 and 41.6 ms with 20 reserved SMs -- the collective is not SM-starved, the
        sweep kernels slow down while 8.6 GB stream through the L2 their field lives in."""
        from . import dist as bdist
        self._dp = bdist
        if exchange == 'auto':
            exchange = bdist.pick_exchange()
        if n_buckets is None:
            n_buckets = 16 if exchange == 'ce' else 8          # measured optima on 2 x B200 (2048^2 x 256)
        self._buckets = self.plan.set_gradient_buckets(n_buckets)
        self._ce = None
        self._hybrid = False
        if exchange == 'ce':
            if self.in_place:
                raise ValueError('the copy-engine exchange needs the gradient in its own buffer (in_place=False)')
            self._ce = bdist.CopyEngineExchange(tuple(self.db.shape), n_buckets=len(self._buckets))
            self.grad = self._ce.grad
            self.plan.set_t_stash(self.grad)
        elif exchange == 'hybrid':
            if self.in_place:
                raise ValueError('the hybrid exchange needs the gradient in its own buffer (in_place=False)')
            self._ce = bdist.CopyEngineExchange(tuple(self.db.shape), n_buckets=len(self._buckets), gather_only=True)
            self.grad = self._ce.grad
            self.plan.set_t_stash(self.grad)
            self._comm_stream = torch.cuda.Stream(device=self.db.device)
            self._hybrid = True
        else:
            self._comm_stream = torch.cuda.Stream(device=self.db.device)
        Z, B, Y, X, _ = self.db.shape
        if sm_reserve < 0:
            sm_reserve = bdist.auto_sm_reserve(B, Y, X) if exchange in ('nccl', 'hybrid') else 0
        check(lib.bdof_set_sm_reserve(int(sm_reserve)))
        names = {'ce': 'copy engines over NVLink peer memory (push partial shards, owner sums, gather)',
                 'hybrid': 'NCCL reduce-scatter (AVG) + copy-engine all-gather over NVLink peer memory',
                 'nccl': 'NCCL all-reduce (AVG) on a communication stream'}
        self.exchange_name = '%s; %d SMs left to the collective' % (names.get(exchange, exchange), sm_reserve)
        return self

    def _exchange(self):
        if self._ce is not None and self._hybrid:
            self._ce.reduce_scatter_gather(self._buckets, self._comm_stream)
            self._ce.finish()
        elif self._ce is not None:
            self._ce.exchange(self._buckets)
            self._ce.finish()
        else:
            works = self._dp.allreduce_gradient(self.grad, average=True, buckets=self._buckets, comm_stream=self._comm_stream)
            self._dp.finish_allreduce(self.grad, works, comm_stream=self._comm_stream)

    def step_device(self, target_dev, accumulate=1):
        """forward + loss + adjoint (+ gradient all-reduce when data parallel) with the target(s) already on the device;
        returns the device loss of this rank.
        accumulate = K > 1: K fields (the projection angles of one minibatch: minibatch_size = 10 in reconstruct_fullfield.py:30)
        are evaluated one after the other and their gradients SUMMED in self.grad before ONE exchange -- the reference's ratio of
        compute to communication.  The sum is fused into the adjoint kernels' gradient stores (bdof_plan_set_grad_accumulate:
        L2 reductions / TMA reduce-stores), so it costs no extra pass; the transmission stash then lives in its own buffer.
        target_dev: [B,Y,X] (shared by the K fields) or [K,B,Y,X].  The stand-in for the K rotated copies of the object is the
        same field K times (same arithmetic and traffic)."""
        dp = getattr(self, '_dp', None) is not None
        if accumulate <= 1:
            self.plan.forward(self.db, self.probe, out=self.exit)
            loss, g = self.plan.loss_mag(self.exit, target_dev)
            self.plan.adjoint(self.db, g, grad_out=None if self.in_place else self.grad)
            if dp:
                self._exchange()
            return loss
        if self.in_place:
            raise ValueError('gradient accumulation needs the gradient in its own buffer (in_place=False)')
        if getattr(self, 'stash', None) is None:
            self.stash = torch.empty_like(self.db)
        self.plan.set_t_stash(self.stash)
        total = None
        try:
            for k in range(accumulate):
                self.plan.forward(self.db, self.probe, out=self.exit)
                loss, g = self.plan.loss_mag(self.exit, target_dev[k] if target_dev.dim() == 4 else target_dev)
                total = loss if total is None else total + loss
                self.plan.set_grad_accumulate(k > 0)
                self.plan.adjoint(self.db, g, grad_out=self.grad)
        finally:
            self.plan.set_grad_accumulate(False)
            self.plan.set_t_stash(self.grad)
        if dp:
            # the buckets' events are those of the LAST adjoint, whose stores complete the sums
            self._exchange()
        return total / accumulate

    def step(self, prj_mag_host, accumulate=1):
        """prj_mag_host: [B,Y,X] float32 in pinned host memory, or [K,B,Y,X] with accumulate = K (one projection per field)."""
        if accumulate > 1 and prj_mag_host.dim() == 4:
            if getattr(self, 'targets', None) is None or self.targets.shape != prj_mag_host.shape:
                self.targets = torch.empty(tuple(prj_mag_host.shape), dtype=torch.float32, device=self.db.device)
            self.targets.copy_(prj_mag_host, non_blocking=True)
            loss = self.step_device(self.targets, accumulate=accumulate)
        else:
            self.target.copy_(prj_mag_host, non_blocking=True)
            loss = self.step_device(self.target, accumulate=accumulate)
        self.loss_host.copy_(loss, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return float(self.loss_host)



class TomographyObjective:
    """One optimiser step of the full-field reconstruction loop (cnn_propagator/fullfield.py:300-360) with the object
    resident on the GPU in the native layout: rotate the object to every angle of the minibatch, multislice forward,
    loss, adjoint, back-rotate and accumulate the gradient, all-reduce over the data-parallel ranks, Adam.

    db_obj: [Z,Y,X,2] float32 CUDA tensor (delta, beta), updated in place by step().
    """

    def __init__(self, db_obj, probe, energy_ev, psize_cm, minibatch_size, free_prop_cm=None, propagate_last=True,
                 step_size=1e-7, deterministic=True, mask=None, shrink_threshold=None, alpha_d=None, alpha_b=None, gamma=0.0,
                 rotation='nearest'):
        Z, Y, X, _ = db_obj.shape
        if rotation not in ('nearest', 'bilinear'):
            raise ValueError("rotation must be 'nearest' (cnn_propagator tables) or 'bilinear' (tf.contrib.image.rotate)")
        self.rotation = rotation
        # finite-support mask [Z,Y,X] float32 (native order), non-negativity and shrink-wrap after every update
        # (cnn_propagator/fullfield.py:359-368); L1 / TV regularisers of the TF driver (fullfield.py:389-396)
        self.mask = None if mask is None else mask.to(db_obj.device, torch.float32).contiguous()
        self.clip = mask is not None or shrink_threshold is not None
        self.shrink_threshold = shrink_threshold
        self.alpha_d, self.alpha_b, self.gamma = alpha_d, alpha_b, gamma
        # back-rotation of the gradient: a gather over inverse lists (default: bit-reproducible run to run and across GPU counts up
        # to the all-reduce order) or fp32 atomic scatter-add (deterministic=False: ~30 % faster back-rotation)
        self.deterministic = bool(deterministic)
        self.shape = (Y, X, Z)
        self.B = int(minibatch_size)
        self.plan = MultislicePlan(Y, X, self.B, Z, energy_ev, psize_cm, free_prop_cm=free_prop_cm,
                                   propagate_last=propagate_last, store_slices=True, device=db_obj.device)
        self.obj = db_obj
        self.probe = probe.to(db_obj.device, torch.complex64).contiguous()
        self.db = torch.empty((Z, self.B, Y, X, 2), dtype=torch.float32, device=db_obj.device)
        self.plan.set_t_stash(self.db)               # rotated copies: rebuilt every step, overwritten in place by the adjoint
        self.grad = torch.zeros_like(db_obj)
        self.m = torch.zeros_like(db_obj)
        self.v = torch.zeros_like(db_obj)
        self.target = torch.empty((self.B, Y, X), dtype=torch.float32, device=db_obj.device)
        self.exit = torch.empty((self.B, Y, X), dtype=torch.complex64, device=db_obj.device)
        self.loss_host = torch.empty((), dtype=torch.float64).pin_memory()
        self.step_size = float(step_size)
        self.i_batch = 0
        self._dp = None
        self._ce = None
        self._n_buckets = 1
        self._comm_stream = None

    def enable_data_parallel(self, exchange='nccl', n_buckets=4):
        """exchange: 'nccl' (default: the gradient of one object is small -- 134 MB at 256^3 -- and exchanged in `n_buckets` z
        buckets, each as soon as its part of the back-rotation is done; measured 36.2 vs 34.4 Gpx*slice/s against 'ce' on two
        B200) or 'ce'."""
        from . import dist as bdist
        self._dp = bdist
        self._ce = None
        self._n_buckets = int(n_buckets)
        self._comm_stream = torch.cuda.Stream(device=self.obj.device)
        self._bucket_events = [torch.cuda.Event() for _ in range(max(1, self.shape[2] // 32 + 1))]
        if exchange == 'auto':
            exchange = 'nccl'
        if exchange == 'ce':
            self._ce = bdist.CopyEngineExchange(tuple(self.grad.shape), n_buckets=1)
            self.grad = self._ce.grad
        return self

    def prepare(self, thetas):
        """Build the rotation tables of all angles up front (the reference writes them to disk once and reads them
        back, save_rotation_lookup / read_all_origin_coords, cnn_propagator/fullfield.py:209-215)."""
        for t in (thetas if self.rotation == 'nearest' else []):
            tab = _rot.device_table(self.shape, float(t), self.obj.device)
            if self.deterministic:
                _rot.device_inverse(tab)
        return self

    def loss_and_grad(self, theta_batch, target_dev):
        dev = self.obj.device
        nearest = self.rotation == 'nearest'
        tabs = [_rot.device_table(self.shape, float(t), dev) for t in theta_batch] if nearest else None
        for b in range(self.B):
            if nearest:
                _rot.rotate_db(self.obj, tabs[b], out=self.db[:, b])
            else:
                _rot.rotate_db_bilinear(self.obj, float(theta_batch[b]), out=self.db[:, b])
        self.plan.forward(self.db, self.probe, out=self.exit)
        loss, g = self.plan.loss_mag(self.exit, target_dev)
        self.plan.adjoint(self.db, g)
        works = None
        if nearest and self.deterministic and self._dp is not None and self._ce is None and self._n_buckets > 1:
            # back-rotation in z buckets: the all-reduce of a bucket runs on the communication stream under the next bucket
            Z = self.shape[2]
            step = max(32, ((Z + self._n_buckets - 1) // self._n_buckets + 31) // 32 * 32)
            buckets = []
            for k, z_lo in enumerate(range(0, Z, step)):
                z_hi = min(Z, z_lo + step)
                _rot.rotate_db_adjoint_batch(self.db, tabs, self.grad, accumulate=False, z_range=(z_lo, z_hi))
                self._bucket_events[k].record()
                buckets.append((z_lo, z_hi, self._bucket_events[k]))
            works = self._dp.allreduce_gradient(self.grad, average=True, buckets=buckets, comm_stream=self._comm_stream)
        elif nearest and self.deterministic:
            _rot.rotate_db_adjoint_batch(self.db, tabs, self.grad, accumulate=False)     # whole minibatch, one pass over the gradient
        else:
            self.grad.zero_()
            for b in range(self.B):
                if nearest:
                    _rot.rotate_db_adjoint(self.db[:, b], tabs[b], self.grad, atomic=True)
                else:
                    _rot.rotate_db_bilinear_adjoint(self.db[:, b], float(theta_batch[b]), self.grad)
        if self._dp is not None:
            if self._ce is not None:
                self._ce.exchange(None)
                self._ce.finish()
            elif works is not None:
                self._dp.finish_allreduce(self.grad, works, comm_stream=self._comm_stream)
            else:
                self._dp.finish_allreduce(self.grad, self._dp.allreduce_gradient(self.grad, average=True))
        # regularisers act on the (replicated) object: added after the exchange, identical on every rank; one fused pass
        # (bdof_regularizers) for L1(delta), L1(beta), TV(delta) and their gradients
        if self.alpha_d or self.alpha_b or self.gamma:
            loss = _rot.regularizers(self.obj, self.grad, loss.clone(), self.alpha_d, self.alpha_b, self.gamma)
        return loss

    def step(self, theta_batch, prj_mag_host):
        """prj_mag_host: [B,Y,X] float32 (pinned host or device); returns this rank's data-fidelity loss."""
        self.target.copy_(prj_mag_host, non_blocking=True)
        loss = self.loss_and_grad(theta_batch, self.target)
        _rot.adam_step(self.obj, self.grad, self.i_batch, self.m, self.v, step_size=self.step_size)
        if self.clip:
            _rot.finite_support(self.obj, self.mask, self.shrink_threshold)
        self.i_batch += 1
        self.loss_host.copy_(loss, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return float(self.loss_host)


class PtychographyObjective:
    """One optimiser step of the ptychography reconstruction loop (cnn_propagator/ptychography.py:286-310, one rotation angle
    per step) with the object resident on the GPU in the native layout: cut this rank's minibatch of probe windows, multislice
    forward to the far field (the resident small-field kernels when the probe is 64 x 64), loss, adjoint, DETERMINISTIC
    accumulation of the window gradients into the object gradient, all-reduce over the data-parallel ranks (positions are
    sharded contiguously, dist.shard_contiguous = ptychography.py:292-297), Adam, clip >= 0 (ptychography.py:307-310).

    db_obj: [Z,Y,X,2] float32 CUDA tensor (delta, beta), updated in place by step().
    """

    def __init__(self, db_obj, probe, probe_size, energy_ev, psize_cm, n_pos_per_step, n_pos_total=None, step_size=1e-7,
                 scale_by_npos=True, clip=True):
        Z, OY, OX, _ = db_obj.shape
        self.obj = db_obj
        self.py, self.px = int(probe_size[0]), int(probe_size[1])
        self.n = int(n_pos_per_step)
        self.half = (np.array([self.py, self.px]) / 2).astype('int')                 # ptychography.py:184
        dev = db_obj.device
        self.plan = MultislicePlan(self.py, self.px, self.n, Z, energy_ev, psize_cm, free_prop_cm='inf', propagate_last=True,
                                   store_slices=True, device=dev)
        self.probe = probe.to(dev, torch.complex64).contiguous()
        self.patches = torch.empty((Z, self.n, self.py, self.px, 2), dtype=torch.float32, device=dev)
        # resident plans (64 x 64 probes) read the object straight through the windows: no window copy of the object is cut and
        # `patches` holds the transmission stash between forward and adjoint and then the per-window gradients; otherwise the
        # windows are cut every step and overwritten in place
        self.windowed = self.plan.is_resident()
        self.plan.set_t_stash(self.patches)
        self.grad = torch.zeros_like(db_obj)
        self.m = torch.zeros_like(db_obj)
        self.v = torch.zeros_like(db_obj)
        self.target = torch.empty((self.n, self.py, self.px), dtype=torch.float32, device=dev)
        self.exit = torch.empty((self.n, self.py, self.px), dtype=torch.complex64, device=dev)
        self.origin = torch.empty((self.n, 2), dtype=torch.int32, device=dev)
        self.origin_host = torch.empty((self.n, 2), dtype=torch.int32).pin_memory()
        self.loss_host = torch.empty((), dtype=torch.float64).pin_memory()
        self.scale = float(n_pos_total if n_pos_total is not None else self.n) if scale_by_npos else 1.0
        self.step_size = float(step_size)
        self.clip = bool(clip)
        self.i_batch = 0
        self._dp = None
        self._n_buckets = 1
        self._comm_stream = None

    def enable_data_parallel(self, n_buckets=1):
        """The object gradient (67 MB at 256 x 256 x 128) is averaged over the ranks with an NCCL all-reduce (comm.Allreduce +
        grads / size, ptychography.py:302-306); with n_buckets > 1 in z buckets, each as soon as its part of the window
        accumulation is done (measured on 8 B200 at config 3: 2.12 ms per update with 4 buckets, 2.10 ms with one -- the
        accumulation of a rank's 128 windows is too short to hide anything, so one collective is the default)."""
        from . import dist as bdist
        self._dp = bdist
        self._n_buckets = int(n_buckets)
        self._comm_stream = torch.cuda.Stream(device=self.obj.device)
        self._bucket_events = [torch.cuda.Event() for _ in range(max(1, self._n_buckets))]
        return self

    def loss_and_grad(self, pos_batch, target_dev):
        """pos_batch: [n, 2] integer (y, x) scan positions of this rank; returns the device loss of this rank's positions."""
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        Z, OY, OX, _ = self.obj.shape
        pos = np.asarray(pos_batch).astype(np.int64)
        assert pos.shape == (self.n, 2)
        self.origin_host.copy_(torch.as_tensor((pos - self.half[None, :]).astype(np.int32)))
        self.origin.copy_(self.origin_host, non_blocking=True)
        if self.windowed:
            self.plan.set_windows((Z, OY, OX), self.origin)
            self.plan.forward(self.obj, self.probe, out=self.exit)
            loss, g = self.plan.loss_mag(self.exit, target_dev, loss_scale=self.scale)
            self.plan.adjoint(self.obj, g, grad_out=self.patches)
        else:
            check(lib.bdof_patch_gather(_ptr(self.obj), Z, OY, OX, _ptr(self.origin), self.n, self.py, self.px, _ptr(self.patches), st))
            self.plan.forward(self.patches, self.probe, out=self.exit)
            loss, g = self.plan.loss_mag(self.exit, target_dev, loss_scale=self.scale)
            self.plan.adjoint(self.patches, g)
        self.grad.zero_()
        if self._dp is not None and self._n_buckets > 1 and Z >= 2 * self._n_buckets:
            # window accumulation in z buckets: the all-reduce of a bucket runs on the communication stream under the next one
            step = (Z + self._n_buckets - 1) // self._n_buckets
            buckets = []
            for k, z_lo in enumerate(range(0, Z, step)):
                z_hi = min(Z, z_lo + step)
                check(lib.bdof_patch_gather_add(_ptr(self.patches[z_lo:z_hi]), z_hi - z_lo, OY, OX, _ptr(self.origin), self.n, self.py, self.px,
                                                _ptr(self.grad[z_lo:z_hi]), st))
                self._bucket_events[k].record()
                buckets.append((z_lo, z_hi, self._bucket_events[k]))
            works = self._dp.allreduce_gradient(self.grad, average=True, buckets=buckets, comm_stream=self._comm_stream)
            self._dp.finish_allreduce(self.grad, works, comm_stream=self._comm_stream)
            return loss
        check(lib.bdof_patch_gather_add(_ptr(self.patches), Z, OY, OX, _ptr(self.origin), self.n, self.py, self.px, _ptr(self.grad), st))
        if self._dp is not None:
            self._dp.finish_allreduce(self.grad, self._dp.allreduce_gradient(self.grad, average=True))
        return loss

    def step(self, pos_batch, prj_mag_host):
        """prj_mag_host: [n, py, px] float32 measured far-field magnitudes of these positions (pinned host or device)."""
        self.target.copy_(prj_mag_host, non_blocking=True)
        loss = self.loss_and_grad(pos_batch, self.target)
        _rot.adam_step(self.obj, self.grad, self.i_batch, self.m, self.v, step_size=self.step_size)
        if self.clip:
            _rot.finite_support(self.obj, None, None)
        self.i_batch += 1
        self.loss_host.copy_(loss, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return float(self.loss_host)
