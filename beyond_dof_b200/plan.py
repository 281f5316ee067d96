"""MultislicePlan: the host-side handle of one multislice problem shape on one GPU.

PyTorch is only the array container here (device memory, streams); all arithmetic happens in
libbdof.so through the C ABI (include/bdof.h).  There is no CPU fallback: constructing a plan
without a CUDA device raises.
"""
import ctypes

import numpy as np
import torch

from . import capi
from .capi import lib, check
from .util import PI, kernel_factors, factor_kernel


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr())


def _hptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _c128(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.complex128))


class MultislicePlan:
    """Forward / adjoint multislice for a fixed [batch, ny, nx, n_slice] shape.

    Semantics switches:
      propagate_last=False  NumPy path, last slice only modulates   (npfuncs.py:35-41)
      propagate_last=True   TF path, every slice propagates          (util.py:464-488)
      free_prop_cm          None | 'inf' | float (cm)                (npfuncs.py:43-61)
      h                     optional caller-supplied centred H [ny,nx] (util.py:459-461)
      factors               optional (phase0, hy, hx): the separable factors of a centred H directly (tiling: windows use the
                            GLOBAL field's frequency grid)
    """

    def __init__(self, ny, nx, batch, n_slice, energy_ev, psize_cm, free_prop_cm=None, propagate_last=False,
                 store_slices=False, z_broadcast=False, h=None, pi=PI, device=None, stepwise=False, factors=None):
        if not torch.cuda.is_available():
            raise RuntimeError('beyond_dof_b200 needs a CUDA device: the multislice path has no CPU fallback')
        self.device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        self.ny, self.nx, self.batch, self.n_slice = int(ny), int(nx), int(batch), int(n_slice)
        self.propagate_last = bool(propagate_last)
        self.store_slices = bool(store_slices)
        self.z_broadcast = bool(z_broadcast)
        if np.ndim(psize_cm) == 0:
            voxel_nm = np.array([psize_cm] * 3, dtype=np.float64) * 1.e7
        else:
            voxel_nm = np.array(psize_cm, dtype=np.float64) * 1.e7
        self.voxel_nm = voxel_nm
        self.lmbda_nm = 1240. / energy_ev
        delta_nm = voxel_nm[-1]
        self.k_dz = 2. * pi * delta_nm / self.lmbda_nm
        flags = 0
        if self.propagate_last:
            flags |= capi.PROPAGATE_LAST
        if self.store_slices:
            flags |= capi.STORE_SLICES
        if self.z_broadcast:
            flags |= capi.Z_BROADCAST
        self.stepwise = bool(stepwise)
        if self.stepwise:
            flags |= capi.STEPWISE          # driven slice by slice (slice_step(..., index=i)): per-pass schedule and tables
        self._h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            self.stream = torch.cuda.current_stream(self.device)
            check(lib.bdof_plan_create(ctypes.byref(self._h), self.ny, self.nx, self.batch, self.n_slice, flags,
                                       ctypes.c_void_p(self.stream.cuda_stream)))
            grid_shape = [self.ny, self.nx, self.n_slice]
            fac = None
            if factors is not None:
                fac = factors                            # (phase0, hy[ny], hx[nx]) centred separable factors supplied by the caller
            elif h is None:
                fac = kernel_factors(delta_nm, self.lmbda_nm, voxel_nm, grid_shape, pi=pi)
            else:
                fac = factor_kernel(h)
            if fac is not None:
                p0, hy, hx = fac
                hy, hx = _c128(hy), _c128(hx)
                check(lib.bdof_set_kernel(self._h, _hptr(hy), _hptr(hx), p0.real, p0.imag, self.k_dz))
            else:
                hh = _c128(h)
                check(lib.bdof_set_kernel_full(self._h, _hptr(hh), self.k_dz))
            if free_prop_cm is None:
                check(lib.bdof_set_free_prop(self._h, capi.FREE_NONE, None, None, 1.0, 0.0))
            elif isinstance(free_prop_cm, str):
                if free_prop_cm != 'inf':
                    raise ValueError("free_prop_cm must be None, 'inf' or a distance in cm")
                check(lib.bdof_set_free_prop(self._h, capi.FREE_INF, None, None, 1.0, 0.0))
            else:
                # npfuncs.py:48-56: the TF/IR choice is computed and then forced to 'TF'
                p0, hy, hx = kernel_factors(free_prop_cm * 1e7, self.lmbda_nm, voxel_nm, grid_shape, pi=pi)
                hy, hx = _c128(hy), _c128(hx)
                check(lib.bdof_set_free_prop(self._h, capi.FREE_TF, _hptr(hy), _hptr(hx), p0.real, p0.imag))

    def __del__(self):
        h = getattr(self, '_h', None)
        if h is not None and h.value:
            lib.bdof_plan_destroy(h)
            self._h = ctypes.c_void_p()

    def use_current_stream(self):
        """Bind the plan to torch's CURRENT stream of its device.  Called at the top of every compute method, so that the
        libbdof kernels run on the same stream as the torch-side work around them (packing, probe set-up, losses) even when a
        cached plan is reused under a different stream than the one it was created on.  The new stream is ordered after
        whatever the plan still has in flight on the old one."""
        cur = torch.cuda.current_stream(self.device)
        if cur.cuda_stream != self.stream.cuda_stream:
            ev = torch.cuda.Event()
            ev.record(self.stream)
            cur.wait_event(ev)
            self.stream = cur
            check(lib.bdof_plan_set_stream(self._h, ctypes.c_void_p(cur.cuda_stream)))
        return self

    # ---- layout helpers -----------------------------------------------------------------
    @property
    def db_shape(self):
        return (1 if self.z_broadcast else self.n_slice, self.batch, self.ny, self.nx, 2)

    def pack(self, grid_delta_batch, grid_beta_batch):
        """[B,Y,X,Z] float32 device tensors -> slice-major interleaved db [Z,B,Y,X,2]."""
        self.use_current_stream()
        d = grid_delta_batch.to(self.device, torch.float32).contiguous()
        b = grid_beta_batch.to(self.device, torch.float32).contiguous()
        B, Y, X, Z = d.shape
        db = torch.empty((Z, B, Y, X, 2), dtype=torch.float32, device=self.device)
        check(lib.bdof_pack_db(_ptr(d), _ptr(b), _ptr(db), B, Y, X, Z, ctypes.c_void_p(self.stream.cuda_stream)))
        return db

    def unpack(self, db):
        self.use_current_stream()
        Z, B, Y, X, _ = db.shape
        d = torch.empty((B, Y, X, Z), dtype=torch.float32, device=self.device)
        b = torch.empty((B, Y, X, Z), dtype=torch.float32, device=self.device)
        check(lib.bdof_unpack_db(_ptr(db), _ptr(d), _ptr(b), B, Y, X, Z, ctypes.c_void_p(self.stream.cuda_stream)))
        return d, b

    # ---- compute ------------------------------------------------------------------------
    def is_resident(self):
        """True when the plan runs the resident small-field kernels (one CTA per 64 x 64 field, one launch per direction; window
        mode available)."""
        return int(lib.bdof_plan_is_resident(self._h)) == 1

    def is_cluster_resident(self):
        """True when the plan runs the cluster-resident kernels (one cluster of 8 CTAs per 256 x 256 field)."""
        return int(lib.bdof_plan_is_resident(self._h)) == 2

    def set_windows(self, obj_shape_zyx, origin):
        """Window mode (resident plans only): the batch elements are windows of one object [Z,OY,OX,2]; origin [B,2] int32 CUDA
        tensor of (y0, x0).  forward()/adjoint() then take the OBJECT instead of db.  origin=None switches it off."""
        if origin is None:
            self._win = None
            check(lib.bdof_plan_set_windows(self._h, 0, 0, None))
            return
        Z, OY, OX = (int(v) for v in obj_shape_zyx)
        assert origin.is_cuda and origin.dtype == torch.int32 and origin.is_contiguous() and tuple(origin.shape) == (self.batch, 2)
        self._win = (Z, OY, OX, origin)
        check(lib.bdof_plan_set_windows(self._h, OY, OX, _ptr(origin)))

    def forward(self, db, probe, out=None):
        """db [Z,B,Y,X,2] float32 (window mode: the object [Z,OY,OX,2]), probe [Y,X] complex64 -> exit wave [B,Y,X] complex64."""
        self.use_current_stream()
        win = getattr(self, '_win', None)
        want = self.db_shape if win is None else ((1 if self.z_broadcast else win[0]), win[1], win[2], 2)
        assert db.is_cuda and db.dtype == torch.float32 and db.is_contiguous() and tuple(db.shape) == want, \
            'db must be a contiguous float32 CUDA tensor of shape %s' % (want,)
        probe = probe.to(self.device, torch.complex64).contiguous()
        assert tuple(probe.shape) == (self.ny, self.nx)
        if out is None:
            out = torch.empty((self.batch, self.ny, self.nx), dtype=torch.complex64, device=self.device)
        check(lib.bdof_forward(self._h, _ptr(db), _ptr(probe), _ptr(out)))
        return out

    def loss_mag(self, exit_wave, target_mag, want_grad=True, loss_scale=1.0):
        """mean((|psi| - |y|)^2) * loss_scale and G = dL/dRe + i dL/dIm (fullfield.py:115)."""
        self.use_current_stream()
        target_mag = target_mag.to(self.device, torch.float32).contiguous()
        loss = torch.empty((), dtype=torch.float64, device=self.device)
        g = torch.empty_like(exit_wave) if want_grad else None
        check(lib.bdof_loss_mag(self._h, _ptr(exit_wave), _ptr(target_mag), float(loss_scale), _ptr(loss),
                                _ptr(g) if want_grad else None))
        return loss, g

    def adjoint(self, db, grad_exit, grad_out=None, want_probe_grad=False):
        """Back-propagate grad_exit.  By default db is overwritten in place with (dL/ddelta, dL/dbeta);
        pass grad_out [Z,B,Y,X,2] to keep db intact (mandatory with z_broadcast)."""
        assert self.store_slices, 'plan was created with store_slices=False'
        self.use_current_stream()
        grad_exit = grad_exit.to(self.device, torch.complex64).contiguous()
        gp = torch.empty((self.ny, self.nx), dtype=torch.complex64, device=self.device) if want_probe_grad else None
        if grad_out is not None:
            assert grad_out.is_contiguous() and tuple(grad_out.shape) == (self.n_slice, self.batch, self.ny, self.nx, 2)
        check(lib.bdof_adjoint(self._h, _ptr(db), _ptr(grad_exit), _ptr(grad_out) if grad_out is not None else None,
                               _ptr(gp) if want_probe_grad else None))
        res = grad_out if grad_out is not None else db
        return (res, gp) if want_probe_grad else res

    def forward_host(self, delta_byxz, beta_byxz, probe, out=None):
        """End-to-end call with HOST float32 arrays in the reference layout [B,Y,X,Z]: H2D copies,
        layout conversion, the forward chain and the D2H copy of the exit wave all happen inside."""
        d = np.ascontiguousarray(delta_byxz, dtype=np.float32)
        b = np.ascontiguousarray(beta_byxz, dtype=np.float32)
        pr = np.ascontiguousarray(probe, dtype=np.complex64)
        if out is None:
            out = np.empty((self.batch, self.ny, self.nx), dtype=np.complex64)
        assert d.shape == (self.batch, self.ny, self.nx, self.n_slice) and b.shape == d.shape
        self.use_current_stream()
        check(lib.bdof_forward_host(self._h, _hptr(d), _hptr(b), _hptr(pr), _hptr(out)))
        return out

    def slice_step(self, field, db_slice, out=None, propagate=True, index=None):
        """One slice: out = P(field * t(db_slice)) (or only the modulation).  field [B,Y,X] complex64,
        db_slice [B,Y,X,2] float32.  The global phase exp(i k dz) is not applied.
        index: position of this slice in a chain stepped in order on a stepwise=True plan (selects that slice's entry of the
        error-feedback multiplier sequence); None = the plain table."""
        self.use_current_stream()
        assert field.is_cuda and field.dtype == torch.complex64 and field.is_contiguous()
        assert db_slice.is_cuda and db_slice.dtype == torch.float32 and db_slice.is_contiguous()
        if out is None:
            out = torch.empty_like(field)
        check(lib.bdof_slice_step_seq(self._h, _ptr(field), _ptr(db_slice), _ptr(out), 1 if propagate else 0,
                                      -1 if index is None else int(index)))
        return out

    def set_t_stash(self, buf):
        """buf: float32 CUDA tensor [Z,B,Y,X,2] (or None).  The forward leaves the transmission t_i of every slice there
        and the next adjoint lands it instead of recomputing exp(k(i delta - beta)).  Pass the tensor the adjoint will write
        its gradient to (grad_out, or db itself for the in-place adjoint): t then costs no memory."""
        if buf is not None:
            assert buf.is_cuda and buf.dtype == torch.float32 and buf.is_contiguous() and \
                tuple(buf.shape) == (self.n_slice, self.batch, self.ny, self.nx, 2)
        self._stash = buf
        check(lib.bdof_plan_set_t_stash(self._h, _ptr(buf) if buf is not None else None))

    def set_grad_accumulate(self, on):
        """While on, adjoint(..., grad_out=acc) ADDS the gradient to acc (fused in the sweep kernels' gradient stores): the sum over
        the fields of a minibatch costs no extra pass.  The transmission stash must live in another buffer than acc."""
        check(lib.bdof_plan_set_grad_accumulate(self._h, 1 if on else 0))

    def set_gradient_buckets(self, n_buckets):
        """Split z into n_buckets (counted from the last slice); returns [(z_lo, z_hi, event)] in the order
        the adjoint sweep completes them.  The events are recorded inside bdof_adjoint."""
        self._bucket_events = [torch.cuda.Event() for _ in range(n_buckets)]
        for ev in self._bucket_events:
            ev.record(self.stream)                       # materialise the cudaEvent_t handle
        arr = (ctypes.c_void_p * max(n_buckets, 1))(*[ctypes.c_void_p(ev.cuda_event) for ev in self._bucket_events])
        check(lib.bdof_plan_set_bucket_events(self._h, n_buckets, arr))
        per = (self.n_slice + n_buckets - 1) // n_buckets if n_buckets else 0
        out = []
        for j in range(n_buckets):
            z_hi = self.n_slice - j * per
            z_lo = max(0, z_hi - per)
            if z_hi > 0:
                out.append((z_lo, z_hi, self._bucket_events[j]))
        return out

    VARIANT_NAMES = ['row_conv_transmit', 'row_conv', 'row_conv_adjoint', 'row_fft', 'row_ifft', 'col_conv', 'col_fft',
                     'col_ifft', 'col_conv_2d', 'col_conv_pipe', 'sweep_forward', 'sweep_adjoint', 'resident_forward', 'resident_adjoint']

    def free_prop(self, field, out=None):
        """The plan's free-space step on its own: [B,Y,X] complex64 -> [B,Y,X]."""
        self.use_current_stream()
        field = field.to(self.device, torch.complex64).contiguous()
        if out is None:
            out = torch.empty_like(field)
        check(lib.bdof_free_prop(self._h, _ptr(field), _ptr(out)))
        return out

    def free_prop_adjoint(self, grad, out=None):
        """Adjoint of the plan's free-space step on its own: [B,Y,X] complex64 -> [B,Y,X]."""
        self.use_current_stream()
        grad = grad.to(self.device, torch.complex64).contiguous()
        if out is None:
            out = torch.empty_like(grad)
        check(lib.bdof_free_prop_adjoint(self._h, _ptr(grad), _ptr(out)))
        return out

    def profile_begin(self):
        check(lib.bdof_profile_begin(self._h))

    def profile_end(self):
        """-> {variant name: (launches, total ms)} for the launches since profile_begin()."""
        n = len(self.VARIANT_NAMES)
        counts = (ctypes.c_int * n)()
        ms = (ctypes.c_double * n)()
        check(lib.bdof_profile_end(self._h, n, counts, ms))
        return {self.VARIANT_NAMES[i]: (int(counts[i]), float(ms[i])) for i in range(n) if counts[i]}

    def last_times(self):
        """Device time of the last forward and the last adjoint call (one event pair around each call's whole launch
        sequence; synchronises): {'forward': (ms, launches), 'adjoint': (ms, launches)}."""
        ms = (ctypes.c_double * 2)()
        n = (ctypes.c_int * 2)()
        check(lib.bdof_plan_last_times(self._h, ms, n))
        return {'forward': (float(ms[0]), int(n[0])), 'adjoint': (float(ms[1]), int(n[1]))}

    def workspace_bytes(self):
        n = ctypes.c_size_t()
        check(lib.bdof_plan_workspace_bytes(self._h, ctypes.byref(n)))
        return n.value
