"""Object rotation and the optimiser step: the stages either side of the multislice hot path in the
cnn_propagator drivers.

  rotation_table      <- save_rotation_lookup   cnn_propagator/util.py:295-336 (one angle)
  apply_rotation      <- apply_rotation         cnn_propagator/util.py:374-402
  rotate_db / rotate_db_adjoint                 the same on the native slice-major object (CUDA gather / scatter-add)
  adam_step           <- apply_gradient_adam    cnn_propagator/util.py:280-291

The lookup table is host logic (a few KB per angle) and is built here exactly as the reference builds
it; applying it to the object and back-projecting the gradient are CUDA kernels of libbdof.
"""
import ctypes

import numpy as np
import torch

from .capi import lib, check
from .plan import _ptr


def rotation_table(array_size, theta):
    """Nearest-neighbour source pixel (x_old, z_old), clipped to the array, of every rotated pixel (x, z) --
    flattened with z fastest -- for a rotation by theta about axis 0 around floor(size / 2)
    (save_rotation_lookup, util.py:295-336).  array_size = [Y, X, Z]; returns int64 [X * Z, 2]."""
    if array_size[1] != array_size[2]:
        # the reference drivers only build tables for [dim_y, dim_x, dim_x] (fullfield.py:213); its reshape of the
        # tiled coordinate vector (util.py:305-307) is not a rotation otherwise
        if theta == 0:
            x = np.repeat(np.arange(array_size[1]), array_size[2])
            return np.stack([x, np.tile(np.arange(array_size[2]), array_size[1])], axis=1)      # identity
        raise ValueError('rotation needs a square (x, z) cross-section, got X=%d, Z=%d' % (array_size[1], array_size[2]))
    cy, cx, cz = (np.floor(v / 2) for v in array_size)
    x = np.repeat(np.arange(array_size[1]), array_size[2]) - cx          # util.py:305-307, 314
    z = np.tile(np.arange(array_size[2]), array_size[1]) - cz            # util.py:303, 315
    coord_new = np.stack([x, z]).astype(np.float32)                      # util.py:318
    m_rot = np.array([[np.cos(theta), -np.sin(theta)],
                      [np.sin(theta), np.cos(theta)]])
    coord_old = np.matmul(m_rot, coord_new)
    x_old = np.clip(np.round(coord_old[0] + cx).astype(int), 0, array_size[1] - 1)
    z_old = np.clip(np.round(coord_old[1] + cz).astype(int), 0, array_size[2] - 1)
    return np.stack([x_old, z_old], axis=1)


_tables = {}


def device_table(array_size, theta, device, coord_old=None):
    """The table of one angle on the device, re-ordered slice-major: int32 [Z, X, 2] = (x_old, z_old) of pixel (z, x)."""
    key = (tuple(int(v) for v in array_size), float(theta), str(device)) if coord_old is None else None
    if key is not None and key in _tables:
        return _tables[key]
    tab = rotation_table(array_size, theta) if coord_old is None else np.asarray(coord_old)
    X, Z = int(array_size[1]), int(array_size[2])
    t = torch.as_tensor(np.ascontiguousarray(tab.reshape(X, Z, 2).transpose(1, 0, 2)).astype(np.int32)).to(device)
    if key is not None:
        if len(_tables) > 512:
            _tables.clear()
        _tables[key] = t
    return t


_inverse = {}


def device_inverse(table_zx):
    """CSR lists of the transpose: for every source pixel (z0, x0) the rotated pixels (z * X + x) that read from it.
    Host logic (one argsort per angle, cached by table); returns (offsets int32 [Z*X + 1], dest int32 [Z*X]) on the device."""
    key = table_zx.data_ptr()
    hit = _inverse.get(key)
    if hit is not None and hit[0] is table_zx:
        return hit[1], hit[2]
    Z, X, _ = table_zx.shape
    t = table_zx.cpu().numpy().astype(np.int64)
    src = (t[..., 1] * X + t[..., 0]).reshape(-1)                 # source cell z_old * X + x_old of rotated pixel z * X + x
    order = np.argsort(src, kind='stable')
    counts = np.bincount(src, minlength=Z * X)
    offsets = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    dev = table_zx.device
    off_d = torch.as_tensor(offsets).to(dev)
    dest_d = torch.as_tensor(order.astype(np.int32)).to(dev)
    if len(_inverse) > 512:
        _inverse.clear()
    _inverse[key] = (table_zx, off_d, dest_d)
    return off_d, dest_d


def rotate_db(db_obj, table_zx, out=None):
    """db_obj [Z, Y, X, 2] float32 (CUDA) -> rotated object, same layout.  `out` may be a [Z, Y, X, 2] view with a
    larger slice stride (one batch element of a plan's [Z, B, Y, X, 2] object)."""
    Z, Y, X, _ = db_obj.shape
    if out is None:
        out = torch.empty_like(db_obj)
    assert out.shape == db_obj.shape and out.stride()[1:] == (X * 2, 2, 1) and out.stride(0) % 2 == 0
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    check(lib.bdof_rotate_gather(_ptr(db_obj), _ptr(table_zx), _ptr(out), out.stride(0) // 2, Y, X, Z, st))
    return out


def rotate_db_adjoint(grad_rot, table_zx, grad_obj, atomic=False):
    """grad_obj [Z, Y, X, 2] += transpose-of-rotation(grad_rot); grad_rot may be a strided batch-element view.
    Default: deterministic gather over the inverse lists of the table; atomic=True: fp32 atomic scatter-add."""
    Z, Y, X, _ = grad_obj.shape
    assert grad_rot.shape == grad_obj.shape and grad_rot.stride()[1:] == (X * 2, 2, 1) and grad_obj.is_contiguous()
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    if atomic:
        check(lib.bdof_rotate_scatter_add(_ptr(grad_rot), grad_rot.stride(0) // 2, _ptr(table_zx), _ptr(grad_obj), Y, X, Z, st))
    else:
        off, dest = device_inverse(table_zx)
        check(lib.bdof_rotate_adjoint_csr(_ptr(grad_rot), grad_rot.stride(0) // 2, _ptr(off), _ptr(dest), _ptr(grad_obj), Y, X, Z, st))
    return grad_obj


def rotate_db_adjoint_batch(grad_rot_db, tables, grad_obj, accumulate=True, z_range=None):
    """grad_obj [Z, Y, X, 2] += (accumulate=False: =) sum over the minibatch of transpose-of-rotation(grad_rot_db[:, b]) with tables[b]; grad_rot_db is a
    plan's db [Z, B, Y, X, 2].  One pass over the object gradient (deterministic: the angles are summed in order in registers).
    z_range = (z_lo, z_hi), z_lo a multiple of 32: only those slices of grad_obj (the back-rotation in buckets, each all-reduced
    while the next is computed)."""
    Z, Y, X, _ = grad_obj.shape
    B = len(tables)
    assert grad_rot_db.shape == (Z, B, Y, X, 2) and grad_rot_db.is_contiguous() and grad_obj.is_contiguous()
    lists = [device_inverse(t) for t in tables]
    offs = (ctypes.c_void_p * B)(*[o.data_ptr() for o, _ in lists])
    dests = (ctypes.c_void_p * B)(*[d.data_ptr() for _, d in lists])
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    z_lo, z_hi = (0, Z) if z_range is None else (int(z_range[0]), int(z_range[1]))
    check(lib.bdof_rotate_adjoint_csr_batch_range(_ptr(grad_rot_db), B * Y * X, Y * X, B, offs, dests, _ptr(grad_obj), 1 if accumulate else 0,
                                                  Y, X, Z, z_lo, z_hi, st))
    return grad_obj


def apply_rotation(obj, coord_old, src_folder=None):
    """Drop-in for cnn_propagator/util.py:374 -- obj [Y, X, Z, C] (NumPy or torch), coord_old the [X*Z, 2] table of one
    angle (read_origin_coords / rotation_table); src_folder is accepted and ignored (the reference re-reads its
    coordinate vectors from there).  Runs on the GPU for C == 2 float32 objects; returns the container kind it was given."""
    is_np = not isinstance(obj, torch.Tensor)
    t = torch.as_tensor(np.ascontiguousarray(obj)) if is_np else obj
    if t.shape[-1] != 2:
        c1 = torch.as_tensor(np.asarray(coord_old)[:, 0].reshape(t.shape[1], t.shape[2]))
        c2 = torch.as_tensor(np.asarray(coord_old)[:, 1].reshape(t.shape[1], t.shape[2]))
        out = t[:, c1, c2]
        return out.numpy() if is_np else out
    dev = torch.device('cuda', torch.cuda.current_device())
    Y, X, Z, _ = t.shape
    db = t.to(dev, torch.float32).permute(2, 0, 1, 3).contiguous()
    rot = rotate_db(db, device_table([Y, X, Z], None, dev, coord_old=coord_old))
    out = rot.permute(1, 2, 0, 3).contiguous()
    return out.cpu().numpy().astype(obj.dtype) if is_np else out.to(obj.dtype)


def adam_step(x, g, i_batch, m=None, v=None, step_size=0.001, b1=0.9, b2=0.999, eps=1e-8):
    """apply_gradient_adam (util.py:280-291) on float32 CUDA tensors, IN PLACE on x, m, v; returns (x, m, v).
    m, v = None start from zero moments (the first minibatch of the reference)."""
    assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous() and g.shape == x.shape
    if m is None or v is None:
        m = torch.zeros_like(x)
        v = torch.zeros_like(x)
    g = g.to(torch.float32).contiguous()
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    check(lib.bdof_adam_step(_ptr(x), _ptr(g), _ptr(m), _ptr(v), x.numel(), int(i_batch), float(step_size), float(b1),
                             float(b2), float(eps), st))
    return x, m, v


def finite_support(db_obj, mask=None, shrink_threshold=None):
    """Finite support, non-negativity and shrink-wrap of cnn_propagator/fullfield.py:359-368 on the native object
    [..., 2] float32 (CUDA), IN PLACE: obj <- clip(obj * mask, 0); with shrink_threshold (the reference: 1e-15) the
    mask (float32, one value per pixel, same leading shape) is multiplied by (delta > threshold)."""
    assert db_obj.is_cuda and db_obj.dtype == torch.float32 and db_obj.is_contiguous() and db_obj.shape[-1] == 2
    if mask is not None:
        assert mask.is_cuda and mask.dtype == torch.float32 and mask.is_contiguous() and mask.numel() == db_obj.numel() // 2
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    check(lib.bdof_finite_support(_ptr(db_obj), None if mask is None else _ptr(mask), db_obj.numel() // 2,
                                  -1.0 if shrink_threshold is None else float(shrink_threshold), st))
    return db_obj


def rotate_db_bilinear(db_obj, theta, out=None):
    """tf.contrib.image.rotate(stack([delta, beta], -1), theta, interpolation='BILINEAR') of the TF drivers
    (tensorflow_recon/fullfield.py:96, ptychography.py:39) on the native object [Z, Y, X, 2] (CUDA float32): bilinear taps,
    zero outside, rotation centre ((X-1)/2, (Z-1)/2).  `out` may be a batch-element view of a plan's [Z, B, Y, X, 2] object."""
    Z, Y, X, _ = db_obj.shape
    if out is None:
        out = torch.empty_like(db_obj)
    assert db_obj.is_contiguous() and out.shape == db_obj.shape and out.stride()[1:] == (X * 2, 2, 1) and out.stride(0) % 2 == 0
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    check(lib.bdof_rotate_bilinear(_ptr(db_obj), _ptr(out), out.stride(0) // 2, float(theta), Y, X, Z, st))
    return out


def rotate_db_bilinear_adjoint(grad_rot, theta, grad_obj):
    """grad_obj [Z, Y, X, 2] += transpose of rotate_db_bilinear applied to grad_rot (what TF's autodiff of the rotation does)."""
    Z, Y, X, _ = grad_obj.shape
    assert grad_rot.shape == grad_obj.shape and grad_rot.stride()[1:] == (X * 2, 2, 1) and grad_obj.is_contiguous()
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    check(lib.bdof_rotate_bilinear_adjoint(_ptr(grad_rot), grad_rot.stride(0) // 2, _ptr(grad_obj), float(theta), Y, X, Z, st))
    return grad_obj


def regularizers(db_obj, grad_db, loss, alpha_d=None, alpha_b=None, gamma=0.0):
    """L1 + 3-D TV regularisers of tensorflow_recon/fullfield.py:389-396 on the native object [Z,Y,X,2] (CUDA), fused in one
    pass (bdof_regularizers): their gradient is ADDED to grad_db [Z,Y,X,2] (None: value only) and their value to the 0-d float64
    device tensor `loss`, in place.  Returns loss."""
    assert db_obj.is_cuda and db_obj.dtype == torch.float32 and db_obj.is_contiguous() and db_obj.dim() == 4 and db_obj.shape[-1] == 2
    assert loss.is_cuda and loss.dtype == torch.float64 and loss.numel() == 1
    if grad_db is not None:
        assert grad_db.is_cuda and grad_db.dtype == torch.float32 and grad_db.is_contiguous() and grad_db.shape == db_obj.shape
    Z, Y, X, _ = db_obj.shape
    work = torch.empty(1184, dtype=torch.float64, device=db_obj.device)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    check(lib.bdof_regularizers(_ptr(db_obj), None if grad_db is None else _ptr(grad_db), Z, Y, X, float(alpha_d or 0.0),
                                float(alpha_b or 0.0), float(gamma or 0.0), _ptr(loss), _ptr(work), st))
    return loss
