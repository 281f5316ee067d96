"""ctypes binding of libbdof.so (include/bdof.h).  No torch types cross this boundary: only raw
device/host pointers and sizes.  Import fails loudly if the library has not been built."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, os.environ.get('BDOF_LIB', 'libbdof.so'))

# plan flags / modes (mirror include/bdof.h)
PROPAGATE_LAST = 1 << 0
STORE_SLICES = 1 << 1
Z_BROADCAST = 1 << 2
STEPWISE = 1 << 3
FREE_NONE, FREE_INF, FREE_TF = 0, 1, 2

# every symbol include/bdof.h declares (checked by tests/test_capi.py)
SYMBOLS = [
    'bdof_version', 'bdof_last_error', 'bdof_launch_count', 'bdof_size_supported', 'bdof_kernel_factors',
    'bdof_plan_create', 'bdof_plan_destroy', 'bdof_set_kernel', 'bdof_set_kernel_full', 'bdof_set_free_prop',
    'bdof_forward', 'bdof_loss_mag', 'bdof_adjoint', 'bdof_pack_db', 'bdof_unpack_db', 'bdof_patch_gather',
    'bdof_patch_scatter_add', 'bdof_cnn_forward', 'bdof_forward_host', 'bdof_plan_workspace_bytes',
    'bdof_free_prop', 'bdof_profile_begin', 'bdof_profile_end', 'bdof_debug_set_buffer', 'bdof_slice_step',
    'bdof_plan_set_bucket_events', 'bdof_set_sm_reserve', 'bdof_rotate_gather', 'bdof_rotate_scatter_add', 'bdof_rotate_adjoint_csr', 'bdof_rotate_adjoint_csr_batch', 'bdof_rotate_adjoint_csr_batch_range', 'bdof_adam_step',
    'bdof_finite_support', 'bdof_plan_set_t_stash', 'bdof_rotate_bilinear', 'bdof_rotate_bilinear_adjoint',
    'bdof_dp_create', 'bdof_dp_destroy', 'bdof_dp_handle_bytes', 'bdof_dp_export', 'bdof_dp_connect', 'bdof_dp_grad_ptr',
    'bdof_dp_bucket', 'bdof_dp_gather', 'bdof_dp_finish', 'bdof_plan_last_times', 'bdof_pack_db_rows', 'bdof_unpack_db_rows', 'bdof_plan_set_stream', 'bdof_debug_fft_gain', 'bdof_debug_rot_split', 'bdof_field_multiply', 'bdof_patch_gather_add', 'bdof_plan_set_windows', 'bdof_plan_is_resident', 'bdof_cnn_forward_store', 'bdof_cnn_adjoint', 'bdof_free_prop_adjoint',
    'bdof_tiles_create', 'bdof_tiles_destroy', 'bdof_tiles_handle_bytes', 'bdof_tiles_export', 'bdof_tiles_connect',
    'bdof_tiles_block_ptr', 'bdof_tiles_cut', 'bdof_tiles_paste', 'bdof_tiles_halo_exchange', 'bdof_slice_step_seq', 'bdof_plan_set_grad_accumulate', 'bdof_regularizers', 'bdof_slice_step_windows',
]


class BdofError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__('libbdof error %d: %s' % (code, msg))
        self.code = code


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError('libbdof.so is missing (%s): build it with `python -m beyond_dof_b200.build`; '
                          'there is no CPU fallback' % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    vp, i32, u32, f64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_uint32, ctypes.c_double
    i64 = ctypes.c_longlong
    lib.bdof_version.restype = i32
    lib.bdof_last_error.restype = ctypes.c_char_p
    lib.bdof_launch_count.restype = ctypes.c_ulonglong
    lib.bdof_size_supported.argtypes = [i32]
    lib.bdof_kernel_factors.argtypes = [f64, f64, vp, i32, i32, f64, vp, vp, vp]
    lib.bdof_plan_create.argtypes = [ctypes.POINTER(vp), i32, i32, i32, i32, u32, vp]
    lib.bdof_plan_destroy.argtypes = [vp]
    lib.bdof_plan_destroy.restype = None
    lib.bdof_set_kernel.argtypes = [vp, vp, vp, f64, f64, f64]
    lib.bdof_set_kernel_full.argtypes = [vp, vp, f64]
    lib.bdof_set_free_prop.argtypes = [vp, i32, vp, vp, f64, f64]
    lib.bdof_forward.argtypes = [vp, vp, vp, vp]
    lib.bdof_loss_mag.argtypes = [vp, vp, vp, f64, vp, vp]
    lib.bdof_adjoint.argtypes = [vp, vp, vp, vp, vp]
    lib.bdof_pack_db.argtypes = [vp, vp, vp, i32, i32, i32, i32, vp]
    lib.bdof_unpack_db.argtypes = [vp, vp, vp, i32, i32, i32, i32, vp]
    lib.bdof_patch_gather.argtypes = [vp, i32, i32, i32, vp, i32, i32, i32, vp, vp]
    lib.bdof_patch_scatter_add.argtypes = [vp, i32, i32, i32, vp, i32, i32, i32, vp, vp]
    lib.bdof_patch_gather_add.argtypes = [vp, i32, i32, i32, vp, i32, i32, i32, vp, vp]
    lib.bdof_cnn_forward.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, vp, i32, f64, vp]
    lib.bdof_forward_host.argtypes = [vp, vp, vp, vp, vp]
    lib.bdof_cnn_forward_store.argtypes = [vp, vp, vp, i32, i32, i32, i32, vp, i32, f64, vp]
    lib.bdof_cnn_adjoint.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, vp, i32, f64, vp]
    lib.bdof_free_prop_adjoint.argtypes = [vp, vp, vp]
    lib.bdof_plan_workspace_bytes.argtypes = [vp, ctypes.POINTER(ctypes.c_size_t)]
    lib.bdof_free_prop.argtypes = [vp, vp, vp]
    lib.bdof_profile_begin.argtypes = [vp]
    lib.bdof_debug_set_buffer.argtypes = [vp]
    lib.bdof_slice_step.argtypes = [vp, vp, vp, vp, i32]
    lib.bdof_slice_step_seq.argtypes = [vp, vp, vp, vp, i32, i32]
    lib.bdof_slice_step_windows.argtypes = [vp, vp, i64, vp, vp, vp, i32]
    lib.bdof_plan_set_bucket_events.argtypes = [vp, i32, vp]
    lib.bdof_set_sm_reserve.argtypes = [i32]
    lib.bdof_profile_end.argtypes = [vp, i32, vp, vp]
    i64 = ctypes.c_longlong
    lib.bdof_rotate_gather.argtypes = [vp, vp, vp, i64, i32, i32, i32, vp]
    lib.bdof_rotate_scatter_add.argtypes = [vp, i64, vp, vp, i32, i32, i32, vp]
    lib.bdof_rotate_adjoint_csr.argtypes = [vp, i64, vp, vp, vp, i32, i32, i32, vp]
    lib.bdof_rotate_adjoint_csr_batch.argtypes = [vp, i64, i64, i32, ctypes.POINTER(vp), ctypes.POINTER(vp), vp, i32, i32, i32, i32, vp]
    lib.bdof_rotate_adjoint_csr_batch_range.argtypes = [vp, i64, i64, i32, ctypes.POINTER(vp), ctypes.POINTER(vp), vp, i32, i32, i32, i32, i32, i32, vp]
    lib.bdof_adam_step.argtypes = [vp, vp, vp, vp, i64, i32, f64, f64, f64, f64, vp]
    lib.bdof_finite_support.argtypes = [vp, vp, i64, f64, vp]
    lib.bdof_regularizers.argtypes = [vp, vp, i32, i32, i32, f64, f64, f64, vp, vp, vp]
    lib.bdof_plan_set_t_stash.argtypes = [vp, vp]
    lib.bdof_plan_set_grad_accumulate.argtypes = [vp, i32]
    lib.bdof_rotate_bilinear.argtypes = [vp, vp, i64, f64, i32, i32, i32, vp]
    lib.bdof_rotate_bilinear_adjoint.argtypes = [vp, i64, vp, f64, i32, i32, i32, vp]
    sz = ctypes.c_size_t
    lib.bdof_dp_create.argtypes = [ctypes.POINTER(vp), i32, i32, sz, i32, i32]
    lib.bdof_dp_destroy.argtypes = [vp]
    lib.bdof_dp_destroy.restype = None
    lib.bdof_dp_handle_bytes.argtypes = []
    lib.bdof_dp_export.argtypes = [vp, vp]
    lib.bdof_dp_connect.argtypes = [vp, vp]
    lib.bdof_dp_grad_ptr.argtypes = [vp, ctypes.POINTER(vp)]
    lib.bdof_dp_bucket.argtypes = [vp, sz, sz, vp]
    lib.bdof_dp_gather.argtypes = [vp, sz, sz, vp]
    lib.bdof_dp_finish.argtypes = [vp, vp]
    lib.bdof_tiles_create.argtypes = [ctypes.POINTER(vp), i32, i32, i32, i32, i32, i32]
    lib.bdof_tiles_destroy.argtypes = [vp]
    lib.bdof_tiles_destroy.restype = None
    lib.bdof_tiles_handle_bytes.argtypes = []
    lib.bdof_tiles_export.argtypes = [vp, vp]
    lib.bdof_tiles_connect.argtypes = [vp, vp]
    lib.bdof_tiles_block_ptr.argtypes = [vp, i32, ctypes.POINTER(vp)]
    lib.bdof_tiles_cut.argtypes = [vp, i32, vp, i32, i32, i32, vp, vp]
    lib.bdof_tiles_paste.argtypes = [vp, i32, vp, vp, vp, i32, i32, i32, vp]
    lib.bdof_tiles_halo_exchange.argtypes = [vp, i32, vp]
    lib.bdof_plan_last_times.argtypes = [vp, vp, vp]
    lib.bdof_plan_set_stream.argtypes = [vp, vp]
    lib.bdof_debug_fft_gain.argtypes = [i32, vp]
    lib.bdof_debug_rot_split.argtypes = [i32, i32, vp, vp]
    lib.bdof_field_multiply.argtypes = [vp, vp, vp, i32, i64, vp]
    lib.bdof_plan_set_windows.argtypes = [vp, i32, i32, vp]
    lib.bdof_plan_is_resident.argtypes = [vp]
    lib.bdof_pack_db_rows.argtypes = [vp, vp, vp, i64, i64, i32, i32, i32, vp]
    lib.bdof_unpack_db_rows.argtypes = [vp, vp, vp, i64, i64, i32, i32, i32, vp]
    for name in SYMBOLS:
        fn = getattr(lib, name)
        if fn.restype is ctypes.c_int and name not in ('bdof_version', 'bdof_size_supported', 'bdof_dp_handle_bytes', 'bdof_plan_is_resident', 'bdof_tiles_handle_bytes'):
            fn.restype = i32
    return lib


lib = _load()


def check(rc):
    if rc != 0:
        raise BdofError(rc, lib.bdof_last_error().decode('utf-8', 'replace'))


def launch_count():
    return int(lib.bdof_launch_count())
