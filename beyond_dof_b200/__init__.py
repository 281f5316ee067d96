"""beyond_dof_b200: B200-native Fresnel multislice engine (drop-in for the multislice hot path of
mdw771/beyond_dof).  Python here is only the host side; the arithmetic lives in libbdof.so
(hand-written sm_100a CUDA behind the C ABI of include/bdof.h).

Submodules are imported lazily so that `python -m beyond_dof_b200.build` works before the
library exists; touching any compute entry point without libbdof.so raises ImportError.
"""
import importlib

__version__ = '0.1.0'

_LAZY = {
    'multislice_propagate_batch_numpy': 'propagation',
    'multislice_propagate_batch': 'propagation',
    'multislice_propagate': 'propagation',
    'multislice_propagate_cnn': 'propagation',
    'cnn_loss_and_grad': 'propagation',
    'MultislicePlan': 'plan',
    'fullfield_loss_and_grad': 'models',
    'ptycho_loss_and_grad': 'models',
    'TomographyObjective': 'models',
    'PtychographyObjective': 'models',
    'FullfieldObjective': 'models',
    'fullfield_loss_and_grad_host': 'models',
    'apply_rotation': 'rotation',
    'rotation_table': 'rotation',
    'adam_step': 'rotation',
    'finite_support': 'rotation',
    'ptycho_position_losses': 'models',
    'dynamic_dropping': 'models',
    'get_kernel': 'util',
    'get_kernel_ir': 'propagation',
    'create_fullfield_data_numpy': 'simulation',
    'create_ptychography_data_batch_numpy': 'simulation',
    'gen_mesh': 'util',
}


def __getattr__(name):
    if name in _LAZY:
        mod = importlib.import_module('.' + _LAZY[name], __name__)
        return getattr(mod, name)
    if name in ('capi', 'plan', 'propagation', 'models', 'util', 'build', 'dist', 'rotation', 'tiling', 'np_funcs', 'simulation'):
        return importlib.import_module('.' + name, __name__)
    raise AttributeError(name)
