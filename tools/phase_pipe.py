"""Developer tool: per-phase cycle breakdown of the pipelined column pass (instrumented build:
BDOF_ALT=9 python -m beyond_dof_b200.build; run with BDOF_LIB=libbdof_alt9.so)."""
import ctypes, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from beyond_dof_b200 import capi
from beyond_dof_b200.plan import MultislicePlan

N = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
Z = 3
plan = MultislicePlan(N, N, 1, Z, 5000, 1e-7, store_slices=False)
db = torch.rand((Z, 1, N, N, 2), device='cuda') * 1e-5
probe = torch.ones((N, N), dtype=torch.complex64, device='cuda')
buf = torch.zeros((10 << 17,), dtype=torch.int64, device='cuda')
for it in range(3):
    plan.forward(db, probe)
capi.check(capi.lib.bdof_debug_set_buffer(ctypes.c_void_p(buf.data_ptr())))
buf.zero_()
plan.forward(db, probe)
torch.cuda.synchronize()
b = buf.cpu().numpy().reshape(10, -1, 32)
r = b[5]
r = r[r[:, 0] != 0]
names = ['start', 'landed', 'read', 'issued', 'conv-A', 'conv-B', 'stored']
print('== pipelined col_conv N=%d: %d warps recorded' % (N, len(r)))
for ti in range(4):
    st = r[:, ti * 8:ti * 8 + 7].astype(np.float64)
    ok = st[:, 0] != 0
    if ok.sum() == 0:
        continue
    st = st[ok]
    d = np.diff(st, axis=1)
    print('  tile %d (%d warps): total %.0f | ' % (ti, ok.sum(), np.mean(st[:, 6] - st[:, 0])) +
          ' | '.join('%s %.0f' % (names[i + 1], np.mean(d[:, i])) for i in range(6)))
first = r[:, 0].min()
last = max(r[:, 8 * k + 6].max() for k in range(4))
print('  span first start -> last recorded store: %.0f cycles' % (last - first))
