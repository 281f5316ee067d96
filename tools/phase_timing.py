"""Developer tool: per-phase cycle breakdown of the line kernels (needs the instrumented build:
BDOF_ALT=9 python -m beyond_dof_b200.build; run with BDOF_LIB=libbdof_alt9.so)."""
import ctypes, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from beyond_dof_b200 import capi
from beyond_dof_b200.plan import MultislicePlan

N = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
Z = 3
plan = MultislicePlan(N, N, 1, Z, 5000, 1e-7, store_slices=True)
db = torch.rand((Z, 1, N, N, 2), device='cuda') * 1e-5
probe = torch.ones((N, N), dtype=torch.complex64, device='cuda')
buf = torch.zeros((9 << 17,), dtype=torch.int64, device='cuda')
for it in range(3):
    psi = plan.forward(db, probe)
    _, g = plan.loss_mag(psi, torch.full((1, N, N), 0.9, device='cuda'))
    if it == 2:
        capi.check(capi.lib.bdof_debug_set_buffer(ctypes.c_void_p(buf.data_ptr())))
    plan.adjoint(db, g, grad_out=torch.empty_like(db)) if it < 2 else None
buf.zero_()
psi = plan.forward(db, probe)
_, g = plan.loss_mag(psi, torch.full((1, N, N), 0.9, device='cuda'))
plan.adjoint(db, g, grad_out=torch.empty_like(db))
torch.cuda.synchronize()
b = buf.cpu().numpy().reshape(9, -1, 32)
names = ['start', 'loaded', 'tables', 'A:bfly1', 'A:exch+tw', 'A:bfly2', 'A:h-mul', 'B:bfly1', 'B:exch+tw', 'B:bfly2', '-', 'stored']
for v, vn in ((0, 'row_conv_transmit'), (5, 'col_conv'), (2, 'row_conv_adjoint')):
    r = b[v]
    r = r[r[:, 0] != 0]
    print('== %s: %d warps recorded' % (vn, len(r)))
    for ti in (0, 1):
        st = r[:, ti * 12:(ti + 1) * 12].astype(np.float64)
        ok = st[:, 0] != 0
        if ok.sum() == 0:
            continue
        st = st[ok]
        t0 = st[:, 0:1]
        cols = [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 11]
        prev = st[:, 0]
        line = []
        for c in cols[1:]:
            cur = st[:, c]
            line.append('%s %.0f' % (names[c], np.mean(cur - prev)))
            prev = cur
        print('  tile %d (%d warps): total %.0f cycles | ' % (ti, ok.sum(), np.mean(st[:, 11] - st[:, 0])) + ' | '.join(line))
    first = r[:, 0].min(); last = max(r[:, 11].max(), r[:, 23].max())
    print('  kernel span (first start -> last store): %.0f cycles' % (last - first))
