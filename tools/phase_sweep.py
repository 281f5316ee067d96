"""Developer tool: per-phase cycle breakdown of the sweep kernels (instrumented build:
BDOF_ALT=9 python -m beyond_dof_b200.build; run with BDOF_LIB=libbdof_alt9.so)."""
import ctypes, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from beyond_dof_b200 import capi
from beyond_dof_b200.plan import MultislicePlan

N = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
Z = 6
plan = MultislicePlan(N, N, 1, Z, 5000, 1e-7, store_slices=True)
db = torch.rand((Z, 1, N, N, 2), device='cuda') * 1e-5
probe = torch.ones((N, N), dtype=torch.complex64, device='cuda')
buf = torch.zeros((4 << 17,), dtype=torch.int64, device='cuda')
tgt = torch.full((1, N, N), 0.9, device='cuda')
gout = torch.empty_like(db)
plan.set_t_stash(gout)            # as in production: the forward leaves tau where the adjoint writes the gradient
for it in range(2):
    psi = plan.forward(db, probe)
    _, g = plan.loss_mag(psi, tgt)
    plan.adjoint(db, g, grad_out=gout)
capi.check(capi.lib.bdof_debug_set_buffer(ctypes.c_void_p(buf.data_ptr())))
buf.zero_()
psi = plan.forward(db, probe)
_, g = plan.loss_mag(psi, tgt)
plan.adjoint(db, g, grad_out=gout)
torch.cuda.synchronize()
b = buf.cpu().numpy().reshape(4, -1, 32)
names = ['start', 'landed', 'read/swap', 'conv1', 'db-wait', 'transmit', 'mid-a', 'slab-wait', 'mid-b', 'conv2', 'stored']
gbase = min(b[k][b[k][:, 27] != 0][:, 27].min() for k in range(4) if (b[k][:, 27] != 0).any())
for k, kn in enumerate(['x forward', 'x adjoint', 'y forward', 'y adjoint']):
    r = b[k]
    r = r[r[:, 0] != 0]
    print('== %s N=%d: %d warps' % (kn, N, len(r)))
    for ti in range(2):
        st = r[:, ti * 16:ti * 16 + 11].astype(np.float64)
        ok = st[:, 0] != 0
        if ok.sum() == 0:
            continue
        st = st[ok]
        out = []
        prev = st[:, 0]
        for c in range(1, 11):
            cur = st[:, c]
            if np.all(cur == 0):
                continue
            out.append('%s %.0f' % (names[c], np.mean(cur - prev)))
            prev = cur
        print('  tile %d (%d warps): total %.0f | ' % (ti, ok.sum(), np.mean(st[:, 10] - st[:, 0])) + ' | '.join(out))

    # kernel-level timeline from %globaltimer (ns): entry, after prologue, after griddepcontrol.wait, exit
    g = r[:, 27:31].astype(np.float64)
    t0 = gbase
    print('  kernel timeline (us, relative to the first recorded CTA entry of any sweep kernel; forward records slices 4 (x) and 5 (y), adjoint 1 (y) and 0 (x): entry %.1f..%.1f | prologue done %.1f..%.1f | dependency wait done %.1f..%.1f | exit %.1f..%.1f'
          % tuple(x / 1e3 for x in ((g[:, 0] - t0).min(), (g[:, 0] - t0).max(), (g[:, 1] - t0).min(), (g[:, 1] - t0).max(),
                                   (g[:, 2] - t0).min(), (g[:, 2] - t0).max(), (g[:, 3] - t0).min(), (g[:, 3] - t0).max())))
