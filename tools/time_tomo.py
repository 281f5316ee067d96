"""Developer tool: where a TomographyObjective step spends its time (CUDA events per phase)."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from beyond_dof_b200.models import TomographyObjective
from beyond_dof_b200 import rotation as rot
n, mb = 256, 10
dev = torch.device('cuda')
obj = torch.rand((n, n, n, 2), device=dev) * 1e-6
tomo = TomographyObjective(obj, torch.ones((n, n), dtype=torch.complex64, device=dev), 5000, 1e-7, mb, free_prop_cm=1e-4)
thetas = np.linspace(0.1, 3.0, mb)
tgt = torch.rand((mb, n, n), device=dev) * 0.1 + 0.9
tabs = [rot.device_table(tomo.shape, float(t), dev) for t in thetas]
for t in tabs: rot.device_inverse(t)
def timed(name, fn, reps=3):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    print('%-28s %.3f ms' % (name, a.elapsed_time(b) / reps))
timed('rotate x10', lambda: [rot.rotate_db(tomo.obj, tabs[b], out=tomo.db[:, b]) for b in range(mb)])
timed('forward', lambda: tomo.plan.forward(tomo.db, tomo.probe, out=tomo.exit))
loss, g = tomo.plan.loss_mag(tomo.exit, tgt)
timed('loss', lambda: tomo.plan.loss_mag(tomo.exit, tgt))
timed('forward+adjoint', lambda: (tomo.plan.forward(tomo.db, tomo.probe, out=tomo.exit), tomo.plan.adjoint(tomo.db, g)))
timed('zero grad', lambda: tomo.grad.zero_())
timed('back-rotate x10 (csr)', lambda: [rot.rotate_db_adjoint(tomo.db[:, b], tabs[b], tomo.grad) for b in range(mb)])
timed('back-rotate x10 (atomic)', lambda: [rot.rotate_db_adjoint(tomo.db[:, b], tabs[b], tomo.grad, atomic=True) for b in range(mb)])
timed('adam', lambda: rot.adam_step(tomo.obj, tomo.grad, 0, tomo.m, tomo.v, step_size=1e-9))
timed('whole step', lambda: tomo.loss_and_grad(thetas, tgt))
