"""Developer tool: event timeline of one TomographyObjective update (BASELINE configs[3] shape): where does the step go?"""
import sys, os, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from beyond_dof_b200.models import TomographyObjective
from beyond_dof_b200 import rotation as _rot

n, mb = 256, 10
dev = torch.device('cuda', 0)
obj = torch.rand((n, n, n, 2), device=dev) * torch.tensor([8.7e-7, 5.1e-8], device=dev)
probe = torch.ones((n, n), dtype=torch.complex64, device=dev)
tomo = TomographyObjective(obj, probe, 5000, 1e-7, minibatch_size=mb, free_prop_cm=1e-4, propagate_last=True, step_size=1e-7)
thetas = np.linspace(0, np.pi, 180)
tomo.prepare(thetas)
prj = (0.9 + 0.1 * torch.rand((mb, n, n))).pin_memory()
for i in range(3):
    tomo.step(thetas[i * mb:(i + 1) * mb], prj)
self = tomo
names = ['h2d', 'rotate', 'forward', 'loss', 'adjoint', 'back-rotate', 'adam']
for it in range(3):
    theta_batch = thetas[(3 + it) * mb:(4 + it) * mb]
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)]
    cpu = [time.perf_counter()]
    torch.cuda.synchronize()
    ev[0].record()
    self.target.copy_(prj, non_blocking=True); ev[1].record(); cpu.append(time.perf_counter())
    tabs = [_rot.device_table(self.shape, float(t), dev) for t in theta_batch]
    for b in range(self.B):
        _rot.rotate_db(self.obj, tabs[b], out=self.db[:, b])
    ev[2].record(); cpu.append(time.perf_counter())
    self.plan.forward(self.db, self.probe, out=self.exit); ev[3].record(); cpu.append(time.perf_counter())
    loss, g = self.plan.loss_mag(self.exit, self.target); ev[4].record(); cpu.append(time.perf_counter())
    self.plan.adjoint(self.db, g); ev[5].record(); cpu.append(time.perf_counter())
    _rot.rotate_db_adjoint_batch(self.db, tabs, self.grad, accumulate=False); ev[6].record(); cpu.append(time.perf_counter())
    _rot.adam_step(self.obj, self.grad, self.i_batch, self.m, self.v, step_size=self.step_size); ev[7].record(); cpu.append(time.perf_counter())
    torch.cuda.synchronize()
    print('iter %d device ms: ' % it + ', '.join('%s %.3f' % (names[k], ev[k].elapsed_time(ev[k + 1])) for k in range(len(names)))
          + ' | total %.3f' % ev[0].elapsed_time(ev[-1]))
    print('        cpu ms:    ' + ', '.join('%s %.3f' % (names[k], (cpu[k + 1] - cpu[k]) * 1e3) for k in range(len(names))))
t0 = time.perf_counter()
for i in range(5):
    tomo.step(thetas[i * mb:(i + 1) * mb], prj)
print('step() wall: %.3f ms' % ((time.perf_counter() - t0) / 5 * 1e3))
