"""Developer probe: back-rotation time against the z-stride of the rotated gradient (TLB reach of the TMA box loads)."""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from beyond_dof_b200 import rotation
dev = torch.device('cuda')
n, B = 256, 10
thetas = np.linspace(0, np.pi, 180)[30:40]
tabs = [rotation.device_table([n, n, n], float(t), dev) for t in thetas]
for t in tabs:
    rotation.device_inverse(t)
out = torch.zeros((n, n, n, 2), device=dev)
def timeit(f, reps=5):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
db = torch.rand((n, B, n, n, 2), device=dev)
print('batch of 10 on [Z,10,Y,X] (z stride %.1f MB): %.3f ms' % (B * n * n * 8 / 1e6, timeit(lambda: rotation.rotate_db_adjoint_batch(db, tabs, out, accumulate=False))))
sep = [torch.rand((n, n, n, 2), device=dev) for _ in range(B)]
def per_angle():
    for b in range(B):
        rotation.rotate_db_adjoint(sep[b], tabs[b], out)
print('10 single-angle calls on contiguous [Z,Y,X] (z stride %.1f MB): %.3f ms' % (n * n * 8 / 1e6, timeit(per_angle)))
def per_angle_view():
    for b in range(B):
        rotation.rotate_db_adjoint(db[:, b], tabs[b], out)
print('10 single-angle calls on db[:, b] views (z stride %.1f MB): %.3f ms' % (B * n * n * 8 / 1e6, timeit(per_angle_view)))
obj = torch.rand((n, n, n, 2), device=dev)
def gathers():
    for b in range(B):
        rotation.rotate_db(obj, tabs[b], out=db[:, b])
print('10 gathers into db[:, b]: %.3f ms' % timeit(gathers))
def gathers_sep():
    for b in range(B):
        rotation.rotate_db(obj, tabs[b], out=sep[b])
print('10 gathers into contiguous arrays: %.3f ms' % timeit(gathers_sep))
