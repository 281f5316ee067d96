"""Developer tool: is a small-field plan launch-bound?  CPU enqueue time vs device time of forward / adjoint."""
import sys, os, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from beyond_dof_b200.plan import MultislicePlan

B, N, Z = (int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (10, 256, 256)))
plan = MultislicePlan(N, N, B, Z, 5000, 1e-7, free_prop_cm=1e-4, propagate_last=True, store_slices=True)
db = torch.rand((Z, B, N, N, 2), device='cuda') * 1e-5
plan.set_t_stash(db)
probe = torch.ones((N, N), dtype=torch.complex64, device='cuda')
tgt = torch.full((B, N, N), 0.9, device='cuda')
ex = torch.empty((B, N, N), dtype=torch.complex64, device='cuda')
for it in range(6):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    plan.forward(db, probe, out=ex)
    t1 = time.perf_counter()
    loss, g = plan.loss_mag(ex, tgt)
    t2 = time.perf_counter()
    plan.adjoint(db, g)
    t3 = time.perf_counter()
    torch.cuda.synchronize()
    t4 = time.perf_counter()
    lt = plan.last_times()
    f_ms, a_ms = lt['forward'][0], lt['adjoint'][0]
    print('iter %d: enqueue forward %.2f ms, loss %.2f ms, adjoint %.2f ms | device forward %.2f ms adjoint %.2f ms | wall %.2f ms'
          % (it, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, f_ms, a_ms, (t4 - t0) * 1e3))
