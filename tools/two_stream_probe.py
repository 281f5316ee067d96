"""Developer probe: small fields are latency-bound per launch -- do two half-batches on two streams overlap?"""
import sys, os, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from beyond_dof_b200.plan import MultislicePlan

N, Z = 256, 256
def make(B):
    plan = MultislicePlan(N, N, B, Z, 5000, 1e-7, free_prop_cm=1e-4, propagate_last=True, store_slices=True)
    db = torch.rand((Z, B, N, N, 2), device='cuda') * 1e-5
    plan.set_t_stash(db)
    return plan, db, torch.full((B, N, N), 0.9, device='cuda'), torch.empty((B, N, N), dtype=torch.complex64, device='cuda')
probe = torch.ones((N, N), dtype=torch.complex64, device='cuda')
def run(p):
    plan, db, tgt, ex = p
    plan.forward(db, probe, out=ex)
    loss, g = plan.loss_mag(ex, tgt)
    plan.adjoint(db, g)
def timeit(f, reps=5):
    for _ in range(2): f()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3
one = make(10)
print('one plan, 10 fields: %.3f ms' % timeit(lambda: run(one)))
for parts in (2, 5):
    ps = [make(10 // parts) for _ in range(parts)]
    ss = [torch.cuda.Stream() for _ in range(parts)]
    def multi():
        cur = torch.cuda.current_stream()
        for p, s in zip(ps, ss):
            s.wait_stream(cur)
            with torch.cuda.stream(s):
                run(p)
        for s in ss:
            cur.wait_stream(s)
    print('%d plans of %d fields on %d streams: %.3f ms' % (parts, 10 // parts, parts, timeit(multi)))
    def serial():
        for p in ps:
            run(p)
    print('%d plans of %d fields on one stream: %.3f ms' % (parts, 10 // parts, timeit(serial)))
