import torch
a = torch.empty(1 << 30, dtype=torch.bfloat16, device='cuda'); b = torch.empty_like(a)
best = 0
for i in range(10):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); b.copy_(a); e1.record(); torch.cuda.synchronize()
    best = max(best, 2 * a.numel() * 2 / (e0.elapsed_time(e1) * 1e-3) / 1e9)
print('copy bandwidth GB/s', round(best, 1))
