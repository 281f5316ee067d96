"""Developer tool: plain NCCL all-reduce bandwidth between the ranks of one box (no other work running)."""
import os, time, torch, torch.distributed as dist
rank = int(os.environ['RANK']); lr = int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(lr)
dist.init_process_group('nccl', device_id=torch.device('cuda', lr))
for mb in (64, 268, 1024):
    x = torch.ones(mb * 1024 * 1024 // 4, device='cuda')
    for _ in range(3):
        dist.all_reduce(x)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 5
    for _ in range(n):
        dist.all_reduce(x)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / n
    if rank == 0:
        print('all_reduce %5d MB: %.2f ms  algbw %.1f GB/s' % (mb, dt * 1e3, mb / 1024 / dt), flush=True)
if rank == 0:
    print(torch.cuda.get_device_name(0), 'p2p 0->1:', torch.cuda.can_device_access_peer(0, 1))
dist.destroy_process_group()
