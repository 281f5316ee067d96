"""torchrun check + timing of the hybrid exchange (NCCL reduce-scatter + copy-engine all-gather) on real multi-GPU."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from beyond_dof_b200.dist import CopyEngineExchange

rank = int(os.environ['RANK']); world = int(os.environ['WORLD_SIZE']); lr = int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(lr)
dist.init_process_group('nccl', device_id=torch.device('cuda', lr))
nb = int(os.environ.get('NB', '16'))
shape = (256, 1, 2048, 2048, 2)
ex = CopyEngineExchange(shape, n_buckets=nb, gather_only=True)
comm = torch.cuda.Stream()
st = torch.cuda.current_stream()
per = 256 // nb
base = torch.arange(2048 * 2048 * 2, device='cuda', dtype=torch.float32).remainder_(977.0).view(1, 2048, 2048, 2)

def run():
    ev = torch.cuda.Event(); ev.record(st)
    ex.reduce_scatter_gather([(256 - (j + 1) * per, 256 - j * per, ev) for j in range(nb)], comm)
    ex.finish()

worst = 0.0
for step in range(3):
    for z in range(0, 256, 32):
        ex.grad[z:z + 32] = base * float(rank + 1 + step) + float(z)
    torch.cuda.synchronize(); dist.barrier()
    run()
    torch.cuda.synchronize()
    mean_scale = sum(r + 1 + step for r in range(world)) / world
    for z in (0, 97, 255):
        want = base[0] * mean_scale + float((z // 32) * 32)
        worst = max(worst, float((ex.grad[z, 0] - want).abs().max() / want.abs().max()))
    dist.barrier()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
a.record()
for _ in range(3):
    run()
b.record(); torch.cuda.synchronize()
t = torch.tensor([worst], device='cuda'); dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print('hybrid world %d, %d buckets: max rel error %.2e; exchange alone %.2f ms' % (world, nb, t.item(), a.elapsed_time(b) / 3), flush=True)
dist.barrier()
ex.close()
dist.destroy_process_group()
