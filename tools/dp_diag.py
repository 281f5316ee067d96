"""torchrun diagnostic of the copy-engine exchange: exchange alone (1 and 8 buckets), one peer copy, the shard sum."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from beyond_dof_b200.dist import CopyEngineExchange

rank = int(os.environ['RANK']); world = int(os.environ['WORLD_SIZE']); lr = int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(lr)
dist.init_process_group('nccl', device_id=torch.device('cuda', lr))
shape = (256, 1, 2048, 2048, 2)
nb = int(os.environ.get('NB', '8'))
ex = CopyEngineExchange(shape, n_buckets=nb)
ex.grad.fill_(1.0)
st = torch.cuda.current_stream()

def timed(fn, reps=3):
    for _ in range(2):
        fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps

def one_bucket():
    ex.exchange(None); ex.finish()

def n_buckets():
    per = 256 // nb
    ev = torch.cuda.Event(); ev.record(st)
    ex.exchange([(256 - (j + 1) * per, 256 - j * per, ev) for j in range(nb)]); ex.finish()

t1 = timed(one_bucket)
t8 = timed(n_buckets)
gb = ex.grad.numel() * 4 / 1e9
# plain local copy and NCCL all-reduce of the same buffer for scale
tmp = torch.empty_like(ex.grad)
tc = timed(lambda: tmp.copy_(ex.grad))
tn = timed(lambda: dist.all_reduce(ex.grad, op=dist.ReduceOp.AVG))
if rank == 0:
    print('world %d  grad %.2f GB: exchange 1 bucket %.2f ms, %d buckets %.2f ms; local copy %.2f ms; NCCL all-reduce %.2f ms'
          % (world, gb, t1, nb, t8, tc, tn), flush=True)
dist.barrier()
ex.close()
dist.destroy_process_group()
