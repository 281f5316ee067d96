TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu"
summ() { python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value %.2f  ms/step %.2f'%(d['value'],d['ms_per_step']), d['config']['parallelism'][-60:])
"; }
echo "G pdl0 reserve 20 ctas 16"; BDOF_PDL=0 $TR 2>/dev/null | summ
echo "H pdl0 reserve 20 no limit"; BDOF_PDL=0 NCCL_MAX_CTAS=64 $TR --sm-reserve 20 2>/dev/null | summ
echo "I pdl0 reserve 0"; BDOF_PDL=0 $TR --sm-reserve 0 2>/dev/null | summ
echo "J pdl0 reserve 20 1 bucket"; BDOF_PDL=0 NCCL_MAX_CTAS=64 $TR --sm-reserve 20 --buckets 1 2>/dev/null | summ
echo "K pdl1 reserve 0 1 bucket"; $TR --sm-reserve 0 --buckets 1 2>/dev/null | summ
