run() { echo "== $*"; env "$@" ; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu"
summ() { python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value %.2f  ms/step %.2f'%(d['value'],d['ms_per_step']), d['config']['parallelism'][-60:])
"; }
echo "A reserve 0"; $TR --sm-reserve 0 2>/dev/null | summ
echo "B auto (20, max ctas 16)"; $TR 2>/dev/null | summ
echo "C reserve 20 no cta limit"; NCCL_MAX_CTAS=64 $TR --sm-reserve 20 2>/dev/null | summ
echo "D auto, 16 buckets"; $TR --buckets 16 2>/dev/null | summ
echo "E reserve 12"; $TR --sm-reserve 12 2>/dev/null | summ
echo "F auto, 4 buckets"; $TR --buckets 4 2>/dev/null | summ
