#!/bin/bash
# One parametrised driver for the bench experiments of DESIGN.md 4.5 / 6 (replaces the round-1 exp*.sh one-offs).
#   tools/sweep_experiments.sh shapes  "1,2048,2048,128 1,4096,4096,48"  [ENV=VAL ...]   single-GPU value + per-kernel times
#   tools/sweep_experiments.sh dp N    "--exchange nccl --buckets 8"     [ENV=VAL ...]   N-GPU step under torchrun
mode=$1; shift
summ() { python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l)
        print('value %.2f  ms/step %.3f' % (d['value'], d['ms_per_step']), {k: round(v['avg_ms'] * 1e3, 1) for k, v in (d.get('kernels') or {}).items()})
    else:
        print(l.strip()[:300])
"; }
case $mode in
shapes)
    shapes=$1; shift
    for shape in $shapes; do
        echo "== shape=$shape $*"
        env "$@" python bench.py --steps 3 --warmup 3 --no-cpu --no-host-object --shape $shape 2>&1 | summ
    done ;;
dp)
    n=$1; flags=$2; shift 2
    echo "== dp$n $flags $*"
    env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
        bench.py --gpus $n --steps 3 --warmup 3 --no-cpu $flags 2>/dev/null | summ ;;
*) echo "usage: $0 shapes|dp ..."; exit 2 ;;
esac
