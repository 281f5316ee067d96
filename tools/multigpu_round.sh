#!/bin/bash
# Multi-GPU measurements of a round (run under `gpurun --gpus N`): N = number of ranks, R = round tag.
#   tools/multigpu_round.sh 8 r02 [extra variants: 1 = also the exchange variants of the default workload]
N=$1; R=${2:-r02}; VAR=${3:-0}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
run() { name=$1; shift; $TR bench.py --gpus $N "$@" > gpurun_out/${R}_bench_${N}gpu_${name}.json 2> gpurun_out/${R}_bench_${N}gpu_${name}.err; tail -c 200 gpurun_out/${R}_bench_${N}gpu_${name}.err; python - <<PY
import json
try:
    d = json.loads([l for l in open('gpurun_out/${R}_bench_${N}gpu_${name}.json') if l.startswith('{')][-1])
    print('${name}: value %.2f  ms/step %.3f  e2e %.2f' % (d['value'], d['ms_per_step'], d['e2e']['value']), d['config']['parallelism'][-90:])
except Exception as ex:
    print('${name}: no line', ex)
PY
}
run default --steps 10 --warmup 3
run config3 --workload config3 --steps 10 --warmup 3
run config4 --workload config4 --steps 10 --warmup 3
run config5 --workload config5 --steps 1 --warmup 1
if [ "$VAR" = "1" ]; then
  run default_nccl_reserve0 --steps 5 --warmup 3 --exchange nccl --sm-reserve 0
  run default_hybrid --steps 5 --warmup 3 --exchange hybrid
  run default_nccl_b16 --steps 5 --warmup 3 --exchange nccl --buckets 16
  run default_k10 --steps 3 --warmup 2 --fields-per-exchange 10
fi
