#!/bin/bash
# Multi-GPU measurements of a round (run under `gpurun --gpus N`): N = number of ranks, R = round tag, MODE = which lines:
#   all      default (K = 10 fields per exchange) + config3 + config4 + config5
#   default  the default workload only
#   variants the exchange variants of the harsh one-field-per-exchange ratio (K = 1)
N=$1; R=${2:-r02}; MODE=${3:-all}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
run() { name=$1; shift; $TR bench.py --gpus $N "$@" > gpurun_out/${R}_bench_${N}gpu_${name}.json 2> gpurun_out/${R}_bench_${N}gpu_${name}.err; python - <<PY
import json
try:
    d = json.loads([l for l in open('gpurun_out/${R}_bench_${N}gpu_${name}.json') if l.startswith('{')][-1])
    print('${name}: value %.2f  ms/step %.3f  e2e %.2f' % (d['value'], d['ms_per_step'], d['e2e']['value']), d['config']['parallelism'][-90:])
except Exception as ex:
    print('${name}: no line', ex)
    print(open('gpurun_out/${R}_bench_${N}gpu_${name}.err').read()[-600:])
PY
}
#   c35      config3 + config5 only
#   c4       config4 only
if [ "$MODE" = "c4" ]; then run config4 --workload config4 --steps 10 --warmup 3; exit 0; fi
if [ "$MODE" = "c3" ]; then run config3 --workload config3 --steps 10 --warmup 3; exit 0; fi
if [ "$MODE" != "c35" ]; then run default --steps 5 --warmup 3; fi
if [ "$MODE" = "c35" ]; then
  run config3 --workload config3 --steps 10 --warmup 3
  run config5 --workload config5 --steps 1 --warmup 1
fi
if [ "$MODE" = "all" ]; then
  run config3 --workload config3 --steps 10 --warmup 3
  run config4 --workload config4 --steps 10 --warmup 3
  run config5 --workload config5 --steps 1 --warmup 1
fi
if [ "$MODE" = "variants" ]; then
  run k1_auto --steps 5 --warmup 3 --fields-per-exchange 1
  run k1_nccl_reserve0 --steps 5 --warmup 3 --fields-per-exchange 1 --exchange nccl --sm-reserve 0
  run k1_hybrid --steps 5 --warmup 3 --fields-per-exchange 1 --exchange hybrid
  run k1_nccl_b16 --steps 5 --warmup 3 --fields-per-exchange 1 --exchange nccl --buckets 16
fi
