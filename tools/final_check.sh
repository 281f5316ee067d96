#!/bin/bash
# End-of-round verification on one GPU: the whole GPU suite, the smoke entry, the config-4 line and the default bench line.
timeout 1300 python -m pytest tests -q -m gpu 2>&1 | tail -6 > gpurun_out/r02_final_suite.txt; cat gpurun_out/r02_final_suite.txt
python __graft_entry__.py --smoke > gpurun_out/r02_smoke.txt 2>&1; tail -1 gpurun_out/r02_smoke.txt
python bench.py --workload config4 > gpurun_out/r02_bench_config4_1gpu_v4.json 2>gpurun_out/c4.err
python -c "import json; d=json.loads(open('gpurun_out/r02_bench_config4_1gpu_v4.json').read().strip().splitlines()[-1]); print('config4', d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['forward_ms'], d['roofline']['adjoint_ms'])"
python bench.py > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err; cut -c1-260 gpurun_out/r02_bench_default.json
python tools/launch_probe.py 40 128 128 | tail -1
BDOF_CLUSTER=0 python tools/launch_probe.py 40 128 128 | tail -1
