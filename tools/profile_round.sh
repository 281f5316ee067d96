#!/bin/bash
# Refresh the committed measurements of a round (run under gpurun, ONE GPU): default bench line (incl. the 4096^2 x 512
# headline record and the cuFFT comparison), reference arm, launch list and one full ncu capture of the sweep kernels at
# 2048^2 and at 4096^2.  Summaries are made afterwards on the CPU box with tools/ncu_summary.py.
R=${1:-r02}
python bench.py > gpurun_out/${R}_bench_default.json 2> gpurun_out/${R}_bench_default.err; tail -c 300 gpurun_out/${R}_bench_default.err
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/${R}_bench_reference_arm.json 2> gpurun_out/${R}_bench_reference_arm.err
python bench.py --steps 5 --warmup 3 --workload config4 > gpurun_out/${R}_bench_config4.json 2> gpurun_out/${R}_bench_config4.err; tail -c 300 gpurun_out/${R}_bench_config4.err
CMD="python bench.py --steps 1 --warmup 1 --no-cpu --no-host-object --no-cufft --no-headline --shape 1,2048,2048,8"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 0 -c 200 --csv --log-file gpurun_out/${R}_launches_sweep_2048x8.csv $CMD > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 21 -c 6 -o gpurun_out/${R}_prof_sweep_2048 -f $CMD > gpurun_out/ncu2.log 2>&1
CMD4="python bench.py --steps 1 --warmup 1 --no-cpu --no-host-object --no-cufft --no-headline --shape 1,4096,4096,8"
$CMD4 > gpurun_out/plain4.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 21 -c 6 -o gpurun_out/${R}_prof_sweep_4096 -f $CMD4 > gpurun_out/ncu3.log 2>&1
tail -2 gpurun_out/ncu2.log gpurun_out/ncu3.log
