# refresh the committed measurements of a round: default bench, reference arm, headline size, launch list, full ncu capture
python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -c 300 gpurun_out/bench_c2.err
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
python bench.py --steps 2 --warmup 3 --no-cpu --workload headline --in-place > gpurun_out/bench_headline.json 2> gpurun_out/bench_headline.err; tail -c 300 gpurun_out/bench_headline.err
python bench.py --steps 5 --warmup 3 --workload config4 > gpurun_out/bench_config4.json 2> gpurun_out/bench_config4.err; tail -c 300 gpurun_out/bench_config4.err
CMD="python bench.py --steps 1 --warmup 1 --no-cpu --no-host-object --shape 1,2048,2048,8"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 0 -c 160 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 21 -c 6 -o gpurun_out/prof_sweep -f $CMD > gpurun_out/ncu2.log 2>&1
tail -2 gpurun_out/ncu2.log
