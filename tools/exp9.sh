for shape in 1,4096,4096,48 1,2048,2048,128; do
    echo "== shape=$shape"
    python bench.py --steps 3 --warmup 3 --no-cpu --shape $shape 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value %.2f'%d['value'], {k:round(v['avg_ms']*1e3,1) for k,v in d['kernels'].items()})
    else: print(l.strip()[:300])
"
done
CMD="python bench.py --steps 1 --warmup 1 --no-cpu --shape 1,4096,4096,4"
ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 10 -c 1 -o gpurun_out/prof_sweep4k -f $CMD > gpurun_out/ncu3.log 2>&1
tail -2 gpurun_out/ncu3.log
