python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for pf in 1 0; do
for shape in 1,4096,4096,48 1,2048,2048,128 4,2048,2048,32; do
    echo "== slab_prefetch=$pf shape=$shape"
    BDOF_SLAB_PREFETCH=$pf python bench.py --steps 3 --warmup 3 --no-cpu --shape $shape 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value %.2f'%d['value'], {k:round(v['avg_ms']*1e3,1) for k,v in d['kernels'].items()})
    else: print(l.strip()[:300])
"
done
done
BDOF_LIB=libbdof_alt9.so python tools/phase_sweep.py 4096
BDOF_LIB=libbdof_alt9.so python tools/phase_sweep.py 2048
